"""The oracle's scoring restatement against the known-answer vectors of SURVEY.md Appendix B
and the golden fixtures recorded from the reference (tests/golden/scoring.npz)."""
import numpy
import pytest

from oracle import scoring_ref as S
from tests import kat
from tests.helpers import golden


def _args(k):
    return (k["en_ids"], k["en"], k["te_ids"], k["te"], k["ndx_models"], k["ndx_segs"], k["trialmask"])


@pytest.mark.parametrize("name", ["plda", "plda_sf05", "twocov", "cosine"])
def test_kat(name):
    k = kat.kat_inputs()
    if name == "plda":
        r = S.fast_plda_scoring(*_args(k), k["mu"], k["F"], k["Sigma"])
    elif name == "plda_sf05":
        r = S.fast_plda_scoring(*_args(k), k["mu"], k["F"], k["Sigma"], scaling_factor=0.5)
    elif name == "twocov":
        r = S.two_covariance_scoring(*_args(k), k["Sigma"], k["F"] @ k["F"].T + 0.1 * numpy.eye(8))
    else:
        r = S.cosine_scoring(*_args(k))
    models, segs, mask, mat = r
    exp = kat.KAT[name]
    assert models.tolist() == kat.KAT_MODELSET and segs.tolist() == kat.KAT_SEGSET
    assert mask.sum() == kat.KAT_MASK_SUM and mask[0].tolist() == kat.KAT_MASK_ROW0
    assert str(mat.dtype) == exp["dtype"]
    tol = 1e-9 if exp["dtype"] == "float64" else 1e-6
    numpy.testing.assert_allclose(mat[0], exp["row0"], atol=tol, rtol=0)
    numpy.testing.assert_allclose(mat.sum(), exp["total"], atol=10 * tol, rtol=0)


@pytest.mark.parametrize("name", ["cosine", "plda", "plda_sf", "plda_open", "twocov", "plda_dup"])
def test_golden_scoring(name):
    g = golden("scoring.npz")
    en_ids = g["en_ids_dup"] if name == "plda_dup" else g["en_ids"]
    a = (en_ids, g["E"], g["te_ids"], g["T"], g["ndx_models"], g["ndx_segs"], g["trialmask"])
    if name == "cosine":
        r = S.cosine_scoring(*a)
    elif name == "plda" or name == "plda_dup":
        r = S.fast_plda_scoring(*a, g["mu"], g["F"], g["Sigma"])
    elif name == "plda_sf":
        r = S.fast_plda_scoring(*a, g["mu"], g["F"], g["Sigma"], scaling_factor=0.7)
    elif name == "plda_open":
        r = S.fast_plda_scoring(*a, g["mu"], g["F"], g["Sigma"], p_known=0.3)
    else:
        r = S.two_covariance_scoring(*a, g["Sigma"], g["B"])
    assert numpy.array_equal(r[0], g[name + "_modelset"])
    assert numpy.array_equal(r[1], g[name + "_segset"])
    assert numpy.array_equal(r[2], g[name + "_mask"])
    assert r[3].dtype == g[name + "_mat"].dtype
    numpy.testing.assert_allclose(r[3], g[name + "_mat"], atol=1e-9 if r[3].dtype == numpy.float64 else 2e-6, rtol=0)


def test_golden_asnorm():
    g = golden("scoring.npz")
    out = S.asnorm(g["asnorm_X"], g["asnorm_cohort"])
    assert out.dtype == g["asnorm_out"].dtype
    numpy.testing.assert_allclose(out, g["asnorm_out"], atol=2e-5, rtol=0)
