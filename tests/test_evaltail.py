"""Evaluation tail: native PAV / ROCCH / EER and the BOSARIS containers against the fixtures
recorded from the real reference (tests/golden/evaltail.npz, oracle/make_golden.py) and against the numpy oracle;
z-/t-/zt-norm (CUDA) in the GPU part."""
import copy

import numpy
import pytest

import sidekit_b200 as sk
from oracle import eval_ref as E
from tests.helpers import golden

CASES = ("sep", "overlap", "ties", "tiny")


def test_pavx_matches_reference_bit_for_bit():
    g = golden("evaltail.npz")
    ghat, width, height = sk.pavx(g["pav_y"])
    assert numpy.array_equal(ghat, g["pav_ghat"])            # including the reference's wrap-around write into ghat[-1]
    assert numpy.array_equal(width, g["pav_width"]) and numpy.array_equal(height, g["pav_height"])
    og, ow, oh = E.pavx(g["pav_y"])
    assert numpy.array_equal(og, ghat) and numpy.array_equal(ow, width) and numpy.array_equal(oh, height)
    with pytest.raises(AssertionError):
        sk.pavx(numpy.zeros((2, 2)))
    with pytest.raises(AssertionError):
        sk.pavx(numpy.zeros(0))


@pytest.mark.parametrize("case", CASES)
def test_rocch_eer_mindcf_match_reference(case):
    g = golden("evaltail.npz")
    tar, non = g[case + "_tar"], g[case + "_non"]
    pm, pf = sk.rocch(tar, non)
    assert numpy.array_equal(pm, g[case + "_pmiss"]) and numpy.array_equal(pf, g[case + "_pfa"])     # bit-exact hull
    assert sk.rocch2eer(pm, pf) == g[case + "_rocch_eer"]
    assert numpy.array_equal(numpy.array(sk.fast_minDCF(tar, non, -2.0, normalize=True), dtype=numpy.float64), g[case + "_mindcf"])
    assert sk.eer(non, tar) == g[case + "_eer"]
    from sidekit_b200.nnet.xvector import eer as eer_xv       # the reference's import path
    assert eer_xv(non, tar) == g[case + "_eer"]


def test_rocch_against_oracle_on_random_and_degenerate_inputs():
    rng = numpy.random.default_rng(5)
    for nt, nn, digits in ((1, 1, 3), (5, 3, 0), (257, 1031, 1), (2000, 7000, 6), (50, 50, 0)):
        tar = numpy.round(rng.normal(0.5, 1.0, nt), digits)
        non = numpy.round(rng.normal(0.0, 1.0, nn), digits)
        pm, pf = sk.rocch(tar, non)
        opm, opf = E.rocch(tar, non)
        assert numpy.array_equal(pm, opm) and numpy.array_equal(pf, opf)
        assert pm[0] == 0.0 and pf[0] == 1.0 and pm[-1] == 1.0 and pf[-1] == 0.0
        assert numpy.all(numpy.diff(pm) >= 0) and numpy.all(numpy.diff(pf) <= 0)          # monotone hull
    # perfectly separated and fully reversed score sets
    assert sk.rocch2eer(*sk.rocch(numpy.array([3., 4.]), numpy.array([1., 2.]))) == 0
    assert sk.rocch2eer(*sk.rocch(numpy.array([1., 2.]), numpy.array([3., 4.]))) == 0.5
    with pytest.raises(RuntimeError):
        sk.rocch(numpy.array([]), numpy.array([1.0]))


def test_key_and_get_tar_non_match_reference():
    g = golden("evaltail.npz")
    key = sk.Key(models=g["key_models"], testsegs=g["key_segs"], trials=g["key_trials"])
    assert numpy.array_equal(key.modelset, g["key_modelset"]) and numpy.array_equal(key.segset, g["key_segset"])
    assert numpy.array_equal(key.tar, g["key_tar"]) and numpy.array_equal(key.non, g["key_non"])
    sc = sk.Scores()
    sc.modelset, sc.segset, sc.scoremat, sc.scoremask = g["sc_modelset"], g["sc_segset"], g["sc_mat"], g["sc_mask"]
    tar, non = sc.get_tar_non(key)                            # different id sets and order -> align_with_ndx path
    assert numpy.array_equal(tar, g["sc_tar"]) and numpy.array_equal(non, g["sc_non"])
    al = sc.align_with_ndx(key.to_ndx())
    assert al.validate() and al.scoremat.shape == key.tar.shape
    al2 = al.align_with_ndx(key)                              # already aligned: same ids -> direct masks
    t2, n2 = al2.get_tar_non(key)
    assert numpy.array_equal(t2, tar) and numpy.array_equal(n2, non)


def _scores(g, nm):
    s = sk.Scores()
    s.modelset, s.segset, s.scoremat = g["zt_%s_modelset" % nm].copy(), g["zt_%s_segset" % nm].copy(), g["zt_%s_mat" % nm].copy()
    s.scoremask = numpy.ones(s.scoremat.shape, dtype=bool)
    return s


@pytest.mark.gpu
def test_znorm_tnorm_ztnorm_match_reference():
    g = golden("evaltail.npz")
    et, ei, it, ii = (_scores(g, n) for n in ("et", "ei", "it", "ii"))
    z = sk.znorm(copy.deepcopy(et), copy.deepcopy(ei))
    assert numpy.allclose(z.scoremat, g["znorm_mat"], rtol=1e-11, atol=1e-12)
    zs = sk.znorm(copy.deepcopy(it), copy.deepcopy(ii), sym=True)
    assert numpy.allclose(zs.scoremat, g["znorm_sym_mat"], rtol=1e-11, atol=1e-12)
    t = sk.tnorm(copy.deepcopy(et), copy.deepcopy(it))
    assert numpy.allclose(t.scoremat, g["tnorm_mat"], rtol=1e-11, atol=1e-12)
    zt = sk.ztnorm(copy.deepcopy(et), copy.deepcopy(ei), copy.deepcopy(it), copy.deepcopy(ii))
    assert numpy.allclose(zt.scoremat, g["ztnorm_mat"], rtol=1e-10, atol=1e-11)
    assert numpy.array_equal(zt.modelset, g["ztnorm_modelset"]) and numpy.array_equal(zt.segset, g["ztnorm_segset"])
    # the oracle restatement on the sorted matrices agrees too, and the reference's shape quirk is kept
    es, eis = copy.deepcopy(et), copy.deepcopy(ei)
    es.sort(); eis.sort()
    assert numpy.allclose(E.znorm_matrix(es.scoremat, eis.scoremat), z.scoremat, rtol=1e-11, atol=1e-12)
    wide = copy.deepcopy(et)
    wide.segset, wide.scoremat, wide.scoremask = wide.segset[:5].copy(), wide.scoremat[:, :5].copy(), wide.scoremask[:, :5].copy()
    with pytest.raises(ValueError):
        sk.znorm(wide, copy.deepcopy(ei))                     # (12, 5) - (12,): numpy cannot broadcast either
    # float32 score matrices (cosine scores) are normalised in float32
    et32, it32 = copy.deepcopy(et), copy.deepcopy(it)
    et32.scoremat, it32.scoremat = et32.scoremat.astype(numpy.float32), it32.scoremat.astype(numpy.float32)
    t32 = sk.tnorm(et32, it32)
    assert t32.scoremat.dtype == numpy.float32 and numpy.allclose(t32.scoremat, g["tnorm_mat"], rtol=2e-5, atol=2e-5)
