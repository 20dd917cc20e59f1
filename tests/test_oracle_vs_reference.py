"""The oracle restatement against the REAL reference, imported from /root/reference (build container only;
skipped on the GPU box where the reference does not exist)."""
import copy

import numpy
import pytest
import torch

from oracle import extract_ref as R, scoring_ref as S
from sidekit_b200 import synth
from tests import kat

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("archi,emb,L", [("halfresnet34", 256, 20000), ("xvector", 512, 40000), ("resnet34", 256, 12000),
                                         ("fastresnet34", 256, 20000)])
def test_extraction_restatement_matches_reference(archi, emb, L):
    from oracle import ref_import
    m = ref_import.build_xtractor(32, archi, emb)
    sd = m.state_dict()
    synth.fill_state_dict(sd, 0)
    m.load_state_dict(sd)
    x = synth.synth_wave(2, L, seed=9)
    with torch.no_grad():
        lo, em = m(x, is_eval=True)
    lo2, em2 = R.forward(sd, x, archi)
    assert (em - em2).abs().max().item() < 2e-6 and (lo - lo2).abs().max().item() < 5e-5
    fb = R.mel_filterbank(513, 90.0, 7600.0, 80, 16000) if archi != "xvector" else R.mel_filterbank(1025, 133.333, 6855.4976, 100, 16000)
    key = "preprocessor.MelSpec.mel_scale.fb" if archi != "xvector" else "preprocessor.MFCC.MelSpectrogram.mel_scale.fb"
    assert torch.equal(fb, sd[key])


def test_resblock_is_unreachable_in_the_reference():
    """`ResBlock` is only instantiated by the YAML / dict branch of Xtractor.__init__ (xvector.py:786-790), and
    that branch cannot construct a model: a `conv2D` entry hits an undefined name (:763) and a `resblock` entry appends
    a bare module to the list an OrderedDict is built from (:786, :792).  No model on the hot path contains a ResBlock;
    the B200 engine therefore has no ResBlock operator (out of scope)."""
    import contextlib, io
    from oracle import ref_import
    ref_import.import_reference()
    from sidekit.nnet.xvector import Xtractor
    base = {"loss": {"type": "aam", "aam_margin": 0.2, "aam_s": 30}, "preprocessor": {"type": "none", "feature_size": 80},
            "activation": "ReLU", "stat_pooling": {"type": "mean_std", "weight_decay": 0.0},
            "before_embedding": {"weight_decay": 0.0}, "embedding": {}, "after_embedding": {"weight_decay": 0.0}}
    res = {"resblock1": {"input_channel": 32, "output_channel": 32}, "weight_decay": 0.0}
    for seg, err in ((dict(res), TypeError), (dict({"conv2D": {}}, **res), NameError)):
        with pytest.raises(err), contextlib.redirect_stdout(io.StringIO()):
            Xtractor(10, dict(base, segmental=seg))


def test_resblock_restatement_matches_the_reference_class():
    """The module itself IS constructible (only the Xtractor branch that would use it is not): the oracle's `res_block`
    follows the reference's `ResBlock.forward` (res_net.py:229-255) for the identity and the widening variant."""
    from oracle import ref_import, extract_ref as R
    from sidekit_b200 import synth
    ref_import.import_reference()
    from sidekit.nnet.res_net import ResBlock
    for cin, cout, first in ((32, 32, False), (32, 64, False), (16, 16, True)):
        torch.manual_seed(cin + cout)
        blk = ResBlock(cin, cout, 1, is_first=first).eval()
        sd = blk.state_dict()
        synth.fill_state_dict(sd, 3)
        blk.load_state_dict(sd)
        x = torch.randn(2, cin, 9, 7)
        with torch.no_grad():
            ref = blk(x.clone())
        got = R.res_block({"b." + k: v for k, v in sd.items()}, "b", x, is_first=first)
        assert (got - ref).abs().max().item() < 1e-5


def test_scoring_kat_recorded_from_reference():
    """Re-derive SURVEY.md Appendix B from the unmodified reference and compare with tests/kat.py and the oracle."""
    from oracle import ref_import
    sidekit = ref_import.import_reference()
    from sidekit.iv_scoring import cosine_scoring, PLDA_scoring, two_covariance_scoring, mahalanobis_scoring
    k = kat.kat_inputs()

    def ss(ids, X):
        s = sidekit.StatServer()
        s.modelset = numpy.array(ids); s.segset = numpy.array(ids)
        s.start = numpy.empty(len(ids), dtype="|O"); s.stop = numpy.empty(len(ids), dtype="|O")
        s.stat0 = numpy.ones((len(ids), 1)); s.stat1 = numpy.array(X, dtype=numpy.float64)
        return s

    def ndx():
        n = sidekit.Ndx()
        n.modelset, n.segset, n.trialmask = k["ndx_models"].copy(), k["ndx_segs"].copy(), k["trialmask"].copy()
        return n

    a = (k["en_ids"], k["en"], k["te_ids"], k["te"], k["ndx_models"], k["ndx_segs"], k["trialmask"])
    B = k["F"] @ k["F"].T + 0.1 * numpy.eye(8)
    cases = {
        "plda": (PLDA_scoring(ss(k["en_ids"], k["en"]), ss(k["te_ids"], k["te"]), ndx(), k["mu"], k["F"], numpy.zeros((8, 0)), k["Sigma"]),
                 S.fast_plda_scoring(*a, k["mu"], k["F"], k["Sigma"])),
        "twocov": (two_covariance_scoring(ss(k["en_ids"], k["en"]), ss(k["te_ids"], k["te"]), ndx(), k["Sigma"], B),
                   S.two_covariance_scoring(*a, k["Sigma"], B)),
        "cosine": (cosine_scoring(ss(k["en_ids"], k["en"]), ss(k["te_ids"], k["te"]), ndx(), device=torch.device("cpu")),
                   S.cosine_scoring(*a)),
    }
    # mahalanobis_scoring (iv_scoring.py:116-156); pinned by tests/kat.py too
    M = numpy.linalg.inv(k["Sigma"]) + 0.05 * numpy.triu(k["F"] @ k["F"].T)          # not symmetric on purpose
    ref = mahalanobis_scoring(ss(k["en_ids"], k["en"]), ss(k["te_ids"], k["te"]), ndx(), M)
    mine = S.mahalanobis_scoring(*a, M)
    assert ref.modelset.tolist() == mine[0].tolist() and ref.segset.tolist() == mine[1].tolist()
    assert numpy.array_equal(ref.scoremask, mine[2]) and ref.scoremat.shape == mine[3].shape
    assert numpy.abs(ref.scoremat - mine[3]).max() < 1e-9
    assert numpy.abs(ref.scoremat[0] - numpy.array(kat.KAT_MAHALANOBIS_ROW0)).max() < 1e-8 and ref.scoremat.shape == (4, 3)
    for name, (ref, mine) in cases.items():
        assert ref.modelset.tolist() == mine[0].tolist() == kat.KAT_MODELSET
        assert ref.segset.tolist() == mine[1].tolist() == kat.KAT_SEGSET
        assert numpy.array_equal(ref.scoremask, mine[2])
        assert ref.scoremat.dtype == mine[3].dtype
        tol = 1e-9 if ref.scoremat.dtype == numpy.float64 else 1e-6
        assert numpy.abs(ref.scoremat - mine[3]).max() < tol
        assert numpy.abs(ref.scoremat[0] - numpy.array(kat.KAT[name]["row0"])).max() < max(tol, 1e-9) * 10


def test_plda_scoring_uncertainty_is_unreachable_in_the_reference():
    """``PLDA_scoring`` hands ``test_uncertainty`` / ``Vtrans`` to ``fast_PLDA_scoring``, which ignores them
    (iv_scoring.py:256-268); ``PLDA_scoring_uncertainty`` itself asserts on an undefined name ``G`` (:513) and raises
    NameError before computing anything, so uncertainty propagation is not part of any runnable path."""
    from oracle import ref_import
    sidekit = ref_import.import_reference()
    from sidekit.iv_scoring import PLDA_scoring_uncertainty
    k = kat.kat_inputs()
    s = sidekit.StatServer()
    s.modelset = numpy.array(k["en_ids"]); s.segset = numpy.array(k["en_ids"])
    s.start = numpy.empty(7, dtype="|O"); s.stop = numpy.empty(7, dtype="|O")
    s.stat0 = numpy.ones((7, 1)); s.stat1 = numpy.array(k["en"], dtype=numpy.float64)
    n = sidekit.Ndx()
    n.modelset, n.segset, n.trialmask = k["en_ids"][:2].copy(), k["en_ids"][:3].copy(), numpy.ones((2, 3), dtype=bool)
    with pytest.raises(NameError):
        PLDA_scoring_uncertainty(s, s, n, k["mu"], k["F"], k["Sigma"], test_uncertainty=numpy.ones((7, 8)), Vtrans=numpy.eye(8))
