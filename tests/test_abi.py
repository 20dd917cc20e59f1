"""CPU checks of the boundary: the library loads, exports every symbol include/sidekit_b200.h declares,
the Python layer mirrors the reference's names / state_dict keys, and the product never imports the oracle."""
import ctypes
import os
import re

import numpy
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from sidekit_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "sidekit_b200.h")).read()
    names = set(re.findall(r"\b(skb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert _lib.lib().skb_version() == 100


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sidekit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_no_cuda_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from tests.models import make_xtractor
    m = make_xtractor("halfresnet34", 8, 256)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 16000), is_eval=True)
    import sidekit_b200 as sk
    en = sk.StatServer.from_embeddings(numpy.array(["a", "b"]), numpy.eye(2, 4))
    ndx = sk.Ndx()
    ndx.modelset, ndx.segset, ndx.trialmask = numpy.array(["a"]), numpy.array(["b"]), numpy.ones((1, 1), bool)
    with pytest.raises(RuntimeError, match="CUDA"):
        sk.cosine_scoring(en, en, ndx)


@pytest.mark.reference
@pytest.mark.parametrize("archi,emb", [("halfresnet34", 256), ("xvector", 512)])
def test_state_dict_contract_matches_reference(archi, emb):
    from oracle import ref_import
    from tests.models import make_xtractor
    ref_model = ref_import.build_xtractor(32, archi, emb)
    ref = ref_model.state_dict()
    mine = make_xtractor(archi, 32, emb)
    sd = mine.state_dict()
    assert list(ref.keys()) == list(sd.keys())
    for k in ref:
        assert ref[k].shape == sd[k].shape, k
        if k.startswith("preprocessor."):
            assert torch.equal(ref[k], sd[k]), k          # window / filterbank / DCT bit-identical to torchaudio's
    mine.load_state_dict(ref, strict=True)
    assert mine.context_size() == ref_model.context_size() == (15 if archi == "xvector" else 3)


def test_index_semantics_host():
    """Ndx.filter / align_* / mean_stat_per_model against the oracle's restatement (duplicates, missing ids)."""
    import sidekit_b200 as sk
    from oracle import scoring_ref as S
    rng = numpy.random.default_rng(0)
    en_ids = numpy.array(["m%d" % i for i in rng.integers(0, 12, 20)])       # duplicates
    te_ids = numpy.array(["s%d" % i for i in rng.permutation(15)])
    E, T = rng.standard_normal((20, 4)), rng.standard_normal((15, 4))
    ndx = sk.Ndx()
    ndx.modelset = numpy.array(["m3", "m99", "m0", "m3", "m7"])
    ndx.segset = numpy.array(["s4", "sX", "s0"])
    ndx.trialmask = rng.random((5, 3)) < 0.5
    en = sk.StatServer.from_embeddings(en_ids, E).mean_stat_per_model()
    uniq, Em = S.mean_per_model(en_ids, E)
    assert numpy.array_equal(en.modelset, uniq) and numpy.allclose(en.stat1, Em, atol=1e-14)
    te = sk.StatServer.from_embeddings(te_ids, T)
    from sidekit_b200.iv_scoring import _check_missing_model
    clean = _check_missing_model(en, te, ndx)
    present = set(en_ids.tolist())
    models, segs, mask, ri, ci = S.check_missing(uniq, te_ids, ndx.modelset, ndx.segset, ndx.trialmask)
    assert numpy.array_equal(clean.modelset, models) and numpy.array_equal(clean.segset, segs)
    assert numpy.array_equal(clean.trialmask, mask)
    assert numpy.allclose(en.stat1, Em[ri]) and numpy.allclose(te.stat1, T[ci])
    assert all(m in present for m in clean.modelset)
