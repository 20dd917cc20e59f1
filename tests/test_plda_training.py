"""PLDA training: oracle restatement and the device implementation against the model recorded
from the real reference's ``FactorAnalyser.plda`` (tests/golden/plda_training.npz)."""
import numpy
import pytest

import sidekit_b200 as sk
from oracle import plda_train_ref as P
from tests.helpers import golden

CASES = (("it3", dict(nb_iter=3)), ("it5_sf", dict(nb_iter=5, scaling_factor=0.5)))


@pytest.mark.parametrize("name,kw", CASES)
def test_oracle_matches_reference_fixture(name, kw):
    g = golden("plda_training.npz")
    mean, F, Sigma = P.plda(g["ids"], g["X"], 6, **kw)
    assert numpy.abs(mean - g[name + "_mean"]).max() < 1e-12
    assert numpy.abs(F - g[name + "_F"]).max() < 1e-9 and numpy.abs(Sigma - g[name + "_Sigma"]).max() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", CASES)
def test_plda_training_matches_reference(name, kw):
    g = golden("plda_training.npz")
    ss = sk.StatServer.from_embeddings(g["ids"], g["X"])
    fa = sk.FactorAnalyser()
    fa.plda(ss, 6, save_final=False, **kw)
    assert numpy.abs(fa.mean - g[name + "_mean"]).max() < 1e-12
    assert numpy.abs(fa.F - g[name + "_F"]).max() < 1e-8 and numpy.abs(fa.Sigma - g[name + "_Sigma"]).max() < 1e-8
    assert ss.stat1.shape == g["X"].shape                          # the caller's StatServer is untouched


@pytest.mark.gpu
def test_trained_plda_feeds_the_scorer():
    """train -> score round trip: a model trained here separates same-speaker from different-speaker trials."""
    rng = numpy.random.default_rng(3)
    D, R, n_spk = 32, 8, 60
    V = rng.standard_normal((R, D))
    spk = rng.standard_normal((n_spk, R)) @ V
    ids = numpy.repeat(numpy.array(["s%02d" % i for i in range(n_spk)]), 6)
    X = numpy.repeat(spk, 6, axis=0) + 0.7 * rng.standard_normal((n_spk * 6, D))
    fa = sk.FactorAnalyser().plda(sk.StatServer.from_embeddings(ids, X), R, nb_iter=5, save_final=False)
    en_ids = numpy.array(["s%02d" % i for i in range(n_spk)])
    enroll = sk.StatServer.from_embeddings(en_ids, X[0::6])
    test = sk.StatServer.from_embeddings(numpy.array(["t%02d" % i for i in range(n_spk)]), X[1::6])
    ndx = sk.Ndx()
    ndx.modelset, ndx.segset, ndx.trialmask = enroll.modelset, test.segset, numpy.ones((n_spk, n_spk), dtype=bool)
    sc = sk.PLDA_scoring(enroll, test, ndx, fa.mean, fa.F, numpy.zeros((D, 0)), fa.Sigma)
    key = sk.Key.create(ndx.modelset, ndx.segset, numpy.eye(n_spk, dtype=bool), ~numpy.eye(n_spk, dtype=bool))
    tar, non = sc.get_tar_non(key)
    assert sk.rocch2eer(*sk.rocch(tar, non)) < 0.05
