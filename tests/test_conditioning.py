"""Embedding conditioning of ``StatServer`` against the outputs of the real reference
(tests/golden/conditioning.npz): covariances, LDA, WCCN, Mahalanobis, whitening, spectral normalisation, sums per model.
Host (numpy float64) code: the reference's per-speaker loops are vectorised, so agreement is to summation order."""
import copy

import numpy
import pytest

import sidekit_b200 as sk
from tests.helpers import golden


def _ss():
    g = golden("plda_training.npz")
    return sk.StatServer.from_embeddings(g["ids"], g["X"])


def _close(a, b, tol=1e-10):
    return numpy.abs(numpy.asarray(a) - numpy.asarray(b)).max() < tol * max(1.0, numpy.abs(b).max())


def test_covariances_lda_wccn_match_reference():
    g, ss = golden("conditioning.npz"), _ss()
    assert _close(ss.get_mean_stat1(), g["mean"]) and _close(ss.get_total_covariance_stat1(), g["total_cov"])
    assert _close(ss.get_within_covariance_stat1(), g["within_cov"]) and _close(ss.get_between_covariance_stat1(), g["between_cov"])
    assert _close(ss.get_mahalanobis_matrix_stat1(), g["mahalanobis"], 1e-8) and _close(ss.get_wccn_choleski_stat1(), g["wccn"], 1e-8)
    lda = ss.get_lda_matrix_stat1(5)
    sign = numpy.sign((lda * g["lda"]).sum(axis=0))              # eigenvectors are defined up to their sign
    assert lda.shape == g["lda"].shape and _close(lda * sign, g["lda"], 1e-6)
    summed, sessions = ss.sum_stat_per_model()
    assert numpy.array_equal(summed.modelset, g["sum_modelset"]) and _close(summed.stat1, g["sum_stat1"])
    assert numpy.array_equal(sessions, g["sum_sessions"]) and summed.validate()


def test_whitening_and_spectral_norm_match_reference():
    g = golden("conditioning.npz")
    w = _ss(); w.whiten_stat1(g["mean"], g["total_cov"]); assert _close(w.stat1, g["whiten_full"], 1e-9)
    w = _ss(); w.whiten_stat1(g["mean"], numpy.diag(g["total_cov"]).copy()); assert _close(w.stat1, g["whiten_diag"])
    w = _ss(); w.whiten_cholesky_stat1(g["mean"], g["total_cov"]); assert _close(w.stat1, g["whiten_chol"], 1e-9)
    for mode in ("efr", "sphNorm"):
        ss = _ss()
        before = ss.stat1.copy()
        means, covs = ss.estimate_spectral_norm_stat1(2, mode)
        assert numpy.array_equal(ss.stat1, before)                # estimation works on a copy
        # iteration 1 is unambiguous; from iteration 2 on the statistics live in the eigenbasis of the previous covariance,
        # whose vectors LAPACK returns up to a sign (a 1e-16 difference in the within-class covariance can flip one), so
        # those quantities agree up to per-coordinate signs
        assert _close(means[0], g["sn_%s_mean" % mode][0], 1e-9) and _close(covs[0], g["sn_%s_cov" % mode][0], 1e-9)
        assert _close(numpy.abs(means[1]), numpy.abs(g["sn_%s_mean" % mode][1]), 1e-8)
        assert _close(numpy.abs(covs[1]), numpy.abs(g["sn_%s_cov" % mode][1]), 1e-8)
        ss.spectral_norm_stat1(means, covs)
        ref = g["sn_%s_out" % mode]
        sign = numpy.sign((ss.stat1 * ref).sum(axis=0))
        assert _close(ss.stat1 * sign, ref, 1e-8)
        assert _close(ss.stat1 @ ss.stat1.T, ref @ ref.T, 1e-8)   # what every scorer sees (inner products) is identical
        assert numpy.allclose(numpy.linalg.norm(ss.stat1, axis=1), 1.0)


@pytest.mark.reference
def test_statserver_accessors_and_merge_match_the_reference():
    """statserver.py:337-388 (merge) and :555-654 (accessors) on the same sessions; merge is compared as a set of sessions
    because the reference's row order is the iteration order of a Python set."""
    import numpy
    import sidekit_b200 as sk
    from oracle import ref_import
    sidekit = ref_import.import_reference()
    rng = numpy.random.default_rng(3)

    def make(cls, models, segs, X):
        s = cls()
        s.modelset, s.segset = numpy.array(models, dtype="|O"), numpy.array(segs, dtype="|O")
        s.start, s.stop = numpy.empty(len(models), dtype="|O"), numpy.empty(len(models), dtype="|O")
        s.stat0, s.stat1 = numpy.ones((len(models), 1)), numpy.array(X, dtype=numpy.float64)
        return s

    Xa, Xb = rng.standard_normal((5, 4)), rng.standard_normal((4, 4))
    ma, sa = ["b", "a", "b", "c", "a"], ["s0", "s1", "s2", "s3", "s4"]
    mb, sb = ["a", "d", "b", "d"], ["s1", "s5", "s6", "s7"]                     # ("a", "s1") repeats a session of the first
    mine_a, mine_b = make(sk.StatServer, ma, sa, Xa), make(sk.StatServer, mb, sb, Xb)
    ref_a, ref_b = make(sidekit.StatServer, ma, sa, Xa), make(sidekit.StatServer, mb, sb, Xb)
    for k in range(3):
        assert numpy.array_equal(mine_a.get_model_stat1_by_index(k), ref_a.get_model_stat1_by_index(k))
        assert numpy.array_equal(mine_a.get_model_stat0_by_index(k), ref_a.get_model_stat0_by_index(k))
    assert numpy.array_equal(mine_a.get_segment_stat1("s3"), ref_a.get_segment_stat1("s3"))
    assert numpy.array_equal(mine_a.get_segment_stat0("s3"), ref_a.get_segment_stat0("s3"))
    assert numpy.array_equal(mine_a.get_segment_stat1_by_index(2), ref_a.get_segment_stat1_by_index(2))
    assert numpy.array_equal(mine_a.get_segment_stat0_by_index(2), ref_a.get_segment_stat0_by_index(2))
    assert numpy.array_equal(mine_a.get_model_segments("b"), ref_a.get_model_segments("b"))
    assert list(mine_a.get_model_segments_by_index(0)) == ["s1", "s4"]
    with pytest.raises(IndexError):
        ref_a.get_model_segments_by_index(0)
    mine = sk.StatServer.merge(mine_a, mine_b)
    ref = sidekit.StatServer.merge(ref_a, ref_b)
    assert mine.validate() and mine.modelset.shape[0] == ref.modelset.shape[0] == 8
    key = lambda s: sorted((m, g, tuple(x)) for m, g, x in zip(s.modelset.tolist(), s.segset.tolist(), s.stat1.tolist()))
    assert key(mine) == key(ref)
    assert list(mine.segset[:5]) == sa and numpy.array_equal(mine.stat1[1], Xb[0])          # first-occurrence order, the later session wins
