"""Seed-based known-answer vectors for the scoring functions.

Recorded from the unmodified reference (sidekit/iv_scoring.py) in the build
container; tests/test_oracle_vs_reference.py re-derives them from the reference
whenever it is present.
"""
import numpy


def kat_inputs():
    rng = numpy.random.default_rng(1234)
    en = rng.standard_normal((7, 8))
    te = rng.standard_normal((5, 8))
    mu = rng.standard_normal(8)
    F = rng.standard_normal((8, 4))
    A = rng.standard_normal((8, 8))
    Sigma = A @ A.T + numpy.eye(8)
    en_ids = numpy.array(["m%d" % i for i in range(7)])
    te_ids = numpy.array(["s%d" % i for i in range(5)])
    ndx_models = numpy.array(["m3", "m0", "mX", "m6", "m1"])
    ndx_segs = numpy.array(["s4", "sY", "s0", "s2"])
    trialmask = (numpy.arange(20).reshape(5, 4) % 3 != 0)
    return dict(en=en, te=te, mu=mu, F=F, Sigma=Sigma, en_ids=en_ids, te_ids=te_ids,
                ndx_models=ndx_models, ndx_segs=ndx_segs, trialmask=trialmask)


KAT_MODELSET = ["m3", "m0", "m6", "m1"]
KAT_SEGSET = ["s4", "s0", "s2"]
KAT_MASK_ROW0 = [False, True, False]
KAT_MASK_SUM = 6
KAT = {
    "plda": dict(row0=[0.5211382104, 0.3902799377, 1.2977190835], total=10.439408472551541, dtype="float64"),
    "plda_sf05": dict(row0=[0.0274658644, -0.037963272, 0.4157563009], total=2.422465346563679, dtype="float64"),
    "twocov": dict(row0=[-1.8489118997, -2.4999501581, -0.2869733666], total=-14.617168376538972, dtype="float64"),
    "cosine": dict(row0=[0.41571918, 0.021552714, -0.10463461], total=1.7147858142852783, dtype="float32"),
}

# mahalanobis_scoring with m = inv(Sigma) + 0.05 * triu(F F') (not symmetric): row 0 of the (4, 3) matrix
KAT_MAHALANOBIS_ROW0 = [-3.6858964595, -5.0085152223, -1.9150206526]
