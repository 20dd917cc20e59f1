"""CPU restatement (numpy / scipy, float64) of ``FactorAnalyser.plda`` (sidekit/factor_analyser.py:830-932 with the
per-class loop of ``fa_model_loop``, :166-205, and the StatServer helpers statserver.py:789-795, :852-884, :920-928,
:1335-1355).  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Pinned by tests/golden/plda_training.npz, recorded
from the imported reference by oracle/make_golden.py.
"""
import numpy
import scipy.linalg


def plda(ids, X, rank_f, nb_iter=10, scaling_factor=1.0):
    X = numpy.asarray(X, dtype=numpy.float64)
    D = X.shape[1]
    mean = X.mean(axis=0)
    C0 = X - X.mean(axis=0)
    sigma_obs = C0.T.dot(C0) / X.shape[0]
    Sigma = sigma_obs.copy()
    models = numpy.unique(ids)
    S0 = numpy.zeros((models.shape[0], 1))
    S1 = numpy.zeros((models.shape[0], D))
    sessions = numpy.zeros(models.shape[0])
    for i, m in enumerate(models):                      # sum_stat_per_model
        sel = ids == m
        S0[i, 0] = sel.sum()
        S1[i] = X[sel].sum(axis=0)
        sessions[i] = sel.sum()
    S0 *= scaling_factor
    S1 *= scaling_factor
    sessions *= scaling_factor
    evals, evecs = scipy.linalg.eigh(sigma_obs)
    idx = numpy.argsort(evals)[::-1]
    F = evecs.real[:, idx[:rank_f]]
    for _ in range(nb_iter):
        ev, evec = scipy.linalg.eigh(Sigma)
        ind = ev.real.argsort()[::-1]
        sqr_inv_sigma = numpy.dot(evec.real[:, ind], numpy.diag(1 / numpy.sqrt(ev.real[ind])))
        local = (S1 - S0 * mean).dot(sqr_inv_sigma)      # whiten_stat1
        F = sqr_inv_sigma.T.dot(F)
        A0 = F.T.dot(F)
        e_h = numpy.zeros((models.shape[0], rank_f))
        e_hh = numpy.zeros((models.shape[0], rank_f, rank_f))
        for i in range(models.shape[0]):                 # fa_model_loop
            inv_lambda = scipy.linalg.inv(S0[i, 0] * A0 + numpy.eye(rank_f))
            e_h[i] = F.T.dot(local[i]).dot(inv_lambda)
            e_hh[i] = inv_lambda + numpy.outer(e_h[i], e_h[i])
        R = e_hh.sum(axis=0) / sessions.shape[0]
        Cm = e_h.T.dot(local).dot(scipy.linalg.inv(sqr_inv_sigma))
        A = numpy.einsum('ijk,i->jk', e_hh, S0.squeeze())
        F = scipy.linalg.solve(A, Cm).T
        Sigma = sigma_obs - F.dot(Cm) / sessions.sum()
        F = F.dot(scipy.linalg.cholesky(R))
    return mean, F, Sigma
