"""Host LAPACK primitives on 256 x 256 float64 (the D x D algebra of the PLDA scorers): which calls are slow on this box?"""
import os, sys, time
import numpy, scipy.linalg
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sidekit_b200 import synth
mu, F, Sigma = synth.synth_plda(256, 256, seed=3)
B = F @ F.T
T = B + Sigma
N = T + 1e-9 * numpy.random.default_rng(0).standard_normal(T.shape)      # not exactly symmetric
def t(name, f, n=9):
    f(); ts = []
    for i in range(n):
        a = time.perf_counter(); f(); ts.append(time.perf_counter() - a)
    print("%-44s min %.4f median %.4f max %.4f" % (name, min(ts), sorted(ts)[n // 2], max(ts)))
t("scipy.linalg.inv(exactly symmetric)", lambda: scipy.linalg.inv(T))
t("scipy.linalg.inv(not exactly symmetric)", lambda: scipy.linalg.inv(N))
t("numpy.linalg.inv(exactly symmetric)", lambda: numpy.linalg.inv(T))
t("numpy.linalg.inv(not exactly symmetric)", lambda: numpy.linalg.inv(N))
t("scipy.linalg.solve(Sigma, F)", lambda: scipy.linalg.solve(Sigma, F))
t("numpy.linalg.solve(Sigma, F)", lambda: numpy.linalg.solve(Sigma, F))
t("scipy.linalg.cholesky", lambda: scipy.linalg.cholesky(T, lower=True))
t("numpy.linalg.cholesky", lambda: numpy.linalg.cholesky(T))
t("numpy.linalg.slogdet", lambda: numpy.linalg.slogdet(T))
t("matmul 256^3", lambda: B @ T)
