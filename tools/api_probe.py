import os, sys, time
import numpy, scipy.linalg, torch
sys.path.insert(0, os.getcwd())
import sidekit_b200 as sk
from sidekit_b200 import synth, iv_scoring as I
mu, F, Sigma = synth.synth_plda(256, 256, seed=3)
def old(F,Sigma,sf):
    invSigma = scipy.linalg.inv(Sigma); I_spk=numpy.eye(F.shape[1])
    K = F.T.dot(invSigma*sf).dot(F); K1=scipy.linalg.inv(K+I_spk); K2=scipy.linalg.inv(2*K+I_spk)
    cst = numpy.linalg.slogdet(K2)[1]/2.0 - numpy.linalg.slogdet(K1)[1]
    Sac=F@F.T; St=Sac+Sigma; Sti=scipy.linalg.inv(St); Tmp=numpy.linalg.inv(St-Sac.dot(Sti).dot(Sac)); Phi=Sti-Tmp; Psi=Sti.dot(Sac).dot(Tmp)
    return Phi, Psi, cst
for name, f in (("new algebra", lambda: I._simplified_plda_terms(F, Sigma, 1.0)), ("old algebra", lambda: old(F, Sigma, 1.0))):
    f(); ts = []
    for i in range(10):
        a = time.perf_counter(); f(); ts.append(time.perf_counter() - a)
    print(name, "min %.4f median %.4f max %.4f" % (min(ts), sorted(ts)[5], max(ts)))
N = 20000
E, T = synth.synth_embeddings(N, 256, seed=6), synth.synth_embeddings(N, 256, seed=7)
ids_e = numpy.array(["m%05d" % i for i in range(N)]); ids_t = numpy.array(["s%05d" % i for i in range(N)])
ndx = sk.Ndx(); ndx.modelset, ndx.segset = ids_e, ids_t; ndx.trialmask = numpy.ones((N, N), dtype=bool)
en, te = sk.StatServer.from_embeddings(ids_e, E), sk.StatServer.from_embeddings(ids_t, T)
import cProfile, pstats
for rep in range(3):
    torch.cuda.synchronize(); a = time.perf_counter()
    sc = sk.PLDA_scoring(en, te, ndx, mu, F, numpy.zeros((256, 0)), Sigma)
    b = time.perf_counter(); torch.cuda.synchronize(); c = time.perf_counter()
    print("PLDA_scoring call %.4f s (+ %.4f s to drain the device)" % (b - a, c - b))
pr = cProfile.Profile(); pr.enable()
sc = sk.PLDA_scoring(en, te, ndx, mu, F, numpy.zeros((256, 0)), Sigma)
pr.disable(); pstats.Stats(pr).sort_stats("cumtime").print_stats(14)
