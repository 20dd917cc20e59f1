"""Score GEMM time against the number of enrol rows (what a rank sees when the 20k x 20k matrix is row-sharded)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import sidekit_b200 as sk
from sidekit_b200 import synth
from sidekit_b200.iv_scoring import PackedEmbeddings

dev = torch.device("cuda", 0)
Nt, D = 20000, 256
T = torch.from_numpy(synth.synth_embeddings(Nt, D, seed=7)).float().to(dev)
Tp = PackedEmbeddings(T)
q = torch.randn(Nt, device=dev)
for Ne in (20000, 10000, 5000, 2500):
    E = torch.from_numpy(synth.synth_embeddings(Ne, D, seed=6)).float().to(dev)
    r = torch.randn(Ne, device=dev)
    out = torch.empty((Ne, Nt), dtype=torch.float32, device=dev)
    for name, tt in (("packed", Tp), ("plain", T)):
        g = lambda i: sk.score_matrix(E, tt, r, q, cst=0.5, alpha=1.0, passes=0, out=out)
        for i in range(5):
            g(i)
        ms = bench.timed(g, 50, False) / 50
        print("Ne %5d %-6s %.4f ms  %.0f GB/s  %.2e trials/s" % (Ne, name, ms, Ne * Nt * 4 / ms / 1e6, Ne * Nt / ms * 1e3))
print("---- column split: all enrol rows against a slice of the test columns")
E = torch.from_numpy(synth.synth_embeddings(20000, D, seed=6)).float().to(dev)
r = torch.randn(20000, device=dev)
for Ntc in (20000, 10000, 5000, 2500):
    Tc = T[:Ntc].contiguous()
    Tpc = PackedEmbeddings(Tc)
    qc = q[:Ntc].contiguous()
    out = torch.empty((20000, Ntc), dtype=torch.float32, device=dev)
    g = lambda i: sk.score_matrix(E, Tpc, r, qc, cst=0.5, alpha=1.0, passes=0, out=out)
    for i in range(5):
        g(i)
    ms = bench.timed(g, 50, False) / 50
    print("Nt %5d packed %.4f ms  %.0f GB/s  %.2e trials/s" % (Ntc, ms, 20000 * Ntc * 4 / ms / 1e6, 20000 * Ntc / ms * 1e3))
