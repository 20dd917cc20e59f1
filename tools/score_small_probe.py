"""One scoring call per enrol size (for an ncu launch list): which kernels cost what when a rank holds 1/8 of the rows."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sidekit_b200 as sk
from sidekit_b200 import synth
from sidekit_b200.iv_scoring import PackedEmbeddings
dev = torch.device("cuda", 0)
Nt, D = 20000, 256
T = torch.from_numpy(synth.synth_embeddings(Nt, D, seed=7)).float().to(dev)
Tp = PackedEmbeddings(T)
q = torch.randn(Nt, device=dev)
for Ne in (20000, 10000, 2500):
    E = torch.from_numpy(synth.synth_embeddings(Ne, D, seed=6)).float().to(dev)
    r = torch.randn(Ne, device=dev)
    out = torch.empty((Ne, Nt), dtype=torch.float32, device=dev)
    for i in range(3):
        sk.score_matrix(E, Tp, r, q, cst=0.5, alpha=1.0, passes=0, out=out)
    torch.cuda.synchronize()
