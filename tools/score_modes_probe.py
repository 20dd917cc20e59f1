"""One call of every score-GEMM output mode at 20k x 20k (for ncu)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sidekit_b200 as sk
from sidekit_b200 import synth
from sidekit_b200.iv_scoring import TrialIndex, score_trials
dev = torch.device("cuda", 0)
N, D = 20000, 256
E = torch.from_numpy(synth.synth_embeddings(N, D, seed=6)).float().to(dev)
T = torch.from_numpy(synth.synth_embeddings(N, D, seed=7)).float().to(dev)
r, q = torch.randn(N, device=dev), torch.randn(N, device=dev)
mask = torch.rand((N, N), device=dev) < 1.7e-3
idx = TrialIndex(mask, dev)
o32 = torch.empty((N, N), dtype=torch.float32, device=dev)
o16 = torch.empty((N, N), dtype=torch.float16, device=dev)
for i in range(2):
    sk.score_matrix(E, T, r, q, 0.5, 1.0, passes=0, out=o32)
    sk.score_matrix(E, T, r, q, 0.5, 1.0, passes=0, out=o16)
    score_trials(E, T, idx, r, q, 0.5, 1.0, passes=0)
torch.cuda.synchronize()
