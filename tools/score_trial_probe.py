"""Steady-state time of the score-GEMM output modes at 20k x 20k x 256 (test side packed once), CUDA events."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import sidekit_b200 as sk
from sidekit_b200 import synth
from sidekit_b200.iv_scoring import TrialIndex, score_trials, PackedEmbeddings
dev = torch.device("cuda", 0)
N, D = 20000, 256
E = torch.from_numpy(synth.synth_embeddings(N, D, seed=6)).float().to(dev)
T = torch.from_numpy(synth.synth_embeddings(N, D, seed=7)).float().to(dev)
Tp = PackedEmbeddings(T)
r, q = torch.randn(N, device=dev), torch.randn(N, device=dev)
g = torch.Generator(device=dev).manual_seed(99)
mask = torch.rand((N, N), device=dev, generator=g) < (37720.0 / 4708.0 ** 2)
idx = TrialIndex(mask, dev)
o32 = torch.empty((N, N), dtype=torch.float32, device=dev)
o16 = torch.empty((N, N), dtype=torch.float16, device=dev)
modes = {"float32 matrix": lambda i: sk.score_matrix(E, Tp, r, q, 0.5, 1.0, passes=0, out=o32),
         "float16 matrix": lambda i: sk.score_matrix(E, Tp, r, q, 0.5, 1.0, passes=0, out=o16),
         "trial list (both sides packed per call)": lambda i: score_trials(E, T, idx, r, q, 0.5, 1.0, passes=0),
         "trial list (test side packed once)": lambda i: score_trials(E, Tp, idx, r, q, 0.5, 1.0, passes=0),
         "trial list, no column term": lambda i: score_trials(E, Tp, idx, r, None, 0.5, 1.0, passes=0)}
for name, f in modes.items():
    for i in range(3):
        f(i)
    ms = min(bench.timed(f, 20, False) for _ in range(3)) / 20
    print("%-42s %.4f ms  %.0f TFLOP/s" % (name, ms, 2.0 * N * N * D / (ms / 1e3) / 1e12))
# the trial list equals matrix[mask]
sk.score_matrix(E, Tp, r, q, 0.5, 1.0, passes=0, out=o32)
tl = score_trials(E, T, idx, r, q, 0.5, 1.0, passes=0)
tl2 = score_trials(E, Tp, idx, r, q, 0.5, 1.0, passes=0)
print("trial list == matrix[mask]:", bool(torch.equal(tl, o32[mask])), bool(torch.equal(tl2, tl)), int(tl.numel()))
