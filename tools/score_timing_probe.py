"""Diagnostic (library built with ./build.sh -DSKB_SCORE_TIMING): where the roles of score_gemm_kernel wait, per output mode."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sidekit_b200 as sk
from sidekit_b200 import synth, _lib
from sidekit_b200.iv_scoring import TrialIndex, score_trials, PackedEmbeddings
dev = torch.device("cuda", 0)
lib = ctypes.CDLL(_lib.LIB_PATH)
N, D = 20000, 256
E = torch.from_numpy(synth.synth_embeddings(N, D, seed=6)).float().to(dev)
T = torch.from_numpy(synth.synth_embeddings(N, D, seed=7)).float().to(dev)
Tp = PackedEmbeddings(T)
r, q = torch.randn(N, device=dev), torch.randn(N, device=dev)
g = torch.Generator(device=dev).manual_seed(99)
idx = TrialIndex(torch.rand((N, N), device=dev, generator=g) < (37720.0 / 4708.0 ** 2), dev)
o32 = torch.empty((N, N), dtype=torch.float32, device=dev)
o16 = torch.empty((N, N), dtype=torch.float16, device=dev)
modes = {"float32 matrix": lambda: sk.score_matrix(E, Tp, r, q, 0.5, 1.0, passes=0, out=o32),
         "float16 matrix": lambda: sk.score_matrix(E, Tp, r, q, 0.5, 1.0, passes=0, out=o16),
         "trial list": lambda: score_trials(E, Tp, idx, r, q, 0.5, 1.0, passes=0)}
buf = (ctypes.c_ulonglong * 16)()
names = ["MMA warp total", "MMA waits b_full (T data)", "MMA waits acc_empty (epilogue)", "MMA waits a_full (E panel)", "producer waits b_empty (ring full)",
         "epilogue warp 2 waits acc_full (MMA)", "epilogue warp 2 total", "producer total",
         "MMA: issue of the MMAs of a stage", "MMA: commit", "MMA: issue section incl. elect + syncwarp"]
for name, f in modes.items():
    for i in range(2):
        f()
    lib.skb_debug_score_timing(buf)
    reps = 5
    for i in range(reps):
        f()
    lib.skb_debug_score_timing(buf)
    print(name)
    for i, n in enumerate(names):
        print("   %-42s %9.0f cycles per CTA and call" % (n, buf[i] / 148.0 / reps))
