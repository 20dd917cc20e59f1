"""Minimal driver for ncu: N forward passes of one workload through the public API (no oracle, no timing).

    python tools/profile_step.py [hr34|tdnn|plda] [n_passes]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "hr34"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    with torch.no_grad():
        if what == "hr34":
            m = bench.build_model("halfresnet34", 256, dev)
            b = bench.make_batches(1, 96, 500, 2.0, 20.0, dev)[0]
            for _ in range(n):
                m.extract_packed(b[1], b[2])
        elif what == "tdnn":
            m = bench.build_model("xvector", 512, dev)
            b = bench.make_batches(1, 512, 900, 2.0, 20.0, dev)[0]
            for _ in range(n):
                m.extract_packed(b[1], b[2])
        else:
            import sidekit_b200 as sk
            from sidekit_b200 import synth
            E = torch.from_numpy(synth.synth_embeddings(20000, 256, seed=6)).float().to(dev)
            T = torch.from_numpy(synth.synth_embeddings(20000, 256, seed=7)).float().to(dev)
            r, q = torch.randn(20000, device=dev), torch.randn(20000, device=dev)
            out = torch.empty((20000, 20000), dtype=torch.float32, device=dev)
            passes = int(os.environ.get("SKB_PASSES", "0"))
            for _ in range(n):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                sk.score_matrix(E, T, r, q, cst=0.5, passes=passes, out=out)
                e1.record()
                torch.cuda.synchronize()
                print("score_matrix 20k x 20k passes=%d: %.3f ms" % (passes, e0.elapsed_time(e1)))
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
