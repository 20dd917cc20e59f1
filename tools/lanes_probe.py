"""Probe: does running independent batches on several LANES (one engine handle + one CUDA stream + one host thread each)
raise extraction throughput?  The convolutions are tensor-pipe bound and the squeeze-excitation / front-end / pooling kernels
between them are HBM- or latency-bound, so kernels of different lanes can share the SMs.

    python tools/lanes_probe.py [--lanes 1 2 3] [--steps 24] [--utts 96]
"""
import argparse
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def run(n_lanes, steps, batches, models):
    streams = [torch.cuda.Stream() for _ in range(n_lanes)]
    n_rot = len(batches)

    def lane(k, count, offset):
        torch.cuda.set_device(0)
        with torch.no_grad(), torch.cuda.stream(streams[k]):
            for i in range(count):
                _, flat, lengths = batches[(offset + i * n_lanes) % n_rot]
                models[k].extract_packed(flat, lengths)

    best = None
    for rep in range(4):                      # rep 0 = warm-up (plans, buffers)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=lane, args=(k, steps // n_lanes, k)) for k in range(n_lanes)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("   rep %d: host %.1f ms, total %.1f ms" % (rep, t_host * 1e3, dt * 1e3), flush=True)
        if rep:
            best = dt if best is None else min(best, dt)
    dt = best
    done = sum(sum(batches[(k + i * n_lanes) % n_rot][2]) for k in range(n_lanes) for i in range(steps // n_lanes)) / 16000.0
    return done / dt, dt / (steps // n_lanes * n_lanes) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lanes", type=int, nargs="+", default=[1, 2, 3])
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--utts", type=int, default=96)
    ap.add_argument("--archi", default="halfresnet34")
    a = ap.parse_args()
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    emb = 256 if a.archi != "xvector" else 512
    batches = bench.make_batches(4, a.utts, seed=500, lo_s=2.0, hi_s=20.0, device=device)
    models = [bench.build_model(a.archi, emb, device) for _ in range(max(a.lanes))]
    ref = None
    for m in models:                          # every lane's engine returns the same embeddings
        out = m.extract_packed(batches[0][1], batches[0][2])
        ref = out if ref is None else ref
        assert torch.equal(out, ref)
    for n in a.lanes:
        v, ms = run(n, a.steps, batches, models)
        print("lanes=%d  %.0f audio-s/s  %.3f ms per batch of %d utterances" % (n, v, ms, a.utts), flush=True)


if __name__ == "__main__":
    main()
