"""Cost of a batch whose geometry has not been seen before (plan build + table upload) against a cached one.

    python tools/plan_probe.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    m = bench.build_model("halfresnet34", 256, dev)
    batches = bench.make_batches(24, 96, 1234, 2.0, 20.0, dev)
    batches.sort(key=lambda b: -sum(b[2]))
    with torch.no_grad():
        m.extract_packed(batches[0][1], batches[0][2])          # buffers sized by the largest batch
        torch.cuda.synchronize()
        for label, order in (("new geometry", range(1, 24)), ("cached (last 8)", list(range(16, 24)) * 2)):
            ts = []
            for i in order:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                m.extract_packed(batches[i][1], batches[i][2])
                t1 = time.perf_counter()
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                ts.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
            print(label, "host+device ms:", " ".join("%.1f" % t[1] for t in ts))
            print("%-18s host %.2f ms  host+device %.2f ms per batch (median of %d)" %
                  (label, sorted(t[0] for t in ts)[len(ts) // 2], sorted(t[1] for t in ts)[len(ts) // 2], len(ts)))
        # back-to-back throughput with new geometries every batch
        fresh = bench.make_batches(24, 96, 4321, 2.0, 20.0, dev)
        fresh.sort(key=lambda b: -sum(b[2]))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        per = []
        for b in fresh:
            t1 = time.perf_counter()
            m.extract_packed(b[1], b[2])
            per.append((time.perf_counter() - t1) * 1e3)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("host ms per call:", " ".join("%.1f" % v for v in per))
        print("24 fresh batches back to back: %.2f ms per batch, %.0f audio-s/s" % (dt / 24 * 1e3, sum(sum(b[2]) for b in fresh) / 16000.0 / dt))
        t0 = time.perf_counter()
        for i in range(24):
            b = fresh[16 + i % 8]
            m.extract_packed(b[1], b[2])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("24 cached batches back to back: %.2f ms per batch" % (dt / 24 * 1e3))


if __name__ == "__main__":
    main()
