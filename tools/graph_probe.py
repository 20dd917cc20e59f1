"""Probe: how much of a step is launch gaps?  Times the same packed batch eagerly and as a replayed CUDA graph."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
with torch.no_grad():
    m = bench.build_model("halfresnet34", 256, dev)
    b = bench.make_batches(1, 96, 500, 2.0, 20.0, dev)[0]
    for _ in range(3):
        m.extract_packed(b[1], b[2])
    torch.cuda.synchronize()

    def timeit(fn, n=20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    print("eager  : %.3f ms / step" % timeit(lambda: m.extract_packed(b[1], b[2])))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        m.extract_packed(b[1], b[2])
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            out = m.extract_packed(b[1], b[2])
        print("graph  : %.3f ms / step" % timeit(g.replay))
    except Exception as e:
        print("graph capture failed:", repr(e)[:300])
