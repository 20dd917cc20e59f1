"""Where does the first pass over never-seen batch geometries lose time?  Per-batch host and device times of two passes."""
import os, sys, time, json
import numpy, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sidekit_b200 import bulk

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
model = bench.build_model("halfresnet34", 256, dev)
K = 20
lengths = bench.config4_lengths(K * 96)
shard = bulk.plan_shards(lengths, 1)[0]
batches = bulk.make_batches_equal_cost(shard, lengths, K)
blens = [[int(lengths[i]) for i in b] for b in batches]
flats = [bench.device_audio(bl, 777000 + k, dev) for k, bl in enumerate(blens)]
use_reserve = os.environ.get("NO_RESERVE") is None
with torch.no_grad():
    t0 = time.perf_counter()
    if use_reserve:
        model.reserve(int(1.1 * max(len(b) for b in blens)), 1.1 * max(sum(bl) for bl in blens) / 16000.0, dev)
    torch.cuda.synchronize()
    print("reserve s", time.perf_counter() - t0)
    wl = bench.config4_lengths(int(3 * 96 * 1.06), seed=9000)
    wb = bulk.make_batches_equal_cost(numpy.argsort(wl, kind="stable"), wl, 3)
    for k in reversed(range(3)):
        wls = [int(wl[i]) for i in wb[k]]
        model.extract_packed(bench.device_audio(wls, 5 + k, dev), wls)
    torch.cuda.synchronize()
    for p in range(3):
        evs, host = [], []
        torch.cuda.synchronize()
        tstart = time.perf_counter()
        for k in reversed(range(K)):
            e = torch.cuda.Event(enable_timing=True); e.record(); evs.append(e)
            h0 = time.perf_counter()
            model.extract_packed(flats[k], blens[k])
            host.append((time.perf_counter() - h0) * 1e3)
        e = torch.cuda.Event(enable_timing=True); e.record(); evs.append(e)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - tstart) * 1e3
        devms = [evs[i].elapsed_time(evs[i + 1]) for i in range(K)]
        print("pass", p, "wall ms/step %.2f" % (wall / K))
        print("  B      ", [len(blens[k]) for k in reversed(range(K))])
        print("  host ms", ["%.1f" % v for v in host])
        print("  dev ms ", ["%.1f" % v for v in devms])
