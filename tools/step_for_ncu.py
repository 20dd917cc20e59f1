"""A few HalfResNet34 steps of the bench workload (config-4 buckets) for ncu launch lists / full captures.
Usage: python tools/step_for_ncu.py [hr34|tdnn] [n_steps]"""
import os, sys
import numpy, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sidekit_b200 import bulk

which = sys.argv[1] if len(sys.argv) > 1 else "hr34"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
with torch.no_grad():
    if which == "hr34":
        model = bench.build_model("halfresnet34", 256, dev)
        K = 20
        lengths = bench.config4_lengths(K * 96)
        shard = bulk.plan_shards(lengths, 1)[0]
        batches = bulk.make_batches_equal_cost(shard, lengths, K)
        pick = [10, 9, 11, 8, 12][:n + 2]                     # mid-length buckets (B ~ 75, 1043 audio-s: the bench step)
        data = [([int(lengths[i]) for i in batches[k]]) for k in pick]
        flats = [bench.device_audio(bl, 100 + j, dev) for j, bl in enumerate(data)]
        model.reserve(int(1.1 * max(len(b) for b in data)), 1.1 * max(sum(b) for b in data) / 16000.0, dev)
        for j in range(2):
            model.extract_packed(flats[j], data[j])           # warm-up (2 steps)
        torch.cuda.synchronize()
        for j in range(2, len(data)):
            model.extract_packed(flats[j], data[j])           # profiled steps: every batch a new geometry
    else:
        model = bench.build_model("xvector", 512, dev)
        tl = bench.config4_lengths(512 * (n + 2), seed=4)
        for j in range(n + 2):
            ls = numpy.sort(tl[j * 512:(j + 1) * 512])
            model.extract_packed(bench.device_audio(ls, 4400 + j, dev), [int(v) for v in ls])
    torch.cuda.synchronize()
