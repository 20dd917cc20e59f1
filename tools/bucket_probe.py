"""How the step time depends on the batch composition (config-4 shards are length buckets: 55 long ... 480 short utterances
per equal-MAC batch) and on whether the geometry plan is cached.  Run on the GPU box: python tools/bucket_probe.py"""
import ctypes
import json
import os
import sys

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sidekit_b200 import _lib, bulk  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    model = bench.build_model("halfresnet34", 256, dev)
    lib = _lib.lib()
    K = 20
    lengths = bench.config4_lengths(K * 96)
    shard = bulk.plan_shards(lengths, 1)[0]
    batches = bulk.make_batches_equal_cost(shard, lengths, K)
    out = {}
    with torch.no_grad():
        for name, ks in (("short_bucket", [0, 1, 2]), ("mid_bucket", [9, 10, 11]), ("long_bucket", [17, 18, 19])):
            data = []
            for k in ks:
                bl = [int(lengths[i]) for i in batches[k]]
                data.append((bench.device_audio(bl, 100 + k, dev), bl))
            for flat, bl in data:
                model.extract_packed(flat, bl)
            f = lambda i: model.extract_packed(data[i % 3][0], data[i % 3][1])
            ms = bench.timed(f, 12, False) / 12
            cat = bench.profile_categories(lib, f, 12) / 12
            aud = numpy.mean([sum(bl) for _, bl in data]) / 16000.0
            out[name] = {"B": [len(bl) for _, bl in data], "audio_s": aud, "ms_cached": ms, "us_per_audio_s": ms * 1e3 / aud,
                         "cat_ms": {"frontend": float(cat[0]), "stem": float(cat[1]), "conv": float(cat[2]), "se": float(cat[3]), "pool": float(cat[4])}}
        # mixed batch (round-1 composition)
        rot = bench.make_batches(3, 96, seed=500, lo_s=2.0, hi_s=20.0, device=dev)
        for r in rot:
            model.extract_packed(r[1], r[2])
        f = lambda i: model.extract_packed(rot[i % 3][1], rot[i % 3][2])
        ms = bench.timed(f, 12, False) / 12
        cat = bench.profile_categories(lib, f, 12) / 12
        aud = numpy.mean([sum(r[2]) for r in rot]) / 16000.0
        out["mixed_96"] = {"B": [96] * 3, "audio_s": aud, "ms_cached": ms, "us_per_audio_s": ms * 1e3 / aud,
                           "cat_ms": {"frontend": float(cat[0]), "stem": float(cat[1]), "conv": float(cat[2]), "se": float(cat[3]), "pool": float(cat[4])}}
        # fresh geometry every step: all 20 batches in a cycle (cache holds 9)
        data = []
        for k in range(K):
            bl = [int(lengths[i]) for i in batches[k]]
            data.append((bench.device_audio(bl, 300 + k, dev), bl))
        for rep in range(2):
            for flat, bl in reversed(data):
                model.extract_packed(flat, bl)
        f = lambda i: model.extract_packed(data[K - 1 - i][0], data[K - 1 - i][1])
        ms = bench.timed(f, K, False) / K
        aud = numpy.mean([sum(bl) for _, bl in data]) / 16000.0
        out["all_buckets_fresh_plan"] = {"audio_s": aud, "ms": ms, "us_per_audio_s": ms * 1e3 / aud}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
