"""Throughput against the batch size (utterances per step) on the config-4 pool: how much of a step is fixed cost?"""
import os, sys
import numpy, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sidekit_b200 import bulk, _lib

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _lib.lib()
model = bench.build_model("halfresnet34", 256, dev)
N = 20 * 96
lengths = bench.config4_lengths(N)
shard = bulk.plan_shards(lengths, 1)[0]
with torch.no_grad():
    for K in (40, 20, 10, 5):
        batches = bulk.make_batches_equal_cost(shard, lengths, K)
        blens = [[int(lengths[i]) for i in b] for b in batches]
        flats = [bench.device_audio(bl, 777000 + k, dev) for k, bl in enumerate(blens)]
        model.reserve(int(1.1 * max(len(b) for b in blens)), 1.1 * max(sum(bl) for bl in blens) / 16000.0, dev)
        for k in range(min(K, 3)):
            model.extract_packed(flats[k], blens[k])
        torch.cuda.synchronize()
        f = lambda i: [model.extract_packed(flats[k], blens[k]) for k in range(K)]
        ms = min(bench.timed(f, 1, False) for _ in range(3))
        audio = sum(sum(bl) for bl in blens) / 16000.0
        cat = bench.profile_categories(lib, f, 1)
        print("K %3d batches of ~%4d utts (%6.0f audio-s each): pool in %7.2f ms = %8.0f audio-s/s;  front-end %.2f stem %.2f conv %.2f SE %.2f pool %.2f  (mem %.1f GB)" % (
            K, N // K, audio / K, ms, audio / (ms / 1e3), cat[0], cat[1], cat[2], cat[3], cat[4], torch.cuda.max_memory_allocated() / 1e9))
        del flats
