"""Device time per equal-MAC bucket of the config-4 pool, with the per-category split: does a step of many short
utterances cost more than one of few long ones?"""
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
K = 20
lengths = bench.config4_lengths(K * 96)
shard = bulk.plan_shards(lengths, 1)[0]
batches = bulk.make_batches_equal_cost(shard, lengths, K)
blens = [[int(lengths[i]) for i in b] for b in batches]
flats = [bench.device_audio(bl, 777000 + k, dev) for k, bl in enumerate(blens)]
with torch.no_grad():
    model.reserve(int(1.1 * max(len(b) for b in blens)), 1.1 * max(sum(bl) for bl in blens) / 16000.0, dev)
    for k in range(K):
        model.extract_packed(flats[k], blens[k])
    torch.cuda.synchronize()
    print("bucket  utts  audio-s  GMAC   ms/step   front-end  stem   conv    SE   pool+head")
    only = [int(v) for v in os.environ.get('SKB_PROBE_BUCKETS', '').split(',') if v] or list(range(K))
    for k in only:
        f = lambda i: model.extract_packed(flats[k], blens[k])
        for i in range(2):
            f(i)
        ms = bench.timed(f, 5, False) / 5
        cat = bench.profile_categories(lib, f, 5) / 5
        print("%4d  %5d  %7.1f  %6.1f  %7.3f   %7.3f %7.3f %7.3f %7.3f %7.3f" % (
            k, len(blens[k]), sum(blens[k]) / 16000.0, float(numpy.sum(bulk.halfresnet34_macs(numpy.asarray(blens[k])))) / 1e9, ms,
            cat[0], cat[1], cat[2], cat[3], cat[4]))
