"""Where does the end-to-end (host batches) path lose time against the device-resident loop?"""
import os, sys, time
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
order = list(reversed(range(K)))
host = [(flats[k].cpu().pin_memory(), blens[k]) for k in order]
out = torch.empty((len(shard), 256), device=dev)
with torch.no_grad():
    model.reserve(int(1.1 * max(len(b) for b in blens)), 1.1 * max(sum(bl) for bl in blens) / 16000.0, dev)
    for k in order[:3]:
        model.extract_packed(flats[k], blens[k])
    model.extract_stream(host[:3])
    model._out_host = torch.empty((len(shard), 256), dtype=torch.float32, pin_memory=True)
    for rep in range(2):
        ms_dev = bench.timed(lambda i: [model.extract_packed(flats[k], blens[k]) for k in order], 1, False) / K
        ms_e2e = bench.timed(lambda i: model.extract_stream(host, device_out=out), 1, False) / K
        # the copies alone
        st = torch.empty(max(h[0].numel() for h in host), device=dev)
        ms_copy = bench.timed(lambda i: [st[:h[0].numel()].copy_(h[0], non_blocking=True) for h in host], 1, False) / K
        print("rep %d: device loop %.3f ms/step, extract_stream %.3f ms/step, H2D alone %.3f ms/step (%.1f GB/s)" % (
            rep, ms_dev, ms_e2e, ms_copy, numpy.mean([h[0].numel() * 4 for h in host]) / ms_copy / 1e6))
    # variants
    devb = [(flats[k], blens[k]) for k in order]
    ms_dd = bench.timed(lambda i: model.extract_stream(devb, device_out=out), 1, False) / K
    print("extract_stream fed with DEVICE tensors (D2D staging copies): %.3f ms/step" % ms_dd)
    side = torch.cuda.Stream(dev)
    def with_side_copies(i):
        for j, k in enumerate(order):
            with torch.cuda.stream(side):
                st[:host[j][0].numel()].copy_(host[j][0], non_blocking=True)     # unrelated H2D traffic beside the compute
            model.extract_packed(flats[k], blens[k])
    ms_side = bench.timed(with_side_copies, 1, False) / K
    print("device loop with an unrelated 67 MB H2D per step on a side stream: %.3f ms/step" % ms_side)
