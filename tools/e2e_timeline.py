"""Timeline of the copy / compute pipeline of Xtractor.extract_stream: when does every H2D copy and every forward
start and end on the device?  Variants: number of staging buffers, and whether the copy of batch i+1 is enqueued
before or after the forward of batch i (host order)."""
import ctypes, os, sys
import numpy, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sidekit_b200 import bulk, _lib

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
n_max = max(h[0].numel() for h in host)


def pipeline(nbuf, ahead, verbose, tables_first=False):
    compute = torch.cuda.current_stream(dev)
    copy_s = torch.cuda.Stream(dev)
    stage = [torch.empty(n_max, device=dev) for _ in range(nbuf)]
    free_ev = [None] * nbuf
    ev = lambda: torch.cuda.Event(enable_timing=True)
    c0, c1, f0, f1 = [ev() for _ in host], [ev() for _ in host], [ev() for _ in host], [ev() for _ in host]
    ready = [None] * len(host)
    torch.cuda.synchronize()
    base = ev(); base.record(compute)

    def enqueue_copy(i):
        k, n = i % nbuf, host[i][0].numel()
        with torch.cuda.stream(copy_s):
            if free_ev[k] is not None:
                copy_s.wait_event(free_ev[k])
            c0[i].record(copy_s)
            stage[k][:n].copy_(host[i][0], non_blocking=True)
            c1[i].record(copy_s)
        ready[i] = c1[i]

    row = 0
    for j in range(min(ahead, len(host))):
        enqueue_copy(j)
    for i in range(len(host)):
        if ahead == 0:
            enqueue_copy(i)
        elif i + ahead < len(host) and i > 0:
            enqueue_copy(i + ahead)                 # before forward(i) in host order
        k, n = i % nbuf, host[i][0].numel()
        compute.wait_event(ready[i])
        f0[i].record(compute)
        model._run(stage[k][:n], host[i][1], True, want_logits=False, emb_out=out[row:row + len(host[i][1])])
        f1[i].record(compute)
        free_ev[k] = f1[i]
        if tables_first:
            _lib.check(_lib.lib().skb_xtractor_wait_tables(model._handle(dev), ctypes.c_void_p(copy_s.cuda_stream)))
        row += len(host[i][1])
        if ahead and i == 0 and ahead < len(host):
            enqueue_copy(ahead)
    end = ev(); end.record(compute)
    torch.cuda.synchronize()
    total = base.elapsed_time(end)
    if verbose:
        for i in range(len(host)):
            print("  batch %2d  copy %7.2f..%7.2f   forward %7.2f..%7.2f  (%.2f)" % (
                i, base.elapsed_time(c0[i]), base.elapsed_time(c1[i]), base.elapsed_time(f0[i]), base.elapsed_time(f1[i]),
                f0[i].elapsed_time(f1[i])))
    return total / len(host)


with torch.no_grad():
    model.reserve(int(1.1 * max(len(b) for b in blens)), 1.1 * max(sum(bl) for bl in blens) / 16000.0, dev)
    for k in order[:3]:
        model.extract_packed(flats[k], blens[k])
    for nbuf, ahead in ((2, 0), (2, 0), (2, 1), (3, 1), (3, 2), (4, 2)):
        ms = pipeline(nbuf, ahead, False)
        print("nbuf %d ahead %d: %.3f ms/step" % (nbuf, ahead, ms))
    for rep in range(3):
        print("nbuf 2 ahead 0, copy stream waits for the tables of the running forward: %.3f ms/step" % pipeline(2, 0, False, True))
    print("timeline of that:")
    pipeline(2, 0, True, True)
    print("timeline nbuf 2 ahead 0 (round-2s order):")
    pipeline(2, 0, True)
    print("timeline nbuf 3 ahead 1:")
    pipeline(3, 1, True)
