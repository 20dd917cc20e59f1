"""Host logic of the multi-GPU path on CPU: world_size-2 gloo run of the sharded extraction (stub extractor),
shard balance / determinism, batch building, row-panel sharding."""
import os
import socket

import numpy
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sidekit_b200 import bulk, synth


def _stub_extract(ws):
    # deterministic "embedding" of a waveform: a few moments, so order mistakes are detected
    return torch.stack([torch.stack([w.sum(), (w * w).sum(), w[0], w[-1], torch.tensor(float(w.numel()))]) for w in ws])


def _worker(rank, world, port, q, n_utt=37):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = synth.synth_lengths(n_utt, 0.2, 2.0, seed=11)
    touched = []

    def wave(i):                         # streamed source: a rank must only ever load ITS utterances
        touched.append(i)
        return synth.synth_wave(1, int(lengths[i]), seed=100 + i)[0]

    out = bulk.extract_embeddings_sharded(_stub_extract, wave, 5, max_audio_seconds=6.0, lengths=lengths)
    mine = bulk.plan_shards(lengths, world)[rank]
    assert sorted(touched) == sorted(mine.tolist())
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_extraction_world2_matches_single_process():
    lengths = synth.synth_lengths(37, 0.2, 2.0, seed=11)
    waves = [synth.synth_wave(1, int(L), seed=100 + i)[0] for i, L in enumerate(lengths)]
    ref = _stub_extract(waves).numpy()
    single = bulk.extract_embeddings_sharded(_stub_extract, waves, 5, max_audio_seconds=6.0).numpy()
    assert numpy.array_equal(single, ref)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert numpy.array_equal(res[0], ref) and numpy.array_equal(res[1], ref)     # bit-for-bit, every rank


def test_sharded_extraction_more_ranks_than_utterances():
    """A rank with an EMPTY shard must still take part in the collective (it used to fall back to a CPU tensor)."""
    lengths = synth.synth_lengths(1, 0.2, 2.0, seed=11)
    ref = _stub_extract([synth.synth_wave(1, int(lengths[0]), seed=100)[0]]).numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, 1)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert numpy.array_equal(res[0], ref) and numpy.array_equal(res[1], ref)


def test_equal_cost_batches():
    lengths = synth.synth_lengths(1000, 2.0, 20.0, seed=5)
    shard = bulk.plan_shards(lengths, 4)[1]
    batches = bulk.make_batches_equal_cost(shard, lengths, 7)
    assert len(batches) == 7 and [i for b in batches for i in b] == shard.tolist()
    loads = [sum(bulk.halfresnet34_macs(int(lengths[i])) for i in b) for b in batches]
    assert max(loads) / (sum(loads) / 7) < 1.05
    assert [len(b) for b in bulk.make_batches_equal_cost(shard[:3], lengths, 5)].count(0) == 2
    assert bulk.make_batches_equal_cost([], lengths, 3) == [[], [], []]
    assert numpy.array_equal(bulk.halfresnet34_macs(lengths), [bulk.halfresnet34_macs(int(v)) for v in lengths])


def test_plan_shards_balanced_and_deterministic():
    lengths = synth.synth_lengths(1000, 2.0, 20.0, seed=5)
    for world in (1, 2, 4, 8):
        shards = bulk.plan_shards(lengths, world)
        allidx = numpy.sort(numpy.concatenate(shards))
        assert numpy.array_equal(allidx, numpy.arange(1000))
        loads = [sum(bulk.halfresnet34_macs(int(lengths[i])) for i in s) for s in shards]
        assert max(loads) / (sum(loads) / world) < 1.01
        again = bulk.plan_shards(lengths, world)
        assert all(numpy.array_equal(a, b) for a, b in zip(shards, again))
        for s in shards:
            assert numpy.all(numpy.diff(lengths[s]) >= 0)                       # length-sorted -> bucketed batches


def test_make_batches_and_row_panels():
    lengths = synth.synth_lengths(100, 2.0, 20.0, seed=6)
    idx = numpy.argsort(lengths)
    batches = bulk.make_batches(idx, lengths, max_audio_seconds=100.0, max_utts=16)
    assert sum(len(b) for b in batches) == 100 and [i for b in batches for i in b] == idx.tolist()
    for b in batches:
        assert len(b) <= 16 and (sum(lengths[i] for i in b) / 16000.0 <= 100.0 or len(b) == 1)
    cover = [bulk.row_panel(20000, r, 8) for r in range(8)]
    assert cover[0][0] == 0 and cover[-1][1] == 20000 and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    assert bulk.halfresnet34_macs(64000) == 9236665216                          # SURVEY.md 8d closed form at 4 s


# ----------------------------------------------------------------------------- as-norm exchange (world 2, gloo)
def _stats_cpu(X, coh, topk):
    top = torch.topk(X @ coh.T, topk, dim=1)[0]
    return top.mean(dim=1), top.std(dim=1)


def _panel_cpu(X, lo, hi, mean, std):
    S = X[lo:hi] @ X.T
    return 0.5 * (S - mean[lo:hi, None]) / std[lo:hi, None] + 0.5 * (S - mean[None, :]) / std[None, :]


def _asnorm_inputs():
    g = torch.Generator().manual_seed(3)
    X = torch.nn.functional.normalize(torch.randn(45, 24, generator=g), dim=1)
    return X, torch.randn(260, 24, generator=g)


def _asnorm_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, coh = _asnorm_inputs()
    lo, hi, panel = bulk.asnorm_sharded(X, coh, topk=200, stats_fn=_stats_cpu, panel_fn=_panel_cpu)
    q.put((rank, lo, hi, panel.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_asnorm_sharded_world2_matches_full_matrix():
    from oracle import scoring_ref as S
    X, coh = _asnorm_inputs()
    full = S.asnorm(X.numpy(), coh.numpy(), 200)
    lo, hi, panel = bulk.asnorm_sharded(X, coh, topk=200, stats_fn=_stats_cpu, panel_fn=_panel_cpu)      # single process
    assert (lo, hi) == (0, 45) and numpy.abs(panel.numpy() - full).max() < 1e-4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_asnorm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 45            # contiguous row panels
    stacked = numpy.concatenate([res[0][3], res[1][3]])
    assert numpy.abs(stacked - full).max() < 1e-4
