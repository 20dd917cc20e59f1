"""The real sharded path on real GPUs (BASELINE config 4): N-rank extraction over NCCL must equal the
1-rank result BIT FOR BIT, a rank must only touch its own utterances, and the native state must be per device.

* ``test_shards_on_one_gpu_are_bit_identical``  -- one GPU: every shard of a 2- / 3-rank plan extracted separately (its own
  length-bucketed batches) and scattered back equals the single-shard run bit for bit (what makes the N-rank run exact).
* ``test_nccl_world2_*``  -- two processes, two GPUs, NCCL (skipped on a 1-GPU box; run with ``gpurun --gpus 2``).
"""
import os
import socket

import numpy
import pytest
import torch

from sidekit_b200 import bulk, synth
from tests.models import make_xtractor

pytestmark = pytest.mark.gpu

N_UTT = 23


def _lengths():
    return synth.synth_lengths(N_UTT, 0.6, 3.0, seed=77)


def _wave(i, lengths):
    return synth.synth_wave(1, int(lengths[i]), seed=500 + i)[0]


def _single_rank_reference(model, lengths):
    waves = [_wave(i, lengths).cuda() for i in range(len(lengths))]
    return model.extract_varlen(waves).cpu().numpy()


def test_shards_on_one_gpu_are_bit_identical():
    model = make_xtractor("halfresnet34", 16, 256).cuda()
    lengths = _lengths()
    with torch.no_grad():
        ref = _single_rank_reference(model, lengths)
        one = bulk.extract_embeddings_sharded(lambda ws: model.extract_varlen([w.cuda() for w in ws]), lambda i: _wave(i, lengths), 256,
                                              max_audio_seconds=9.0, lengths=lengths).cpu().numpy()
        assert numpy.array_equal(one, ref)
        for world in (2, 3):
            shards = bulk.plan_shards(lengths, world)
            out = numpy.zeros_like(ref)
            for s in shards:
                parts = [model.extract_varlen([_wave(i, lengths).cuda() for i in b]) for b in bulk.make_batches_equal_cost(s, lengths, 2) if b]
                out[s] = torch.cat(parts).cpu().numpy()
            assert numpy.array_equal(out, ref), "world %d" % world


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    model = make_xtractor("halfresnet34", 16, 256).to(dev)
    lengths = _lengths()
    touched = []

    def wave(i):
        touched.append(i)
        return _wave(i, lengths)

    with torch.no_grad():
        out = bulk.extract_embeddings_sharded(lambda ws: model.extract_varlen([w.to(dev) for w in ws]), wave, 256,
                                              max_audio_seconds=9.0, lengths=lengths)
        # a shard plan with an EMPTY rank (one utterance, two ranks) must not hang the collective
        one = bulk.extract_embeddings_sharded(lambda ws: model.extract_varlen([w.to(dev) for w in ws]), lambda i: _wave(i, lengths),
                                              256, lengths=lengths[:1])
        # the as-norm exchange on the real kernels
        X = torch.nn.functional.normalize(torch.from_numpy(synth.synth_embeddings(301, 64, seed=3, unit_norm=False)).float(), dim=1).to(dev)
        coh = torch.from_numpy(synth.synth_embeddings(450, 64, seed=4, unit_norm=False)).float().to(dev)
        lo, hi, panel = bulk.asnorm_sharded(X, coh, 200)
    mine = bulk.plan_shards(lengths, world)[rank]
    q.put((rank, out.cpu().numpy(), one.cpu().numpy(), sorted(touched) == sorted(mine.tolist()), lo, hi, panel.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_nccl_world2_equals_single_rank_bit_for_bit():
    import torch.multiprocessing as mp
    import sidekit_b200 as sk
    lengths = _lengths()
    model = make_xtractor("halfresnet34", 16, 256).cuda()
    with torch.no_grad():
        ref = _single_rank_reference(model, lengths)
        X = torch.nn.functional.normalize(torch.from_numpy(synth.synth_embeddings(301, 64, seed=3, unit_norm=False)).float(), dim=1).cuda()
        coh = torch.from_numpy(synth.synth_embeddings(450, 64, seed=4, unit_norm=False)).float().cuda()
        full = sk.asnorm(X, coh, None)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, out, one, only_mine, lo, hi, panel in res:
        assert numpy.array_equal(out, ref), "rank %d: N-rank embeddings differ from the 1-rank run" % rank
        assert numpy.array_equal(one, ref[:1])
        assert only_mine, "rank %d loaded utterances outside its shard" % rank
    assert res[0][4] == 0 and res[0][5] == res[1][4] and res[1][5] == 301
    stacked = numpy.concatenate([res[0][6], res[1][6]])
    assert numpy.array_equal(stacked, full), "row panels of the sharded as-norm differ from the single-GPU matrix"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_native_state_is_per_device():
    """workspaces and the >48 KB shared-memory opt-in are per device: the same process scores and extracts on
    cuda:0, then on cuda:1, then on cuda:0 again."""
    import sidekit_b200 as sk
    E, T = synth.synth_embeddings(300, 256, seed=1), synth.synth_embeddings(200, 256, seed=2)
    ids_e = numpy.array(["m%d" % i for i in range(300)])
    ids_t = numpy.array(["s%d" % i for i in range(200)])
    ndx = sk.Ndx()
    ndx.modelset, ndx.segset, ndx.trialmask = ids_e, ids_t, numpy.ones((300, 200), dtype=bool)
    en, te = sk.StatServer.from_embeddings(ids_e, E), sk.StatServer.from_embeddings(ids_t, T)
    x = synth.synth_wave(2, 16000, seed=5)
    model = make_xtractor("halfresnet34", 16, 256)
    mats, embs = [], []
    with torch.no_grad():
        for d in (0, 1, 0):
            mats.append(sk.cosine_scoring(en, te, ndx, device="cuda:%d" % d).scoremat.copy())
            model = model.to("cuda:%d" % d)
            embs.append(model(x.to("cuda:%d" % d), is_eval=True)[1].cpu().numpy())
    assert numpy.array_equal(mats[0], mats[1]) and numpy.array_equal(mats[0], mats[2])
    assert numpy.array_equal(embs[0], embs[1]) and numpy.array_equal(embs[0], embs[2])
