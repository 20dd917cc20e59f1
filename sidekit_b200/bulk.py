"""Bulk extraction / scoring across GPUs: one process per GPU, utterances sharded by length
bucket with greedy FLOP balancing, embeddings all-gathered with NCCL (the path's only collective); score
matrices are sharded by row panel with no collective at all.  Host logic only -- the per-rank work goes through
``Xtractor.extract_packed`` / ``score_matrix`` (CUDA).  Mirrors what the reference does one utterance at a time in
``extract_embeddings`` (sidekit/nnet/xvector.py:1796-1916) and returns the same ``StatServer`` layout.
"""
import numpy
import torch

from .statserver import StatServer


def halfresnet34_macs(n_samples):
    """Algorithmic MACs of one HalfResNet34 embedding (SURVEY.md 8d closed form); used as the balancing weight."""
    t1 = 1 + n_samples // 160
    t2 = (t1 - 1) // 2 + 1
    t3 = (t2 - 1) // 2 + 1
    t4 = (t3 - 1) // 2 + 1
    return 80 * t1 * 56608 + 40 * t2 * 278528 + 20 * t3 * 1703936 + 10 * t4 * 3276800 + 1310720 * t4 + 1350016


def plan_shards(lengths, world_size, cost=halfresnet34_macs):
    """Deterministic assignment of utterances to ranks: longest-first greedy on the MAC count.
    Returns ``world_size`` index arrays, each sorted by length (so consecutive batches are length buckets)."""
    lengths = numpy.asarray(lengths, dtype=numpy.int64)
    order = numpy.argsort(-lengths, kind="stable")
    load = numpy.zeros(world_size, dtype=numpy.float64)
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = int(numpy.argmin(load))            # ties -> lowest rank: deterministic
        shards[r].append(int(i))
        load[r] += float(cost(int(lengths[i])))
    return [numpy.array(sorted(s, key=lambda i: (int(lengths[i]), i)), dtype=numpy.int64) for s in shards]


def make_batches(indices, lengths, max_audio_seconds=1200.0, max_utts=256, sample_rate=16000):
    """Split a length-sorted index list into packed batches bounded by total audio (activation memory) and count."""
    batches, cur, cur_s = [], [], 0.0
    for i in indices:
        s = float(lengths[i]) / sample_rate
        if cur and (cur_s + s > max_audio_seconds or len(cur) >= max_utts):
            batches.append(cur)
            cur, cur_s = [], 0.0
        cur.append(int(i))
        cur_s += s
    if cur:
        batches.append(cur)
    return batches


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def extract_embeddings_sharded(extract_fn, waves, embedding_size, max_audio_seconds=1200.0, device=None):
    """Embeddings of ``waves`` (list of 1-D tensors) on every rank, in input order.

    ``extract_fn(list_of_waves) -> (n, E) tensor`` is ``model.extract_varlen`` in production.  Each rank extracts
    its shard in length-bucketed packed batches; one ``all_gather`` of the padded (n_max, E) blocks plus the index
    vectors restores the global order.  With a single process this is just the batched local extraction.
    """
    dist, rank, world = _dist()
    lengths = numpy.array([int(w.shape[-1]) for w in waves], dtype=numpy.int64)
    shards = plan_shards(lengths, world)
    mine = shards[rank]
    out_dev = device
    batches = make_batches(mine, lengths, max_audio_seconds)
    parts = [None] * len(batches)
    for k in reversed(range(len(batches))):        # largest batch first: the engine sizes its work buffers once
        e = extract_fn([waves[i] for i in batches[k]])
        out_dev = e.device if out_dev is None else out_dev
        parts[k] = e.to(torch.float32)
    if out_dev is None:
        out_dev = torch.device("cpu")
    local = torch.cat(parts) if parts else torch.zeros((0, embedding_size), dtype=torch.float32, device=out_dev)
    n = len(waves)
    result = torch.empty((n, embedding_size), dtype=torch.float32, device=out_dev)
    if world == 1:
        result[torch.as_tensor(mine, device=out_dev)] = local
        return result
    n_max = max(len(s) for s in shards)
    block = torch.zeros((n_max, embedding_size), dtype=torch.float32, device=out_dev)
    block[: local.shape[0]] = local
    gathered = torch.empty((world * n_max, embedding_size), dtype=torch.float32, device=out_dev)
    dist.all_gather_into_tensor(gathered, block)
    for r, s in enumerate(shards):             # shards are a pure function of (lengths, world): no index exchange needed
        result[torch.as_tensor(s, device=out_dev)] = gathered[r * n_max: r * n_max + len(s)]
    return result


def extract_embeddings(ids, waves, model, max_audio_seconds=1200.0):
    """In-memory counterpart of the reference's ``extract_embeddings``: returns a ``StatServer`` whose ``stat1`` holds
    the embeddings and ``stat0`` ones (xvector.py:1905-1914)."""
    emb = extract_embeddings_sharded(lambda ws: model.extract_varlen([w.to(next(model.parameters()).device) for w in ws]),
                                     waves, model.embedding_size, max_audio_seconds)
    return StatServer.from_embeddings(numpy.asarray(ids), emb.cpu().numpy())


def row_panel(n_rows, rank=None, world=None):
    """Contiguous block of enrol rows scored by this rank (score-matrix sharding; no collective)."""
    dist, r, w = _dist()
    rank = r if rank is None else rank
    world = w if world is None else world
    per = (n_rows + world - 1) // world
    return min(rank * per, n_rows), min((rank + 1) * per, n_rows)


def _asnorm_stats_cuda(X, cohort_normalised, topk):
    from . import _lib
    N, D = X.shape
    mean = torch.empty((N,), dtype=torch.float32, device=X.device)
    std = torch.empty((N,), dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        _lib.check(_lib.lib().skb_asnorm_stats(X.data_ptr(), cohort_normalised.data_ptr(), N, cohort_normalised.shape[0], D, int(topk),
                                               mean.data_ptr(), std.data_ptr(), _lib.stream_ptr()))
    return mean, std


def _asnorm_panel_cuda(X, lo, hi, mean, std):
    from . import _lib
    N, D = X.shape
    out = torch.empty((hi - lo, N), dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        _lib.check(_lib.lib().skb_asnorm_apply_panel(X.data_ptr(), N, D, lo, hi - lo, mean.data_ptr(), std.data_ptr(), out.data_ptr(),
                                                     out.stride(0), _lib.stream_ptr()))
    return out


def asnorm_sharded(enrol_xv, cohort_xv, topk=200, stats_fn=_asnorm_stats_cuda, panel_fn=_asnorm_panel_cuda):
    """Adaptive s-norm across ranks: every rank computes the top-k cohort statistics of ITS row panel,
    one ``all_gather`` shares the (N,) mean / std vectors (the symmetric formula needs mu_j, sigma_j of every column),
    then every rank normalises its rows.  Returns ``(row_begin, row_end, panel)`` with ``panel`` the (rows, N) block of
    the matrix ``asnorm`` returns; the result stays row-sharded.  ``stats_fn`` / ``panel_fn`` are the CUDA routines in
    production and injectable so the exchange can be tested on CPU."""
    dist, rank, world = _dist()
    X = enrol_xv.to(torch.float32).contiguous()
    coh = torch.nn.functional.normalize(cohort_xv.to(X.device, torch.float32), dim=1).contiguous()
    N = X.shape[0]
    lo, hi = row_panel(N, rank, world)
    per = (N + world - 1) // world
    mean = torch.zeros((world * per,), dtype=torch.float32, device=X.device)
    std = torch.ones((world * per,), dtype=torch.float32, device=X.device)
    if hi > lo:
        m, s = stats_fn(X[lo:hi].contiguous(), coh, topk)
        mean[lo:hi], std[lo:hi] = m, s
    if world > 1:
        both = torch.stack([mean[rank * per:(rank + 1) * per], std[rank * per:(rank + 1) * per]]).contiguous()
        gathered = torch.empty((world * 2, per), dtype=torch.float32, device=X.device)
        dist.all_gather_into_tensor(gathered, both)
        gathered = gathered.view(world, 2, per)
        mean, std = gathered[:, 0, :].reshape(-1).contiguous(), gathered[:, 1, :].reshape(-1).contiguous()
    mean, std = mean[:N].contiguous(), std[:N].contiguous()
    panel = panel_fn(X, lo, hi, mean, std) if hi > lo else torch.zeros((0, N), dtype=torch.float32, device=X.device)
    return lo, hi, panel
