#!/usr/bin/env python
"""Benchmark of the speaker-verification inference hot path (BASELINE.json metric:
audio-seconds embedded per second + PLDA trials per second, with % of roofline).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU implementation

One JSON line on stdout.

Headline = BASELINE config 4, the real sharded path: ONE pool of ``N * K * utts`` utterances (2-20 s, ``default_rng(5)``,
audio generated on the device, generation untimed) is sharded over the N ranks by ``bulk.plan_shards`` (length-sorted,
MAC-balanced), every rank extracts its shard in K length-bucketed packed batches of equal MAC count (a "step" = one
batch per GPU, every batch a never-seen length composition, so the geometry planning is inside the timed region), and
ONE NCCL all-gather of the embeddings ends the timed region.  ``value`` = all audio-seconds / device time (max over
ranks); ``e2e`` = the same job from pinned HOST batches through ``Xtractor.extract_stream`` (every batch's H2D and the
D2H of its embeddings inside the timed region).  ``extra`` carries the other configs of BASELINE.json (TDNN extraction,
20k x 20k PLDA scoring, the as-norm exchange, the vox1-O-shaped pipeline) with their own rooflines.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SPK = 7205                 # egs/voxceleb12_train/cfg/model.yaml:3
WORKLOAD = ("HalfResNet34 x-vector extraction (256-d, random init) of 16 kHz utterances 2-20 s drawn with default_rng(5) "
            "(BASELINE config 4 law), sharded by length bucket")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def load_traffic(kernel):
    """Measured DRAM bytes per launch of a kernel (ncu --set full, summarised in profiles/traffic.json) or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return float(json.load(open(path))[kernel]["bytes_per_launch"])
    except Exception:
        return None


def hr34_macs(L):
    """Algorithmic MACs of one HalfResNet34 embedding of L samples (SURVEY.md 8d closed form): (trunk, whole network)."""
    T1 = 1 + L // 160
    T2 = (T1 - 1) // 2 + 1
    T3 = (T2 - 1) // 2 + 1
    T4 = (T3 - 1) // 2 + 1
    trunk = 80 * T1 * 56608 + 40 * T2 * 278528 + 20 * T3 * 1703936 + 10 * T4 * 3276800
    return trunk, trunk + 1310720 * T4 + 1350016


def tdnn_macs(L):
    T = 1 + L // 512
    return 204800 * (T - 4) + 786432 * (T - 8) + 786432 * (T - 14) + 262144 * (T - 14) + 786432 * (T - 14) + 1572864


def config4_lengths(n, seed=5):
    """L_i = round(16000 * U[2, 20]) with numpy default_rng(5)."""
    return numpy.round(16000 * numpy.random.default_rng(seed).uniform(2.0, 20.0, size=n)).astype(numpy.int64)


def device_audio(lengths, seed, device):
    """Counter-based Gaussian audio x 0.1 generated ON THE DEVICE (704 GB of config-4 audio cannot be staged)."""
    g = torch.Generator(device=device).manual_seed(int(seed))
    return torch.randn(int(sum(lengths)), generator=g, device=device) * 0.1


def make_batches(n_batches, n_utt, seed, lo_s, hi_s, device):
    """Rotating synthetic batches (lengths ~ U[lo, hi] s sorted inside the batch) for the side measurements."""
    from sidekit_b200 import synth
    out = []
    for i in range(n_batches):
        lengths = numpy.sort(synth.synth_lengths(n_utt, lo_s, hi_s, seed=seed + i))
        g = torch.Generator(device="cpu").manual_seed(seed * 1000 + i)
        flat = torch.randn(int(lengths.sum()), generator=g) * 0.1
        out.append((flat.pin_memory(), flat.to(device), [int(v) for v in lengths]))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Start of the timed region: only samples taken from here on are reported.  (The sampler itself is started well before:
        nvidia-smi's start-up takes the driver lock for tens of milliseconds, which would otherwise land in the timed region.)"""
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        rows = self.rows[max(0, getattr(self, "first", 0) - 1):]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(numpy.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model(archi, emb, device):
    import contextlib, io
    from sidekit_b200 import synth
    from sidekit_b200.nnet import Xtractor
    with contextlib.redirect_stdout(io.StringIO()):
        m = Xtractor(N_SPK, archi, loss="aam", embedding_size=emb)
    sd = m.state_dict()
    synth.fill_state_dict(sd, 0)
    m.load_state_dict(sd)
    return m.eval().to(device)


def timed(fn, steps, dist_on):
    """`steps` calls bracketed by barrier + synchronize; device time from CUDA events; max over ranks."""
    import torch.distributed as dist
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if dist_on:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


def profile_categories(lib, fn, steps):
    """Per-category device time (CUDA events around each kernel group, skb_profile_*) of `steps` calls of fn."""
    lib.skb_profile_enable(1)
    for i in range(steps):
        fn(i)
    cat = numpy.zeros(8, dtype=numpy.float32)
    lib.skb_profile_read(cat.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 8)
    lib.skb_profile_enable(0)
    return cat


def roofline(kernel, bound, achieved, peaks, traffic=None, **more):
    """`frac` against the SUSTAINED peak (the kernels are timed inside a long step); `frac_burst` against the burst figure."""
    if bound == "tensor":
        peak, burst, unit = peaks["tf_sustained"], peaks["tf_burst"], "TFLOP/s"
    else:
        peak, burst, unit = peaks["hbm_gbs"], peaks["hbm_gbs"], "GB/s"
    out = {"kernel": kernel, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
           "peak_burst": burst, "frac_burst": achieved / burst, "traffic": traffic,
           "peak_source": peaks["source"] + (" (sustained bf16; burst beside it)" if bound == "tensor" else " (copy bandwidth)")}
    out.update(more)
    return out


# --------------------------------------------------------------------------------------------- reference arm (CPU)
def _reference_model():
    """(model, kind): the UNMODIFIED reference's Xtractor (oracle/_ref or /root/reference, patches P1/P2 of SURVEY.md 8c)
    with the synthetic weights, or None when it cannot be imported here (then the oracle port is timed)."""
    try:
        from oracle import ref_import
        if not ref_import.available():
            return None
        from tests.models import synthetic_state_dict
        m = ref_import.build_xtractor(N_SPK, "halfresnet34", 256)
        sd = synthetic_state_dict("halfresnet34", N_SPK, 256)
        m.load_state_dict({k: v for k, v in sd.items()}, strict=True)
        return m.eval()
    except Exception as e:                                    # pragma: no cover
        print("reference import failed: %r" % (e,), file=sys.stderr)
        return None


def cpu_reference_extraction(budget_s, n_threads, model=None, offset=0):
    """The reference's CPU implementation on a bounded sample of the same workload: batch-1 loop over config-4
    utterances (what extract_embeddings / extract_xvectors.py do, xvector.py:1839-1893), fp32, all host threads."""
    from sidekit_b200 import synth
    torch.set_num_threads(n_threads)
    lengths = config4_lengths(4096)
    done_s, n = 0.0, 0
    if model is None:
        from oracle import extract_ref as R
        from tests.models import synthetic_state_dict
        if not hasattr(cpu_reference_extraction, "sd"):
            cpu_reference_extraction.sd = synthetic_state_dict("halfresnet34", N_SPK, 256)
        sd = cpu_reference_extraction.sd
        fwd = lambda w: R.forward(sd, w, "halfresnet34")
        kind = "port"
    else:
        fwd = lambda w: model(w, is_eval=True)
        kind = "reference"
    with torch.no_grad():
        fwd(synth.synth_wave(1, 32000, seed=1))                                 # warm-up
        t0 = time.perf_counter()
        while True:
            L = int(lengths[(offset + n) % len(lengths)])
            fwd(synth.synth_wave(1, L, seed=10 + offset + n))
            done_s += L / 16000.0
            n += 1
            if time.perf_counter() - t0 > budget_s:
                break
    dt = time.perf_counter() - t0
    return done_s / dt, dt * 1e3, kind, "%d utterances (%.0f audio-s) of the config-4 law, batch-1 loop, %.1f s of CPU" % (n, done_s, dt)


def gpu_eager_reference_extraction(model, budget_s):
    """SURVEY.md 8d "reference GPU path" / BASELINE.md 1 "stock PyTorch eager on the same B200": the reference model (or
    its restatement) as stock PyTorch eager ops on this GPU (cuDNN / cuBLAS / cuFFT) -- (a) the batch-1 loop with a D2H per
    utterance that the reference's extractors run, fp32 and fp16 autocast; (b) BATCHED 64 x 4 s (config 1), fp16 autocast,
    channels_last (the reference's trunk converts to channels_last itself), best case for eager.  None of this repo's
    kernels run here."""
    from sidekit_b200 import synth
    dev = torch.device("cuda", 0)
    if model is None:
        from oracle import extract_ref as R
        from tests.models import synthetic_state_dict
        sd = {k: v.to(dev) for k, v in synthetic_state_dict("halfresnet34", N_SPK, 256).items()}
        fwd = lambda w: R.forward(sd, w, "halfresnet34")[1]
    else:
        model = model.to(dev)
        fwd = lambda w: model(w, is_eval=True)[1]
    lengths = config4_lengths(64)
    waves = [synth.synth_wave(1, int(L), seed=10 + i).to(dev) for i, L in enumerate(lengths)]
    batch = synth.synth_wave(64, 64000, seed=3).to(dev)
    out = {}
    for name, amp in (("batch1_fp32", False), ("batch1_fp16_autocast", True)):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            for w in waves[:3]:
                fwd(w)
            torch.cuda.synchronize()
            t0, done, n = time.perf_counter(), 0.0, 0
            while time.perf_counter() - t0 < budget_s:
                w = waves[n % len(waves)]
                fwd(w).cpu()                                               # per-utterance D2H like extract_xvectors.py
                done += w.shape[1] / 16000.0
                n += 1
            dt = time.perf_counter() - t0
        out[name] = {"value": done / dt, "unit": "audio-s/s", "sample": "%d utterances, batch-1 loop, %.1f s" % (n, dt)}
    for name, amp in (("batched_64x4s_fp32", False), ("batched_64x4s_fp16_autocast_channels_last", True)):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            for _ in range(3):
                fwd(batch)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 10
            for _ in range(n):
                fwd(batch)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
        out[name] = {"value": 64 * 4.0 / (ms / 1e3), "unit": "audio-s/s", "ms_per_batch": ms,
                     "sample": "64 x 4 s per forward (BASELINE config 1 shape), %d forwards, device-timed" % n}
    if model is not None:
        model.to("cpu")
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cores = os.cpu_count() or 1
    model = _reference_model()
    vals, ms, sample, kind = [], [], "", "port"
    budget = max(3.0, 45.0 / max(1, args.steps + args.warmup))
    for i in range(args.warmup + args.steps):
        v, step_ms, kind, sample = cpu_reference_extraction(budget, cores, model, offset=64 * i)
        if i >= args.warmup:
            vals.append(v)
            ms.append(step_ms)
    value = float(numpy.mean(vals))
    line = {"impl": "reference", "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(numpy.mean(ms)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "step": "a bounded sample of that workload per step (%s)" % sample},
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    extra = {}
    try:
        v64, _, _, s64 = cpu_reference_batched(model, cores)
        extra["cpu_batched_64x4s"] = {"value": v64, "unit": "audio-s/s", "sample": s64}
    except Exception as e:                                        # informational only
        extra["cpu_batched_64x4s"] = {"unavailable": repr(e)[:200]}
    if torch.cuda.is_available():
        try:
            extra["torch_eager_on_this_gpu"] = gpu_eager_reference_extraction(model, 3.0)
        except Exception as e:                                    # informational only
            extra["torch_eager_on_this_gpu"] = {"unavailable": repr(e)[:200]}
    line["extra"] = extra
    return line


def cpu_reference_batched(model, cores):
    """BASELINE config 1 as the reference would run it when handed a batch: 64 x 4 s in batches of 8 on the host cores."""
    from sidekit_b200 import synth
    torch.set_num_threads(cores)
    if model is None:
        from oracle import extract_ref as R
        from tests.models import synthetic_state_dict
        sd = synthetic_state_dict("halfresnet34", N_SPK, 256)
        fwd = lambda w: R.forward(sd, w, "halfresnet34")
    else:
        fwd = lambda w: model(w, is_eval=True)
    x = synth.synth_wave(16, 64000, seed=3)
    with torch.no_grad():
        fwd(x[:8])
        t0 = time.perf_counter()
        fwd(x[:8]); fwd(x[8:])
        dt = time.perf_counter() - t0
    return 16 * 4.0 / dt, dt * 1e3, None, "16 of the 64 x 4 s utterances in batches of 8, %.1f s of CPU" % dt


# --------------------------------------------------------------------------------------------- this repo's arm
def run_ours(args):
    import torch.distributed as dist
    from sidekit_b200 import _lib, bulk
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist_on = world > 1
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if dist_on:
        dist.init_process_group("nccl", device_id=device)
    peaks = load_peaks()
    lib = _lib.lib()
    K, W = args.steps, max(3, args.warmup)

    # ------------------------------------------------------------------ HalfResNet34 extraction, config 4 (headline)
    model = build_model("halfresnet34", 256, device)
    n_total = world * K * args.utts
    lengths = config4_lengths(n_total)
    shards = bulk.plan_shards(lengths, world)
    mine = shards[rank]
    batches = bulk.make_batches_equal_cost(mine, lengths, K)
    blens = [[int(lengths[i]) for i in b] for b in batches]
    offs = numpy.concatenate([[0], numpy.cumsum([len(b) for b in batches])]).astype(numpy.int64)
    flats = [device_audio(bl, 777000 + 1000 * rank + k, device) for k, bl in enumerate(blens)]       # generation untimed
    local_emb = torch.empty((len(mine), 256), dtype=torch.float32, device=device)
    # warm-up on OTHER utterances of the same law (6 % more audio per batch, so every work buffer is sized here)
    wl = config4_lengths(int(W * args.utts * 1.06), seed=9000 + rank)
    wb = bulk.make_batches_equal_cost(numpy.argsort(wl, kind="stable"), wl, W)
    order = list(reversed(range(K)))                        # longest utterances first, as extract_embeddings_sharded does

    step_events = []

    def job_dev(_):
        for k in order:
            model.extract_packed(flats[k], blens[k], out=local_emb[offs[k]:offs[k + 1]])
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            step_events.append(ev)
        return bulk.gather_embeddings(local_emb, shards, 256, device)          # the one collective of the path

    # ONE sampler per job (rank 0's GPU): every nvidia-smi poll takes the driver lock, and eight of them at 10 Hz beside eight
    # ranks are a perturbation of their own
    sampler = ClockSampler(local)
    if not args.no_clock_sampler and rank == 0:
        sampler.start()
    with torch.no_grad():
        # the bulk run's batch budget is known up front (what bulk.make_batches is given): every work buffer and plan-cache
        # slot is allocated now, none inside the run
        model.reserve(int(1.1 * max(len(b) for b in blens + [wb[k] for k in range(W)])),
                      1.1 * max(sum(bl) for bl in blens) / 16000.0, device)
        for k in reversed(range(W)):
            wls = [int(wl[i]) for i in wb[k]]
            model.extract_packed(device_audio(wls, 5 + k, device), wls)
        bulk.gather_embeddings(local_emb, shards, 256, device)     # warm-up of the collective / the scatter kernels too
        # nvidia-smi must be UP before the timed region (its start-up holds the driver lock for tens of milliseconds: 9.3 instead
        # of 8.5 ms per step when it lands inside), and the GPU must not idle while we wait for it (the clocks would drop):
        # keep running warm-up batches until the sampler has delivered its first rows
        # (all ranks keep their GPUs busy until rank 0's sampler is up)
        t_wait = time.perf_counter()
        while True:
            ready = 0 if (sampler.proc is not None and len(sampler.rows) < 2 and time.perf_counter() - t_wait < 5.0) else 1
            if dist_on:
                flag = torch.tensor([ready], device=device, dtype=torch.int32)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                ready = int(flag.item())
            if ready:
                break
            wls = [int(wl[i]) for i in wb[0]]
            model.extract_packed(device_audio(wls, 5, device), wls)
            torch.cuda.synchronize()
        sampler.mark()
        l0 = lib.skb_kernel_launches()
        ms = timed(job_dev, 1, dist_on)
        launches = lib.skb_kernel_launches() - l0
        clocks = sampler.stop()
        step_ms = [step_events[i].elapsed_time(step_events[i + 1]) for i in range(len(step_events) - 1)]
        model.check_overflow()
        # end to end: the same job from pinned HOST batches -- every batch's waveforms cross PCIe inside the timed region
        # (overlapped with the previous batch's compute on a second stream), its embeddings come back to the host, the
        # shard's embeddings are all-gathered and rank 0 reads the (N, 256) result
        host_batches = [(flats[k].cpu().pin_memory(), blens[k]) for k in order]
        e2e_out = torch.empty_like(local_emb)

        def job_e2e(_):
            model.extract_stream(host_batches, device_out=e2e_out)
            allemb = bulk.gather_embeddings(e2e_out, shards, 256, device)
            if rank == 0 and dist_on:
                allemb.cpu()

        big = sorted(range(K), key=lambda i: -host_batches[i][0].numel())[:2]        # the two largest batches size the staging buffers
        model.extract_stream([host_batches[i] for i in big] + host_batches[:1], device_out=e2e_out[:sum(len(host_batches[i][1]) for i in big) + len(host_batches[0][1])])
        model._out_host = torch.empty((len(mine), 256), dtype=torch.float32, pin_memory=True)
        ms_e2e = timed(job_e2e, 1, dist_on)
        # per-category device time (separate pass with event brackets) for the roofline of the dominant kernel
        cat = profile_categories(lib, lambda i: model.extract_packed(flats[order[i]], blens[order[i]]), K)
    my_audio = float(sum(sum(bl) for bl in blens)) / 16000.0
    my_trunk = float(sum(hr34_macs(L)[0] for bl in blens for L in bl))
    my_all = float(sum(hr34_macs(L)[1] for bl in blens for L in bl))
    tot = torch.tensor([my_audio, my_all, my_trunk], device=device, dtype=torch.float64)
    mx = tot.clone()
    if dist_on:
        dist.all_reduce(tot)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    all_audio, all_macs = tot[0].item(), tot[1].item()
    value = all_audio / (ms / 1e3)
    e2e_value = all_audio / (ms_e2e / 1e3)
    conv_ms = float(cat[2])
    conv_tf = 2.0 * my_trunk / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    n_conv_launches = K * 36
    rl = roofline("conv_umma_kernel (tcgen05 shift-GEMM conv, 36 launches per step)", "tensor", conv_tf, peaks,
                  traffic=load_traffic("conv_umma_kernel"),
                  traffic_unit="bytes per launch (ncu dram read+write, profiles/traffic.json)",
                  avg_launch_ms=conv_ms / n_conv_launches if n_conv_launches else None,
                  flops_per_launch=2.0 * my_trunk / n_conv_launches,
                  whole_step_tflops_per_gpu=2.0 * all_macs / world / (ms / 1e3) / 1e12,
                  whole_step_frac_sustained=2.0 * all_macs / world / (ms / 1e3) / 1e12 / peaks["tf_sustained"],
                  whole_step_frac_burst=2.0 * all_macs / world / (ms / 1e3) / 1e12 / peaks["tf_burst"],
                  mac_balance_max_over_mean=mx[1].item() / (all_macs / world),
                  device_ms_by_category={"frontend": float(cat[0]), "stem": float(cat[1]), "conv": conv_ms, "se": float(cat[3]),
                                         "pooling_head": float(cat[4])})
    fe_bytes = float(sum(4 * L + 80 * (1 + L // 160) * 4 for bl in blens for L in bl))
    rl["frontend_roofline"] = roofline("frontend_kernel<512,8> + cmvn (log-Mel)", "hbm", fe_bytes / (float(cat[0]) / 1e3) / 1e9 if cat[0] > 0 else 0.0,
                                       peaks, traffic=load_traffic("frontend_kernel_logmel"))
    h2d = int(numpy.mean([b[0].numel() * 4 for b in host_batches]))
    d2h = int(numpy.mean([len(bl) for bl in blens]) * 256 * 4)
    del host_batches, flats

    extra = {}
    with torch.no_grad():
        extra.update(run_extras(args, device, peaks, dist_on, rank, world, lib))
        if world == 1:
            extra.update(run_single_gpu_extras(args, model, device, peaks))

    if rank == 0:
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, _, kind, sample = cpu_reference_extraction(15.0, cores, _reference_model())
            cpu = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample}
        line = {"impl": "ours", "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": WORKLOAD,
                           "step": "one packed batch per GPU: the pool of n_gpus x steps x %d utterances is sharded with bulk.plan_shards "
                                   "(MAC-balanced, length-sorted) and each shard cut into `steps` equal-MAC length buckets; one NCCL "
                                   "all-gather of the embeddings closes the timed region" % args.utts,
                           "utterances": int(n_total), "audio_s_per_step_per_gpu": all_audio / world / K,
                           "utterances_per_step_rank0": [len(blens[k]) for k in order],
                           "device_ms_between_steps_rank0": [round(v, 3) for v in step_ms],
                           "accumulate": "fp32", "parallelism": "utterance-sharded x%d, one NCCL all_gather of the embeddings per shard" % world,
                           "l2": "every batch is new data (%.0f MB of audio per step) and every step streams >1 GB of activations, so "
                                 "nothing survives in the 126 MB L2 between steps; every batch is also a never-seen length composition "
                                 "(its geometry plan is built inside the timed region)" % (h2d / 1e6)},
                "roofline": rl, "cpu_baseline": cpu, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / K},
                "gpu_launches": int(launches), "extra": extra}
    else:
        line = None
    if dist_on:
        dist.barrier()
        dist.destroy_process_group()
    return line


def run_extras(args, device, peaks, dist_on, rank, world, lib):
    """BASELINE configs 2 (TDNN extraction), 3 (20k x 20k PLDA scoring) and the as-norm exchange of config 5, on every
    rank: value + roofline each."""
    import torch.distributed as dist
    import sidekit_b200 as sk
    from sidekit_b200 import synth, bulk
    from sidekit_b200.iv_scoring import PackedEmbeddings
    out = {}
    # ---- TDNN 'xvector' (512-d), config 2 law (default_rng(4)): front-end dominated, HBM-bound as a whole
    tdnn = build_model("xvector", 512, device)
    tl = config4_lengths(2 * 512 * world, seed=4)
    tb = []
    for i in range(2):
        ls = numpy.sort(tl[(2 * rank + i) * 512:(2 * rank + i + 1) * 512])
        tb.append((device_audio(ls, 4400 + 2 * rank + i, device), [int(v) for v in ls]))
    f = lambda i: tdnn.extract_packed(tb[i % 2][0], tb[i % 2][1])
    for i in range(3):
        f(i)
    ms = timed(f, args.steps, dist_on)
    cat = profile_categories(lib, f, args.steps)
    aud = sum(sum(tb[i % 2][1]) for i in range(args.steps)) / 16000.0 * world
    bytes_alg = float(sum(4 * L + 80 * (1 + L // 512) * 4 for i in range(args.steps) for L in tb[i % 2][1]))
    macs = float(sum(tdnn_macs(L) for i in range(args.steps) for L in tb[i % 2][1]))
    out["tdnn_xvector"] = {
        "metric": "audio_seconds_per_second", "value": aud / (ms / 1e3), "unit": "audio-s/s",
        "workload": "TDNN xvector 512-d, 512 utterances 2-20 s (default_rng(4)) per GPU per step (BASELINE config 2 shape)",
        "ms_per_step": ms / args.steps,
        "roofline": roofline("whole TDNN step (SURVEY 8d assigns config 2 the HBM roofline: 4 L bytes read + 80 T 4 bytes of features)",
                             "hbm", bytes_alg / (ms / 1e3) / 1e9, peaks, tensor_tflops=2.0 * macs / (ms / 1e3) / 1e12,
                             tensor_frac_sustained=2.0 * macs / (ms / 1e3) / 1e12 / peaks["tf_sustained"]),
        "frontend_roofline": roofline("frontend_kernel<1024,16> + cmvn (MFCC)", "hbm",
                                      bytes_alg / (float(cat[0]) / 1e3) / 1e9 if cat[0] > 0 else 0.0, peaks,
                                      traffic=load_traffic("frontend_kernel_mfcc")),
        "conv_roofline": roofline("conv_umma_kernel (5 TDNN layers)", "tensor",
                                  2.0 * (macs - 1572864.0 * 512 * args.steps) / (float(cat[2]) / 1e3) / 1e12 if cat[2] > 0 else 0.0, peaks),
        "device_ms_by_category": {"frontend": float(cat[0]), "conv": float(cat[2]), "pooling_head": float(cat[4])}}
    del tdnn, tb
    # ---- PLDA-form scoring, 20k x 20k x 256 (config 3).  The embeddings arrive row-sharded (as extraction leaves them);
    # the test side is all-gathered over NVLink and packed ONCE (one-time cost, reported), then every step scores this
    # rank's enrol row panel against it: pack of the panel + GEMM, no collective in the step.
    Ne = Nt = 20000
    D = 256
    lo, hi = bulk.row_panel(Ne, rank, world)
    rows = hi - lo
    E = torch.from_numpy(synth.synth_embeddings(Ne, D, seed=6)).float()[lo:hi].to(device)
    T_local = torch.from_numpy(synth.synth_embeddings(Nt, D, seed=7)).float()[lo:hi].to(device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    T_all = bulk.gather_rows(T_local, Nt, device).contiguous()
    Tp = PackedEmbeddings(T_all)
    torch.cuda.synchronize()
    prep_ms = (time.perf_counter() - t0) * 1e3
    r = torch.randn(rows, device=device)
    q = torch.randn(Nt, device=device)
    outm = torch.empty((rows, Nt), dtype=torch.float32, device=device)
    g = lambda i: sk.score_matrix(E, Tp, r, q, cst=0.5, alpha=1.0, passes=0, out=outm)
    for i in range(3):
        g(i)
    ms = timed(g, args.steps, dist_on)
    trials = float(Ne) * Nt * args.steps
    gbs = (rows * Nt * 4 + 4 * (rows + Nt) * D) * args.steps / (ms / 1e3) / 1e9
    out["plda_20k"] = {"metric": "trials_per_second", "value": trials / (ms / 1e3), "unit": "trials/s", "scaling": "strong",
                       "workload": "PLDA-form scoring 20k x 20k x 256 (BASELINE config 3), fp32 score panel resident in HBM, enrol rows "
                                   "sharded over %d GPU(s), test side all-gathered + packed once (%.2f ms, outside the step)" % (world, prep_ms),
                       "ms_per_step": ms / args.steps, "test_prepare_ms": prep_ms,
                       "roofline": roofline("score_gemm_kernel (+ absmax / pack_split of the enrol panel)", "hbm", gbs, peaks,
                                            traffic=load_traffic("score_gemm_kernel") if world == 1 else None,
                                            tensor_tflops=2.0 * rows * Nt * D * args.steps / (ms / 1e3) / 1e12,
                                            tensor_frac_sustained=2.0 * rows * Nt * D * args.steps / (ms / 1e3) / 1e12 / peaks["tf_sustained"])}
    if world == 1:
        # consumers that keep the matrix out of HBM: (i) trial-list mode -- only the trials of a sparse mask
        # are written (config-5 density: 37 720 trials per 4 708^2 pairs), the GEMM is then bound by the tensor pipe / the L2
        # operand stream; (ii) float16 output -- half the write
        from sidekit_b200.iv_scoring import TrialIndex, score_trials
        gm = torch.Generator(device=device).manual_seed(99)
        mask = torch.rand((Ne, Nt), device=device, generator=gm) < (37720.0 / 4708.0 ** 2)
        t0 = time.perf_counter()
        tidx = TrialIndex(mask, device)
        torch.cuda.synchronize()
        tidx_ms = (time.perf_counter() - t0) * 1e3
        del mask
        gt = lambda i: score_trials(E, Tp, tidx, r, q, cst=0.5, alpha=1.0, passes=0)       # test side packed once, like the matrix modes
        for i in range(3):
            gt(i)
        ms_t = timed(gt, args.steps, False)
        tf = 2.0 * Ne * Nt * D * args.steps / (ms_t / 1e3) / 1e12
        out["plda_20k_trial_list"] = {"metric": "trials_per_second", "value": float(Ne) * Nt * args.steps / (ms_t / 1e3), "unit": "trials/s",
                                      "workload": "the same 20k x 20k GEMM (test side packed once), only the %d trials of a sparse mask written "
                                                  "(row-major, as scoremat[trialmask]); trial index built once in %.2f ms" % (tidx.n_trials, tidx_ms),
                                      "ms_per_step": ms_t / args.steps,
                                      "roofline": roofline("score_gemm_kernel, trial-list epilogue (+ operand packing)", "tensor", tf, peaks)}
        out16 = torch.empty((rows, Nt), dtype=torch.float16, device=device)
        gh = lambda i: sk.score_matrix(E, Tp, r, q, cst=0.5, alpha=1.0, passes=0, out=out16)
        for i in range(3):
            gh(i)
        ms_h = timed(gh, args.steps, False)
        gbs_h = (rows * Nt * 2 + 4 * (rows + Nt) * D) * args.steps / (ms_h / 1e3) / 1e9
        out["plda_20k_fp16_out"] = {"metric": "trials_per_second", "value": trials / (ms_h / 1e3), "unit": "trials/s",
                                    "workload": "the same 20k x 20k scoring with a float16 score matrix (2 bytes per trial)",
                                    "ms_per_step": ms_h / args.steps,
                                    "roofline": roofline("score_gemm_kernel, float16 epilogue", "hbm", gbs_h, peaks,
                                                         tensor_tflops=2.0 * rows * Nt * D * args.steps / (ms_h / 1e3) / 1e12,
                                                         tensor_frac_sustained=2.0 * rows * Nt * D * args.steps / (ms_h / 1e3) / 1e12 / peaks["tf_sustained"])}
        del out16, tidx
    del outm, Tp, T_all
    # ---- as-norm across ranks (config 5 shape: N = 4708 unit-norm embeddings, cohort 7205): top-200 cohort statistics of
    # this rank's rows, ONE all-gather of the (N,) mean / std vectors, then the rank's rows of the normalised matrix
    X = torch.nn.functional.normalize(torch.from_numpy(synth.synth_embeddings(4708, D, seed=21, unit_norm=False)).float(), dim=1).to(device)
    coh = torch.from_numpy(synth.synth_embeddings(N_SPK, D, seed=22, unit_norm=False)).float().to(device)
    h = lambda i: bulk.asnorm_sharded(X, coh, 200)
    for i in range(3):
        h(i)
    ms = timed(h, args.steps, dist_on)
    out["asnorm_sharded"] = {"metric": "trials_per_second", "value": 4708.0 * 4708.0 * args.steps / (ms / 1e3), "unit": "trials/s",
                             "scaling": "strong", "ms_per_step": ms / args.steps,
                             "workload": "as-norm of 4708 x 4708 scores, cohort 7205, top-200 (BASELINE config 5 shape), row panels over "
                                         "%d GPU(s), one NCCL all-gather of the statistics" % world}
    return out


def run_single_gpu_extras(args, model, device, peaks):
    import sidekit_b200 as sk
    from sidekit_b200 import synth
    out = {}
    # the r01 measurement for continuity: 4 rotating batches whose geometry plans are cached after the warm-up
    rot = make_batches(4, args.utts, seed=500, lo_s=2.0, hi_s=20.0, device=device)
    f = lambda i: model.extract_packed(rot[i % 4][1], rot[i % 4][2])
    for i in range(4):
        f(i)
    ms = timed(f, args.steps, False)
    aud = sum(sum(rot[i % 4][2]) for i in range(args.steps)) / 16000.0
    out["cached_geometry"] = {"metric": "audio_seconds_per_second", "unit": "audio-s/s", "value": aud / (ms / 1e3),
                              "ms_per_step": ms / args.steps,
                              "workload": "round-1 headline workload: 4 rotating batches of %d utterances 2-20 s (mixed lengths inside a "
                                          "batch), geometry plans cached" % args.utts}
    del rot
    # feed path: 44.1 kHz files brought to the model's 16 kHz on the device (torchaudio.transforms.Resample of
    # xsets.py:435 / extract_xvectors.py:144), 256 utterances 2-20 s; HBM-bound: 4 B read + 4 B written per sample
    from sidekit_b200.nnet.preprocessor import Resample
    rs = Resample(44100, 16000)
    rl = numpy.round(44100 * numpy.random.default_rng(11).uniform(2.0, 20.0, size=256)).astype(numpy.int64)
    rx = [torch.randn(int(rl.sum()), device=device) * 0.1 for _ in range(2)]          # 2 x 0.5 GB: larger than L2
    for i in range(3):
        ry = rs.resample_packed(rx[i % 2], rl)
    ms_r = timed(lambda i: rs.resample_packed(rx[i % 2], rl), args.steps, False)
    gbs = (rx[0].numel() + ry.numel()) * 4 * args.steps / (ms_r / 1e3) / 1e9
    out["resample_44k1_to_16k"] = {"metric": "audio_seconds_per_second", "unit": "audio-s/s",
                                   "value": float(rl.sum()) / 44100.0 * args.steps / (ms_r / 1e3), "ms_per_step": ms_r / args.steps,
                                   "workload": "256 utterances 2-20 s at 44.1 kHz -> 16 kHz, packed ragged batch",
                                   "roofline": roofline("resample_kernel", "hbm", gbs, peaks, traffic=load_traffic("resample_kernel"))}
    del rx, ry
    # the other trunks that run on the HalfResNet34 kernels
    ob = make_batches(2, 32, seed=950, lo_s=2.0, hi_s=20.0, device=device)
    for archi in ("resnet34", "fastresnet34"):
        om = build_model(archi, 256, device)
        fo = lambda i: om.extract_packed(ob[i % 2][1], ob[i % 2][2])
        for i in range(3):
            fo(i)
        ms = timed(fo, args.steps, False)
        aud = sum(sum(ob[i % 2][2]) for i in range(args.steps)) / 16000.0
        out[archi] = {"metric": "audio_seconds_per_second", "value": aud / (ms / 1e3), "unit": "audio-s/s",
                      "workload": "%s 256-d, 32 utterances 2-20 s per step" % archi, "ms_per_step": ms / args.steps}
        del om
    del ob
    # config 3 end to end through the reference-shaped API: StatServer / Ndx in, Scores (float64 numpy on the host) out
    Ne = Nt = 20000
    D = 256
    mu, F, Sigma = synth.synth_plda(D, D, seed=8)
    ids_e = numpy.array(["m%06d" % i for i in range(Ne)])
    ids_t = numpy.array(["s%06d" % i for i in range(Nt)])
    en = sk.StatServer.from_embeddings(ids_e, synth.synth_embeddings(Ne, D, seed=6))
    te = sk.StatServer.from_embeddings(ids_t, synth.synth_embeddings(Nt, D, seed=7))
    ndx = sk.Ndx()
    ndx.modelset, ndx.segset, ndx.trialmask = ids_e, ids_t, numpy.ones((Ne, Nt), dtype=bool)
    runs = []
    for rep in range(3):                                   # the first call also pays one-time pinned-pool / workspace allocations
        t0 = time.perf_counter()
        sc = sk.PLDA_scoring(en, te, ndx, mu, F, numpy.zeros((D, 0)), Sigma)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        mat = sc.scoremat
        t2 = time.perf_counter()
        runs.append({"api_call_s": t1 - t0, "d2h_s": t2 - t1, "total_s": t2 - t0, "dtype": str(mat.dtype), "bytes": int(mat.nbytes)})
        del sc, mat
    best = min(runs[1:], key=lambda r: r["total_s"])
    out["plda_20k_api"] = {"metric": "trials_per_second", "value": float(Ne) * Nt / best["total_s"], "unit": "trials/s",
                           "workload": "sk.PLDA_scoring(StatServer 20k, StatServer 20k, Ndx all-true) -> Scores.scoremat on the host "
                                       "(id matching, D x D algebra, H2D, scoring, D2H)", "first_call": runs[0], "steady_state": best}
    if not args.no_cpu_baseline:
        out["plda_20k_api"]["cpu_baseline"] = cpu_reference_plda(20000, 256)
        out["vox1o_pipeline"] = run_vox1o_pipeline(device)
    return out


def cpu_reference_plda(n, D):
    """The reference's own fast-PLDA scoring (numpy float64, id matching included) at the full n x n size; the oracle port
    when the reference is not importable here."""
    from sidekit_b200 import synth
    E, T = synth.synth_embeddings(n, D, seed=6), synth.synth_embeddings(n, D, seed=7)
    mu, F, Sigma = synth.synth_plda(D, D, seed=8)
    ids_e = numpy.array(["m%06d" % i for i in range(n)])
    ids_t = numpy.array(["s%06d" % i for i in range(n)])
    mask = numpy.ones((n, n), dtype=bool)
    kind = "port"
    try:
        from oracle import ref_import
        if not ref_import.available():
            raise RuntimeError("no reference")
        sidekit = ref_import.import_reference()

        def stat(ids, X):
            s = sidekit.StatServer()
            s.modelset, s.segset = ids.copy(), ids.copy()
            s.start = numpy.empty(len(ids), dtype="|O")
            s.stop = numpy.empty(len(ids), dtype="|O")
            s.stat0, s.stat1 = numpy.ones((len(ids), 1)), X.copy()
            return s
        ndx = sidekit.Ndx()
        ndx.modelset, ndx.segset, ndx.trialmask = ids_e, ids_t, mask
        en, te = stat(ids_e, E), stat(ids_t, T)
        t0 = time.perf_counter()
        sidekit.iv_scoring.PLDA_scoring(en, te, ndx, mu, F, numpy.zeros((D, 0)), Sigma)
        dt = time.perf_counter() - t0
        kind = "reference"
    except Exception as e:
        print("reference PLDA_scoring unavailable (%r): timing the oracle port" % (e,), file=sys.stderr)
        from oracle import scoring_ref as S
        t0 = time.perf_counter()
        S.fast_plda_scoring(ids_e, E, ids_t, T, ids_e, ids_t, mask, mu, F, Sigma)
        dt = time.perf_counter() - t0
    return {"value": n * n / dt, "unit": "trials/s", "cores": os.cpu_count() or 1, "kind": kind,
            "sample": "%d x %d x %d trials (the full config), numpy float64, BLAS threads as configured, %.2f s" % (n, n, D, dt)}


def run_vox1o_pipeline(device):
    """BASELINE config 5: vox1-O-shaped pipeline on one GPU -- 4 708 utterances -> log-Mel -> HalfResNet34 embeddings ->
    cosine + PLDA scoring of 37 720 trials -> as-norm (cohort = the 7 205 margin-head rows) -> EER / minDCF.
    Wall-clock per stage (host + device, synchronised), everything through the public API."""
    import sidekit_b200 as sk
    from sidekit_b200 import synth
    N, n_trials = 4708, 37720
    rng = numpy.random.default_rng(55)
    model = build_model("halfresnet34", 256, device)
    lengths = numpy.sort(synth.synth_lengths(N, 4.0, 12.0, seed=21))[::-1].copy()     # longest first: work buffers are sized once
    audio_s = float(lengths.sum()) / 16000.0
    batches = []
    for lo in range(0, N, 128):                                    # 128 utterances (~1000 audio-s) per packed batch, pinned host memory
        ls = [int(v) for v in lengths[lo:lo + 128]]
        g = torch.Generator(device="cpu").manual_seed(7000 + lo)
        batches.append(((torch.randn(sum(ls), generator=g) * 0.1).pin_memory(), ls))
    out = {"workload": "vox1-O-shaped: %d utterances 4-12 s (%.0f audio-s), %d trials, cohort 7205 (BASELINE config 5)" % (N, audio_s, n_trials)}
    sync = torch.cuda.synchronize
    with torch.no_grad():
        model.extract_stream(batches[:2])                          # warm-up (plans of later batches are built inside the timed call)
        sync(); t0 = time.perf_counter()
        emb = torch.cat(model.extract_stream(batches)).numpy().copy()
        sync(); t1 = time.perf_counter()
    out["extract_s"] = t1 - t0
    out["extract_rate"] = {"value": audio_s / (t1 - t0), "unit": "audio-s/s", "note": "37 distinct batch geometries, each planned once"}
    ids = numpy.array(["utt%05d" % i for i in range(N)])
    # trial list: 18 860 "target" + 18 860 "non-target" pairs over the same 4 708 files (labels are arbitrary for random weights)
    mi, si = rng.integers(0, N, n_trials), rng.integers(0, N, n_trials)
    labels = numpy.where(numpy.arange(n_trials) % 2 == 0, "target", "nontarget")
    t0 = time.perf_counter()
    key = sk.Key(models=ids[mi], testsegs=ids[si], trials=labels)
    ndx = key.to_ndx()
    enroll = sk.StatServer.from_embeddings(ids, emb)
    test = sk.StatServer.from_embeddings(ids, emb)
    out["key_ndx_s"] = time.perf_counter() - t0
    cohort = model.after_speaker_embedding.weight.detach()

    def tail():
        t = {}
        t0 = time.perf_counter()
        cos = sk.cosine_scoring(enroll, test, ndx)
        tar, non = cos.get_tar_non(key)
        sync(); t["cosine_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        S = sk.asnorm(torch.from_numpy(emb).to(device), cohort, None)
        sync(); t["asnorm_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        sn = sk.Scores()
        sn.modelset, sn.segset, sn.scoremat, sn.scoremask = ids, ids, S, numpy.ones(S.shape, dtype=bool)
        atar, anon = sn.get_tar_non(key)
        t["asnorm_get_tar_non_s"] = time.perf_counter() - t0
        res = {"cosine": sk.fast_minDCF(tar, non, numpy.log(0.01 / 0.99), normalize=True),
               "asnorm": sk.fast_minDCF(atar, anon, numpy.log(0.01 / 0.99), normalize=True)}
        t["eer_mindcf_2x_s"] = time.perf_counter() - t0
        return t, tar, non, res

    tail()                                                         # first call: one-time allocations
    t, tar, non, res = tail()
    out.update(t)
    out["tail_s"] = t["cosine_s"] + t["asnorm_s"] + t["eer_mindcf_2x_s"]
    train_ss = sk.StatServer.from_embeddings(numpy.array(["spk%04d" % (i % 1177) for i in range(N)]), emb)
    for name in ("plda_train_first_call_s", "plda_train_s"):       # the first call also pays LAPACK's thread-pool start-up (eigh of the init)
        t0 = time.perf_counter()
        fa = sk.FactorAnalyser().plda(train_ss, 128, nb_iter=5, save_final=False)
        sync(); out[name] = time.perf_counter() - t0
    t0 = time.perf_counter()
    pl = sk.PLDA_scoring(enroll, test, ndx, fa.mean, fa.F, numpy.zeros((256, 0)), fa.Sigma)
    ptar, pnon = pl.get_tar_non(key)
    res["plda"] = sk.fast_minDCF(ptar, pnon, numpy.log(0.01 / 0.99), normalize=True)
    sync(); out["plda_s"] = time.perf_counter() - t0
    out["eer"] = {k: float(v[4]) for k, v in res.items()}
    out["n_trials_scored"] = int(tar.shape[0] + non.shape[0])
    out["total_s"] = out["extract_s"] + out["key_ndx_s"] + out["tail_s"] + out["plda_train_s"] + out["plda_s"]
    # the same tail with the reference's algorithm on the host (oracle port): id matching + cosine + as-norm + 2 x ROCCH
    from oracle import scoring_ref as SR, eval_ref as ER
    t0 = time.perf_counter()
    SR.cosine_scoring(ids, emb.astype(numpy.float64), ids, emb.astype(numpy.float64), ndx.modelset, ndx.segset, ndx.trialmask)
    SR.asnorm(emb, cohort.cpu().numpy(), 200)
    ER.rocch(tar.astype(numpy.float64), non.astype(numpy.float64))
    ER.rocch(tar.astype(numpy.float64), non.astype(numpy.float64))
    out["cpu_baseline"] = {"value": time.perf_counter() - t0, "unit": "s", "cores": os.cpu_count() or 1, "kind": "port",
                           "sample": "the same scoring tail on the host: id matching + cosine scoring of the full trial list + as-norm "
                                     "of all %d rows + 2 x ROCCH of all %d trials (compare with tail_s)" % (N, tar.shape[0] + non.shape[0])}
    return out


class StdoutToStderr:
    """Everything but the final JSON line goes to stderr (NCCL and friends print banners on stdout)."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utts", type=int, default=96, help="mean utterances per GPU per step (96 x 11 s: just under the 1200 audio-s batch "
                                                           "budget that bulk.make_batches uses by default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clock-sampler", action="store_true", help="diagnostic: do not run nvidia-smi beside the timed region")
    args = ap.parse_args()
    with StdoutToStderr():
        line = run_reference(args) if args.impl == "reference" else run_ours(args)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
