#!/usr/bin/env python
"""Benchmark of the speaker-verification inference hot path (BASELINE.json metric:
audio-seconds embedded per second + PLDA trials per second, with % of roofline).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port)

One JSON line on stdout.  Headline `value` = HalfResNet34 x-vector extraction throughput (audio-s/s, inputs
resident in HBM); `e2e` = the same through the public API with host buffers (H2D of the waveforms and D2H
of the embeddings inside the timed region).  `extra` carries the two other configs of BASELINE.json
(TDNN extraction, 20k x 20k PLDA scoring) with their own rooflines.  A "step" = one packed batch of
variable-length utterances per GPU (the per-GPU shard shape of BASELINE config 4, utterances bucketed by length).
"""
import argparse
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
    """Algorithmic MACs of one HalfResNet34 embedding of L samples (SURVEY.md 8d closed form)."""
    T1 = 1 + L // 160
    T2 = (T1 - 1) // 2 + 1
    T3 = (T2 - 1) // 2 + 1
    T4 = (T3 - 1) // 2 + 1
    trunk = 80 * T1 * 56608 + 40 * T2 * 278528 + 20 * T3 * 1703936 + 10 * T4 * 3276800
    return trunk, trunk + 1310720 * T4 + 1350016


def tdnn_macs(L):
    T = 1 + L // 512
    return 204800 * (T - 4) + 786432 * (T - 8) + 786432 * (T - 14) + 262144 * (T - 14) + 786432 * (T - 14) + 1572864


def make_batches(n_batches, n_utt, seed, lo_s, hi_s, device):
    """Synthetic 16 kHz Gaussian audio; lengths ~ U[lo, hi] s, sorted inside the batch (length bucketing)."""
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
                                          "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
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
    """K steps bracketed by barrier + synchronize; device time from CUDA events; max over ranks."""
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


def cpu_reference_extraction(budget_s, n_threads):
    """The reference's algorithm (oracle port, torch CPU fp32) on a bounded sample of the same workload."""
    from oracle import extract_ref as R
    from sidekit_b200 import synth
    from tests.models import synthetic_state_dict
    torch.set_num_threads(n_threads)
    sd = synthetic_state_dict("halfresnet34", N_SPK, 256)
    lengths = synth.synth_lengths(64, 2.0, 20.0, seed=4)
    done_s, t0, n = 0.0, time.perf_counter(), 0
    with torch.no_grad():
        R.forward(sd, synth.synth_wave(1, 32000, seed=1), "halfresnet34")     # warm-up
        t0 = time.perf_counter()
        for i, L in enumerate(lengths):                                       # batch-1 loop, like the reference's extractors
            R.forward(sd, synth.synth_wave(1, int(L), seed=10 + i), "halfresnet34")
            done_s += L / 16000.0
            n += 1
            if time.perf_counter() - t0 > budget_s:
                break
    dt = time.perf_counter() - t0
    cpu_reference_extraction.last_ms = dt * 1e3
    return done_s / dt, "%d utterances (%.0f audio-s) of the 2-20 s workload, batch-1 loop, %.1f s of CPU" % (n, done_s, dt)


def gpu_eager_reference_extraction(budget_s):
    """SURVEY.md 8d "reference GPU path": the reference's algorithm as stock PyTorch eager ops on this B200 (cuDNN /
    cuBLAS / cuFFT), batch-1 loop like the reference's extractors, fp32 and fp16 autocast.  Reported next to the CPU
    number in the `--impl reference` line; none of this repo's kernels run here."""
    from oracle import extract_ref as R
    from sidekit_b200 import synth
    from tests.models import synthetic_state_dict
    dev = torch.device("cuda", 0)
    sd = {k: v.to(dev) for k, v in synthetic_state_dict("halfresnet34", N_SPK, 256).items()}
    lengths = synth.synth_lengths(256, 2.0, 20.0, seed=4)
    waves = [synth.synth_wave(1, int(L), seed=10 + i).to(dev) for i, L in enumerate(lengths[:64])]
    out = {}
    for name, amp in (("fp32", False), ("fp16_autocast", True)):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            for w in waves[:3]:
                R.forward(sd, w, "halfresnet34")
            torch.cuda.synchronize()
            t0, done, n = time.perf_counter(), 0.0, 0
            while time.perf_counter() - t0 < budget_s:
                w = waves[n % len(waves)]
                R.forward(sd, w, "halfresnet34")[1].cpu()          # per-utterance D2H like extract_xvectors.py
                done += w.shape[1] / 16000.0
                n += 1
            dt = time.perf_counter() - t0
        out[name] = {"value": done / dt, "unit": "audio-s/s", "sample": "%d utterances, batch-1 loop, %.1f s" % (n, dt)}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cores = os.cpu_count() or 1
    vals, ms, sample = [], [], ""
    for i in range(args.warmup + args.steps):
        v, sample = cpu_reference_extraction(max(4.0, 40.0 / max(1, args.steps + args.warmup)), cores)
        if i >= args.warmup:
            vals.append(v)
            ms.append(cpu_reference_extraction.last_ms)
    value = float(numpy.mean(vals))
    line = {"impl": "reference", "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(numpy.mean(ms)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "HalfResNet34 x-vector extraction, utterances 2-20 s (BASELINE config 4 shard shape)"},
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if torch.cuda.is_available():
        try:
            line["extra"] = {"torch_eager_on_this_gpu": gpu_eager_reference_extraction(4.0)}
        except Exception as e:                                        # informational only
            line["extra"] = {"torch_eager_on_this_gpu": {"unavailable": repr(e)[:200]}}
    return line


def run_ours(args):
    import torch.distributed as dist
    from sidekit_b200 import _lib
    import sidekit_b200 as sk
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

    # ------------------------------------------------------------------ HalfResNet34 extraction (headline)
    model = build_model("halfresnet34", 256, device)
    n_rot = 4
    batches = make_batches(n_rot, args.utts, seed=500 + 97 * rank, lo_s=2.0, hi_s=20.0, device=device)
    audio_s = [sum(b[2]) / 16000.0 for b in batches]
    gathered = [torch.empty((world * args.utts, 256), device=device)] if dist_on else None

    def step_dev(i):
        _, flat, lengths = batches[i % n_rot]
        emb = model.extract_packed(flat, lengths)
        if dist_on:                       # the one collective of the path: embeddings all-gathered over NVLink
            dist.all_gather_into_tensor(gathered[0], emb)

    with torch.no_grad():
        for i in range(max(3, args.warmup, n_rot)):      # every rotating batch once: its geometry plan is built (and cached) here
            step_dev(i)
        sampler = ClockSampler(local)
        sampler.start()
        l0 = lib.skb_kernel_launches()
        ms = timed(step_dev, args.steps, dist_on)
        launches = lib.skb_kernel_launches() - l0
        clocks = sampler.stop()
        # end to end: the public bulk call on pinned HOST batches -- every step's waveforms cross PCIe inside the timed
        # region (overlapped with the previous step's compute on a second stream) and its embeddings come back to the host
        host_batches = [(batches[i % n_rot][0], batches[i % n_rot][2]) for i in range(args.steps)]
        model.extract_stream(host_batches[:n_rot])
        ms_e2e = timed(lambda i: model.extract_stream(host_batches) if i == 0 else None, 1, dist_on)
        # per-category device time (separate pass with event brackets) for the roofline of the dominant kernel
        lib.skb_profile_enable(1)
        for i in range(args.steps):
            step_dev(i)
        cat = (torch.zeros(8).numpy()).astype(numpy.float32)
        import ctypes
        lib.skb_profile_read(cat.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 8)
        lib.skb_profile_enable(0)
    done_audio = sum(audio_s[i % n_rot] for i in range(args.steps))
    tot = torch.tensor([done_audio], device=device)
    if dist_on:
        dist.all_reduce(tot)
    value = tot.item() / (ms / 1e3)
    e2e_value = tot.item() / (ms_e2e / 1e3)
    trunk_macs = sum(hr34_macs(L)[0] for i in range(args.steps) for L in batches[i % n_rot][2])
    all_macs = sum(hr34_macs(L)[1] for i in range(args.steps) for L in batches[i % n_rot][2])
    conv_ms = float(cat[2])
    conv_tf = 2.0 * trunk_macs / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    n_conv_launches = args.steps * 36
    roofline = {"kernel": "conv_umma_kernel (tcgen05 shift-GEMM conv, 36 launches per step)", "bound": "tensor",
                "achieved": conv_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": conv_tf / peaks["tf_sustained"],
                "traffic": load_traffic("conv_umma_kernel"), "traffic_unit": "bytes per launch (ncu dram read+write, profiles/traffic.json)",
                "peak_source": peaks["source"] + " sustained bf16",
                "avg_launch_ms": conv_ms / n_conv_launches if n_conv_launches else None,
                "flops_per_launch": 2.0 * trunk_macs / n_conv_launches,
                "whole_step_tflops": 2.0 * all_macs / (ms / 1e3) / 1e12 / max(world, 1),
                "device_ms_by_category": {"frontend": float(cat[0]), "stem": float(cat[1]), "conv": conv_ms, "se": float(cat[3]),
                                          "pooling_head": float(cat[4])}}
    h2d = int(numpy.mean([b[0].numel() * 4 for b in batches]))
    d2h = args.utts * 256 * 4

    extra = {}
    if rank == 0 or dist_on:
        extra = run_extras(args, device, peaks, dist_on, rank, world)
    if world == 1:
        # the same step on batches whose length composition has never been seen (no cached geometry plan): what a bulk
        # extraction over a real corpus pays -- the plan is built on the host while the previous batch runs on the GPU
        with torch.no_grad():
            fresh = make_batches(13, args.utts, seed=7700, lo_s=2.0, hi_s=20.0, device=device)
            fresh.sort(key=lambda b: -sum(b[2]))
            model.extract_packed(fresh[0][1], fresh[0][2])      # the largest one sizes the work buffers (untimed), as the
            fresh = fresh[1:]                                   # first batch of a length-sorted bulk run does
            ms_f = timed(lambda i: model.extract_packed(fresh[i][1], fresh[i][2]), len(fresh), False)
        extra["fresh_geometry"] = {"metric": "audio_seconds_per_second", "unit": "audio-s/s",
                                   "value": sum(sum(b[2]) for b in fresh) / 16000.0 / (ms_f / 1e3), "ms_per_step": ms_f / len(fresh),
                                   "workload": "%d batches of %d utterances, every batch a new length composition" % (len(fresh), args.utts)}
        del fresh
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
        extra["resample_44k1_to_16k"] = {"metric": "audio_seconds_per_second", "unit": "audio-s/s",
                                         "value": float(rl.sum()) / 44100.0 * args.steps / (ms_r / 1e3), "ms_per_step": ms_r / args.steps,
                                         "workload": "256 utterances 2-20 s at 44.1 kHz -> 16 kHz, packed ragged batch",
                                         "roofline": {"kernel": "resample_kernel", "bound": "hbm", "achieved": gbs,
                                                      "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                                                      "traffic": load_traffic("resample_kernel")}}
        del rx, ry

    if rank == 0:
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, sample = cpu_reference_extraction(15.0, cores)
            cpu = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample}
        line = {"metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup, n_rot), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": "HalfResNet34 x-vector extraction (256-d, random init), %d utterances 2-20 s per GPU per step, "
                                       "length-bucketed packed batch (BASELINE config 4 shard shape)" % args.utts,
                           "utts_per_step_per_gpu": args.utts, "audio_s_per_step_per_gpu": float(numpy.mean(audio_s)),
                           "accumulate": "fp32", "parallelism": "utterance-sharded x%d, NCCL all_gather of embeddings" % world,
                           "l2": "inputs rotate over %d batches (%.0f MB) and every step streams >1 GB of activations, "
                                 "so nothing survives in the 126 MB L2 between steps" % (n_rot, n_rot * h2d / 1e6)},
                "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches), "extra": extra}
    else:
        line = None
    if dist_on:
        dist.barrier()
        dist.destroy_process_group()
    return line


def run_extras(args, device, peaks, dist_on, rank, world):
    """BASELINE configs 2 (TDNN extraction) and 3 (20k x 20k PLDA scoring): value + roofline each."""
    import torch.distributed as dist
    import sidekit_b200 as sk
    from sidekit_b200 import synth
    out = {}
    with torch.no_grad():
        # ---- TDNN 'xvector' (512-d): front-end dominated, HBM-bound
        tdnn = build_model("xvector", 512, device)
        tb = make_batches(2, 512, seed=900 + rank, lo_s=2.0, hi_s=20.0, device=device)
        f = lambda i: tdnn.extract_packed(tb[i % 2][1], tb[i % 2][2])
        for i in range(3):
            f(i)
        ms = timed(f, args.steps, dist_on)
        aud = sum(sum(tb[i % 2][2]) for i in range(args.steps)) / 16000.0 * world
        bytes_alg = sum(4 * L + 80 * (1 + L // 512) * 4 for i in range(args.steps) for L in tb[i % 2][2])
        out["tdnn_xvector"] = {"metric": "audio_seconds_per_second", "value": aud / (ms / 1e3), "unit": "audio-s/s",
                               "workload": "TDNN xvector 512-d, 512 utterances 2-20 s per GPU per step (BASELINE config 2 shape)",
                               "ms_per_step": ms / args.steps,
                               "tflops": 2.0 * sum(tdnn_macs(L) for i in range(args.steps) for L in tb[i % 2][2]) / (ms / 1e3) / 1e12}
        del tdnn, tb
        # ---- the other trunks that run on the HalfResNet34 kernels: ResNet34 (128/256 channels) and
        # FastResNet34 (16..128 channels, frequency axis halved by the stem), single GPU only
        if world == 1:
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
        # ---- PLDA two-covariance scoring, 20k x 20k x 256 (rows sharded over ranks, no collective)
        Ne = Nt = 20000
        D = 256
        rows = Ne // world
        E = torch.from_numpy(synth.synth_embeddings(Ne, D, seed=6)).float()[rank * rows:(rank + 1) * rows].to(device)
        T = torch.from_numpy(synth.synth_embeddings(Nt, D, seed=7)).float().to(device)
        mu, F, Sigma = synth.synth_plda(D, D, seed=8)
        r = torch.randn(rows, device=device)
        q = torch.randn(Nt, device=device)
        outm = torch.empty((rows, Nt), dtype=torch.float32, device=device)
        g = lambda i: sk.score_matrix(E, T, r, q, cst=0.5, alpha=1.0, passes=0, out=outm)
        for i in range(3):
            g(i)
        ms = timed(g, args.steps, dist_on)
        trials = float(Ne) * Nt * args.steps
        gbs = (rows * Nt * 4 + 4 * (rows + Nt) * D) * args.steps / (ms / 1e3) / 1e9
        out["plda_20k"] = {"metric": "trials_per_second", "value": trials / (ms / 1e3), "unit": "trials/s",
                           "workload": "PLDA-form scoring 20k x 20k x 256, fp32 score matrix resident in HBM (BASELINE config 3)",
                           "ms_per_step": ms / args.steps,
                           "roofline": {"kernel": "score_gemm_kernel", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"],
                                        "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                                        "traffic": load_traffic("score_gemm_kernel") if world == 1 else None,
                                        "tensor_tflops": 2.0 * rows * Nt * D * args.steps / (ms / 1e3) / 1e12}}
        if world == 1:
            # end to end through the reference-shaped API: StatServer / Ndx in, Scores (float64 numpy on the host) out
            ids_e = numpy.array(["m%06d" % i for i in range(Ne)])
            ids_t = numpy.array(["s%06d" % i for i in range(Nt)])
            en = sk.StatServer.from_embeddings(ids_e, synth.synth_embeddings(Ne, D, seed=6))
            te = sk.StatServer.from_embeddings(ids_t, synth.synth_embeddings(Nt, D, seed=7))
            ndx = sk.Ndx()
            ndx.modelset, ndx.segset, ndx.trialmask = ids_e, ids_t, numpy.ones((Ne, Nt), dtype=bool)
            t0 = time.perf_counter()
            sc = sk.PLDA_scoring(en, te, ndx, mu, F, numpy.zeros((D, 0)), Sigma)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            mat = sc.scoremat
            t2 = time.perf_counter()
            out["plda_20k"]["e2e"] = {"value": float(Ne) * Nt / (t2 - t0), "unit": "trials/s", "api_call_s": t1 - t0,
                                      "d2h_float64_s": t2 - t1, "d2h_bytes": int(mat.nbytes)}
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        out["plda_20k"]["cpu_baseline"] = cpu_reference_plda(2000, 256)
        out["vox1o_pipeline"] = run_vox1o_pipeline(device)
    return out


def cpu_reference_plda(n, D):
    """The reference's fast-PLDA algorithm (oracle port, numpy float64, id matching included) on a bounded n x n sample."""
    from oracle import scoring_ref as S
    from sidekit_b200 import synth
    E, T = synth.synth_embeddings(n, D, seed=6), synth.synth_embeddings(n, D, seed=7)
    mu, F, Sigma = synth.synth_plda(D, D, seed=8)
    ids_e = numpy.array(["m%06d" % i for i in range(n)])
    ids_t = numpy.array(["s%06d" % i for i in range(n)])
    mask = numpy.ones((n, n), dtype=bool)
    t0 = time.perf_counter()
    S.fast_plda_scoring(ids_e, E, ids_t, T, ids_e, ids_t, mask, mu, F, Sigma)
    dt = time.perf_counter() - t0
    return {"value": n * n / dt, "unit": "trials/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "%d x %d x %d trials, numpy float64 (BLAS threads as configured), %.2f s" % (n, n, D, dt)}


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
    t0 = time.perf_counter()
    cos = sk.cosine_scoring(enroll, test, ndx)
    tar, non = cos.get_tar_non(key)
    sync(); out["cosine_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    fa = sk.FactorAnalyser().plda(sk.StatServer.from_embeddings(numpy.array(["spk%04d" % (i % 1177) for i in range(N)]), emb), 128,
                                  nb_iter=5, save_final=False)
    sync(); out["plda_train_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    pl = sk.PLDA_scoring(enroll, test, ndx, fa.mean, fa.F, numpy.zeros((256, 0)), fa.Sigma)
    ptar, pnon = pl.get_tar_non(key)
    sync(); out["plda_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    cohort = model.after_speaker_embedding.weight.detach()
    S = sk.asnorm(torch.from_numpy(emb).to(device), cohort, None)
    sync(); out["asnorm_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    sn = sk.Scores()
    sn.modelset, sn.segset, sn.scoremat, sn.scoremask = ids, ids, S, numpy.ones(S.shape, dtype=bool)
    atar, anon = sn.get_tar_non(key)
    res = {"cosine": sk.fast_minDCF(tar, non, numpy.log(0.01 / 0.99), normalize=True), "plda": sk.fast_minDCF(ptar, pnon, numpy.log(0.01 / 0.99), normalize=True),
           "asnorm": sk.fast_minDCF(atar, anon, numpy.log(0.01 / 0.99), normalize=True)}
    out["eer_mindcf_s"] = time.perf_counter() - t0
    out["eer"] = {k: float(v[4]) for k, v in res.items()}
    out["n_trials_scored"] = int(tar.shape[0] + non.shape[0])
    out["total_s"] = sum(v for k, v in out.items() if k.endswith("_s") and isinstance(v, float))
    # the same tail with the reference's algorithm on the host (oracle port): id matching + cosine + as-norm + ROCCH
    from oracle import scoring_ref as SR, eval_ref as ER
    t0 = time.perf_counter()
    SR.cosine_scoring(ids, emb.astype(numpy.float64), ids, emb.astype(numpy.float64), ndx.modelset, ndx.segset, ndx.trialmask)
    SR.asnorm(emb, cohort.cpu().numpy(), 200)
    ER.rocch(tar.astype(numpy.float64), non.astype(numpy.float64))
    out["cpu_baseline"] = {"value": time.perf_counter() - t0, "unit": "s", "cores": os.cpu_count() or 1, "kind": "port",
                           "sample": "the same scoring tail on the host: id matching + cosine scoring of the full trial list + as-norm "
                                     "of all %d rows + ROCCH of all %d trials (compare with cosine_s + asnorm_s + eer_mindcf_s / 3)" % (N, tar.shape[0] + non.shape[0])}
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
    ap.add_argument("--utts", type=int, default=96, help="utterances per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    with StdoutToStderr():
        line = run_reference(args) if args.impl == "reference" else run_ours(args)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
