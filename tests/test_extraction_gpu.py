"""GPU parity of the extraction path (through the C ABI) against the oracle and the golden fixtures.

Tolerances (BASELINE.json north_star): embeddings within 1e-3 relative L2 and cosine >= 0.9999 of the fp32
reference; the front-end is fp32 arithmetic and is held to 1e-4 absolute on unit-variance features.
"""
import numpy
import pytest
import torch

from oracle import extract_ref as R
from sidekit_b200 import synth
from tests.helpers import golden, min_cosine, rel_l2
from tests.models import make_xtractor

pytestmark = pytest.mark.gpu

STAGES = ["stem"] + ["layer%d.%d" % (l + 1, b) for l, n in enumerate((3, 4, 6, 3)) for b in range(n)]


@pytest.fixture(scope="module")
def hr34():
    m = make_xtractor("halfresnet34", 32, 256).cuda()
    return m, {k: v.cpu() for k, v in m.state_dict().items()}


@pytest.fixture(scope="module")
def tdnn():
    m = make_xtractor("xvector", 32, 512).cuda()
    return m, {k: v.cpu() for k, v in m.state_dict().items()}


def test_logmel_frontend_matches_oracle(hr34):
    m, sd = hr34
    for L, seed in ((16000, 1), (24160, 2), (11111, 3), (64000, 4)):
        x = synth.synth_wave(2, L, seed=seed)
        got = m.preprocessor(x.cuda(), is_eval=True).cpu()
        ref = R.logmel_frontend(sd, x)
        assert got.shape == ref.shape
        assert (got - ref).abs().max().item() < 1e-4


def test_real_audio_frontend_golden(hr34):
    m, _ = hr34
    g = golden("extraction.npz")
    x = torch.from_numpy(g["real_pcm16"].astype(numpy.float32) / 32768.0)
    got = m.preprocessor(x.cuda(), is_eval=True).cpu().numpy()[0]
    assert numpy.abs(got - g["real_feats"]).max() < 1e-3        # digital-silence bins are ill-conditioned under CMVN
    emb = m(x.cuda(), is_eval=True)[1].cpu()
    assert rel_l2(emb, g["real_emb"]) < 1e-3


def test_mfcc_frontend_matches_oracle(tdnn):
    m, sd = tdnn
    for L, seed in ((32000, 1), (51234, 2)):
        x = synth.synth_wave(2, L, seed=seed)
        got = m.preprocessor(x.cuda(), is_eval=True).cpu()
        ref = R.mfcc_frontend(sd, x)
        assert got.shape == ref.shape
        assert (got - ref).abs().max().item() < 2e-4


@pytest.mark.parametrize("init,tol", [("default", 2e-3), ("he", 6e-3)])
def test_every_block_matches_oracle(init, tol):
    """Per-stage activations (debug hook) for a packed batch of three different lengths; 'he' = stress weights."""
    m = make_xtractor("halfresnet34", 32, 256, init=init).cuda()
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    waves = [synth.synth_wave(1, L, seed=100 + i)[0] for i, L in enumerate((16000, 24160, 11111))]
    refs = []
    for w in waves:
        col = {}
        R.halfresnet34_forward(sd, w.unsqueeze(0), collect=col)
        refs.append(col)
    cw = [w.cuda() for w in waves]
    for st in STAGES:
        out = m.debug_stage(cw, st).cpu()
        for i, r in enumerate(refs):
            ref = r[st][0]
            H = ref.shape[1]
            got = out[i, :, :H, :]
            err = ((got.double() - ref.double()).norm() / ref.double().norm()).item()
            assert err < tol, (st, i, err)
            assert out[i, :, H:, :].abs().max().item() == 0.0 if out.shape[2] > H else True
    pooled = m.debug_stage(cw, "pooled").cpu()
    assert rel_l2(pooled, torch.cat([r["pooled"] for r in refs])) < tol


def test_embeddings_and_logits_match_oracle_and_golden(hr34):
    m, sd = hr34
    g = golden("extraction.npz")
    waves = [synth.synth_wave(1, int(L), seed=int(s))[0] for L, s in zip(g["hr_lengths"], g["hr_seeds"])]
    logits, emb = m.extract_varlen([w.cuda() for w in waves], want_logits=True)
    ref = [R.forward(sd, w.unsqueeze(0), "halfresnet34") for w in waves]
    ref_emb, ref_logits = torch.cat([r[1] for r in ref]), torch.cat([r[0] for r in ref])
    assert rel_l2(emb.cpu(), ref_emb) < 1e-3 and min_cosine(emb.cpu(), ref_emb) >= 0.9999
    assert rel_l2(emb.cpu(), g["hr_emb"]) < 1e-3                 # recorded from the real reference
    assert (logits.cpu() - ref_logits).abs().max().item() < 3e-2  # logits = 30 * cos
    assert numpy.abs(logits.cpu().numpy() - g["hr_logits"]).max() < 3e-2
    # the reference-style dense batch call returns the same pair
    lo2, em2 = m(torch.stack(waves[:2]).cuda(), is_eval=True)
    assert rel_l2(em2.cpu(), ref_emb[:2]) < 1e-3 and lo2.shape == (2, 32)
    assert torch.allclose(emb.norm(dim=1).cpu(), torch.ones(4), atol=1e-5)


def _stages(table):
    return ["stem"] + ["layer%d.%d" % (l + 1, b) for l, (n, _) in enumerate(table) for b in range(n)]


@pytest.mark.parametrize("archi,prefix,table,real_c0", [("resnet34", "r34", R.RESNET34_STAGES, 128),
                                                        ("fastresnet34", "f34", R.FASTRESNET34_STAGES, 16)])
def test_other_resnet_trunks_blocks_embeddings_and_packing(archi, prefix, table, real_c0):
    """The trunks that share the HalfResNet34 kernels -- ResNet34 (128/256 channels, 7 layers) and FastResNet34 (7x7
    stride-(1,2) stem, 16..128 channels, layer4 widening without stride, pooling without global context): every block
    against the oracle, embeddings/logits against the oracle and the recorded reference outputs, packing invariance."""
    m = make_xtractor(archi, 32, 256).cuda()
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    g = golden("extraction_resnet34.npz")
    waves = [synth.synth_wave(1, int(L), seed=int(s))[0] for L, s in zip(g[prefix + "_lengths"], g[prefix + "_seeds"])]
    refs, outs = [], []
    for w in waves:
        col = {}
        outs.append(R.forward(sd, w.unsqueeze(0), archi, collect=col))
        refs.append(col)
    cw = [w.cuda() for w in waves]
    for st in _stages(table):
        out = m.debug_stage(cw, st).cpu()
        for i, r in enumerate(refs):
            ref = r[st][0]
            C, H = ref.shape[0], ref.shape[1]
            assert out.shape[1] >= C and out.shape[3] == ref.shape[2], (st, out.shape, ref.shape)
            err = ((out[i, :C, :H, :].double() - ref.double()).norm() / ref.double().norm()).item()
            assert err < 2e-3, (st, i, err)
            assert out[i, C:].abs().max().item() == 0.0 if out.shape[1] > C else True      # channel padding stays zero
    assert rel_l2(m.debug_stage(cw, "pooled").cpu(), torch.cat([r["pooled"] for r in refs])) < 2e-3
    logits, emb = m.extract_varlen(cw, want_logits=True)
    ref_emb, ref_logits = torch.cat([o[1] for o in outs]), torch.cat([o[0] for o in outs])
    assert rel_l2(emb.cpu(), ref_emb) < 1e-3 and min_cosine(emb.cpu(), ref_emb) >= 0.9999
    assert rel_l2(emb.cpu(), g[prefix + "_emb"]) < 1e-3
    assert (logits.cpu() - ref_logits).abs().max().item() < 3e-2
    assert numpy.abs(logits.cpu().numpy() - g[prefix + "_logits"]).max() < 3e-2
    solo = torch.cat([m.extract_varlen([w]) for w in cw])
    assert torch.equal(emb, solo) and torch.equal(emb, m.extract_varlen(cw[::-1]).flip(0))
    lo2, em2 = m(torch.stack(waves[:2]).cuda(), is_eval=True)
    assert rel_l2(em2.cpu(), ref_emb[:2]) < 1e-3 and lo2.shape == (2, 32)


def test_packing_invariance_is_bit_exact(hr34):
    """Packed variable-length batch == one-by-one == any other packing, bit for bit (integer SE sums)."""
    m, _ = hr34
    waves = [synth.synth_wave(1, L, seed=300 + i)[0].cuda() for i, L in enumerate((16000, 9000, 20321, 16000, 12345))]
    packed = m.extract_varlen(waves)
    solo = torch.cat([m.extract_varlen([w]) for w in waves])
    rev = m.extract_varlen(waves[::-1]).flip(0)
    assert torch.equal(packed, solo)
    assert torch.equal(packed, rev)
    assert torch.equal(packed, m.extract_varlen(waves))           # run-to-run determinism


@pytest.mark.parametrize("which", ["hr34", "tdnn"])
def test_packing_invariance_holds_across_batch_sizes(which, hr34, tdnn):
    """An utterance's embedding must not depend on HOW MANY utterances share its batch (the dense head's K split, the
    span tables and the work partition all change with the batch size): 150 utterances packed == the same ones in small
    batches, bit for bit."""
    m, _ = hr34 if which == "hr34" else tdnn
    lo = 8000 if which == "hr34" else 12000
    waves = [synth.synth_wave(1, lo + 37 * i, seed=700 + i)[0].cuda() for i in range(150)]
    big = m.extract_varlen(waves)
    small = torch.cat([m.extract_varlen(waves[i:i + 7]) for i in range(0, 150, 7)])
    assert torch.equal(big, small)
    assert torch.equal(big[[0, 77, 149]], torch.cat([m.extract_varlen([waves[i]]) for i in (0, 77, 149)]))


def test_baseline_config1_dense_batch(hr34):
    """BASELINE.json configs[0]: 64 synthetic 4 s utterances through ``Xtractor(...)(x, is_eval=True)`` (the reference's own
    CPU-runnable case; SURVEY 8d: ``torch.randn(64, 64000, generator=seed 3) * 0.1``), every row against the oracle."""
    m, sd = hr34
    x = torch.randn(64, 64000, generator=torch.Generator().manual_seed(3)) * 0.1
    logits, emb = m(x.cuda(), is_eval=True)
    ref_lo, ref_em = R.forward(sd, x, "halfresnet34")
    assert emb.shape == (64, 256) and logits.shape == (64, 32)
    assert rel_l2(emb.cpu(), ref_em) < 1e-3 and min_cosine(emb.cpu(), ref_em) >= 0.9999
    assert (logits.cpu() - ref_lo).abs().max().item() < 3e-2
    one = m(x[5].cuda(), is_eval=True)[1]                      # 1-D input, as extract_xvectors.py feeds it
    assert torch.equal(one[0], emb[5])


def test_extreme_lengths_in_one_batch(hr34):
    """The shortest utterance with two lines on the last level (1 280 samples: 9 frames) next to a 45 s one (4 501 frames)
    in the same packed batch, against the oracle.  With a single pooled frame the reference's global-context std is NaN
    (unbiased estimator over one sample, pooling.py:68): the engine returns NaN for that utterance too, and only for it."""
    m, sd = hr34
    waves = [synth.synth_wave(1, L, seed=900 + i)[0] for i, L in enumerate((1280, 720000, 1500))]
    emb = m.extract_varlen([w.cuda() for w in waves]).cpu()
    ref = torch.cat([R.forward(sd, w.unsqueeze(0), "halfresnet34")[1] for w in waves])
    assert rel_l2(emb, ref) < 1e-3 and min_cosine(emb, ref) >= 0.9999
    one = synth.synth_wave(1, 513, seed=903)[0]
    assert not torch.isfinite(R.forward(sd, one.unsqueeze(0), "halfresnet34")[1]).any()
    both = m.extract_varlen([one.cuda(), waves[0].cuda()]).cpu()
    assert not torch.isfinite(both[0]).any() and torch.equal(both[1], emb[0])
    with pytest.raises(RuntimeError):
        m.extract_varlen([synth.synth_wave(1, 512, seed=1)[0].cuda()])        # torch.stft(reflect) needs L > n_fft / 2


def test_plan_cache_switches_and_evictions_leave_no_stale_state(hr34):
    """Geometries alternate (the engine keeps 8 plans and never clears its activation buffers): results stay bit-identical."""
    m, _ = hr34
    sets = [[synth.synth_wave(1, 8000 + 1777 * i + 531 * j, seed=900 + 10 * i + j)[0].cuda() for j in range(1 + i % 3)] for i in range(11)]
    first = [m.extract_varlen(ws) for ws in sets]                 # 11 distinct plans: the oldest ones are evicted
    for i in (0, 10, 3, 0, 7):                                    # cache hits, rebuilt plans, big-after-small and back
        assert torch.equal(m.extract_varlen(sets[i]), first[i])


def test_experimental_fused_tap_kernel_parity():
    """conv3_umma_kernel (SKB_FUSED_TAPS=1, read once per process -> subprocess): oracle parity and packing invariance."""
    import os
    import subprocess
    import sys
    code = (
        "import torch\n"
        "from oracle import extract_ref as R\n"
        "from sidekit_b200 import synth\n"
        "from tests.models import make_xtractor\n"
        "from tests.helpers import rel_l2\n"
        "m = make_xtractor('halfresnet34', 32, 256).cuda()\n"
        "sd = {k: v.cpu() for k, v in m.state_dict().items()}\n"
        "ws = [synth.synth_wave(1, L, seed=700 + i)[0] for i, L in enumerate((16000, 9000, 20321))]\n"
        "emb = m.extract_varlen([w.cuda() for w in ws])\n"
        "ref = torch.cat([R.forward(sd, w.unsqueeze(0), 'halfresnet34')[1] for w in ws])\n"
        "assert rel_l2(emb.cpu(), ref) < 1e-3, rel_l2(emb.cpu(), ref)\n"
        "solo = torch.cat([m.extract_varlen([w.cuda()]) for w in ws])\n"
        "assert torch.equal(emb, solo)\n"
        "print('fused-tap ok')\n")
    env = dict(os.environ, SKB_FUSED_TAPS="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "fused-tap ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_fused_channel_totals_equal_the_separate_pass(tmp_path):
    """conv1's epilogue accumulates the SE channel totals on the 32-channel layers (fixed point, per-thread accumulators
    flushed with 64-bit atomics); SKB_NO_FUSED_SUMS=1 (read once per process -> subprocess) runs the separate plane_sum
    pass instead.  Same integers, so the embeddings must agree bit for bit -- also when the first blocks of layers 1 and 2
    are scaled up so that most of their activations exceed 128 (the range a cheaper 32-bit variant could not hold)."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, torch\n"
        "from sidekit_b200 import synth\n"
        "from tests.models import make_xtractor\n"
        "outs = []\n"
        "for scale in (1.0, 64.0):\n"
        "    m = make_xtractor('halfresnet34', 32, 256)\n"
        "    sd = m.state_dict()\n"
        "    for blk in ('layer1.0', 'layer2.1'):\n"
        "        for k in ('weight', 'bias'):\n"
        "            sd['sequence_network.%s.bn1.%s' % (blk, k)] *= scale\n"
        "    m.load_state_dict(sd)\n"
        "    m = m.cuda()\n"
        "    ws = [synth.synth_wave(1, L, seed=1700 + i)[0].cuda() for i, L in enumerate((16000, 9000, 40321, 8000, 25000))]\n"
        "    outs.append(m.extract_varlen(ws).cpu())\n"
        "    outs.append(torch.cat([m.extract_varlen([w]) for w in ws]).cpu())\n"
        "torch.save(outs, sys.argv[1])\n"
        "print('sums ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for name, env in (("fused", {}), ("separate", {"SKB_NO_FUSED_SUMS": "1"})):
        path = os.path.join(tmp_path, name + ".pt")
        out = subprocess.run([sys.executable, "-c", code, path], cwd=root, env=dict(os.environ, **env), capture_output=True, text=True,
                             timeout=300)
        assert out.returncode == 0 and "sums ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
        res[name] = torch.load(path)
    for a, b in zip(res["fused"], res["separate"]):
        assert torch.isfinite(a).all() and torch.equal(a, b)
    assert torch.equal(res["fused"][0], res["fused"][1]) and torch.equal(res["fused"][2], res["fused"][3])      # packing invariance
    assert not torch.equal(res["fused"][0], res["fused"][2])


def test_extract_stream_overlapped_copies_equal_packed_calls(hr34):
    """extract_stream (H2D of batch i+1 overlapped with the forward of batch i) == extract_packed batch by batch."""
    m, _ = hr34
    batches = []
    for i, lens in enumerate(((16000, 9000), (20321, 16000, 12345), (8000,), (30000, 11111), (9000, 16000))):
        ws = [synth.synth_wave(1, L, seed=950 + 10 * i + j)[0] for j, L in enumerate(lens)]
        batches.append((torch.cat(ws).pin_memory(), list(lens)))
    outs = m.extract_stream(batches)
    for (flat, lens), o in zip(batches, outs):
        assert not o.is_cuda and torch.equal(o, m.extract_packed(flat.cuda(), lens).cpu())
    again = m.extract_stream(batches[::-1])                       # staging buffers are reused, other order
    for (flat, lens), o in zip(batches[::-1], again):
        assert torch.equal(o, m.extract_packed(flat.cuda(), lens).cpu())


def test_host_buffer_entry_point(hr34):
    m, _ = hr34
    x = synth.synth_wave(3, 16000, seed=7)
    lo_d, em_d = m(x.cuda(), is_eval=True)
    lo_h, em_h = m(x, is_eval=True)                               # CPU tensors -> skb_xtractor_forward_host
    assert not em_h.is_cuda and torch.equal(em_h, em_d.cpu()) and torch.equal(lo_h, lo_d.cpu())


def test_bf16_operand_mode_meets_cosine_gate():
    m = make_xtractor("halfresnet34", 32, 256, compute_dtype="bf16").cuda()
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    x = synth.synth_wave(2, 16000, seed=11)
    emb = m(x.cuda(), is_eval=True)[1].cpu()
    ref = R.forward(sd, x, "halfresnet34")[1]
    assert min_cosine(emb, ref) >= 0.9999 and rel_l2(emb, ref) < 2e-2


def test_tdnn_embeddings_match_oracle_and_golden(tdnn):
    m, sd = tdnn
    g = golden("extraction.npz")
    waves = [synth.synth_wave(1, int(L), seed=int(s))[0] for L, s in zip(g["td_lengths"], g["td_seeds"])]
    logits, emb = m.extract_varlen([w.cuda() for w in waves], want_logits=True)
    ref = [R.forward(sd, w.unsqueeze(0), "xvector") for w in waves]
    ref_emb = torch.cat([r[1] for r in ref])
    assert rel_l2(emb.cpu(), ref_emb) < 1e-3 and min_cosine(emb.cpu(), ref_emb) >= 0.9999
    assert rel_l2(emb.cpu(), g["td_emb"]) < 1e-3
    assert numpy.abs(logits.cpu().numpy() - g["td_logits"]).max() < 6e-2   # logits = 64 * cos
    solo = torch.cat([m.extract_varlen([w.cuda()]) for w in waves])
    assert torch.equal(emb, solo)


def test_bulk_extract_embeddings_statserver(hr34):
    """bulk.extract_embeddings (length-bucketed batches) == one packed call, bit for bit; StatServer layout as the reference's."""
    from sidekit_b200 import bulk
    m, _ = hr34
    lengths = (16000, 9000, 20321, 16000, 12345, 8000, 30000)
    waves = [synth.synth_wave(1, L, seed=400 + i)[0] for i, L in enumerate(lengths)]
    ids = numpy.array(["utt%d" % i for i in range(len(waves))])
    ss = bulk.extract_embeddings(ids, waves, m, max_audio_seconds=2.5)           # forces several batches
    ref = m.extract_varlen([w.cuda() for w in waves]).cpu().numpy()
    assert ss.validate() and ss.stat1.shape == (7, 256) and ss.stat0.shape == (7, 1)
    assert numpy.array_equal(ss.modelset, ids) and numpy.array_equal(ss.segset, ids)
    assert numpy.array_equal(ss.stat1.astype(numpy.float32), ref)


def test_meanstd_pooling_op():
    from sidekit_b200.nnet import MeanStdPooling
    x = torch.randn(3, 40, 77, generator=torch.Generator().manual_seed(0))
    got = MeanStdPooling()(x.cuda()).cpu()
    assert torch.allclose(got, R.mean_std_pooling(x), atol=1e-5)
    x4 = torch.randn(2, 6, 31, 5, generator=torch.Generator().manual_seed(1))
    assert torch.allclose(MeanStdPooling()(x4.cuda()).cpu(), R.mean_std_pooling(x4), atol=1e-5)


def test_errors_are_loud(hr34):
    m, _ = hr34
    with pytest.raises(RuntimeError, match="too short"):
        m(torch.zeros(1, 300).cuda(), is_eval=True)
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 16000).cuda(), is_eval=False)


def test_halfresnet34_return_conventions_for_every_loss(hr34):
    """xvector.py:585-593, :896-907: 'aam' -> (s*cos logits, normalised emb); 'aps' -> (Linear logits, normalised emb);
    None / 'cce' build no after_speaker_embedding and return the bare (l2-normalised) embedding."""
    import contextlib, io
    from sidekit_b200.nnet import Xtractor
    m_aam, sd = hr34
    x = synth.synth_wave(3, 20000, seed=9)
    ref_emb = R.forward(sd, x, "halfresnet34")[1]
    for loss in (None, "cce", "aps"):
        with contextlib.redirect_stdout(io.StringIO()):
            m = Xtractor(32, "halfresnet34", loss=loss, embedding_size=256)
        own = m.state_dict()
        for k in own:
            if k in sd:
                own[k] = sd[k].clone()
        synth.fill_state_dict({k: v for k, v in own.items() if k.startswith("after_speaker_embedding.cce_backend")}, 0)
        m.load_state_dict(own)
        m = m.eval().cuda()
        out = m(x.cuda(), is_eval=True)
        if loss == "aps":
            logits, emb = out
            W, b = own["after_speaker_embedding.cce_backend.linear8.weight"], own["after_speaker_embedding.cce_backend.linear8.bias"]
            assert logits.shape == (3, 32)
            assert (logits.cpu() - (ref_emb @ W.T + b)).abs().max().item() < 2e-3
        else:
            assert torch.is_tensor(out) and out.shape == (3, 256)
            assert not hasattr(m, "after_speaker_embedding")
            emb = out
        assert rel_l2(emb.cpu(), ref_emb) < 1e-3


def test_fp16_range_guard_fails_loudly_and_bf16_runs():
    """SURVEY.md 7 "Precision": weights whose activations exceed 65504 must not yield silent inf / NaN embeddings."""
    x = synth.synth_wave(2, 16000, seed=5)
    for dtype in ("fp16", "bf16"):
        m = make_xtractor("halfresnet34", 16, 256, compute_dtype=dtype)
        sd = m.state_dict()
        sd["sequence_network.conv1.weight"] *= 3.0e5             # stem output far beyond the fp16 range
        m.load_state_dict(sd)
        m = m.cuda()
        if dtype == "fp16":
            with pytest.raises(OverflowError):
                m(x.cuda(), is_eval=True)
            m.extract_packed(x.reshape(-1).cuda(), [16000, 16000])       # the bulk primitive defers the check ...
            with pytest.raises(OverflowError):
                m.check_overflow()                                       # ... to the caller
        else:
            emb = m(x.cuda(), is_eval=True)[1]
            assert torch.isfinite(emb).all()
    ok = make_xtractor("halfresnet34", 16, 256).cuda()
    ok(x.cuda(), is_eval=True)                                           # ordinary weights: no error
    ok.check_overflow()


def test_pre_emphasis_coefficient_comes_from_the_checkpoint(hr34):
    """PreEmphasis.flipped_filter of the state_dict is honoured (it was hard-coded to 0.97)."""
    m, sd = hr34
    m2 = make_xtractor("halfresnet34", 32, 256)
    sd2 = {k: v.clone() for k, v in sd.items()}
    sd2["preprocessor.PreEmphasis.flipped_filter"] = torch.tensor([[[-0.5, 1.0]]])
    m2.load_state_dict(sd2)
    m2 = m2.cuda()
    x = synth.synth_wave(2, 16000, seed=1)
    got = m2.preprocessor(x.cuda(), is_eval=True).cpu()
    y = R.pre_emphasis(x, 0.5)
    spec = R.power_spectrogram(y, 1024, 400, 160)
    fb = sd["preprocessor.MelSpec.mel_scale.fb"]
    ref = R.instance_norm(torch.log(torch.matmul(spec.transpose(1, 2), fb).transpose(1, 2) + 1e-6))
    assert (got - ref).abs().max().item() < 1e-4
