"""Feed path: IdMap / IdMapSet / extract_embeddings with the reference's signatures, on wav files
written by the test.  Segment logic against oracle/feed_ref.py (parity unpinned: see its header); embeddings against
the packed extraction of the same segments (bit-exact) and the CPU oracle."""
import os
import wave

import numpy
import pytest
import torch

import sidekit_b200 as sk
from oracle import extract_ref as R
from oracle import feed_ref as F
from sidekit_b200 import synth
from sidekit_b200.nnet import xsets
from tests.helpers import rel_l2
from tests.models import make_xtractor


def _write_wavs(tmp_path, lengths):
    pcm = {}
    for i, L in enumerate(lengths):
        x = (synth.synth_wave(1, L, seed=800 + i)[0].numpy() * 32768.0).clip(-32768, 32767).astype(numpy.int16)
        with wave.open(os.path.join(tmp_path, "f%d.wav" % i), "wb") as f:
            f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000)
            f.writeframes(x.tobytes())
        pcm["f%d" % i] = x
    return pcm


def _idmap(rows):
    im = sk.IdMap()
    im.leftids = numpy.array([r[0] for r in rows], dtype="|O")
    im.rightids = numpy.array([r[1] for r in rows], dtype="|O")
    im.start = numpy.array([r[2] for r in rows], dtype="|O")
    im.stop = numpy.array([r[3] for r in rows], dtype="|O")
    assert im.validate()
    return im


def test_idmapset_segments_match_oracle(tmp_path):
    pcm = _write_wavs(str(tmp_path), (64000, 80000, 50000))
    rows = [("spk0", "f0", None, None), ("spk1", "f1", 100, 450), ("spk1", "f1", 200, 260), ("spk2", "f2", 0, 20), ("spk0", "f0", 50, None)]
    ds = xsets.IdMapSet(_idmap(rows), str(tmp_path), "wav", sample_rate=16000, min_duration=1.0)
    assert len(ds) == 5
    for i, (m, f, a, b) in enumerate(rows):
        speech, left, right, start, stop = ds[i]
        seg, ostart, ostop = F.cut_segment(pcm[f], a, b, 16000, 1.0)
        assert (left, right, start, stop) == (m, f, ostart, ostop)
        assert numpy.array_equal(speech.numpy(), seg)
    dsw = xsets.IdMapSet(_idmap(rows[:2]), str(tmp_path), "wav", sliding_window=True, window_len=1.0, window_shift=0.5, min_duration=1.0)
    w = dsw[1][0]
    assert numpy.array_equal(w.numpy(), F.windows(F.cut_segment(pcm["f1"], 100, 450, 16000, 1.0)[0], 1.0, 0.5))
    with pytest.raises(NotImplementedError):
        xsets.IdMapSet(_idmap(rows), str(tmp_path), "wav", transform_pipeline={"add_noise": {}})


@pytest.mark.gpu
def test_extract_embeddings_from_wav_files(tmp_path):
    pcm = _write_wavs(str(tmp_path), (64000, 80000, 50000))
    m = make_xtractor("halfresnet34", 32, 256).cuda()
    rows = [("spk0", "f0", None, None), ("spk1", "f1", 100, 450), ("spk2", "f2", 0, 20), ("spk1", "f1", 0, None)]
    ss = sk.nnet.extract_embeddings(_idmap(rows), m, str(tmp_path), torch.device("cuda"), win_duration=1.0)
    assert ss.validate() and ss.stat1.shape == (4, 256) and ss.stat0.shape == (4, 1)
    assert list(ss.modelset) == [r[0] for r in rows] and list(ss.segset) == [r[1] for r in rows]
    segs = [F.cut_segment(pcm[f], a, b, 16000, 1.0) for (_, f, a, b) in rows]
    ostart, ostop = F.bookkeeping([(None, s[1], s[0].shape[0]) for s in segs], False, 1.0, 1.5)
    assert numpy.array_equal(ss.start, ostart) and numpy.array_equal(ss.stop, ostop)
    ref = torch.cat([m.extract_varlen([torch.from_numpy(s[0]).cuda()]) for s in segs]).cpu().numpy()
    assert numpy.array_equal(ss.stat1, ref)                                   # packed batches == one-by-one, bit for bit
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    cpu = R.forward(sd, torch.from_numpy(segs[2][0]).unsqueeze(0), "halfresnet34")[1]
    assert rel_l2(torch.from_numpy(ss.stat1[2:3]), cpu) < 1e-3
    # sliding windows: one embedding per 1 s window every 0.5 s
    sw = sk.nnet.extract_embeddings(_idmap(rows[:2]), m, str(tmp_path), torch.device("cuda"), sliding_window=True, win_duration=1.0,
                                    win_shift=0.5)
    wins = [F.windows(F.cut_segment(pcm[f], a, b, 16000, 1.0)[0], 1.0, 0.5) for (_, f, a, b) in rows[:2]]
    assert sw.stat1.shape[0] == sum(w.shape[0] for w in wins) == sw.modelset.shape[0]
    ostart, ostop = F.bookkeeping([(w.shape[0], F.cut_segment(pcm[f], a, b, 16000, 1.0)[1], None) for w, (_, f, a, b) in zip(wins, rows[:2])],
                                  True, 1.0, 0.5)
    assert numpy.array_equal(sw.start, ostart) and numpy.array_equal(sw.stop, ostop)
    ref = m.extract_varlen([torch.from_numpy(w).cuda() for ws in wins for w in ws]).cpu().numpy()
    assert numpy.array_equal(sw.stat1, ref)
    # a StatServer from the feed path goes straight into the scorer
    ndx = sk.Ndx()
    ndx.modelset, ndx.segset = numpy.unique(ss.modelset), numpy.unique(ss.segset)
    ndx.trialmask = numpy.ones((ndx.modelset.shape[0], ndx.segset.shape[0]), dtype=bool)
    en = ss.mean_stat_per_model()
    te = sk.StatServer.from_embeddings(ss.segset, ss.stat1)
    te.modelset = te.segset
    sc = sk.cosine_scoring(en, te, ndx)
    assert sc.validate() and numpy.isfinite(sc.scoremat).all()


@pytest.mark.reference
def test_reference_checkpoint_loads_strictly(tmp_path):
    """A checkpoint in the layout the reference's training loop saves (xvector.py: ``speaker_number``, ``model_archi``,
    ``model_state_dict``), built from the REFERENCE's own Xtractor, loads with strict=True into ours (no GPU needed:
    the modules only hold the weights until the first forward)."""
    from oracle import ref_import
    ref = ref_import.build_xtractor(32, "halfresnet34", 256)
    sd = ref.state_dict()
    synth.fill_state_dict(sd, 3)
    path = os.path.join(tmp_path, "best_model.pt")
    torch.save({"speaker_number": 32, "model_archi": {"model_type": "halfresnet34", "loss": {"type": "aam"}},
                "model_state_dict": sd, "epoch": 7, "accuracy": 0.5}, path)
    m = xsets.load_checkpoint(path, torch.device("cpu"))
    assert m.speaker_number == 32 and m.embedding_size == 256 and not m.training
    mine = m.state_dict()
    assert list(mine.keys()) == list(sd.keys())
    assert all(torch.equal(mine[k].cpu(), sd[k]) for k in sd)


@pytest.mark.gpu
def test_extract_embeddings_from_checkpoint_file(tmp_path):
    _write_wavs(str(tmp_path), (40000, 52000))
    m = make_xtractor("halfresnet34", 32, 256).cuda()
    path = os.path.join(tmp_path, "model.pt")
    torch.save({"speaker_number": 32, "model_archi": {"model_type": "halfresnet34", "loss": {"type": "aam"}, "embedding_size": 256},
                "model_state_dict": {k: v.cpu() for k, v in m.state_dict().items()}}, path)
    rows = [("a", "f0", None, None), ("b", "f1", 20, 250)]
    a = sk.nnet.extract_embeddings(_idmap(rows), path, str(tmp_path), torch.device("cuda"), win_duration=1.0)
    b = sk.nnet.extract_embeddings(_idmap(rows), m, str(tmp_path), torch.device("cuda"), win_duration=1.0)
    assert numpy.array_equal(a.stat1, b.stat1) and a.stat1.shape == (2, 256)


def test_kaldi_ark_scp_round_trip(tmp_path):
    """x-vector tables in Kaldi's binary ark / scp layout (what extract_xvectors.py writes through kaldiio): header bytes
    as published, scp offsets pointing at the binary marker, row matrices and vectors, speaker means."""
    import numpy
    from sidekit_b200 import kaldi_io
    from sidekit_b200.statserver import StatServer
    rng = numpy.random.default_rng(0)
    X = rng.standard_normal((5, 16)).astype(numpy.float32)
    ss = StatServer.from_embeddings(numpy.array(["utt%d" % i for i in range(5)]), X)
    scp = str(tmp_path / "xv.scp")
    ark = kaldi_io.write_xvectors(ss, scp)
    raw = open(ark, "rb").read()
    assert raw.startswith(b"utt0 \0BFM \4" + (1).to_bytes(4, "little") + b"\4" + (16).to_bytes(4, "little") + X[0].tobytes())
    lines = open(scp).read().split("\n")
    assert lines[0] == "utt0 %s:5" % ark and lines[1].startswith("utt1 %s:" % ark)
    got = list(kaldi_io.read_scp(scp))
    assert [k for k, _ in got] == list(ss.segset) and all(a.shape == (1, 16) for _, a in got)
    assert numpy.array_equal(numpy.concatenate([a for _, a in got]), X)
    assert [k for k, _ in kaldi_io.read_ark(ark)] == list(ss.segset)
    kaldi_io.speaker_means(scp, {"spkA": ["utt0", "utt1"], "spkB": ["utt4"]}, str(tmp_path / "spk.scp"))
    spk = dict(kaldi_io.read_scp(str(tmp_path / "spk.scp")))
    m = X[:2].mean(0, keepdims=True)
    assert numpy.allclose(spk["spkA"], m / numpy.linalg.norm(m)) and abs(numpy.linalg.norm(spk["spkB"]) - 1) < 1e-6
    with kaldi_io.ArkScpWriter(str(tmp_path / "v.ark"), str(tmp_path / "v.scp")) as w:
        w("vec", X[3])
    assert numpy.array_equal(dict(kaldi_io.read_scp(str(tmp_path / "v.scp")))["vec"], X[3])
    import pytest
    with pytest.raises(ValueError):
        kaldi_io.ArkScpWriter(str(tmp_path / "bad.ark"))("two words", X[0])


def _xv_setup(tmp_path):
    """Three utterances: a plain 16 kHz path, an 8 kHz file (resampled on the device) and a `cat file |` pipe entry."""
    import wave as wavmod
    d = str(tmp_path)
    pcm = {}
    for name, rate, n, seed in (("a", 16000, 30000, 1), ("b", 8000, 12000, 2), ("c", 16000, 21000, 3)):
        x = (synth.synth_wave(1, n, seed=1800 + seed)[0].numpy() * 32768.0).clip(-32768, 32767).astype(numpy.int16)
        with wavmod.open(os.path.join(d, name + ".wav"), "wb") as f:
            f.setnchannels(1); f.setsampwidth(2); f.setframerate(rate)
            f.writeframes(x.tobytes())
        pcm[name] = (x.astype(numpy.float32) / 32768.0, rate)
    with open(os.path.join(d, "wav.scp"), "w") as f:
        f.write("utt-a %s/a.wav\nutt-b %s/b.wav\nutt-c cat %s/c.wav |\n" % (d, d, d))
    with open(os.path.join(d, "spk2utt"), "w") as f:
        f.write("spk1 utt-a utt-c\nspk2 utt-b\n")
    return d, pcm


def test_wav_scp_entries_paths_and_pipes(tmp_path):
    from sidekit_b200 import extract_xvectors as X
    d, pcm = _xv_setup(tmp_path)
    scp = X.read_wav_scp(os.path.join(d, "wav.scp"))
    assert list(scp.keys()) == ["utt-a", "utt-b", "utt-c"] and scp["utt-c"][0] == "cat" and scp["utt-c"][-1] == "|"
    for key, name in (("utt-a", "a"), ("utt-b", "b"), ("utt-c", "c")):
        sig, sr = X.prepare(scp[key])
        assert sr == pcm[name][1] and numpy.array_equal(sig.numpy(), pcm[name][0])
    with pytest.raises(IOError):
        X.prepare(["cat", os.path.join(d, "missing.wav"), "|"])


@pytest.mark.gpu
def test_extract_xvectors_main_writes_kaldi_tables(tmp_path):
    """extract_xvectors.py:93-173 end to end: checkpoint file -> load_model -> wav.scp (path / other rate / pipe) -> ark + scp
    + speaker means; vectors equal the direct extraction of the same (resampled) signals bit for bit."""
    from oracle import resample_ref as RR
    from sidekit_b200 import extract_xvectors as X, kaldi_io
    d, pcm = _xv_setup(tmp_path)
    m = make_xtractor("halfresnet34", 32, 256).cuda()
    ckpt = os.path.join(d, "model.pt")
    torch.save({"speaker_number": 32, "model_archi": {"model_type": "halfresnet34", "loss": {"type": "aam"}},
                "model_state_dict": {k: v.cpu() for k, v in m.state_dict().items()}}, ckpt)
    model, cfg = X.load_model(ckpt, "cuda")
    assert cfg["speaker_number"] == 32
    out_scp, spk_scp = os.path.join(d, "xv.scp"), os.path.join(d, "spk_xv.scp")
    X.main(model, os.path.join(d, "wav.scp"), out_scp, "cuda", False, 2000, 1500, 16000, spk_scp, os.path.join(d, "spk2utt"))
    table = dict(kaldi_io.read_scp(out_scp))
    assert list(table.keys()) == ["utt-a", "utt-b", "utt-c"] and all(v.shape == (1, 256) for v in table.values())
    b16 = sk.nnet.Resample(8000, 16000)(torch.from_numpy(pcm["b"][0]).cuda())
    assert numpy.abs(b16.cpu().numpy() - RR.resample(pcm["b"][0], 8000, 16000)).max() < 2e-6
    direct = m.extract_varlen([torch.from_numpy(pcm["a"][0]).cuda(), b16, torch.from_numpy(pcm["c"][0]).cuda()]).cpu().numpy()
    assert numpy.array_equal(numpy.concatenate([table[k] for k in ("utt-a", "utt-b", "utt-c")]), direct)
    spk = dict(kaldi_io.read_scp(spk_scp))
    mean = direct[[0, 2]].mean(axis=0)
    assert numpy.allclose(spk["spk1"], mean / numpy.linalg.norm(mean), atol=1e-6) and abs(numpy.linalg.norm(spk["spk2"]) - 1) < 1e-6
    # streamed in windows (the corpus is never resident): one utterance per window gives the same tables
    out2 = os.path.join(d, "xv_windowed.scp")
    X.main(model, os.path.join(d, "wav.scp"), out2, "cuda", window_audio_seconds=0.01)
    table2 = dict(kaldi_io.read_scp(out2))
    assert list(table2.keys()) == list(table.keys()) and all(numpy.array_equal(table2[k], table[k]) for k in table)
    with pytest.raises(NotImplementedError):
        X.main(model, os.path.join(d, "wav.scp"), out_scp, "cuda", True)


# ----------------------------------------------------------------------------- IdMapSet against the reference's recorded outputs
FEED_CASES = [("whole", "f16k_a", None, None, False), ("whole_start_only", "f16k_a", 50, None, False),
              ("segment", "f16k_a", 30, 150, False), ("too_short_recentred", "f16k_a", 100, 105, False),
              ("too_short_at_file_start", "f16k_b", 0, 4, False), ("whole_8k_resampled", "f8k", None, None, False),
              ("sliding", "f16k_b", None, None, True), ("sliding_segment", "f16k_b", 20, 330, True)]
FEED_RATES = {"f16k_a": 16000, "f16k_b": 16000, "f8k": 8000}


from tests.helpers import golden  # noqa: E402


def _feed_files(tmp_path, g):
    import wave as wavmod
    for name, rate in FEED_RATES.items():
        with wavmod.open(os.path.join(str(tmp_path), name + ".wav"), "wb") as f:
            f.setnchannels(1); f.setsampwidth(2); f.setframerate(rate)
            f.writeframes(g["pcm_" + name].tobytes())


def _feed_item(tmp_path, fname, start, stop, sliding):
    im = sk.IdMap()
    im.leftids, im.rightids = numpy.array(["spk"], dtype="|O"), numpy.array([fname], dtype="|O")
    im.start, im.stop = numpy.array([start], dtype="|O"), numpy.array([stop], dtype="|O")
    ds = sk.IdMapSet(im, str(tmp_path), "wav", sliding_window=sliding, window_len=1.0, window_shift=0.5, sample_rate=16000,
                     min_duration=0.165)
    return ds[0]


def test_idmapset_items_match_the_reference_recordings(tmp_path):
    """tests/golden/feed_path.npz = the reference's IdMapSet.__getitem__ (xsets.py:419-464) on these files: whole file,
    start only, segment, too-short segment (recentred min_duration window, also clamped at the file start), sliding windows."""
    from oracle import feed_ref as FR
    g = golden("feed_path.npz")
    _feed_files(tmp_path, g)
    for name, fname, start, stop, sliding in FEED_CASES:
        if FEED_RATES[fname] != 16000:
            continue                                             # needs the device resampler: GPU test below
        speech, left, right, s0, s1 = _feed_item(tmp_path, fname, start, stop, sliding)
        assert (left, right) == ("spk", fname)
        assert [int(s0), int(s1)] == g[name + "_bounds"].tolist(), name
        assert numpy.array_equal(speech.numpy(), g[name + "_speech"]), name
        # and the oracle's restatement of the same arithmetic
        seg, o0, o1 = FR.cut_segment(g["pcm_" + fname], start, stop, 16000, 0.165)
        if sliding:
            seg = FR.windows(seg, 1.0, 0.5)
        assert [o0, o1] == g[name + "_bounds"].tolist() and numpy.array_equal(seg, g[name + "_speech"]), name


@pytest.mark.gpu
def test_idmapset_resamples_like_the_reference(tmp_path):
    g = golden("feed_path.npz")
    _feed_files(tmp_path, g)
    speech, _, _, s0, s1 = _feed_item(tmp_path, "f8k", None, None, False)
    assert [int(s0), int(s1)] == g["whole_8k_resampled_bounds"].tolist()
    assert speech.shape == g["whole_8k_resampled_speech"].shape
    assert numpy.abs(speech.cpu().numpy() - g["whole_8k_resampled_speech"]).max() < 2e-6
