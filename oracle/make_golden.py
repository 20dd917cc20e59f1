"""Regenerate tests/golden/*.npz from the REAL reference (build container only).

    python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  The reference is imported by oracle/ref_import.py
(stub modules + patches P1/P2); weights come from
``sidekit_b200.synth.fill_state_dict`` (key-hash seeded, so they can be rebuilt
on the GPU box without a checkpoint).  Only inputs that cannot be regenerated
(the real-audio snippet) and the reference's OUTPUTS are stored.
"""
import copy
import os
import sys
import wave as wavmod

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import                      # noqa: E402
from sidekit_b200 import synth                     # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _ref_model(archi, emb, n_spk=32, seed=0):
    m = ref_import.build_xtractor(n_spk, archi, emb)
    sd = m.state_dict()
    synth.fill_state_dict(sd, seed)
    m.load_state_dict(sd)
    return m


def golden_extraction():
    torch.set_num_threads(8)
    out = {}
    # HalfResNet34: two equal-length utterances (batched) + two of other lengths (run one by one, like
    # the reference's extractors do) -> exercises every T parity through the three stride-2 stages.
    m = _ref_model("halfresnet34", 256)
    lengths = [16000, 16000, 24160, 11111]
    waves = [synth.synth_wave(1, L, seed=100 + i)[0] for i, L in enumerate(lengths)]
    with torch.no_grad():
        lo01, em01 = m(torch.stack(waves[:2]), is_eval=True)
        res = [m(w, is_eval=True) for w in waves[2:]]
        feats0 = m.preprocessor(waves[0], True)
    out["hr_lengths"] = numpy.array(lengths)
    out["hr_seeds"] = numpy.array([100, 101, 102, 103])
    out["hr_emb"] = torch.cat([em01] + [r[1] for r in res]).numpy()
    out["hr_logits"] = torch.cat([lo01] + [r[0] for r in res]).numpy()
    out["hr_feats0"] = feats0[0].numpy()
    # TDNN
    m = _ref_model("xvector", 512)
    lengths = [32000, 32000, 51234]
    waves = [synth.synth_wave(1, L, seed=200 + i)[0] for i, L in enumerate(lengths)]
    with torch.no_grad():
        lo01, em01 = m(torch.stack(waves[:2]), is_eval=True)
        lo2, em2 = m(waves[2].unsqueeze(0), is_eval=True)
        feats0 = m.preprocessor(waves[0].unsqueeze(0), True)
    out["td_lengths"] = numpy.array(lengths)
    out["td_seeds"] = numpy.array([200, 201, 202])
    out["td_emb"] = torch.cat([em01, em2]).numpy()
    out["td_logits"] = torch.cat([lo01, lo2]).numpy()
    out["td_feats0"] = feats0[0].numpy()
    # real audio: first 1.5 s of one of the reference's example wavs through the log-Mel front-end
    wdir = os.path.join(ref_import.REFERENCE_ROOT, "egs", "examples_decode")
    wavs = sorted(f for f in os.listdir(wdir) if f.endswith(".wav"))
    with wavmod.open(os.path.join(wdir, wavs[0]), "rb") as f:
        assert f.getframerate() == 16000 and f.getnchannels() == 1 and f.getsampwidth() == 2
        pcm = numpy.frombuffer(f.readframes(24000), dtype=numpy.int16).copy()
    x = torch.from_numpy(pcm.astype(numpy.float32) / 32768.0)
    m = _ref_model("halfresnet34", 256)
    with torch.no_grad():
        feats = m.preprocessor(x, True)
        lo, em = m(x, is_eval=True)
    out["real_pcm16"] = pcm
    out["real_feats"] = feats[0].numpy()
    out["real_emb"] = em.numpy()
    numpy.savez_compressed(os.path.join(GOLD, "extraction.npz"), **out)
    print("extraction.npz", {k: v.shape for k, v in out.items()})


def golden_extraction_resnet34():
    """The full-width ResNet34 trunk (xvector.py:515-537): a batched pair and one odd length, outputs only."""
    torch.set_num_threads(8)
    out = {}
    m = _ref_model("resnet34", 256)
    lengths = [16000, 16000, 20321]
    waves = [synth.synth_wave(1, L, seed=400 + i)[0] for i, L in enumerate(lengths)]
    with torch.no_grad():
        lo01, em01 = m(torch.stack(waves[:2]), is_eval=True)
        lo2, em2 = m(waves[2], is_eval=True)
    out["r34_lengths"] = numpy.array(lengths)
    out["r34_seeds"] = numpy.array([400, 401, 402])
    out["r34_emb"] = torch.cat([em01, em2]).numpy()
    out["r34_logits"] = torch.cat([lo01, lo2]).numpy()
    out["r34_blocks"] = numpy.array([len(getattr(m.sequence_network, "layer%d" % i)) for i in range(1, 8)])
    # FastResNet34 (xvector.py:539-567): 7x7 stride-(1,2) stem, 16/32/64/128 channels, attentive pooling without global context
    m = _ref_model("fastresnet34", 256)
    lengths = [16000, 16000, 24160, 11111]
    waves = [synth.synth_wave(1, L, seed=500 + i)[0] for i, L in enumerate(lengths)]
    with torch.no_grad():
        lo01, em01 = m(torch.stack(waves[:2]), is_eval=True)
        res = [m(w, is_eval=True) for w in waves[2:]]
    out["f34_lengths"] = numpy.array(lengths)
    out["f34_seeds"] = numpy.array([500, 501, 502, 503])
    out["f34_emb"] = torch.cat([em01] + [r[1] for r in res]).numpy()
    out["f34_logits"] = torch.cat([lo01] + [r[0] for r in res]).numpy()
    out["f34_shortcuts"] = numpy.array([int(len(getattr(m.sequence_network, "layer%d" % i)[0].shortcut) > 0) for i in range(1, 5)])
    numpy.savez_compressed(os.path.join(GOLD, "extraction_resnet34.npz"), **out)
    print("extraction_resnet34.npz", {k: v.shape for k, v in out.items()}, out["r34_blocks"])


def _statserver(sidekit, ids, X):
    s = sidekit.StatServer()
    s.modelset = numpy.array(ids)
    s.segset = numpy.array(ids)
    s.start = numpy.empty(len(ids), dtype="|O")
    s.stop = numpy.empty(len(ids), dtype="|O")
    s.stat0 = numpy.ones((len(ids), 1))
    s.stat1 = numpy.array(X, dtype=numpy.float64)
    return s


def golden_scoring():
    sidekit = ref_import.import_reference()
    from sidekit.iv_scoring import cosine_scoring, PLDA_scoring, fast_PLDA_scoring, two_covariance_scoring
    from sidekit.score_normalization import asnorm
    rng = numpy.random.default_rng(4321)
    D, R = 16, 10
    Ne, Nt = 37, 29
    en_ids = ["m%02d" % i for i in range(Ne)]
    te_ids = ["s%02d" % i for i in range(Nt)]
    te_ids[7] = te_ids[3]                                   # duplicate test id -> first match wins
    E = rng.standard_normal((Ne, D))
    T = rng.standard_normal((Nt, D))
    mu = 0.1 * rng.standard_normal(D)
    F = 0.5 * rng.standard_normal((D, R))
    A = 0.3 * rng.standard_normal((D, D))
    Sigma = A @ A.T + numpy.eye(D)
    B = F @ F.T + 0.1 * numpy.eye(D)
    ndx = sidekit.Ndx()
    perm_m = rng.permutation(Ne)[:30]
    perm_s = rng.permutation(Nt)[:25]
    ndx.modelset = numpy.array([en_ids[i] for i in perm_m] + ["missing_model", en_ids[perm_m[0]]])
    ndx.segset = numpy.array([te_ids[i] for i in perm_s] + ["missing_seg"])
    ndx.trialmask = rng.random((ndx.modelset.shape[0], ndx.segset.shape[0])) < 0.6
    out = dict(en_ids=numpy.array(en_ids), te_ids=numpy.array(te_ids), E=E, T=T, mu=mu, F=F, Sigma=Sigma, B=B,
               ndx_models=ndx.modelset, ndx_segs=ndx.segset, trialmask=ndx.trialmask)

    def pack(name, sc):
        out[name + "_modelset"] = numpy.array(sc.modelset)
        out[name + "_segset"] = numpy.array(sc.segset)
        out[name + "_mask"] = numpy.array(sc.scoremask)
        out[name + "_mat"] = numpy.array(sc.scoremat)

    mk = lambda: (_statserver(sidekit, en_ids, E), _statserver(sidekit, te_ids, T), copy.deepcopy(ndx))
    pack("cosine", cosine_scoring(*mk(), device=torch.device("cpu")))
    pack("plda", PLDA_scoring(*mk(), mu, F, numpy.zeros((D, 0)), Sigma))
    pack("plda_sf", PLDA_scoring(*mk(), mu, F, numpy.zeros((D, 0)), Sigma, scaling_factor=0.7))
    pack("plda_open", fast_PLDA_scoring(*mk(), mu, F, Sigma, p_known=0.3))
    pack("twocov", two_covariance_scoring(*mk(), Sigma, B))
    # duplicate enrol models -> averaged per sorted-unique model (fast-PLDA path)
    en_dup = list(en_ids)
    en_dup[5] = en_dup[2]
    en_dup[11] = en_dup[2]
    out["en_ids_dup"] = numpy.array(en_dup)
    pack("plda_dup", fast_PLDA_scoring(_statserver(sidekit, en_dup, E), _statserver(sidekit, te_ids, T),
                                       copy.deepcopy(ndx), mu, F, Sigma))
    # as-norm
    N, C, Dx = 40, 260, 32
    X = rng.standard_normal((N, Dx)).astype(numpy.float32)
    X /= numpy.linalg.norm(X, axis=1, keepdims=True)
    coh = rng.standard_normal((C, Dx)).astype(numpy.float32)
    out["asnorm_X"] = X
    out["asnorm_cohort"] = coh
    out["asnorm_out"] = asnorm(torch.from_numpy(X), torch.from_numpy(coh), None)
    numpy.savez_compressed(os.path.join(GOLD, "scoring.npz"), **out)
    print("scoring.npz", sorted(out.keys()))


def golden_scoring_full():
    """full_PLDA_scoring (channel subspace G) of the real reference: closed set, scaled, open set."""
    sidekit = ref_import.import_reference()
    from sidekit.iv_scoring import PLDA_scoring, full_PLDA_scoring
    rng = numpy.random.default_rng(2468)
    D, R, Rc = 16, 8, 5
    Ne, Nt = 23, 31
    en_ids = ["m%02d" % i for i in range(Ne)]
    te_ids = ["s%02d" % i for i in range(Nt)]
    E, T = rng.standard_normal((Ne, D)), rng.standard_normal((Nt, D))
    mu = 0.1 * rng.standard_normal(D)
    F = 0.5 * rng.standard_normal((D, R))
    G = 0.4 * rng.standard_normal((D, Rc))
    A = 0.3 * rng.standard_normal((D, D))
    Sigma = A @ A.T + numpy.eye(D)
    ndx = sidekit.Ndx()
    ndx.modelset = numpy.array([en_ids[i] for i in rng.permutation(Ne)[:20]] + ["missing_model"])
    ndx.segset = numpy.array([te_ids[i] for i in rng.permutation(Nt)[:27]])
    ndx.trialmask = rng.random((ndx.modelset.shape[0], ndx.segset.shape[0])) < 0.7
    out = dict(en_ids=numpy.array(en_ids), te_ids=numpy.array(te_ids), E=E, T=T, mu=mu, F=F, G=G, Sigma=Sigma,
               ndx_models=ndx.modelset, ndx_segs=ndx.segset, trialmask=ndx.trialmask)
    mk = lambda: (_statserver(sidekit, en_ids, E), _statserver(sidekit, te_ids, T), copy.deepcopy(ndx))
    for name, sc in (("full", PLDA_scoring(*mk(), mu, F, G, Sigma, full_model=True)),
                     ("full_sf", full_PLDA_scoring(*mk(), mu, F, G, Sigma, scaling_factor=0.6)),
                     ("full_open", full_PLDA_scoring(*mk(), mu, F, G, Sigma, p_known=0.25))):
        out[name + "_modelset"], out[name + "_segset"] = numpy.array(sc.modelset), numpy.array(sc.segset)
        out[name + "_mask"], out[name + "_mat"] = numpy.array(sc.scoremask), numpy.array(sc.scoremat)
    numpy.savez_compressed(os.path.join(GOLD, "scoring_full.npz"), **out)
    print("scoring_full.npz", sorted(out.keys()))


def golden_plda_training():
    """FactorAnalyser.plda of the real reference on a seeded synthetic StatServer (no file output)."""
    sidekit = ref_import.import_reference()
    from sidekit.factor_analyser import FactorAnalyser
    rng = numpy.random.default_rng(1357)
    D, R, n_spk = 24, 6, 40
    spk_means = rng.standard_normal((n_spk, R)) @ (0.8 * rng.standard_normal((R, D)))
    counts = rng.integers(2, 7, size=n_spk)
    ids, X = [], []
    for s_, n in enumerate(counts):
        for _ in range(int(n)):
            ids.append("spk%02d" % s_)
            X.append(spk_means[s_] + 0.5 * rng.standard_normal(D) + 0.2)
    perm = rng.permutation(len(ids))
    ids, X = numpy.array(ids)[perm], numpy.array(X)[perm]
    ss = _statserver(sidekit, ids, X)
    ss.segset = numpy.array(["seg%03d" % i for i in range(len(ids))])
    out = dict(ids=ids, X=X)
    for name, kw in (("it3", dict(nb_iter=3)), ("it5_sf", dict(nb_iter=5, scaling_factor=0.5))):
        fa = FactorAnalyser()
        fa.plda(copy.deepcopy(ss), R, save_final=False, **kw)
        out[name + "_mean"], out[name + "_F"], out[name + "_Sigma"] = fa.mean, fa.F, fa.Sigma
    numpy.savez_compressed(os.path.join(GOLD, "plda_training.npz"), **out)
    print("plda_training.npz", sorted(out.keys()))


def golden_conditioning():
    """StatServer conditioning functions of the real reference (covariances, LDA, WCCN, whitening, spectral norm)."""
    sidekit = ref_import.import_reference()
    g = numpy.load(os.path.join(GOLD, "plda_training.npz"))
    ids, X = g["ids"], g["X"]
    ss = _statserver(sidekit, ids, X)
    out = {}
    out["mean"] = ss.get_mean_stat1()
    out["total_cov"] = ss.get_total_covariance_stat1()
    out["within_cov"] = ss.get_within_covariance_stat1()
    out["between_cov"] = ss.get_between_covariance_stat1()
    out["lda"] = ss.get_lda_matrix_stat1(5)
    out["mahalanobis"] = ss.get_mahalanobis_matrix_stat1()
    out["wccn"] = ss.get_wccn_choleski_stat1()
    summed, sessions = ss.sum_stat_per_model()
    out["sum_modelset"], out["sum_stat1"], out["sum_sessions"] = summed.modelset, summed.stat1, sessions
    w = copy.deepcopy(ss); w.whiten_stat1(out["mean"], out["total_cov"]); out["whiten_full"] = w.stat1
    w = copy.deepcopy(ss); w.whiten_stat1(out["mean"], numpy.diag(out["total_cov"]).copy()); out["whiten_diag"] = w.stat1
    w = copy.deepcopy(ss); w.whiten_cholesky_stat1(out["mean"], out["total_cov"]); out["whiten_chol"] = w.stat1
    for mode in ("efr", "sphNorm"):
        m_, c_ = copy.deepcopy(ss).estimate_spectral_norm_stat1(2, mode)
        out["sn_%s_mean" % mode], out["sn_%s_cov" % mode] = numpy.stack(m_), numpy.stack(c_)
        w = copy.deepcopy(ss); w.spectral_norm_stat1(m_, c_); out["sn_%s_out" % mode] = w.stat1
    numpy.savez_compressed(os.path.join(GOLD, "conditioning.npz"), **out)
    print("conditioning.npz", sorted(out.keys()))


def golden_evaltail():
    """PAV / ROCCH / EER / minDCF / Key / z-t-norm outputs of the real reference on seeded scores."""
    sidekit = ref_import.import_reference()
    from sidekit.bosaris.detplot import pavx, rocch, rocch2eer, fast_minDCF
    from sidekit.nnet.xvector import eer
    from sidekit.score_normalization import znorm, tnorm, ztnorm
    from sidekit.bosaris import Key, Scores
    rng = numpy.random.default_rng(777)
    out = {}
    cases = {"sep": (rng.normal(2.0, 1.0, 400), rng.normal(-1.0, 1.2, 1500)),
             "overlap": (rng.normal(0.3, 1.0, 300), rng.normal(0.0, 1.0, 900)),
             "ties": (numpy.round(rng.normal(1.0, 1.0, 250), 1), numpy.round(rng.normal(0.0, 1.0, 700), 1)),
             "tiny": (numpy.array([0.2, 0.9, 0.5]), numpy.array([0.1, 0.6, 0.4, 0.55]))}
    for name, (tar, non) in cases.items():
        out[name + "_tar"], out[name + "_non"] = tar, non
        pm, pf = rocch(tar, non)
        out[name + "_pmiss"], out[name + "_pfa"] = pm, pf
        out[name + "_rocch_eer"] = numpy.array(rocch2eer(pm, pf), dtype=numpy.float64)
        out[name + "_mindcf"] = numpy.array(fast_minDCF(tar, non, -2.0, normalize=True), dtype=numpy.float64)
        out[name + "_eer"] = numpy.array(eer(non, tar), dtype=numpy.float64)
    y = numpy.concatenate([rng.random(60), numpy.array([0.5, 0.5, 0.25, 0.75, 0.75])])
    g, w, h = pavx(y)
    out["pav_y"], out["pav_ghat"], out["pav_width"], out["pav_height"] = y, g, w, h
    # Key built from trial lists (with a repeated pair: the last listing wins) and get_tar_non through align_with_ndx
    models = numpy.array(["m%d" % (i % 7) for i in range(60)] + ["m0"])
    segs = numpy.array(["s%d" % ((i * 5) % 11) for i in range(60)] + ["s0"])
    trials = numpy.array(["target" if (i % 3 == 0) else "nontarget" for i in range(60)] + ["nontarget"])
    key = Key(models=models, testsegs=segs, trials=trials)
    out["key_models"], out["key_segs"], out["key_trials"] = models, segs, trials
    out["key_modelset"], out["key_segset"], out["key_tar"], out["key_non"] = key.modelset, key.segset, key.tar, key.non
    sc = Scores()
    sc.modelset = numpy.array(["m%d" % i for i in (3, 0, 6, 1, 9, 2, 5)])       # m4 missing, m9 extra, other order
    sc.segset = numpy.array(["s%d" % i for i in (10, 2, 7, 0, 5, 1, 3, 8, 6, 4)])   # s9 missing
    sc.scoremat = rng.standard_normal((7, 10))
    sc.scoremask = rng.random((7, 10)) < 0.8
    out["sc_modelset"], out["sc_segset"], out["sc_mat"], out["sc_mask"] = sc.modelset, sc.segset, sc.scoremat, sc.scoremask
    t, n = sc.get_tar_non(key)
    out["sc_tar"], out["sc_non"] = t, n
    # z / t / zt-norm on square score sets (the reference's znorm broadcasts per-model statistics against the last axis)
    def scores(M, S, seed, mpre, spre):
        r = numpy.random.default_rng(seed)
        s_ = Scores()
        s_.modelset = numpy.array(["%s%02d" % (mpre, i) for i in r.permutation(M)])
        s_.segset = numpy.array(["%s%02d" % (spre, i) for i in r.permutation(S)])
        s_.scoremat = r.standard_normal((M, S)) + 0.1 * numpy.arange(S)
        s_.scoremask = numpy.ones((M, S), dtype=bool)
        return s_
    n_ = 12
    et, ei, it, ii = scores(n_, n_, 1, "e", "t"), scores(n_, n_, 2, "e", "i"), scores(n_, n_, 3, "i", "t"), scores(n_, n_, 4, "i", "i")
    for nm, s_ in (("et", et), ("ei", ei), ("it", it), ("ii", ii)):
        out["zt_%s_modelset" % nm], out["zt_%s_segset" % nm], out["zt_%s_mat" % nm] = s_.modelset.copy(), s_.segset.copy(), s_.scoremat.copy()
    out["znorm_mat"] = znorm(copy.deepcopy(et), copy.deepcopy(ei)).scoremat
    out["znorm_sym_mat"] = znorm(copy.deepcopy(it), copy.deepcopy(ii), sym=True).scoremat
    out["tnorm_mat"] = tnorm(copy.deepcopy(et), copy.deepcopy(it)).scoremat
    zt = ztnorm(copy.deepcopy(et), copy.deepcopy(ei), copy.deepcopy(it), copy.deepcopy(ii))
    out["ztnorm_mat"], out["ztnorm_modelset"], out["ztnorm_segset"] = zt.scoremat, zt.modelset, zt.segset
    numpy.savez_compressed(os.path.join(GOLD, "evaltail.npz"), **out)
    print("evaltail.npz", sorted(out.keys()))


def text_io_reference(out_dir, objects):
    """Write ``objects`` (tests/test_text_io.py:_objects, sidekit_b200 containers) with the REFERENCE's writers and read the
    files back with its readers -> dict of arrays.  ``check_path_existance`` (sidekit_wrappers.py:73-89) drops the
    return value of the decorated readers, so the undecorated functions are taken from the wrapper's closure."""
    ref_import.import_reference()
    from sidekit.bosaris import IdMap, Key, Ndx, Scores
    key, ndx, sc, sc32, im = objects

    def raw(method):
        f = getattr(method, "__func__", method)
        return f.__closure__[0].cell_contents if f.__closure__ else f

    def fill(dst, src, names):
        for n in names:
            setattr(dst, n, getattr(src, n))
        return dst

    p = lambda n: os.path.join(out_dir, n)
    fill(Key(), key, ("modelset", "segset", "tar", "non")).write_txt(p("key.txt"))
    fill(Ndx(), ndx, ("modelset", "segset", "trialmask")).save_txt(p("ndx.txt"))
    fill(Scores(), sc, ("modelset", "segset", "scoremask", "scoremat")).write_txt(p("scores64.txt"))
    fill(Scores(), sc32, ("modelset", "segset", "scoremask", "scoremat")).write_txt(p("scores32.txt"))
    rim = fill(IdMap(), im, ("leftids", "rightids", "start", "stop"))
    rim.write_txt(p("idmap4.txt"))
    rim.start = numpy.array([None] * len(im.leftids), dtype="|O")
    rim.stop = numpy.array([None] * len(im.leftids), dtype="|O")
    rim.write_txt(p("idmap_none.txt"))
    with open(p("idmap2.txt"), "w") as f:
        f.writelines("%s %s\n" % (a, b) for a, b in zip(im.leftids, im.rightids))
    out = {}
    k = Key.read_txt(p("key.txt"))
    out.update(key_modelset=k.modelset.astype("U"), key_segset=k.segset.astype("U"), key_tar=k.tar, key_non=k.non)
    n = raw(Ndx.read_txt)(Ndx, p("ndx.txt"))
    out.update(ndx_modelset=n.modelset.astype("U"), ndx_segset=n.segset.astype("U"), ndx_trialmask=n.trialmask)
    for name in ("scores64", "scores32"):
        s_ = raw(Scores.read_txt)(Scores, p(name + ".txt"))
        out.update({name + "_modelset": s_.modelset.astype("U"), name + "_segset": s_.segset.astype("U"),
                    name + "_scoremask": s_.scoremask, name + "_scoremat": s_.scoremat})
    i4 = raw(IdMap.read_txt)(IdMap, p("idmap4.txt"))
    out.update(idmap4_leftids=i4.leftids.astype("U"), idmap4_rightids=i4.rightids.astype("U"), idmap4_start=i4.start, idmap4_stop=i4.stop)
    i2 = raw(IdMap.read_txt)(IdMap, p("idmap2.txt"))
    assert all(v is None for v in i2.start)
    out.update(idmap2_leftids=i2.leftids.astype("U"))
    return out


def golden_text_io():
    sys.path.insert(0, os.path.join(ROOT))
    from tests.test_text_io import _objects
    d = os.path.join(GOLD, "text_io")
    os.makedirs(d, exist_ok=True)
    out = text_io_reference(d, _objects())
    numpy.savez_compressed(os.path.join(d, "read_back.npz"), **out)
    print("text_io/", sorted(os.listdir(d)))


def golden_resample():
    """torchaudio.transforms.Resample outputs (the third-party call of xsets.py:435 / extract_xvectors.py:144) on seeded
    Gaussian audio: the rate pairs a 16 kHz model meets in practice plus an upsampling and an awkward ratio."""
    import torchaudio
    out = {"torchaudio_version": numpy.array(torchaudio.__version__)}
    cases = [(44100, 16000, 3001), (48000, 16000, 2500), (8000, 16000, 1999), (22050, 16000, 4410), (16000, 8000, 777),
             (11025, 16000, 1500), (44100, 16000, 5)]
    for i, (fo, fn, n) in enumerate(cases):
        x = synth.synth_wave(2, n, seed=300 + i)
        y = torchaudio.transforms.Resample(fo, fn)(x)
        out["case%d_rates" % i] = numpy.array([fo, fn, n])
        out["case%d_y" % i] = y.numpy()
    out["n_cases"] = numpy.array(len(cases))
    numpy.savez_compressed(os.path.join(GOLD, "resample.npz"), **out)
    print("resample.npz", len(cases), "cases, torchaudio", torchaudio.__version__)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["scoring", "scoring_full", "extraction", "extraction_resnet34", "evaltail", "plda_training"]
    if "plda_training" in which:
        golden_plda_training()
    if "extraction_resnet34" in which:
        golden_extraction_resnet34()
    if "conditioning" in which or "plda_training" in which:
        golden_conditioning()
    if "scoring_full" in which:
        golden_scoring_full()
    if "scoring" in which:
        golden_scoring()
    if "extraction" in which:
        golden_extraction()
    if "evaltail" in which:
        golden_evaltail()
    if "resample" in which:
        golden_resample()
    if "text_io" in which:
        golden_text_io()
