"""Every kernel family once at CI-size shapes, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Small enough to finish in minutes under the sanitizer's slowdown."""
import os
import sys

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sidekit_b200 as sk  # noqa: E402
from sidekit_b200 import synth  # noqa: E402
from sidekit_b200.iv_scoring import PackedEmbeddings, TrialIndex, score_matrix, score_trials  # noqa: E402
from sidekit_b200.nnet import res_net, Resample  # noqa: E402
from tests.models import make_xtractor  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.cuda.set_device(0)
with torch.no_grad():
    if which in ("all", "extract"):
        for archi, emb in (("halfresnet34", 256), ("xvector", 512), ("resnet34", 256), ("fastresnet34", 256)):
            m = make_xtractor(archi, 16, emb).cuda()
            waves = [synth.synth_wave(1, L, seed=i)[0].cuda() for i, L in enumerate((9000, 12345, 16000))]
            e = m.extract_varlen(waves, want_logits=True)
            m.reserve(8, 4.0)
            m.extract_packed(torch.cat(waves), [9000, 12345, 16000])
            print(archi, float(e[1].abs().sum()))
    if which in ("all", "modules"):
        blk = res_net.BasicBlock(32, 64, (2, 2)).eval().cuda()
        print("basicblock", float(blk(torch.randn(2, 32, 9, 20).cuda()).sum()))
        rb = res_net.ResBlock(32, 64, 1).eval().cuda()
        print("resblock", float(rb(torch.randn(2, 32, 7, 10).cuda()).sum()))
        m = make_xtractor("halfresnet34", 16, 256).cuda()
        print("attpool", float(m.stat_pooling(torch.randn(2, 256, 9, 10).cuda()).sum()))
        print("margin", float(m.after_speaker_embedding(torch.randn(3, 256).cuda()).sum()))
        print("resample", float(Resample(44100, 16000)(torch.randn(20000).cuda()).sum()))
    if which in ("all", "scoring"):
        E = torch.from_numpy(synth.synth_embeddings(300, 256, seed=1)).float().cuda()
        T = torch.from_numpy(synth.synth_embeddings(333, 256, seed=2, unit_norm=False)).float().cuda()
        r, q = torch.randn(300, device="cuda"), torch.randn(333, device="cuda")
        for passes in (1, 3):
            for dt in (torch.float32, torch.float64, torch.float16):
                score_matrix(E, T, r, q, 0.5, 1.5, passes=passes, out_dtype=dt)
        score_matrix(E, PackedEmbeddings(T), r, q, 0.5, 1.5)
        mask = numpy.random.default_rng(0).random((300, 333)) < 0.1
        print("trials", float(score_trials(E, T, TrialIndex(mask), r, q, 0.5, 1.5).sum()))
        X = torch.nn.functional.normalize(E, dim=1)
        print("asnorm", float(numpy.abs(sk.asnorm(X, torch.randn(260, 256).cuda(), None)).sum()))
        ids_e = numpy.array(["m%d" % i for i in range(300)])
        ids_t = numpy.array(["s%d" % i for i in range(333)])
        ndx = sk.Ndx()
        ndx.modelset, ndx.segset, ndx.trialmask = ids_e, ids_t, mask
        mu, F, Sigma = synth.synth_plda(256, 256, seed=3)
        sc = sk.PLDA_scoring(sk.StatServer.from_embeddings(ids_e, E.cpu().numpy()), sk.StatServer.from_embeddings(ids_t, T.cpu().numpy()),
                             ndx, mu, F, numpy.zeros((256, 0)), Sigma)
        print("plda", float(sc.scoremat.sum()))
        rng = numpy.random.default_rng(3)
        ids = numpy.repeat(numpy.array(["s%02d" % i for i in range(20)]), 4)
        fa = sk.FactorAnalyser().plda(sk.StatServer.from_embeddings(ids, rng.standard_normal((80, 24))), 6, nb_iter=2, save_final=False)
        print("plda_train", float(fa.F.sum()))
torch.cuda.synchronize()
print("sanitize_smoke done")
