"""GPU diagnostic (not collected by pytest): per-stage error of the CUDA path against the oracle.

    python -m tests.diag_gpu [halfresnet34|xvector] > gpurun_out/diag.log

Prints, for every stage of the network, the relative L2 error and max abs error against the fp32
oracle so a wrong kernel can be localised from one remote run.
"""
import sys
import time

import torch

from oracle import extract_ref as R
from sidekit_b200 import synth
from tests.models import make_xtractor


def err(name, got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    rel = ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()
    mx = (got - ref).abs().max().item()
    print("%-12s rel_l2 %.3e  max_abs %.3e  ref_rms %.3e  shape %s" % (name, rel, mx, ref.pow(2).mean().sqrt().item(),
                                                                       tuple(ref.shape)), flush=True)
    return rel


def diag_hr34(lengths=(16000, 24160, 11111)):
    m = make_xtractor("halfresnet34", 32, 256).cuda()
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    waves = [synth.synth_wave(1, L, seed=100 + i)[0] for i, L in enumerate(lengths)]
    refs = []
    for w in waves:
        col = {}
        with torch.no_grad():
            lo, em = R.halfresnet34_forward(sd, w.unsqueeze(0), collect=col)
        col["logits"], col["emb"] = lo, em
        refs.append(col)
    cw = [w.cuda() for w in waves]
    # front-end (equal-length API) on the first utterance
    f = m.preprocessor(cw[0], is_eval=True)
    err("feats", f[0], refs[0]["feats"][0])
    stages = ["stem"] + ["layer%d.%d" % (l + 1, b) for l, n in enumerate((3, 4, 6, 3)) for b in range(n)]
    for st in stages:
        out = m.debug_stage(cw, st)            # (B, C, Hmax, W)
        worst = 0.0
        for i, r in enumerate(refs):
            ref = r[st][0]                     # (C, H, W)
            H = ref.shape[1]
            got = out[i, :, :H, :]
            rel = ((got.double().cpu() - ref.double()).norm() / ref.double().norm()).item()
            worst = max(worst, rel)
            if out.shape[2] > H:
                assert out[i, :, H:, :].abs().max().item() == 0.0
        print("%-12s worst rel_l2 over utts %.3e" % (st, worst), flush=True)
    pooled = m.debug_stage(cw, "pooled")
    err("pooled", pooled, torch.cat([r["pooled"] for r in refs]))
    logits, emb = m.extract_varlen(cw, want_logits=True)
    err("emb", emb, torch.cat([r["emb"] for r in refs]))
    err("logits", logits, torch.cat([r["logits"] for r in refs]))
    cos = torch.nn.functional.cosine_similarity(emb.double().cpu(), torch.cat([r["emb"] for r in refs]).double()).min().item()
    print("min cosine %.7f" % cos)
    # one-by-one vs packed
    solo = torch.cat([m(w, is_eval=True)[1] for w in cw])
    err("packed-vs-1", emb, solo)


def diag_tdnn(lengths=(32000, 51234, 20000)):
    m = make_xtractor("xvector", 32, 512).cuda()
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    waves = [synth.synth_wave(1, L, seed=200 + i)[0] for i, L in enumerate(lengths)]
    cw = [w.cuda() for w in waves]
    f = m.preprocessor(cw[0], is_eval=True)
    err("mfcc", f[0], R.mfcc_frontend(sd, waves[0])[0])
    refs = [R.tdnn_forward(sd, w.unsqueeze(0)) for w in waves]
    logits, emb = m.extract_varlen(cw, want_logits=True)
    err("emb", emb, torch.cat([r[1] for r in refs]))
    err("logits", logits, torch.cat([r[0] for r in refs]))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "halfresnet34"
    t0 = time.time()
    with torch.no_grad():
        if which == "halfresnet34":
            diag_hr34()
        else:
            diag_tdnn()
    torch.cuda.synchronize()
    print("done in %.1f s" % (time.time() - t0))
