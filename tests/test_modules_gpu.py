"""The reference's modules on their own: BasicBlock / SELayer / ResBlock / PreHalfResNet34 / AttentivePooling /
ArcMarginProduct ``forward`` through the stand-alone operators (``skb_conv2d_bn_act`` & co), against the oracle.
16-bit tensor-core operands, fp32 accumulation: activations within 2e-3 relative L2 per block (the gate of the engine's
per-block test), fp32 operators within 1e-4."""
import numpy
import pytest
import torch

from oracle import extract_ref as R
from sidekit_b200 import synth
from sidekit_b200.nnet import res_net, functional as Fn, AttentivePooling, ArcMarginProduct
from tests.models import make_xtractor

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


def _filled(module, seed):
    sd = module.state_dict()
    synth.fill_state_dict(sd, seed)
    module.load_state_dict(sd)
    return module.eval().cuda(), {k: v.cpu() for k, v in sd.items()}


def test_conv2d_bn_act_operator_shapes_strides_and_fusions():
    g = torch.Generator().manual_seed(0)
    for cin, cout, k, stride, H, W in ((32, 32, 3, 1, 13, 80), (32, 64, 3, 2, 13, 80), (64, 128, 1, 2, 12, 39), (1, 32, 3, 1, 9, 80),
                                       (24, 40, 3, 1, 6, 10), (128, 256, 3, 2, 7, 20), (256, 256, 3, 1, 5, 10)):
        x = torch.randn(3, cin, H, W, generator=g)
        w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
        b = torch.randn(cout, generator=g) * 0.1
        ref = torch.nn.functional.conv2d(x, w, b, stride=stride, padding=k // 2)
        got = Fn.conv2d_bn_act(x.cuda(), w.numpy(), b.numpy(), stride, 1.0)
        assert got.shape == ref.shape and _rel(got, ref) < 2e-3, (cin, cout, k, stride)
        got = Fn.conv2d_bn_act(x.cuda(), w.numpy(), b.numpy(), stride, 0.0)
        assert _rel(got, torch.relu(ref)) < 2e-3
        res = torch.randn(ref.shape, generator=g)
        sc = torch.rand(3, cout, generator=g)
        got = Fn.conv2d_bn_act(x.cuda(), w.numpy(), b.numpy(), stride, 0.01, se_scale=sc.cuda(), residual=res.cuda())
        assert _rel(got, torch.nn.functional.leaky_relu(ref * sc[:, :, None, None] + res, 0.01)) < 2e-3
    Fn.check_overflow("cuda")


def test_selayer_and_basicblock_forward_match_the_oracle():
    g = torch.Generator().manual_seed(1)
    se, sd = _filled(res_net.SELayer(64), 2)
    x = torch.randn(2, 64, 11, 20, generator=g)
    assert _rel(se(x.cuda()), R.se_layer({"b." + k: v for k, v in sd.items()}, "b", x)) < 1e-5
    for cin, cout, stride in ((32, 32, 1), (32, 32, (1, 1)), (32, 64, (2, 2)), (128, 256, (2, 2)), (256, 256, 1)):
        blk, sd = _filled(res_net.BasicBlock(cin, cout, stride), 3)
        x = torch.relu(torch.randn(2, cin, 13, 20, generator=g))
        ref = R.basic_block({"b." + k: v for k, v in sd.items()}, "b", x, stride[0] if isinstance(stride, tuple) else stride)
        got = blk(x.cuda())
        assert got.shape == ref.shape and _rel(got, ref) < 2e-3, (cin, cout, stride)


def test_resblock_forward_matches_the_oracle():
    """the pre-activation ResBlock (res_net.py:186-255) on the tcgen05 convolution kernel."""
    g = torch.Generator().manual_seed(2)
    for cin, cout, first in ((32, 32, False), (32, 64, False), (16, 16, True), (64, 128, False)):
        blk, sd = _filled(res_net.ResBlock(cin, cout, 1, is_first=first), 4)
        x = torch.randn(2, cin, 12, 17, generator=g)
        ref = R.res_block({"b." + k: v for k, v in sd.items()}, "b", x, is_first=first)
        got = blk(x.cuda())
        assert got.shape == ref.shape and _rel(got, ref) < 2e-3, (cin, cout, first)
        assert _rel(blk(x.cuda(), compute_dtype="bf16"), ref) < 1.5e-2


def test_prehalfresnet34_forward_matches_the_oracle_and_the_engine():
    m = make_xtractor("halfresnet34", 16, 256).cuda()
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    x = synth.synth_wave(2, 12000, seed=6)
    feats = R.logmel_frontend(sd, x)                                    # (B, 80, T)
    ref = R.halfresnet34_trunk(sd, feats)                               # (B, 256, T4, 10)
    got = m.sequence_network(feats.cuda())
    assert got.shape == ref.shape and _rel(got, ref) < 3e-3
    eng = m.debug_stage([w.cuda() for w in x], "layer4.2")
    assert _rel(got, eng) < 3e-3                                        # the fused engine computes the same trunk


def test_attentive_pooling_and_margin_head_forward():
    m = make_xtractor("halfresnet34", 40, 256).cuda()
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    x = torch.relu(torch.randn(3, 256, 17, 10, generator=g)) * 0.1      # (B, C, T, F) as the trunk emits
    ref = R.attentive_pooling(sd, x)
    got = m.stat_pooling(x.cuda())
    assert got.shape == ref.shape == (3, 5120) and _rel(got, ref) < 1e-4
    pool = AttentivePooling(8, 10, attention_channels=16, global_context=False)
    pool, psd = _filled(pool, 7)
    x2 = torch.randn(2, 8, 9, 10, generator=g)
    ref2 = R.attentive_pooling({"stat_pooling." + k: v for k, v in psd.items()}, x2, global_context=False)
    assert _rel(pool(x2.cuda()), ref2) < 1e-4
    e = torch.randn(5, 256, generator=g)
    head = m.after_speaker_embedding
    assert isinstance(head, ArcMarginProduct)
    got = head(e.cuda())
    ref = R.arc_margin_eval(sd["after_speaker_embedding.weight"], e, 30.0)
    assert got.shape == (5, 40) and (got.cpu() - ref).abs().max().item() < 1e-3
    with pytest.raises(NotImplementedError):
        head(e.cuda(), target=torch.zeros(5, dtype=torch.long))
