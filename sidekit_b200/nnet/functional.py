"""Stand-alone module operators: thin ctypes wrappers over ``skb_conv2d_bn_act`` / ``skb_channel_mean`` /
``skb_se_gate`` / ``skb_scale_residual_act`` / ``skb_l2_normalize`` / ``skb_attentive_pool`` (csrc/module_ops.cuh), which
let the reference's modules (``BasicBlock``, ``SELayer``, ``ResBlock``, ``PreHalfResNet34``, ``AttentivePooling``,
``ArcMarginProduct``) run on their own with dense fp32 CUDA tensors.  Same kernels as the fused engine, plus a
dense <-> plane conversion per call: correct and native, not the fast path (that is ``Xtractor.forward``).
"""
import ctypes

import numpy
import torch

from .. import _lib

_COMPUTE = {"fp16": 0, "bf16": 1}
_ovf_seen = {}


def _need_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("sidekit_b200 has no CPU path: move the tensor to a CUDA device")
    return x.contiguous().float()


def fold_conv_bn(conv, bn=None):
    """Conv2d followed by an eval-mode BatchNorm2d as one affine map, folded in float64 on the host:
    ``w' = w * g / sqrt(var + eps)``, ``b' = (b - mean) * g / sqrt(var + eps) + beta``.  Returns fp32 numpy arrays."""
    w = conv.weight.detach().cpu().double()
    b = conv.bias.detach().cpu().double() if conv.bias is not None else torch.zeros(w.shape[0], dtype=torch.float64)
    if bn is not None:
        s = bn.weight.detach().cpu().double() / torch.sqrt(bn.running_var.detach().cpu().double() + bn.eps)
        w = w * s.view(-1, 1, 1, 1)
        b = (b - bn.running_mean.detach().cpu().double()) * s + bn.bias.detach().cpu().double()
    return numpy.ascontiguousarray(w.float().numpy()), numpy.ascontiguousarray(b.float().numpy())


def bn_affine(bn, device):
    """Eval-mode BatchNorm as per-channel (scale, shift) device tensors."""
    s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    t = bn.bias.detach().double() - bn.running_mean.detach().double() * s
    return s.float().to(device).contiguous(), t.float().to(device).contiguous()


def check_overflow(device):
    """fp16 range guard of the stand-alone operators (see ``Xtractor.check_overflow``)."""
    count = ctypes.c_int64(0)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().skb_ops_overflow_count(_lib.stream_ptr(), ctypes.byref(count)))
    key = torch.device(device).index
    seen = _ovf_seen.get(key, 0)
    if count.value > seen:
        _ovf_seen[key] = count.value
        raise OverflowError("sidekit_b200: activations beyond the fp16 range (65504) in a stand-alone operator; "
                            "use compute_dtype='bf16'")


def conv2d_bn_act(x, w, bias, stride=1, act_slope=0.0, se_scale=None, residual=None, pre=None, pre_slope=0.01,
                  compute_dtype="fp16"):
    """``act(conv2d(x, w) + bias)`` (+ the fused BasicBlock tail ``act(conv * se_scale + residual)``) on the tcgen05
    convolution kernel.  ``w`` (Cout, Cin, k, k) / ``bias`` (Cout): fp32 numpy with the BatchNorm folded in
    (``fold_conv_bn``); k = 3 (padding 1) or 1 (padding 0); ``stride`` 1 or 2; ``pre`` = (scale, shift) device tensors of a
    pre-activation BatchNorm + LeakyReLU(``pre_slope``) applied to ``x`` first (ResBlock)."""
    x = _need_cuda(x)
    B, Cin, H, W = x.shape
    Cout, k = int(w.shape[0]), int(w.shape[2])
    assert w.shape[1] == Cin and w.shape[2] == w.shape[3] and k in (1, 3), "conv2d_bn_act: 3x3 or 1x1 kernels"
    stride = int(stride[0]) if isinstance(stride, (tuple, list)) else int(stride)
    Ho, Wo = ((H - 1) // 2 + 1, (W - 1) // 2 + 1) if stride == 2 else (H, W)
    y = torch.empty((B, Cout, Ho, Wo), dtype=torch.float32, device=x.device)
    w = numpy.ascontiguousarray(w, dtype=numpy.float32)
    bias = numpy.ascontiguousarray(bias, dtype=numpy.float32)
    se = None if se_scale is None else se_scale.contiguous().float()
    res = None if residual is None else residual.contiguous().float()
    if res is not None:
        assert tuple(res.shape) == tuple(y.shape), "residual shape %s != output shape %s" % (tuple(res.shape), tuple(y.shape))
    ps, pt = (None, None) if pre is None else pre
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().skb_conv2d_bn_act(
            x.data_ptr(), B, Cin, H, W, w.ctypes.data, bias.ctypes.data, Cout, k, stride,
            None if ps is None else ps.data_ptr(), None if pt is None else pt.data_ptr(), float(pre_slope), float(act_slope),
            None if se is None else se.data_ptr(), None if res is None else res.data_ptr(), _COMPUTE[compute_dtype],
            y.data_ptr(), _lib.stream_ptr()))
    return y


def channel_mean(x):
    x = _need_cuda(x)
    B, C = x.shape[:2]
    hw = x[0, 0].numel()
    out = torch.empty((B, C), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().skb_channel_mean(x.data_ptr(), B, C, hw, out.data_ptr(), _lib.stream_ptr()))
    return out


def se_gate(mean, fc1, fc2):
    """``sigmoid(fc2 @ relu(fc1 @ mean))`` per row of ``mean`` (B, C); ``fc1`` (R, C), ``fc2`` (C, R)."""
    mean = _need_cuda(mean)
    B, C = mean.shape
    fc1, fc2 = fc1.detach().to(mean.device).contiguous().float(), fc2.detach().to(mean.device).contiguous().float()
    out = torch.empty_like(mean)
    with torch.cuda.device(mean.device):
        _lib.check(_lib.lib().skb_se_gate(mean.data_ptr(), fc1.data_ptr(), fc2.data_ptr(), B, C, int(fc1.shape[0]), out.data_ptr(),
                                          _lib.stream_ptr()))
    return out


def scale_residual_act(y, scale=None, residual=None, slope=1.0):
    """``act(y * scale[b][c] + residual)``, ``act(v) = max(v, slope * v)``."""
    y = _need_cuda(y)
    B, C = y.shape[:2]
    hw = y[0, 0].numel()
    out = torch.empty_like(y)
    sc = None if scale is None else scale.contiguous().float()
    res = None if residual is None else residual.contiguous().float()
    with torch.cuda.device(y.device):
        _lib.check(_lib.lib().skb_scale_residual_act(y.data_ptr(), None if sc is None else sc.data_ptr(),
                                                     None if res is None else res.data_ptr(), B, C, hw, float(slope), out.data_ptr(),
                                                     _lib.stream_ptr()))
    return out


def l2_normalize(x, eps=1e-12):
    x = _need_cuda(x)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().skb_l2_normalize(x.data_ptr(), x.shape[0], x.shape[1], float(eps), out.data_ptr(), _lib.stream_ptr()))
    return out


def attentive_pool(x, w1, b1, bn, w2, b2, global_context):
    """``AttentivePooling.forward`` on ``x`` (B, D, T)."""
    x = _need_cuda(x)
    B, D, T = x.shape
    dev = x.device
    f = lambda t: t.detach().to(dev).contiguous().float()
    w1, b1, w2, b2 = f(w1).reshape(w1.shape[0], -1), f(b1), f(w2).reshape(w2.shape[0], -1), f(b2)
    bn_s, bn_t = bn_affine(bn, dev)
    out = torch.empty((B, 2 * D), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().skb_attentive_pool(x.data_ptr(), B, D, T, w1.data_ptr(), b1.data_ptr(), bn_s.data_ptr(), bn_t.data_ptr(),
                                                 w2.data_ptr(), b2.data_ptr(), int(w1.shape[0]), int(bool(global_context)),
                                                 out.data_ptr(), _lib.stream_ptr()))
    return out
