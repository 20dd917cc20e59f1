"""CPU restatement of ``Xtractor.forward`` (torch CPU ops, fp32 or fp64).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Every function cites the
reference lines it restates (paths relative to /root/reference/sidekit).  The
functions are purely functional over a ``state_dict`` that uses the reference's
key names, so the same weights drive the reference,
this oracle and the CUDA path.

Pinned by tests/test_oracle_vs_reference.py (against the imported reference,
build container only) and tests/golden/*.npz (generated from the reference by
oracle/make_golden.py).
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- front-end constants
def hann_periodic(win_length, dtype=torch.float64):
    """torch.hann_window(periodic=True): 0.5 - 0.5 cos(2 pi n / N)."""
    n = torch.arange(win_length, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * n / win_length)).to(dtype)


def mel_filterbank(n_freqs, f_min, f_max, n_mels, sample_rate, dtype=torch.float32):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') restated.

    Called from torchaudio MelScale, which the reference instantiates at
    nnet/preprocessor.py:253-261 (log-Mel) and :104-109 (MFCC).
    """
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.clamp(torch.min(down, up), min=0.0)
    return fb.to(dtype)                                      # (n_freqs, n_mels)


def dct_matrix(n_mfcc, n_mels, dtype=torch.float32):
    """torchaudio.functional.create_dct(norm='ortho') -> (n_mels, n_mfcc)."""
    n = torch.arange(float(n_mels))
    k = torch.arange(float(n_mfcc)).unsqueeze(1)
    dct = torch.cos(math.pi / float(n_mels) * (n + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / float(n_mels))
    return dct.t().to(dtype)


# ----------------------------------------------------------------------------- front-end
def pre_emphasis(x, coef=0.97):
    """nnet/augmentation.py:63-74: y[t] = x[t] - coef*x[t-1], x[-1] := x[1] (reflect pad)."""
    prev = torch.cat([x[:, 1:2], x[:, :-1]], dim=1)
    return x - coef * prev


def power_spectrogram(y, n_fft, win_length, hop, window=None):
    """torch.stft(center=True, pad_mode='reflect', onesided) -> |X|^2, (B, n_fft/2+1, T).

    The reference reaches it through torchaudio Spectrogram(power=2)
    (nnet/preprocessor.py:253-261, :279).  Window = periodic Hann of
    ``win_length`` zero-padded (centred) to ``n_fft``.
    """
    B, L = y.shape
    pad = n_fft // 2
    yp = F.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    T = 1 + L // hop
    if window is None:
        window = hann_periodic(win_length, y.dtype)
    w = torch.zeros(n_fft, dtype=y.dtype, device=y.device)
    left = (n_fft - win_length) // 2
    w[left:left + win_length] = window.to(device=y.device, dtype=y.dtype)
    frames = yp.unfold(1, n_fft, hop)[:, :T, :]              # (B, T, n_fft)
    spec = torch.fft.rfft(frames * w, dim=2)                 # (B, T, n_fft/2+1)
    return (spec.real ** 2 + spec.imag ** 2).transpose(1, 2)


def instance_norm(z, eps=1e-5):
    """torch.nn.InstanceNorm1d (no affine, biased variance) -- preprocessor.py:263, :281."""
    mean = z.mean(dim=2, keepdim=True)
    var = ((z - mean) ** 2).mean(dim=2, keepdim=True)
    return (z - mean) / torch.sqrt(var + eps)


def logmel_frontend(sd, x):
    """MelSpecFrontEnd.forward(is_eval=True), nnet/preprocessor.py:267-285."""
    if x.dim() == 1:
        x = x.unsqueeze(0)
    fb = sd.get("preprocessor.MelSpec.mel_scale.fb")
    if fb is None:
        fb = mel_filterbank(513, 90.0, 7600.0, 80, 16000)
    y = pre_emphasis(x)
    p = power_spectrogram(y, 1024, 400, 160, sd.get("preprocessor.MelSpec.spectrogram.window"))
    mel = torch.matmul(p.transpose(1, 2), fb.to(x.dtype)).transpose(1, 2)      # (B, 80, T)
    return instance_norm(torch.log(mel + 1e-6))


def mfcc_frontend(sd, x):
    """MfccFrontEnd.forward, nnet/preprocessor.py:113-124 (+ torchaudio MFCC, log_mels=True)."""
    if x.dim() == 1:
        x = x.unsqueeze(0)
    fb = sd.get("preprocessor.MFCC.MelSpectrogram.mel_scale.fb")
    if fb is None:
        fb = mel_filterbank(1025, 133.333, 6855.4976, 100, 16000)
    dct = sd.get("preprocessor.MFCC.dct_mat")
    if dct is None:
        dct = dct_matrix(80, 100)
    y = pre_emphasis(x)
    p = power_spectrogram(y, 2048, 1024, 512, sd.get("preprocessor.MFCC.MelSpectrogram.spectrogram.window"))
    mel = torch.matmul(p.transpose(1, 2), fb.to(x.dtype)).transpose(1, 2)      # (B, 100, T)
    logmel = torch.log(mel + 1e-6)
    mfcc = torch.matmul(logmel.transpose(1, 2), dct.to(x.dtype)).transpose(1, 2)
    return instance_norm(mfcc)


# ----------------------------------------------------------------------------- HalfResNet34 trunk
def _bn(sd, prefix, x, eps=1e-5):
    shape = [1, -1] + [1] * (x.dim() - 2)
    w = sd[prefix + ".weight"].to(x.dtype).view(shape)
    b = sd[prefix + ".bias"].to(x.dtype).view(shape)
    m = sd[prefix + ".running_mean"].to(x.dtype).view(shape)
    v = sd[prefix + ".running_var"].to(x.dtype).view(shape)
    return (x - m) / torch.sqrt(v + eps) * w + b


def se_layer(sd, prefix, x):
    """SELayer.forward, nnet/res_net.py:272-281."""
    y = x.mean(dim=(2, 3))
    y = F.relu(F.linear(y, sd[prefix + ".fc.0.weight"].to(x.dtype)))
    y = torch.sigmoid(F.linear(y, sd[prefix + ".fc.2.weight"].to(x.dtype)))
    return x * y[:, :, None, None]


def basic_block(sd, prefix, x, stride):
    """BasicBlock.forward, nnet/res_net.py:309-320; shortcut rule :301-307 (the shortcut conv uses the block's stride)."""
    out = F.conv2d(x, sd[prefix + ".conv1.weight"].to(x.dtype), stride=stride, padding=1)
    out = F.relu(_bn(sd, prefix + ".bn1", out))
    out = F.conv2d(out, sd[prefix + ".conv2.weight"].to(x.dtype), stride=1, padding=1)
    out = _bn(sd, prefix + ".bn2", out)
    out = se_layer(sd, prefix + ".se", out)
    if prefix + ".shortcut.0.weight" in sd:
        sc = F.conv2d(x, sd[prefix + ".shortcut.0.weight"].to(x.dtype), stride=stride)
        sc = _bn(sd, prefix + ".shortcut.1", sc)
    else:
        sc = x
    return F.relu(out + sc)


def res_block(sd, prefix, x, is_first=False, slope=0.01):
    """ResBlock.forward, nnet/res_net.py:229-255 (eval mode, stride 1): pre-activation BN + LeakyReLU unless ``is_first``;
    conv1 (bias) -> batch_norm2 -> LeakyReLU -> conv2 (bias) -> the SAME batch_norm2 -> + identity (1x1 conv with bias +
    BN when the width changes, :201-208) -> LeakyReLU."""
    identity = x
    out = x if is_first else F.leaky_relu(_bn(sd, prefix + ".batch_norm1", x), slope)
    out = F.conv2d(out, sd[prefix + ".conv1.weight"], sd[prefix + ".conv1.bias"], padding=1)
    out = F.leaky_relu(_bn(sd, prefix + ".batch_norm2", out), slope)
    out = F.conv2d(out, sd[prefix + ".conv2.weight"], sd[prefix + ".conv2.bias"], padding=1)
    out = _bn(sd, prefix + ".batch_norm2", out)
    if prefix + ".resample.0.weight" in sd:
        identity = _bn(sd, prefix + ".resample.1", F.conv2d(x, sd[prefix + ".resample.0.weight"], sd[prefix + ".resample.0.bias"]))
    return F.leaky_relu(out + identity, slope)


HALFRESNET34_STAGES = ((3, 1), (4, 2), (6, 2), (3, 2))       # (num_blocks, first stride), res_net.py:520-523
# PreResNet34, res_net.py:455-462: layer7 is built with num_blocks[5], i.e. ONE block
RESNET34_STAGES = ((3, 1), (1, 2), (3, 1), (1, 2), (5, 1), (1, 2), (1, 1))
FASTRESNET34_STAGES = ((3, 1), (4, 2), (6, 2), (3, 1))       # PreFastResNet34, res_net.py:575-578


def halfresnet34_trunk(sd, feats, collect=None, stages=HALFRESNET34_STAGES, stem_stride=1, stem_pad=1):
    """PreHalfResNet34.forward, nnet/res_net.py:539-554 (PreResNet34.forward, :476-498, with ``RESNET34_STAGES``).
    feats (B, 80, T) -> (B, 256, T4, 10).  Whether a block has a 1x1 shortcut is read off the state_dict."""
    x = feats.unsqueeze(1).permute(0, 1, 3, 2)                # (B, 1, T, F)
    p = "sequence_network"
    x = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"].to(x.dtype), stride=stem_stride, padding=stem_pad)))
    if collect is not None:
        collect["stem"] = x
    for li, (nb, stride) in enumerate(stages, start=1):
        for bi in range(nb):
            x = basic_block(sd, "%s.layer%d.%d" % (p, li, bi), x, stride if bi == 0 else 1)
            if collect is not None:
                collect["layer%d.%d" % (li, bi)] = x
    return x


# ----------------------------------------------------------------------------- pooling + heads
def mean_std_pooling(x):
    """MeanStdPooling.forward, nnet/pooling.py:55-70 (std is the unbiased estimator)."""
    if x.dim() == 4:
        x = x.permute(0, 1, 3, 2).flatten(1, 2)
    return torch.cat([x.mean(dim=2), x.std(dim=2)], dim=1)


def attentive_pooling(sd, x, prefix="stat_pooling", global_context=True):
    """AttentivePooling(C, 10, global_context=...).forward, nnet/pooling.py:151-171."""
    if x.dim() == 4:
        x = x.permute(0, 1, 3, 2).flatten(1, 2)               # (B, C*F, T), channel = c*F + f
    h = x
    if global_context:
        gc = mean_std_pooling(x).unsqueeze(2).repeat(1, 1, x.shape[-1])
        h = torch.cat([x, gc], dim=1)
    h = F.conv1d(h, sd[prefix + ".attention.0.weight"].to(x.dtype), sd[prefix + ".attention.0.bias"].to(x.dtype))
    h = torch.tanh(_bn(sd, prefix + ".attention.2", F.relu(h)))
    h = F.conv1d(h, sd[prefix + ".attention.4.weight"].to(x.dtype), sd[prefix + ".attention.4.bias"].to(x.dtype))
    w = torch.softmax(h, dim=2)
    mu = torch.sum(x * w, dim=2)
    rh = torch.sqrt((torch.sum((x ** 2) * w, dim=2) - mu ** 2).clamp(min=1e-9))
    return torch.cat((mu, rh), 1)


def l2_norm(x):
    """nnet/loss.py:91-100 (no epsilon)."""
    return x / torch.norm(x, 2, 1, True)


def arc_margin_eval(weight, x, s):
    """ArcMarginProduct.forward(target=None), nnet/loss.py:299-310."""
    return F.linear(F.normalize(x), F.normalize(weight.to(x.dtype))) * s


def halfresnet34_forward(sd, wave, norm_embedding=True, s=30.0, collect=None):
    """Xtractor('halfresnet34', loss='aam').forward(x, is_eval=True), nnet/xvector.py:876-907."""
    feats = logmel_frontend(sd, wave)
    if collect is not None:
        collect["feats"] = feats
    x = halfresnet34_trunk(sd, feats, collect)
    x = attentive_pooling(sd, x)
    if collect is not None:
        collect["pooled"] = x
    x = F.linear(x, sd["before_speaker_embedding.lin_be.weight"].to(x.dtype))
    x = _bn(sd, "before_speaker_embedding.bn_be", x)
    if norm_embedding:
        x = l2_norm(x)
    logits = arc_margin_eval(sd["after_speaker_embedding.weight"], x, s)
    return logits, F.normalize(x, dim=1)


def resnet34_forward(sd, wave, norm_embedding=True, s=30.0, collect=None):
    """Xtractor('resnet34').forward(x, is_eval=True), nnet/xvector.py:516-540 + :876-907 (pooling patched to
    AttentivePooling(256, 10, global_context=True), like halfresnet34)."""
    feats = logmel_frontend(sd, wave)
    if collect is not None:
        collect["feats"] = feats
    x = halfresnet34_trunk(sd, feats, collect, RESNET34_STAGES)
    x = attentive_pooling(sd, x)
    if collect is not None:
        collect["pooled"] = x
    x = F.linear(x, sd["before_speaker_embedding.weight"].to(x.dtype), sd["before_speaker_embedding.bias"].to(x.dtype))
    if norm_embedding:
        x = l2_norm(x)
    logits = arc_margin_eval(sd["after_speaker_embedding.weight"], x, s)
    return logits, F.normalize(x, dim=1)


def fastresnet34_forward(sd, wave, norm_embedding=True, s=30.0, collect=None):
    """Xtractor('fastresnet34', loss='aam').forward(x, is_eval=True), nnet/xvector.py:539-567 + :876-907 (pooling patched
    to AttentivePooling(128, 10, global_context=False)).  Stem: 7x7, stride (1, 2), padding 3 (res_net.py:566-571)."""
    feats = logmel_frontend(sd, wave)
    if collect is not None:
        collect["feats"] = feats
    x = halfresnet34_trunk(sd, feats, collect, FASTRESNET34_STAGES, stem_stride=(1, 2), stem_pad=3)
    x = attentive_pooling(sd, x, global_context=False)
    if collect is not None:
        collect["pooled"] = x
    x = F.linear(x, sd["before_speaker_embedding.weight"].to(x.dtype), sd["before_speaker_embedding.bias"].to(x.dtype))
    if norm_embedding:
        x = l2_norm(x)
    logits = arc_margin_eval(sd["after_speaker_embedding.weight"], x, s)
    return logits, F.normalize(x, dim=1)


# ----------------------------------------------------------------------------- TDNN ("xvector")
TDNN_LAYERS = ((5, 1), (3, 2), (3, 3), (1, 1), (1, 1))       # (kernel, dilation), xvector.py:467-483


def tdnn_trunk(sd, feats):
    """sequence_network of model_archi='xvector': conv -> LeakyReLU(0.2) -> BN, x5 (xvector.py:467-483)."""
    x = feats
    for i, (_, dil) in enumerate(TDNN_LAYERS, start=1):
        p = "sequence_network"
        x = F.conv1d(x, sd["%s.conv%d.weight" % (p, i)].to(x.dtype), sd["%s.conv%d.bias" % (p, i)].to(x.dtype),
                     dilation=dil)
        x = _bn(sd, "%s.batch_norm%d" % (p, i), F.leaky_relu(x, 0.2))
    return x


def tdnn_forward(sd, wave, norm_embedding=True, s=64.0):
    """Xtractor('xvector', loss='aam').forward(x, is_eval=True) (xvector.py:453-498, :876-907)."""
    feats = mfcc_frontend(sd, wave)
    x = tdnn_trunk(sd, feats)
    x = mean_std_pooling(x)
    x = F.linear(x, sd["before_speaker_embedding.linear6.weight"].to(x.dtype),
                 sd["before_speaker_embedding.linear6.bias"].to(x.dtype))
    if norm_embedding:
        x = l2_norm(x)
    logits = arc_margin_eval(sd["after_speaker_embedding.weight"], x, s)
    return logits, F.normalize(x, dim=1)


def forward(sd, wave, model_archi, **kw):
    with torch.no_grad():
        if model_archi == "halfresnet34":
            return halfresnet34_forward(sd, wave, **kw)
        if model_archi == "xvector":
            return tdnn_forward(sd, wave, **kw)
        if model_archi == "resnet34":
            return resnet34_forward(sd, wave, **kw)
        if model_archi == "fastresnet34":
            return fastresnet34_forward(sd, wave, **kw)
    raise NotImplementedError(model_archi)


def state_dict_to(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
