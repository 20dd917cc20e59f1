"""TEST INFRASTRUCTURE ONLY -- numpy restatement of ``torchaudio.transforms.Resample`` (sinc_interp_hann), the call the
reference makes on every file whose rate differs from the model's (sidekit/nnet/xsets.py:435, :452;
sidekit/bin/extract_xvectors.py:144).

The arithmetic lives in a third-party dependency that is not vendored: torchaudio (the reference pins 0.8.2 in
install.sh:34-36; this container has 2.11.0).  This file restates ``torchaudio.functional._get_sinc_resample_kernel``
and ``_apply_sinc_resample_kernel`` of the installed version -- filter bank in float64, cast to float32, zero padding
``(width, width + orig)``, strided FIR, output cut to ``ceil(new * L / orig)`` -- with the FIR accumulated in float64
(the full, un-compacted bank: every tap, including the clamped ones the CUDA kernel drops).  Pinned by
``tests/golden/resample.npz``, recorded from torchaudio itself by ``oracle/make_golden.py resample``.
"""
import math

import numpy


def sinc_bank(orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """-> (orig_r, new_r, width, kernels float32 (new_r, 2 * width + orig_r))."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig_r, new_r = int(orig_freq) // g, int(new_freq) // g
    base = min(orig_r, new_r) * rolloff
    width = math.ceil(lowpass_filter_width * orig_r / base)
    idx = numpy.arange(-width, width + orig_r, dtype=numpy.float64)[None, :] / orig_r
    # torchaudio divides an int64 arange by new_freq: torch's true division yields FLOAT32 phase offsets, which are then
    # promoted to float64 by the addition -- reproduced here, it moves the taps by up to 6e-8 of a sample
    phase = (numpy.arange(0, -new_r, -1).astype(numpy.float32) / numpy.float32(new_r)).astype(numpy.float64)
    t = phase[:, None] + idx
    t = numpy.clip(t * base, -lowpass_filter_width, lowpass_filter_width)
    window = numpy.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with numpy.errstate(invalid="ignore", divide="ignore"):
        kernels = numpy.where(t == 0, 1.0, numpy.sin(t) / t)
    kernels = kernels * window * (base / orig_r)
    return orig_r, new_r, width, kernels.astype(numpy.float32)


def resample(x, orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """x: (L,) or (n, L) float32 -> resampled float32, ``ceil(new * L / orig)`` samples per row."""
    x = numpy.asarray(x, dtype=numpy.float32)
    if orig_freq == new_freq:
        return x
    squeeze = x.ndim == 1
    if x.shape[-1] == 0:
        return x                                   # nothing to resample (torchaudio returns an empty tensor as well)
    x2 = x.reshape(-1, x.shape[-1])
    orig_r, new_r, width, kern = sinc_bank(orig_freq, new_freq, lowpass_filter_width, rolloff)
    L = x2.shape[1]
    xp = numpy.pad(x2.astype(numpy.float64), ((0, 0), (width, width + orig_r)))
    n_q = (xp.shape[1] - kern.shape[1]) // orig_r + 1
    # frames[q] = xp[q * orig : q * orig + taps]; y[q * new + ph] = kern[ph] . frames[q]
    frames = numpy.lib.stride_tricks.sliding_window_view(xp, kern.shape[1], axis=1)[:, ::orig_r][:, :n_q]
    y = numpy.einsum("nqk,pk->nqp", frames, kern.astype(numpy.float64)).reshape(x2.shape[0], -1)
    target = int(math.ceil(new_r * L / orig_r))
    y = y[:, :target].astype(numpy.float32)
    return y[0] if squeeze else y
