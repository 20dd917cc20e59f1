"""Deterministic synthetic weights and audio (no checkpoints / corpora offline).

The reference publishes its numbers on a checkpoint that cannot be fetched, so
every parity test and benchmark runs on random-init weights of the reference
architecture.  Default-initialised eval-mode BatchNorm is the
identity, which would hide BN-folding bugs, so ``fill_state_dict`` randomises
every BN statistic ("variant B" of SURVEY.md 8d).  The fill depends only on the
key name, the tensor shape and the seed, so the reference model, the oracle and
the CUDA path all see bit-identical parameters without shipping a checkpoint.
"""
import hashlib
import math

import numpy
import torch


def _gen(key, seed):
    h = hashlib.sha256(("%s|%d" % (key, seed)).encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)
    return g


def fill_state_dict(state_dict, seed=0, init="default"):
    """Overwrite every learnable tensor / BN statistic of ``state_dict`` in place.

    Front-end buffers (window, mel filterbank, DCT, pre-emphasis filter) are
    left untouched: they are constants of the architecture.  ``init='default'``
    draws conv / linear weights from torch's default-initialisation distribution
    (what a freshly constructed reference model has); ``init='he'`` is a stress
    variant (He-normal, 6x the variance) whose residual branches dominate, so a
    wrong convolution or a precision loss shows up much more strongly.
    """
    for key in sorted(state_dict.keys()):
        t = state_dict[key]
        if key.startswith("preprocessor.") or key.endswith("num_batches_tracked"):
            continue
        g = _gen(key, seed)
        shape = tuple(t.shape)
        if key.endswith("running_mean"):
            v = torch.randn(shape, generator=g) * 0.1
        elif key.endswith("running_var"):
            v = torch.rand(shape, generator=g) + 0.5
        elif key.endswith(".weight") and t.dim() == 1 and key[:-6] + "running_var" in state_dict:   # BN scale
            v = torch.rand(shape, generator=g) * 0.4 + 0.8
        elif t.dim() == 1:                       # biases (conv / linear / BN)
            v = torch.randn(shape, generator=g) * 0.1
        else:
            # conv / linear weights: U(-1/sqrt(fan_in), 1/sqrt(fan_in)), the distribution torch's default
            # reset_parameters() (kaiming_uniform_, a=sqrt(5)) gives the reference's freshly built modules
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if init == "he":
                v = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
            else:
                bound = 1.0 / math.sqrt(fan_in)
                v = (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
        t.copy_(v.to(t.dtype))
    return state_dict


def synth_wave(n_utt, n_samples, seed=3, scale=0.1):
    """(n_utt, n_samples) fp32 Gaussian audio, the C1 input of SURVEY.md 8d."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(n_utt, n_samples, generator=g) * scale


def synth_lengths(n_utt, lo_s=2.0, hi_s=20.0, seed=4, sample_rate=16000):
    """Utterance lengths (samples) ~ U[lo_s, hi_s] seconds (C2 / C4 of SURVEY.md 8d)."""
    rng = numpy.random.default_rng(seed)
    return numpy.round(sample_rate * rng.uniform(lo_s, hi_s, size=n_utt)).astype(numpy.int64)


def synth_embeddings(n, dim=256, seed=6, unit_norm=True):
    rng = numpy.random.default_rng(seed)
    x = rng.standard_normal((n, dim))
    if unit_norm:
        x /= numpy.linalg.norm(x, axis=1, keepdims=True)
    return x


def synth_plda(dim=256, rank=256, seed=7):
    """(mu, F, Sigma) as in SURVEY.md 8d: mu=0.01 N, F=0.3 N (dim x rank), Sigma=AA'+I, A=0.1 N."""
    rng = numpy.random.default_rng(seed)
    mu = 0.01 * rng.standard_normal(dim)
    F = 0.3 * rng.standard_normal((dim, rank))
    A = 0.1 * rng.standard_normal((dim, dim))
    Sigma = A @ A.T + numpy.eye(dim)
    return mu, F, Sigma
