"""Front-end modules with the reference's names, constructor arguments and state_dict keys
(sidekit/nnet/preprocessor.py:61-124, :212-285; sidekit/nnet/augmentation.py:49-74).

The modules only hold the constant buffers (pre-emphasis filter, Hann window, Mel filterbank, DCT
matrix -- computed with the same float32 formulas as torchaudio so they are bit-identical to the
reference's buffers); the arithmetic runs in the fused CUDA front-end kernel (csrc/frontend.cu).
"""
import math

import torch


def melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate):
    """HTK triangular filterbank, norm=None (torchaudio.functional.melscale_fbanks) -> (n_freqs, n_mels)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down_slopes = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up_slopes = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down_slopes, up_slopes))


def create_dct(n_mfcc, n_mels):
    """Orthonormal DCT-II matrix (torchaudio.functional.create_dct(norm='ortho')) -> (n_mels, n_mfcc)."""
    n = torch.arange(float(n_mels))
    k = torch.arange(float(n_mfcc)).unsqueeze(1)
    dct = torch.cos(math.pi / float(n_mels) * (n + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / float(n_mels))
    return dct.t().contiguous()


class PreEmphasis(torch.nn.Module):
    """y[t] = x[t] - coef * x[t-1] with reflect padding (augmentation.py:49-74); buffer only."""

    def __init__(self, coef: float = 0.97):
        super().__init__()
        self.coef = coef
        self.register_buffer("flipped_filter", torch.FloatTensor([-self.coef, 1.]).unsqueeze(0).unsqueeze(0))


class _Spectrogram(torch.nn.Module):
    def __init__(self, win_length):
        super().__init__()
        self.register_buffer("window", torch.hann_window(win_length))


class _MelScale(torch.nn.Module):
    def __init__(self, n_freqs, f_min, f_max, n_mels, sample_rate):
        super().__init__()
        self.register_buffer("fb", melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate))


class _MelSpectrogram(torch.nn.Module):
    """Container mirroring torchaudio.transforms.MelSpectrogram's buffers (.spectrogram.window, .mel_scale.fb)."""

    def __init__(self, sample_rate, n_fft, win_length, f_min, f_max, n_mels):
        super().__init__()
        self.spectrogram = _Spectrogram(win_length)
        self.mel_scale = _MelScale(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate)


class _MFCC(torch.nn.Module):
    def __init__(self, sample_rate, n_mfcc, melkwargs):
        super().__init__()
        self.register_buffer("dct_mat", create_dct(n_mfcc, melkwargs["n_mels"]))
        self.MelSpectrogram = _MelSpectrogram(sample_rate, melkwargs["n_fft"], melkwargs["win_length"],
                                              melkwargs["f_min"], melkwargs["f_max"], melkwargs["n_mels"])


class _FrontEndBase(torch.nn.Module):
    _owner = None     # set by Xtractor: the native front-end lives in the extractor's engine

    def _run(self, x):
        if x.dim() == 1:
            x = x.unsqueeze(0)
        owner = self.__dict__.get("_owner")
        if owner is None:
            raise RuntimeError("the front-end runs inside the fused CUDA engine: call it through Xtractor.preprocessor")
        return owner()._frontend(x)


class MelSpecFrontEnd(_FrontEndBase):
    """Log-Mel front-end of the ResNet archis (preprocessor.py:212-285); inference path only."""

    def __init__(self, pre_emphasis=0.97, sample_rate=16000, n_fft=1024, f_min=90, f_max=7600, win_length=400,
                 window_fn=torch.hann_window, hop_length=160, power=2.0, n_mels=80):
        super().__init__()
        if (n_fft, win_length, hop_length, power) != (1024, 400, 160, 2.0) or window_fn is not torch.hann_window:
            raise NotImplementedError("the CUDA front-end implements n_fft=1024, win=400, hop=160, power=2, Hann")
        self.pre_emphasis, self.sample_rate, self.n_fft, self.f_min, self.f_max = pre_emphasis, sample_rate, n_fft, f_min, f_max
        self.win_length, self.hop_length, self.power, self.n_mels = win_length, hop_length, power, n_mels
        self.PreEmphasis = PreEmphasis(self.pre_emphasis)
        self.MelSpec = _MelSpectrogram(sample_rate, n_fft, win_length, f_min, f_max, n_mels)

    def forward(self, x, is_eval=False):
        if not is_eval:
            raise NotImplementedError("training-time SpecAugment masking is out of scope (inference hot path only)")
        return self._run(x)


class MfccFrontEnd(_FrontEndBase):
    """MFCC front-end of the TDNN 'xvector' archi (preprocessor.py:61-124).

    ``forward`` accepts (and ignores) ``is_eval``: the reference's Xtractor passes it (xvector.py:885)
    although the reference's own MfccFrontEnd.forward lacks the parameter (SURVEY.md finding 4).
    """

    def __init__(self, pre_emphasis=0.97, sample_rate=16000, n_fft=2048, f_min=133.333, f_max=6855.4976,
                 win_length=1024, window_fn=torch.hann_window, hop_length=512, power=2.0, n_mels=100, n_mfcc=80):
        super().__init__()
        if (n_fft, win_length, hop_length, power) != (2048, 1024, 512, 2.0) or window_fn is not torch.hann_window:
            raise NotImplementedError("the CUDA front-end implements n_fft=2048, win=1024, hop=512, power=2, Hann")
        self.pre_emphasis, self.sample_rate, self.n_fft, self.f_min, self.f_max = pre_emphasis, sample_rate, n_fft, f_min, f_max
        self.win_length, self.hop_length, self.power, self.n_mels, self.n_mfcc = win_length, hop_length, power, n_mels, n_mfcc
        self.PreEmphasis = PreEmphasis(self.pre_emphasis)
        self.melkwargs = {"n_fft": n_fft, "f_min": f_min, "f_max": f_max, "win_length": win_length,
                          "hop_length": hop_length, "power": power, "n_mels": n_mels}
        self.MFCC = _MFCC(sample_rate, n_mfcc, self.melkwargs)

    def forward(self, x, is_eval=True):
        return self._run(x)
