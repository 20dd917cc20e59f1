"""Sample-rate conversion of the feed path: ``torchaudio.transforms.Resample`` as
sidekit/nnet/xsets.py:435 / sidekit/bin/extract_xvectors.py:144 apply it.  CPU tests pin the oracle restatement and
the compact filter bank against outputs recorded from torchaudio (tests/golden/resample.npz); the GPU tests run the
CUDA kernel through the C ABI against both."""
import os
import wave

import numpy
import pytest
import torch

import sidekit_b200 as sk
from oracle import resample_ref as RR
from sidekit_b200 import synth
from sidekit_b200.nnet.preprocessor import Resample, sinc_resample_bank
from tests.helpers import golden

TOL = 2e-6          # float32 FIR of <= 475 taps on 0.1-scale audio; torchaudio itself accumulates in float32


def _cases():
    g = golden("resample.npz")
    for i in range(int(g["n_cases"])):
        fo, fn, n = (int(v) for v in g["case%d_rates" % i])
        yield i, fo, fn, n, synth.synth_wave(2, n, seed=300 + i), g["case%d_y" % i]


def test_oracle_matches_torchaudio_golden():
    for i, fo, fn, n, x, y in _cases():
        r = RR.resample(x.numpy(), fo, fn)
        assert r.shape == y.shape, (i, r.shape, y.shape)
        assert numpy.abs(r - y).max() < TOL, (i, numpy.abs(r - y).max())


def test_compact_bank_is_the_full_bank_without_clamped_taps():
    for fo, fn in ((44100, 16000), (48000, 16000), (8000, 16000), (22050, 16000), (11025, 16000), (16000, 8000)):
        orig_r, new_r, width, bank, start = sinc_resample_bank(fo, fn)
        o2, n2, w2, full = RR.sinc_bank(fo, fn)
        assert (orig_r, new_r, width) == (o2, n2, w2)
        assert bank.shape[1] == new_r and start.shape == (new_r,)
        rebuilt = numpy.zeros_like(full)
        for ph in range(new_r):
            rebuilt[ph, start[ph]:start[ph] + bank.shape[0]] = bank[:, ph]
        assert numpy.abs(rebuilt - full).max() < 1e-30          # the dropped taps carry cos(pi/2)^2 ~ 4e-33
        assert bank.shape[0] <= 2 * width + 2


def test_out_length_and_identity():
    r = Resample(44100, 16000)
    for n in (0, 1, 5, 440, 441, 442, 44100):
        assert r.out_length(n) == int(numpy.ceil(160 * n / 441))
    x = torch.zeros(3, 7)
    assert Resample(16000, 16000)(x) is x
    with pytest.raises(Exception):
        Resample(16000.5, 8000)


@pytest.mark.gpu
def test_kernel_matches_golden_and_oracle():
    for i, fo, fn, n, x, y in _cases():
        out = Resample(fo, fn)(x.cuda()).cpu().numpy()
        assert out.shape == y.shape
        assert numpy.abs(out - y).max() < TOL, (i, numpy.abs(out - y).max())
        assert numpy.abs(out - RR.resample(x.numpy(), fo, fn)).max() < TOL


@pytest.mark.gpu
def test_kernel_ragged_packed_batch_and_long_waves():
    # utterances of very different lengths back to back (several CTAs per wave, waves shorter than one filter)
    lengths = [3, 44100 * 7 + 13, 0, 1000, 441, 200001, 1]
    waves = [synth.synth_wave(1, L, seed=900 + k)[0] for k, L in enumerate(lengths)]
    r = Resample(44100, 16000)
    out = r.resample_packed(torch.cat(waves).cuda(), lengths).cpu().numpy()
    off = 0
    for w, L in zip(waves, lengths):
        ref = RR.resample(w.numpy(), 44100, 16000)
        got = out[off:off + r.out_length(L)]
        assert got.shape == ref.shape
        assert L == 0 or numpy.abs(got - ref).max() < TOL, (L, numpy.abs(got - ref).max())
        off += r.out_length(L)
    assert off == out.shape[0]
    assert r.resample_packed(torch.empty(0, device="cuda"), [0, 0]).numel() == 0          # nothing to do: no launch
    with pytest.raises(RuntimeError):
        r.resample_packed(torch.zeros(10), [10])                                          # no CPU path
    # leading dimensions are kept, like torchaudio
    x = synth.synth_wave(6, 2000, seed=7).reshape(2, 3, 2000)
    y = Resample(8000, 16000)(x.cuda())
    assert y.shape == (2, 3, 4000)
    assert numpy.abs(y.cpu().numpy().reshape(6, -1) - RR.resample(x.reshape(6, -1).numpy(), 8000, 16000)).max() < TOL


@pytest.mark.gpu
def test_idmapset_resamples_files_at_another_rate(tmp_path):
    # a 44.1 kHz file: IdMapSet hands the extractor the 16 kHz signal torchaudio's Resample would produce (xsets.py:433-435)
    x = (synth.synth_wave(1, 44100 * 2, seed=55)[0].numpy() * 32768.0).clip(-32768, 32767).astype(numpy.int16)
    with wave.open(os.path.join(tmp_path, "a.wav"), "wb") as f:
        f.setnchannels(1); f.setsampwidth(2); f.setframerate(44100)
        f.writeframes(x.tobytes())
    im = sk.IdMap()
    im.leftids, im.rightids = numpy.array(["spk"], dtype="|O"), numpy.array(["a"], dtype="|O")
    im.start, im.stop = numpy.array([None], dtype="|O"), numpy.array([None], dtype="|O")
    from sidekit_b200.nnet import xsets
    speech, left, right, start, stop = xsets.IdMapSet(im, str(tmp_path), "wav")[0]
    ref = RR.resample(x.astype(numpy.float32) / 32768.0, 44100, 16000)
    assert speech.is_cuda and speech.shape == ref.shape and (start, stop) == (0, ref.shape[0])
    assert numpy.abs(speech.cpu().numpy() - ref).max() < TOL
