"""Golden fixture for the feed path: outputs of the REFERENCE's ``IdMapSet.__getitem__``
(sidekit/nnet/xsets.py:419-464) recorded in the build container -> tests/golden/feed_path.npz.

TEST INFRASTRUCTURE ONLY.  The reference decodes audio with ``torchaudio.load`` / ``torchaudio.info``, which need
torchcodec (absent here).  Only that decoder is replaced, by a stdlib ``wave`` reader with torchaudio's conventions
((channels, frames) float32 scaled by 1 / 32768, ``frame_offset`` / ``num_frames``); the segment arithmetic, the
too-short-segment recentring, the sliding-window unfold and the resampling call all run unmodified.  The PCM of the test
files is stored in the fixture, so the GPU box regenerates the same wav files.

    python oracle/make_golden_feed.py
"""
import os
import sys
import tempfile
import types
import wave

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from sidekit_b200 import synth  # noqa: E402

CASES = [  # (name, file, start_cs, stop_cs, sliding_window)
    ("whole", "f16k_a", None, None, False),
    ("whole_start_only", "f16k_a", 50, None, False),
    ("segment", "f16k_a", 30, 150, False),
    ("too_short_recentred", "f16k_a", 100, 105, False),
    ("too_short_at_file_start", "f16k_b", 0, 4, False),
    ("whole_8k_resampled", "f8k", None, None, False),
    ("sliding", "f16k_b", None, None, True),
    ("sliding_segment", "f16k_b", 20, 330, True),
]
FILES = {"f16k_a": (16000, 28000, 1), "f16k_b": (16000, 56000, 2), "f8k": (8000, 9000, 3)}


def pcm_of(name):
    rate, n, seed = FILES[name]
    return (synth.synth_wave(1, n, seed=2600 + seed)[0].numpy() * 32768.0).clip(-32768, 32767).astype(numpy.int16), rate


def write_wavs(d):
    for name in FILES:
        x, rate = pcm_of(name)
        with wave.open(os.path.join(d, name + ".wav"), "wb") as f:
            f.setnchannels(1); f.setsampwidth(2); f.setframerate(rate)
            f.writeframes(x.tobytes())


def _wave_load(path, frame_offset=0, num_frames=-1, **kw):
    with wave.open(path, "rb") as f:
        rate, total = f.getframerate(), f.getnframes()
        frame_offset = min(max(0, int(frame_offset)), total)
        f.setpos(frame_offset)
        n = total - frame_offset if num_frames is None or num_frames < 0 else min(int(num_frames), total - frame_offset)
        a = numpy.frombuffer(f.readframes(n), dtype=numpy.int16)
    return torch.from_numpy(a.astype(numpy.float32) / numpy.float32(32768.0)).unsqueeze(0), rate


def _wave_info(path, **kw):
    with wave.open(path, "rb") as f:
        return types.SimpleNamespace(sample_rate=f.getframerate(), num_frames=f.getnframes(), num_channels=f.getnchannels())


def main():
    sidekit = ref_import.import_reference()
    from sidekit.nnet import xsets
    from sidekit.bosaris.idmap import IdMap
    xsets.torchaudio.load = _wave_load          # the decoder only (see the module docstring)
    xsets.torchaudio.info = _wave_info
    out = {}
    with tempfile.TemporaryDirectory() as d:
        write_wavs(d)
        for name, fname, start, stop, sliding in CASES:
            im = IdMap()
            im.leftids = numpy.array(["spk"], dtype="|O")
            im.rightids = numpy.array([fname], dtype="|O")
            im.start = numpy.array([start], dtype="|O")
            im.stop = numpy.array([stop], dtype="|O")
            ds = xsets.IdMapSet(im, d, "wav", transform_pipeline={}, sliding_window=sliding, window_len=1.0, window_shift=0.5,
                                sample_rate=16000, min_duration=0.165)
            speech, left, right, s0, s1 = ds[0]
            out[name + "_speech"] = speech.numpy()
            out[name + "_bounds"] = numpy.array([s0, s1], dtype=numpy.int64)
            assert left == "spk" and right == fname
    for fname in FILES:
        out["pcm_" + fname] = pcm_of(fname)[0]
    path = os.path.join(ROOT, "tests", "golden", "feed_path.npz")
    numpy.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("_speech")})


if __name__ == "__main__":
    main()
