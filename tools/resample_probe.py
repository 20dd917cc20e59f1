"""Resample kernel bandwidth: 44.1 kHz -> 16 kHz (and 48 k, 8 k) on a packed ragged batch; CUDA events on the launching stream.

    python tools/resample_probe.py
"""
import os
import sys

import numpy
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sidekit_b200.nnet.preprocessor import Resample  # noqa: E402


def main():
    torch.cuda.set_device(0)
    rng = numpy.random.default_rng(0)
    for fo, fn in ((44100, 16000), (48000, 16000), (8000, 16000), (22050, 16000)):
        r = Resample(fo, fn)
        lengths = numpy.round(fo * rng.uniform(2.0, 20.0, size=256)).astype(numpy.int64).tolist()
        x = torch.randn(int(sum(lengths)), device="cuda") * 0.1
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            y = r.resample_packed(x, lengths)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = r.resample_packed(x, lengths)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(numpy.median(ts))
        gb = (x.numel() + y.numel()) * 4 / 1e9
        print("%d -> %d: %d waves, %.1f audio-s, %.3f ms per call (incl. the table upload), %.0f GB/s algorithmic, %.2e audio-s/s"
              % (fo, fn, len(lengths), sum(lengths) / fo, ms, gb / (ms / 1e3), sum(lengths) / fo / (ms / 1e3)))


if __name__ == "__main__":
    main()
