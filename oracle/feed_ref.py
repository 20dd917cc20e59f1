"""CPU restatement of the segment / sliding-window logic of the reference's feed path
(sidekit/nnet/xsets.py:419-464 ``IdMapSet.__getitem__`` and sidekit/nnet/xvector.py:1877-1914).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Pinned by tests/golden/feed_path.npz: outputs of the reference's own
``IdMapSet.__getitem__`` recorded by oracle/make_golden_feed.py with only the file decoder replaced (``torchaudio.load``
needs torchcodec, absent here; a stdlib ``wave`` reader with the same conventions stands in).  Works on in-memory int16
arrays.
"""
import numpy


def cut_segment(pcm, start_cs, stop_cs, sample_rate=16000, min_duration=3.0):
    """-> (samples float32, start, stop) of one id-map entry (start / stop in centiseconds or None)."""
    x = pcm.astype(numpy.float32) / numpy.float32(32768.0)
    start = 0 if start_cs is None else int(start_cs * 0.01 * sample_rate)
    if stop_cs is None:
        duration = int(x.shape[0] - start)            # the whole file is returned, only the bookkeeping uses start
        seg = x
    else:
        duration = int(stop_cs * 0.01 * sample_rate) - start
        if duration <= min_duration * sample_rate:
            middle = start + duration // 2
            start = int(max(0, int(middle - (min_duration * sample_rate / 2))))
            duration = int(min_duration * sample_rate)
        seg = x[start:start + duration]
    return seg, start, start + duration


def windows(seg, win_duration, win_shift, sample_rate=16000):
    """torch.Tensor.unfold(0, window_len, window_shift) on a numpy vector."""
    wl, ws = int(win_duration * sample_rate), int(win_shift * sample_rate)
    n = (seg.shape[0] - wl) // ws + 1
    return numpy.stack([seg[i * ws:i * ws + wl] for i in range(n)])


def bookkeeping(entries, sliding_window, win_duration, win_shift, sample_rate=16000):
    """start / stop vectors of the returned StatServer.  ``entries`` = [(n_windows or None, start, n_samples)]."""
    starts, stops, last = [], [], 1
    for n, start, n_samples in entries:
        if sliding_window:
            starts.extend((numpy.arange(0, n * win_shift, win_shift) * sample_rate + start).tolist())
            split = n // max(1, n // 100)
            last = -(-n // split)
        else:
            starts.append(start)
            stops.append(n_samples)
    start = numpy.array(starts).squeeze()
    return start, start + (last if sliding_window else numpy.array(stops).squeeze())
