"""ROC convex hull, EER and minDCF (sidekit/bosaris/detplot.py:289-511), the numeric part only (no plotting).

``pavx`` and ``rocch`` run natively (``csrc/evaltail.cu``: the reference's Python loop over every trial and its
per-vertex re-summation become O(N log N) C++ with the same floating-point operations, so the hull is bit-identical);
``rocch2eer`` and ``fast_minDCF`` work on the handful of hull vertices and stay in numpy.
"""
import ctypes

import numpy

from . import _lib


def _f64(a):
    return numpy.ascontiguousarray(numpy.asarray(a, dtype=numpy.float64))


def pavx(y):
    """detplot.py:289-347: returns ``(ghat, width, height)``."""
    y = numpy.asarray(y)
    assert y.ndim == 1, 'Argument should be a 1-D array'
    assert y.shape[0] > 0, 'Input array is empty'
    y = _f64(y)
    n = y.shape[0]
    ghat = numpy.empty(n)
    width = numpy.empty(n, dtype=numpy.int64)
    height = numpy.empty(n)
    nb = ctypes.c_int64(0)
    _lib.check(_lib.lib().skb_pavx(y.ctypes.data, n, ghat.ctypes.data, width.ctypes.data, height.ctypes.data, ctypes.byref(nb)))
    return ghat, width[:nb.value].copy(), height[:nb.value].copy()


def rocch(tar_scores, nontar_scores):
    """detplot.py:391-441: ``(pmiss, pfa)`` of the vertices of the ROC convex hull."""
    tar, non = _f64(tar_scores), _f64(nontar_scores)
    n = tar.shape[0] + non.shape[0]
    pmiss, pfa = numpy.empty(n + 1), numpy.empty(n + 1)
    npts = ctypes.c_int64(0)
    _lib.check(_lib.lib().skb_rocch(tar.ctypes.data, tar.shape[0], non.ctypes.data, non.shape[0], pmiss.ctypes.data,
                                    pfa.ctypes.data, ctypes.byref(npts)))
    return pmiss[:npts.value].copy(), pfa[:npts.value].copy()


def _single_threaded_lapack():
    """The hull has a few dozen vertices and every one costs a 2 x 2 ``numpy.linalg.solve``: on a many-core host a threaded
    BLAS wakes its pool for each of them (measured on the GPU box: 0.14 s per ``fast_minDCF`` against 9 ms single-threaded).
    Same LAPACK routine, same result."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=1)
    except Exception:                                    # threadpoolctl missing: run as is
        import contextlib
        return contextlib.nullcontext()


def rocch2eer(pmiss, pfa):
    """detplot.py:350-388: the EER is the largest intersection of a hull segment with the diagonal."""
    with _single_threaded_lapack():
        return _rocch2eer(pmiss, pfa)


def _rocch2eer(pmiss, pfa):
    eer = 0
    for i in range(pfa.shape[0] - 1):
        xx, yy = pfa[i:i + 2], pmiss[i:i + 2]
        assert (xx[1] <= xx[0]) & (yy[0] <= yy[1]), 'pmiss and pfa have to be sorted'
        XY = numpy.column_stack((xx, yy))
        dd = numpy.dot(numpy.array([1, -1]), XY)
        if numpy.min(numpy.abs(dd)) == 0:
            eerseg = 0
        else:
            seg = numpy.linalg.solve(XY, numpy.array([[1], [1]]))     # seg' [x; y] = 1 on the segment's line
            eerseg = 1 / (numpy.sum(seg))
        eer = max([eer, eerseg])
    return eer


def sigmoid(log_odds):
    """detplot.py:444-456."""
    return 1 / (1 + numpy.exp(-log_odds))


def fast_minDCF(tar, non, plo, normalize=False):
    """detplot.py:459-511: ``(minDCF, Pmiss, Pfa, prbep, eer)`` at prior log-odds ``plo``."""
    tar, non = numpy.asarray(tar), numpy.asarray(non)
    Pmiss, Pfa = rocch(tar, non)
    prbep = rocch2eer(Pmiss * tar.shape[0], Pfa * non.shape[0])
    eer = rocch2eer(Pmiss, Pfa)
    Ptar, Pnon = sigmoid(plo), sigmoid(-plo)
    cdet = numpy.dot(numpy.array([[Ptar, Pnon]]), numpy.vstack((Pmiss, Pfa)))
    ii = numpy.argmin(cdet, axis=1)
    minDCF = cdet[0, ii][0]
    if normalize:
        minDCF = minDCF / min([Ptar, Pnon])
    return minDCF, Pmiss[ii][0], Pfa[ii][0], prbep, eer


def eer(negatives, positives):
    """sidekit/nnet/xvector.py:101-209 (``sidekit.nnet.xvector.eer``): bisection EER, native."""
    neg, pos = _f64(negatives), _f64(positives)
    out = ctypes.c_double(0.0)
    rc = _lib.lib().skb_eer(neg.ctypes.data, neg.shape[0], pos.ctypes.data, pos.shape[0], ctypes.byref(out))
    if rc != 0:
        raise IndexError(_lib.lib().skb_last_error().decode())
    return out.value
