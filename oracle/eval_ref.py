"""CPU restatement (numpy) of the evaluation tail: PAV, ROC convex hull, EER, minDCF, z-/t-norm.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Follows sidekit/bosaris/detplot.py:289-511,
sidekit/nnet/xvector.py:101-209 and sidekit/score_normalization.py:44-117 step by step (including the
reference's quirks, which the product reproduces).  Pinned by tests/golden/evaltail.npz, recorded from the
imported reference by oracle/make_golden.py.
"""
import numpy


def pavx(y):
    """detplot.py:289-347.  Quirk kept: the last fill pass writes ghat[-1] (Python wrap-around)."""
    n = y.shape[0]
    index = numpy.zeros(n, dtype=int)
    length = numpy.zeros(n, dtype=int)
    ghat = numpy.zeros(n)
    ci = 0
    length[0] = 1
    ghat[0] = y[0]
    for j in range(1, n):
        ci += 1
        index[ci] = j + 1
        length[ci] = 1
        ghat[ci] = y[j]
        while ci >= 1 and ghat[ci - 1] >= ghat[ci]:
            nw = length[ci - 1] + length[ci]
            ghat[ci - 1] = ghat[ci - 1] + (length[ci] / nw) * (ghat[ci] - ghat[ci - 1])
            length[ci - 1] = nw
            ci -= 1
    height = ghat[:ci + 1].copy()
    width = length[:ci + 1].copy()
    while n >= 0:
        for j in range(int(index[ci]), int(n + 1)):
            ghat[j - 1] = ghat[ci]
        n = index[ci] - 1
        ci -= 1
    return ghat, width, height


def rocch(tar, non):
    """detplot.py:391-441."""
    nt, nn = tar.shape[0], non.shape[0]
    n = nt + nn
    scores = numpy.concatenate((tar, non))
    ideal = numpy.concatenate((numpy.ones(nt), numpy.zeros(nn)))
    ideal = ideal[numpy.argsort(scores, kind='mergesort')]
    _, width, _ = pavx(ideal)
    nbins = width.shape[0]
    pmiss, pfa = numpy.zeros(nbins + 1), numpy.zeros(nbins + 1)
    left, fa, miss = 0, nn, 0
    for i in range(nbins):
        pmiss[i] = miss / nt
        pfa[i] = fa / nn
        left = int(left + width[i])
        miss = numpy.sum(ideal[:left])
        fa = n - left - numpy.sum(ideal[left:])
    pmiss[nbins] = miss / nt
    pfa[nbins] = fa / nn
    return pmiss, pfa


def znorm_matrix(test, imp, sym=False):
    """score_normalization.py:62-70 on already sorted matrices."""
    if sym:
        mean = (imp.sum(1) - numpy.diag(imp)) / (imp.shape[1] - 1)
        tmp = numpy.square(imp - mean)
        std = (tmp.sum(1) - numpy.diag(tmp)) / (tmp.shape[1] - 1)
    else:
        mean, std = imp.mean(1), imp.std(1)
    return (test - mean) / std


def tnorm_matrix(test, imp):
    """score_normalization.py:91-93."""
    return (test - imp.mean(0)) / imp.std(0)
