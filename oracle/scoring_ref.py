"""numpy restatement of the reference's trial scoring (float64, like the reference).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Inputs are plain arrays:
``en_ids`` / ``te_ids`` (1-D str arrays = StatServer.modelset / .segset),
``E`` / ``T`` (stat1), ``ndx_models`` / ``ndx_segs`` / ``trialmask``.  Each
function returns ``(modelset, segset, scoremask, scoremat)`` -- the four fields
of the reference's ``Scores`` object (bosaris/scores.py:69-72).

Pinned by the known-answer vectors of SURVEY.md Appendix B
(tests/test_oracle_kat.py) and, in the build container, against the imported
reference (tests/test_oracle_vs_reference.py).
"""
import numpy
import scipy.linalg


def first_index(have, wanted):
    """``numpy.argwhere(have == v)[0][0]`` for every v (statserver.py:656-684): FIRST match wins."""
    pos = {}
    for i, v in enumerate(have.tolist()):
        if v not in pos:
            pos[v] = i
    return numpy.array([pos[v] for v in wanted.tolist()], dtype=numpy.int64)


def check_missing(en_ids, te_ids, ndx_models, ndx_segs, trialmask):
    """iv_scoring.py:52-60 + Ndx.filter(keep=True) (bosaris/ndx.py:128-165).

    Rows follow ``ndx_models`` order restricted to ids present in ``en_ids``
    (every occurrence in the ndx is kept), columns likewise; returns the row /
    column gather indices into the enrol / test matrices.
    """
    en_set, te_set = set(en_ids.tolist()), set(te_ids.tolist())
    keepm = numpy.array([m in en_set for m in ndx_models.tolist()], dtype=bool)
    keeps = numpy.array([s in te_set for s in ndx_segs.tolist()], dtype=bool)
    models, segs = ndx_models[keepm], ndx_segs[keeps]
    mask = trialmask[keepm, :][:, keeps]
    return models, segs, mask, first_index(en_ids, models), first_index(te_ids, segs)


def mean_per_model(ids, X):
    """StatServer.mean_stat_per_model (statserver.py:1357-1374): sorted unique ids, row means."""
    uniq = numpy.unique(ids)
    out = numpy.zeros((uniq.shape[0], X.shape[1]), dtype=numpy.float64)
    for i, m in enumerate(uniq):
        out[i] = X[ids == m].mean(axis=0)
    return uniq, out


def cosine_scoring(en_ids, E, te_ids, T, ndx_models, ndx_segs, trialmask, wccn=None):
    """iv_scoring.py:63-113; returns float32 like the reference's FloatTensor einsum."""
    models, segs, mask, ri, ci = check_missing(en_ids, te_ids, ndx_models, ndx_segs, trialmask)
    E, T = E[ri].astype(numpy.float64), T[ci].astype(numpy.float64)
    if wccn is not None:
        E, T = E @ wccn, T @ wccn
    E = (E.T / numpy.clip(numpy.linalg.norm(E, axis=1), 1e-8, numpy.inf)).T      # statserver.py:797-800
    T = (T.T / numpy.clip(numpy.linalg.norm(T, axis=1), 1e-8, numpy.inf)).T
    S = E.astype(numpy.float32) @ T.astype(numpy.float32).T
    return models, segs, mask, S


def plda_matrices(F, Sigma, scaling_factor=1.0):
    """The D x D algebra of fast_PLDA_scoring, iv_scoring.py:429-448 -> (Phi, Psi, plda_cst)."""
    invSigma = scipy.linalg.inv(Sigma)
    I_spk = numpy.eye(F.shape[1], dtype="float")
    K = F.T.dot(invSigma * scaling_factor).dot(F)
    K1 = scipy.linalg.inv(K + I_spk)
    K2 = scipy.linalg.inv(2 * K + I_spk)
    alpha1 = numpy.linalg.slogdet(K1)[1]
    alpha2 = numpy.linalg.slogdet(K2)[1]
    plda_cst = alpha2 / 2.0 - alpha1
    Sigma_ac = numpy.dot(F, F.T)
    Sigma_tot = Sigma_ac + Sigma
    Sigma_tot_inv = scipy.linalg.inv(Sigma_tot)
    Tmp = numpy.linalg.inv(Sigma_tot - Sigma_ac.dot(Sigma_tot_inv).dot(Sigma_ac))
    Phi = Sigma_tot_inv - Tmp
    Psi = Sigma_tot_inv.dot(Sigma_ac).dot(Tmp)
    return Phi, Psi, plda_cst


def fast_plda_scoring(en_ids, E, te_ids, T, ndx_models, ndx_segs, trialmask, mu, F, Sigma,
                      p_known=0.0, scaling_factor=1.0):
    """fast_PLDA_scoring, iv_scoring.py:370-477 (simplified PLDA log-likelihood ratio)."""
    E = E.astype(numpy.float64)
    T = T.astype(numpy.float64)
    if numpy.unique(en_ids).shape != en_ids.shape:                        # :409-411
        en_ids, E = mean_per_model(en_ids, E)
    models, segs, mask, ri, ci = check_missing(en_ids, te_ids, ndx_models, ndx_segs, trialmask)
    E, T = E[ri] - mu, T[ci] - mu                                          # center_stat1, stat0 == 1
    if numpy.unique(models).shape != models.shape:
        # :422-425 -- the reference averages AGAIN after alignment; when the ndx lists a model twice the
        # score matrix then has one row per sorted-unique model while `modelset` keeps the ndx order
        # (an inconsistent Scores object; reproduced verbatim).
        _, E = mean_per_model(models, E)
    Phi, Psi, cst = plda_matrices(F, Sigma, scaling_factor)
    model_part = 0.5 * numpy.einsum("ij, ji->i", E.dot(Phi), E.T)
    seg_part = 0.5 * numpy.einsum("ij, ji->i", T.dot(Phi), T.T)
    S = model_part[:, numpy.newaxis] + seg_part + cst
    S += E.dot(Psi).dot(T.T)
    S *= scaling_factor
    if p_known != 0:                                                       # :467-475
        N = S.shape[0]
        tmp = numpy.exp(S)
        out = numpy.empty(S.shape)
        for ii in range(N):
            out[ii, :] = S[ii, :] - numpy.log(p_known * tmp[~(numpy.arange(N) == ii)].sum(axis=0) / (N - 1)
                                               + (1 - p_known))
        S = out
    return models, segs, mask, S


def full_plda_scoring(en_ids, E, te_ids, T, ndx_models, ndx_segs, trialmask, mu, F, G, Sigma, p_known=0.0, scaling_factor=1.0):
    """full_PLDA_scoring, iv_scoring.py:272-368 (speaker subspace F, channel subspace G); loops as in the reference."""
    E, T = E.astype(numpy.float64), T.astype(numpy.float64)
    models, segs, mask, ri, ci = check_missing(en_ids, te_ids, ndx_models, ndx_segs, trialmask)
    E, T = E[ri] - mu, T[ci] - mu
    invSigma = scipy.linalg.inv(Sigma)
    I_iv, I_ch, I_spk = numpy.eye(mu.shape[0]), numpy.eye(G.shape[1]), numpy.eye(F.shape[1])
    A = numpy.linalg.inv(G.T.dot(invSigma * scaling_factor).dot(G) + I_ch)
    B = F.T.dot(invSigma * scaling_factor).dot(I_iv - G.dot(A).dot(G.T).dot(invSigma * scaling_factor))
    K = B.dot(F)
    K1 = scipy.linalg.inv(K + I_spk)
    K2 = scipy.linalg.inv(2 * K + I_spk)
    constant = numpy.linalg.slogdet(K2)[1] / 2.0 - numpy.linalg.slogdet(K1)[1]
    test_tmp, enroll_tmp = B.dot(T.T), B.dot(E.T)
    S1 = numpy.array([test_tmp[:, j].dot(K1).dot(test_tmp[:, j]) / 2. for j in range(T.shape[0])])
    S = numpy.zeros((E.shape[0], T.shape[0]))
    S2 = numpy.empty(E.shape[0])
    for i in range(E.shape[0]):
        both = test_tmp + enroll_tmp[:, i:i + 1]
        S2[i] = enroll_tmp[:, i].dot(K1).dot(enroll_tmp[:, i]) / 2.
        S[i, :] = numpy.einsum("ij, ji->i", both.T.dot(K2), both) / 2.
    S += constant - (S1 + S2[:, numpy.newaxis])
    S *= scaling_factor
    if p_known != 0:
        N = S.shape[0]
        tmp = numpy.exp(S)
        out = numpy.empty(S.shape)
        for ii in range(N):
            out[ii, :] = S[ii, :] - numpy.log(p_known * tmp[~(numpy.arange(N) == ii)].sum(axis=0) / (N - 1) + (1 - p_known))
        S = out
    return models, segs, mask, S


def mahalanobis_scoring(en_ids, E, te_ids, T, ndx_models, ndx_segs, trialmask, m):
    """mahalanobis_scoring, iv_scoring.py:116-156 (check_missing=True: both sets aligned with the cleaned ndx)."""
    E = E.astype(numpy.float64)
    T = T.astype(numpy.float64)
    if numpy.unique(en_ids).shape != en_ids.shape:
        en_ids, E = mean_per_model(en_ids, E)
    models, segs, mask, ri, ci = check_missing(en_ids, te_ids, ndx_models, ndx_segs, trialmask)
    E, T = E[ri], T[ci]
    S = numpy.zeros((E.shape[0], T.shape[0]))
    for i in range(E.shape[0]):
        diff = E[i, :] - T
        S[i, :] = -0.5 * numpy.sum(numpy.dot(diff, m) * diff, axis=1)
    return models, segs, mask, S


def two_covariance_scoring(en_ids, E, te_ids, T, ndx_models, ndx_segs, trialmask, W, B):
    """two_covariance_scoring, iv_scoring.py:159-212 (no centring)."""
    E = E.astype(numpy.float64)
    T = T.astype(numpy.float64)
    if numpy.unique(en_ids).shape != en_ids.shape:
        en_ids, E = mean_per_model(en_ids, E)
    models, segs, mask, ri, ci = check_missing(en_ids, te_ids, ndx_models, ndx_segs, trialmask)
    E, T = E[ri], T[ci]
    iW = scipy.linalg.inv(W)
    iB = scipy.linalg.inv(B)
    G = iW @ scipy.linalg.inv(iB + 2 * iW) @ iW
    H = iW @ scipy.linalg.inv(iB + iW) @ iW
    s2 = numpy.sum(numpy.dot(E, H) * E, axis=1)
    s3 = numpy.sum(numpy.dot(T, H) * T, axis=1)
    S = numpy.zeros((E.shape[0], T.shape[0]))
    for ii in range(E.shape[0]):
        A = E[ii, :] + T
        S[ii, :] = numpy.sum(numpy.dot(A, G) * A, axis=1) - s3 - s2[ii]
    return models, segs, mask, S


def asnorm(X, cohort, topk=200):
    """score_normalization.py:120-140 on float32 inputs (X unit-norm rows, cohort raw)."""
    X = X.astype(numpy.float32)
    C = cohort.astype(numpy.float32)
    C = C / numpy.maximum(numpy.linalg.norm(C, axis=1, keepdims=True), 1e-12)
    S = X @ X.T
    Cs = X @ C.T
    top = -numpy.sort(-Cs, axis=1)[:, :topk]
    mean = top.mean(axis=1)
    std = top.std(axis=1, ddof=1)
    return 0.5 * ((S.T - mean) / std).T + 0.5 * (S - mean) / std
