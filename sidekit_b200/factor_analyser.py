"""PLDA training: ``FactorAnalyser.plda`` of sidekit/factor_analyser.py:830-932, the simplified
PLDA (speaker subspace ``F``, full residual covariance ``Sigma``) whose ``(mean, F, Sigma)`` the scorers consume.

The reference loops over the speakers in Python (``fa_model_loop``, :166-205).  Here the E-step is closed-form linear
algebra over ALL classes at once: with ``L_n = (n F'F + I)^-1`` (one inverse per distinct session count ``n``),
``E[h_i] = (s_i F) L_{n_i}`` and ``E[h_i h_i'] = L_{n_i} + E[h_i] E[h_i]'``, so the accumulators are

    R = (sum_i L_{n_i} + Eh' Eh) / C        A = sum_i n_i L_{n_i} + Eh' diag(n) Eh        C = Eh' S (sqrt(Sigma)^-1)^-1

i.e. three GEMMs over the (classes x D) statistics -- float64 on the device (plain library GEMMs: there is nothing to
fuse) -- and D x D / R x R factorizations on the host exactly as the reference does them (scipy).  HDF5 output
(``write`` / ``save_partial`` / ``save_final``) is out of scope.
"""
import logging

import numpy
import scipy.linalg
import torch


def _dev(a):
    if not torch.cuda.is_available():
        raise RuntimeError("sidekit_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.from_numpy(numpy.ascontiguousarray(a, dtype=numpy.float64)).cuda()


def _sqr_inv(sigma):
    """sqrt(Sigma)^-1 as the reference builds it (statserver.py:872-878): eigenvectors scaled by 1/sqrt(eigenvalue)."""
    eigen_values, eigen_vectors = scipy.linalg.eigh(sigma)
    ind = eigen_values.real.argsort()[::-1]
    eigen_values = eigen_values.real[ind]
    eigen_vectors = eigen_vectors.real[:, ind]
    return numpy.dot(eigen_vectors, numpy.diag(1 / numpy.sqrt(eigen_values.real)))


class FactorAnalyser:
    """factor_analyser.py:207-262: ``mean``, ``F``, ``G``, ``H``, ``Sigma`` (file IO out of scope)."""

    def __init__(self, input_file_name=None, mean=None, F=None, G=None, H=None, Sigma=None):
        if input_file_name is not None:
            raise NotImplementedError("FactorAnalyser HDF5 IO is out of scope")
        self.mean, self.F, self.G, self.H, self.Sigma = mean, F, G, H, Sigma

    def plda(self, stat_server, rank_f, nb_iter=10, scaling_factor=1., output_file_name=None, save_partial=False,
             save_final=True, num_thread=1):
        """factor_analyser.py:830-932.  Sets ``self.mean``, ``self.F`` (D, rank_f) and ``self.Sigma`` (D, D)."""
        if output_file_name is not None or save_partial:
            logging.warning("sidekit_b200: FactorAnalyser.plda does not write HDF5 files; the model stays in memory")
        stat1 = numpy.asarray(stat_server.stat1, dtype=numpy.float64)
        X = _dev(stat1)
        n_sess = X.shape[0]
        # mean and total covariance of the training data (statserver.py:789-795, :920-928)
        mean_d = X.mean(dim=0)
        Cd = X - mean_d
        sigma_obs = (Cd.T @ Cd / n_sess).cpu().numpy()
        self.mean = mean_d.cpu().numpy()
        self.Sigma = sigma_obs.copy()
        # statistics summed per (sorted unique) model (statserver.py:1335-1355)
        modelset, inv = numpy.unique(stat_server.modelset, return_inverse=True)
        class_nb = modelset.shape[0]
        inv_d = torch.from_numpy(inv.astype(numpy.int64)).cuda()
        S1 = torch.zeros((class_nb, X.shape[1]), dtype=torch.float64, device=X.device).index_add_(0, inv_d, X)
        sessions = numpy.bincount(inv, minlength=class_nb).astype(numpy.float64)
        S1 = S1 * scaling_factor
        stat0 = sessions * scaling_factor                      # model_shifted_stat.stat0 (one Gaussian) == scaled session count
        session_per_model = sessions * scaling_factor
        n_d = _dev(stat0)
        # eigenvoice initialisation: leading eigenvectors of the total covariance (:862-866)
        evals, evecs = scipy.linalg.eigh(sigma_obs)
        idx = numpy.argsort(evals)[::-1]
        self.F = evecs.real[:, idx[:rank_f]][:, :rank_f]
        uniq_n, n_index = numpy.unique(stat0, return_inverse=True)
        counts = numpy.bincount(n_index, minlength=uniq_n.shape[0]).astype(numpy.float64)
        n_index_d = torch.from_numpy(n_index.astype(numpy.int64)).cuda()
        for it in range(nb_iter):
            logging.info('Estimate between class covariance, it %d / %d', it + 1, nb_iter)
            # whiten the statistics and the eigenvoice matrix with the current (mean, Sigma) (:878-893)
            sqr_inv_sigma = _sqr_inv(self.Sigma)
            W = _dev(sqr_inv_sigma)
            local = (S1 - n_d[:, None] * _dev(self.mean)[None, :]) @ W                  # center_stat1 (stat0-weighted) + rotate
            F = sqr_inv_sigma.T.dot(self.F)
            # E-step over all classes (fa_model_loop, :166-205)
            A0 = F.T.dot(F)
            inv_lambda = numpy.stack([scipy.linalg.inv(n * A0 + numpy.eye(rank_f)) for n in uniq_n])       # (U, R, R)
            L = _dev(inv_lambda)
            aux = local @ _dev(F)                                                        # (C, R)
            e_h = torch.bmm(aux.unsqueeze(1), L[n_index_d]).squeeze(1)                   # aux_i . L_{n_i}
            sum_L = numpy.einsum('u,ujk->jk', counts, inv_lambda)
            sum_nL = numpy.einsum('u,ujk->jk', counts * uniq_n, inv_lambda)
            _R = (sum_L + (e_h.T @ e_h).cpu().numpy()) / session_per_model.shape[0]
            _C = (e_h.T @ local).cpu().numpy().dot(scipy.linalg.inv(sqr_inv_sigma))
            _A = sum_nL + (e_h.T @ (e_h * n_d[:, None])).cpu().numpy()
            # M-step, residual covariance, minimum divergence (:914-922)
            self.F = scipy.linalg.solve(_A, _C).T
            self.Sigma = sigma_obs - self.F.dot(_C) / session_per_model.sum()
            self.F = self.F.dot(scipy.linalg.cholesky(_R))
        return self
