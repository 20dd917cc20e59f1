"""PLDA training: ``FactorAnalyser.plda`` of sidekit/factor_analyser.py:830-932, the simplified
PLDA (speaker subspace ``F``, full residual covariance ``Sigma``) whose ``(mean, F, Sigma)`` the scorers consume.

The reference whitens the statistics with ``sqrt(Sigma)^-1`` (an eigendecomposition per iteration), loops over the
speakers in Python (``fa_model_loop``, :166-205) and inverts one ``R x R`` matrix per speaker.  Here the whole EM loop
runs on the device in float64 (``csrc/plda_train.cu``: ``skb_plda_stats`` + ``skb_plda_em``) with no host round trip
inside it: the E- and M-step only depend on the whitening through ``Sigma^-1``, so there is one Cholesky factorisation
of ``Sigma`` per iteration, one of ``n F' Sigma^-1 F + I`` per DISTINCT session count ``n`` (batched), GEMMs over all
classes at once and Cholesky solves for the M-step.  The host only groups the sessions by class and computes the
eigenvoice initialisation (``scipy.linalg.eigh`` of the total covariance, as the reference does: the trained ``F``
depends on the sign convention of those eigenvectors).  HDF5 output (``write`` / ``save_partial`` / ``save_final``) is
out of scope.
"""
import logging

import numpy
import scipy.linalg
import torch

from . import _lib


def _dev(a, dtype=numpy.float64):
    if not torch.cuda.is_available():
        raise RuntimeError("sidekit_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.from_numpy(numpy.ascontiguousarray(a, dtype=dtype)).cuda()


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
        lib = _lib.lib()
        X = _dev(stat_server.stat1)
        n_sess, D = X.shape
        # sessions grouped by (sorted unique) model, in input order inside a class (sum_stat_per_model, statserver.py:1335-1355)
        modelset, inv = numpy.unique(stat_server.modelset, return_inverse=True)
        n_cls = modelset.shape[0]
        order = numpy.argsort(inv, kind="stable").astype(numpy.int32)
        sessions = numpy.bincount(inv, minlength=n_cls)
        cls_ptr = numpy.concatenate([[0], numpy.cumsum(sessions)]).astype(numpy.int32)
        mean = torch.empty((D,), dtype=torch.float64, device=X.device)
        sigma_obs = torch.empty((D, D), dtype=torch.float64, device=X.device)
        S1 = torch.empty((n_cls, D), dtype=torch.float64, device=X.device)
        ptr_d, rows_d = _dev(cls_ptr, numpy.int32), _dev(order, numpy.int32)
        _lib.check(lib.skb_plda_stats(X.data_ptr(), n_sess, D, ptr_d.data_ptr(), rows_d.data_ptr(), n_cls, float(scaling_factor),
                                      mean.data_ptr(), sigma_obs.data_ptr(), S1.data_ptr(), _lib.stream_ptr()))
        # eigenvoice initialisation: leading eigenvectors of the total covariance (:862-866), on the host like the reference
        evals, evecs = scipy.linalg.eigh(sigma_obs.cpu().numpy())
        idx = numpy.argsort(evals)[::-1]
        F = _dev(evecs.real[:, idx[:rank_f]][:, :rank_f])
        # model_shifted_stat.stat0 (one Gaussian) == scaled session count; one R x R factorisation per DISTINCT count
        stat0 = sessions.astype(numpy.float64) * scaling_factor
        uniq_n, n_index = numpy.unique(stat0, return_inverse=True)
        counts = numpy.bincount(n_index, minlength=uniq_n.shape[0]).astype(numpy.float64)
        Sigma = sigma_obs.clone()
        n_d, u_d, uq_d, cnt_d = _dev(stat0), _dev(n_index, numpy.int32), _dev(uniq_n), _dev(counts)
        logging.info('Estimate between class covariance, %d iterations on the device', nb_iter)
        _lib.check(lib.skb_plda_em(S1.data_ptr(), n_d.data_ptr(), u_d.data_ptr(), n_cls, uq_d.data_ptr(), cnt_d.data_ptr(),
                                   int(uniq_n.shape[0]), mean.data_ptr(), sigma_obs.data_ptr(), D, int(rank_f), int(nb_iter),
                                   float(stat0.sum()), F.data_ptr(), Sigma.data_ptr(), _lib.stream_ptr()))
        self.mean, self.F, self.Sigma = mean.cpu().numpy(), F.cpu().numpy(), Sigma.cpu().numpy()
        return self
