"""Adaptive symmetric score normalisation (sidekit/score_normalization.py:120-140)."""
import torch

from . import _lib


def asnorm(enrol_xv, cohort_xv, ndx=None, topk=200, return_device=False):
    """Same contract as the reference: ``enrol_xv`` (N, D) unit-norm embeddings, ``cohort_xv`` (C, D) raw cohort
    vectors (normalised here, :124), ``ndx`` unused; returns the (N, N) float32 normalised score matrix as numpy."""
    if not torch.cuda.is_available():
        raise RuntimeError("sidekit_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = enrol_xv.device if enrol_xv.is_cuda else torch.device("cuda", torch.cuda.current_device())
    X = enrol_xv.to(dev, torch.float32).contiguous()
    coh = torch.nn.functional.normalize(cohort_xv.to(dev, torch.float32), dim=1).contiguous()
    N, D = X.shape
    C = coh.shape[0]
    mean = torch.empty((N,), dtype=torch.float32, device=dev)
    std = torch.empty((N,), dtype=torch.float32, device=dev)
    out = torch.empty((N, N), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        l = _lib.lib()
        _lib.check(l.skb_asnorm_stats(X.data_ptr(), coh.data_ptr(), N, C, D, int(topk), mean.data_ptr(), std.data_ptr(),
                                      _lib.stream_ptr()))
        _lib.check(l.skb_asnorm_apply(X.data_ptr(), N, D, mean.data_ptr(), std.data_ptr(), out.data_ptr(), _lib.stream_ptr()))
    return out if return_device else out.cpu().numpy()
