"""Margin head and normalisation helpers (sidekit/nnet/loss.py:91-100, :256-326); inference branch only."""
import math

import torch


def l2_norm(input, axis=1):
    norm = torch.norm(input, 2, axis, True)
    return torch.div(input, norm)


class ArcMarginProduct(torch.nn.Module):
    """``forward(x, target=None)`` = s * cos(x, W) (sidekit/nnet/loss.py:299-310).  Inside an ``Xtractor`` the engine's
    head GEMM computes it; on its own it runs ``skb_l2_normalize`` + the split-precision tcgen05 score GEMM."""

    def __init__(self, in_features, out_features, s=30.0, m=0.50, easy_margin=False):
        super().__init__()
        self.in_features, self.out_features, self.s, self.m = in_features, out_features, s, m
        self.weight = torch.nn.Parameter(torch.FloatTensor(out_features, in_features))
        torch.nn.init.xavier_uniform_(self.weight)
        self.easy_margin = easy_margin
        self.cos_m, self.sin_m = math.cos(self.m), math.sin(self.m)
        self.th = math.cos(math.pi - self.m)
        self.mm = math.sin(math.pi - self.m) * self.m

    def forward(self, input, target=None):
        if target is not None:
            raise NotImplementedError("training-time margin is out of scope (inference hot path only)")
        from . import functional as Fn
        from ..iv_scoring import score_matrix
        x = Fn.l2_normalize(input)
        w = Fn.l2_normalize(self.weight.detach().to(x.device))
        return score_matrix(x, w, alpha=float(self.s), passes=3)


class SoftmaxAngularProto(torch.nn.Module):
    """Parameter container of the 'aps' head (sidekit/nnet/loss.py:329-373); ``forward(x, target=None)`` =
    ``cce_backend(x)``, a plain Linear on the embedding, runs in the engine's head GEMM."""

    def __init__(self, spk_count, emb_dim=256, init_w=10.0, init_b=-5.0, **kwargs):
        super().__init__()
        from collections import OrderedDict
        self.test_normalize = True
        self.w = torch.nn.Parameter(torch.tensor(init_w))
        self.b = torch.nn.Parameter(torch.tensor(init_b))
        self.cce_backend = torch.nn.Sequential(OrderedDict([("linear8", torch.nn.Linear(emb_dim, spk_count))]))

    def forward(self, x, target=None):
        if target is not None:
            raise NotImplementedError("the angular-prototypical training loss is out of scope (inference hot path only)")
        raise RuntimeError("the 'aps' head runs inside the fused CUDA engine: call Xtractor.forward(x, is_eval=True)")
