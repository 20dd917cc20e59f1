"""Temporal pooling modules (sidekit/nnet/pooling.py:44-70, :123-171)."""
import ctypes

import torch

from .. import _lib


class MeanStdPooling(torch.nn.Module):
    """Mean and unbiased standard deviation over time, concatenated (pooling.py:55-70)."""

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("sidekit_b200 has no CPU path: move the tensor to a CUDA device")
        if x.dim() == 4:                        # [B, C, T, F] -> [B, C*F, T]
            x = x.permute(0, 1, 3, 2).flatten(start_dim=1, end_dim=2)
        x = x.contiguous().float()
        B, D, T = x.shape
        out = torch.empty((B, 2 * D), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().skb_meanstd_pool(x.data_ptr(), B, D, T, out.data_ptr(), _lib.stream_ptr()))
        return out


class AttentivePooling(torch.nn.Module):
    """Attentive statistics pooling (global-context term hoisted to a per-utterance bias, softmax over time, weighted
    mean / std)."""

    def __init__(self, num_channels, num_freqs=10, attention_channels=128, global_context=False):
        super().__init__()
        in_factor = 3 if global_context else 1
        self.attention = torch.nn.Sequential(
            torch.nn.Conv1d(num_channels * num_freqs * in_factor, attention_channels, kernel_size=1),
            torch.nn.ReLU(),
            torch.nn.BatchNorm1d(attention_channels),
            torch.nn.Tanh(),
            torch.nn.Conv1d(attention_channels, num_channels * num_freqs, kernel_size=1),
            torch.nn.Softmax(dim=2))
        self.global_context = global_context
        self.gc = MeanStdPooling()

    def forward(self, x):
        """pooling.py:151-171 (eval mode): ``x`` (B, C, T, F) or (B, C*F, T) -> (B, 2*C*F) = [weighted mean ; weighted std].
        Inside an ``Xtractor`` the fused engine runs this from its plane layout; on its own it goes through
        ``skb_attentive_pool``."""
        from . import functional as Fn
        if len(x.shape) == 4:
            x = x.permute(0, 1, 3, 2).flatten(start_dim=1, end_dim=2)
        a = self.attention
        return Fn.attentive_pool(x, a[0].weight, a[0].bias, a[2], a[4].weight, a[4].bias, self.global_context)
