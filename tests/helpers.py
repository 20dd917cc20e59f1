"""Shared helpers for the test-suite (oracle-side model building, golden loading)."""
import os

import numpy
import torch

from sidekit_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return numpy.load(os.path.join(GOLD, name), allow_pickle=False)


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).norm(dim=-1) / b.norm(dim=-1)).max().item()


def min_cosine(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return torch.nn.functional.cosine_similarity(a, b, dim=-1).min().item()
