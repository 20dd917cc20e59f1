"""Model builders shared by the tests: same key-hash seeded weights for the oracle and the CUDA path."""
import contextlib
import io

import torch

from sidekit_b200 import synth


def make_xtractor(archi, n_spk, emb, seed=0, compute_dtype="fp16", init="default"):
    from sidekit_b200.nnet import Xtractor
    with contextlib.redirect_stdout(io.StringIO()):
        m = Xtractor(n_spk, archi, loss="aam", embedding_size=emb, compute_dtype=compute_dtype)
    sd = m.state_dict()
    synth.fill_state_dict(sd, seed, init)
    m.load_state_dict(sd)
    m.eval()
    return m


def synthetic_state_dict(archi, n_spk, emb, seed=0, init="default"):
    return {k: v.clone() for k, v in make_xtractor(archi, n_spk, emb, seed, init=init).state_dict().items()}
