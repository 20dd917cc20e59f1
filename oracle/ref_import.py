"""Import the REAL reference (``/root/reference``) in the build container.

Test infrastructure only (see oracle/__init__.py).  The reference cannot
travel to the GPU box, so everything here is used solely to (i) validate the
restatement in ``extract_ref.py`` / ``scoring_ref.py`` and (ii) generate the
committed fixtures under ``tests/golden/``.

Recipe:
  * ``h5py``, ``matplotlib(.pyplot/.mlab)``, ``soundfile`` are not installed;
    they are only needed for file IO / plotting, so empty stub modules are
    registered before ``import sidekit``.
  * P1: ``Xtractor("halfresnet34")`` ships ``AttentivePooling(256, 80, ...)``
    which cannot consume the trunk's 2560 channels (sidekit/nnet/xvector.py:583
    vs the 5120-wide ``lin_be`` at :579) -> replaced by
    ``AttentivePooling(256, 10, global_context=True)``.
  * P2: ``MfccFrontEnd.forward(self, x)`` lacks the ``is_eval`` argument that
    ``Xtractor.forward`` passes (preprocessor.py:113 vs xvector.py:885).
"""
import contextlib
import io
import os
import sys
import types

# /root/reference in the build container; on the GPU box the unmodified install made by oracle/build_ref.py (git-ignored,
# travels with the snapshot)
_INSTALLED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("SIDEKIT_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isdir("/root/reference/sidekit") else _INSTALLED)


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sidekit"))


def import_reference():
    """Return the imported ``sidekit`` package of the reference."""
    if "sidekit" in sys.modules and getattr(sys.modules["sidekit"], "__file__", "").startswith(REFERENCE_ROOT):
        return sys.modules["sidekit"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("h5py", "matplotlib", "matplotlib.pyplot", "matplotlib.mlab", "soundfile"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                if name == "h5py":
                    m.File = object
                if name == "matplotlib":
                    m.use = lambda *a, **k: None
                sys.modules[name] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torch
    rng_state = torch.get_rng_state()
    with contextlib.redirect_stdout(io.StringIO()):
        import sidekit  # noqa
    torch.set_rng_state(rng_state)   # preprocessor.py:55-58 reseeds at import
    torch.backends.cudnn.deterministic = False
    return sidekit


def build_xtractor(speaker_number, model_archi, embedding_size, loss="aam"):
    """Patched reference Xtractor in eval mode (P1 / P2 applied)."""
    sidekit = import_reference()
    from sidekit.nnet.xvector import Xtractor
    from sidekit.nnet.pooling import AttentivePooling
    from sidekit.nnet import preprocessor as _pp
    with contextlib.redirect_stdout(io.StringIO()):
        model = Xtractor(speaker_number, model_archi, loss=loss, embedding_size=embedding_size)
    if model_archi in ("halfresnet34", "resnet34"):
        model.stat_pooling = AttentivePooling(256, 10, global_context=True)       # P1 (same defect in both constructors)
    if model_archi == "fastresnet34":
        model.stat_pooling = AttentivePooling(128, 10, global_context=False)      # P1: shipped num_freqs=80 -> 10240 inputs
    if model_archi == "xvector":
        fe = model.preprocessor
        orig = _pp.MfccFrontEnd.forward

        def fwd(x, is_eval=False, _fe=fe, _orig=orig):                            # P2
            return _orig(_fe, x)
        fe.forward = fwd
        if not hasattr(model, "embedding_size"):
            model.embedding_size = embedding_size
    model.eval()
    return model
