"""Inference-path subset of ``sidekit.nnet`` (sidekit/nnet/__init__.py:31-47)."""
from .xvector import Xtractor
from .pooling import MeanStdPooling, AttentivePooling
from .res_net import PreHalfResNet34, PreResNet34, PreFastResNet34, BasicBlock, SELayer, ResBlock
from . import functional
from .preprocessor import MfccFrontEnd, MelSpecFrontEnd, PreEmphasis, Resample
from .loss import ArcMarginProduct, l2_norm
from .xsets import IdMap, IdMapSet, extract_embeddings, load_checkpoint, read_wav  # noqa: F401,E402
