"""B200-native (sm_100a) open-set classification head and evaluation path.

Drop-in for the hot path of AIML-IfI/openset-imagenet: ``losses.EntropicOpensetLoss`` and the
softmax / garbage-class cross-entropy variants, ``metrics.confidence``, ``util.calculate_oscr`` and the
CCR@FPR lookup, plus ``model.FusedOpenSetHead`` (tcgen05/TMEM linear + loss + backward).  Everything
computes in hand-written CUDA kernels behind the C ABI of ``include/osb.h``; there is no CPU fallback.
"""
from . import losses, metrics, train, util  # noqa: F401
from .losses import CrossEntropyLoss, EntropicOpensetLoss, make_loss  # noqa: F401
from .metrics import StreamingConfidence, confidence  # noqa: F401
from .util import calculate_oscr, ccr_at_fpr  # noqa: F401

__all__ = ["losses", "metrics", "train", "util", "StreamingConfidence", "EntropicOpensetLoss", "CrossEntropyLoss", "make_loss",
           "confidence", "calculate_oscr", "ccr_at_fpr"]
