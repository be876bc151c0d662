"""ktf.layers -- same names as the reference's kaldi_tflite/lib/layers/__init__.py:20-35."""

from .base import Layer
from .dsp import Framing, FramedSignal, Windowing, FilterBank, DCT, MFCC, VAD
from .normalization import CMVN, BatchNorm
from .tdnn import TDNN, ReLU, reshapeKaldiTdnnWeights
from .stats import StatsPooling
from .plda import PLDA

__all__ = ["Layer", "Framing", "FramedSignal", "Windowing", "FilterBank", "DCT", "MFCC", "VAD", "CMVN",
           "BatchNorm", "TDNN", "ReLU", "StatsPooling", "PLDA", "reshapeKaldiTdnnWeights"]
