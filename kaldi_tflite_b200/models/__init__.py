"""ktf.models -- same names as the reference's kaldi_tflite/lib/models/__init__.py:20-21."""

from .sequential import Sequential, SequentialFromConfig, cfg2layers
from .xvector_extractor import XvectorExtractor, XvectorExtractorFromConfig

__all__ = ["Sequential", "SequentialFromConfig", "cfg2layers", "XvectorExtractor",
           "XvectorExtractorFromConfig"]
