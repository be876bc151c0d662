"""Kaldi file-format readers (host side, load-time only)."""

from .object_reader import KaldiObjReader
from .nnet3_reader import KaldiNnet3Reader
from .plda_reader import KaldiPldaReader
from .array_reader import ReadKaldiArray

__all__ = ["KaldiObjReader", "KaldiNnet3Reader", "KaldiPldaReader", "ReadKaldiArray"]
