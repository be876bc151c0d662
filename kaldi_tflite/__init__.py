"""Drop-in alias: `import kaldi_tflite as ktf` resolves to the B200 implementation
(the reference package exposes ktf.layers / ktf.models / ktf.io / ktf.kaldi_numpy,
kaldi_tflite/__init__.py:18-23)."""

from kaldi_tflite_b200 import io, kaldi_numpy, layers, models, parallel  # noqa: F401
from kaldi_tflite_b200 import KtfNativeError, __version__  # noqa: F401
