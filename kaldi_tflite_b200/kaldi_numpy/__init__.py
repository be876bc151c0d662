"""
Host-side NumPy helpers for Kaldi's `snip-edges=false` framing (the reference keeps the same
helpers in kaldi_tflite/lib/kaldi_numpy/frame_extraction.py:28-125; the layers never pad).
"""

import numpy as np

__all__ = ["MirrorPad", "PadWaveform", "ExtractFrames"]


def MirrorPad(x, left_pad, right_pad):
    """Reflects `left_pad` / `right_pad` samples about the first / last sample (edge repeated)."""
    x = np.asarray(x)
    left = x[..., :left_pad][..., ::-1]
    right = x[..., x.shape[-1] - right_pad:][..., ::-1] if right_pad > 0 else x[..., :0]
    return np.concatenate([left, x, right], axis=-1)


def PadWaveform(x, frameSize, frameShift):
    """Pads so that un-padded framing yields Kaldi's snip-edges=false frames: round(N / shift) of them."""
    n = np.asarray(x).shape[-1]
    frames = (n + frameShift // 2) // frameShift
    needed = (frames - 1) * frameShift + frameSize
    left = (frameSize - frameShift) // 2
    return MirrorPad(x, left, abs(n - needed) - left)


def ExtractFrames(samples, frameSizeMs, frameShiftMs, sampleFreq, snipEdges):
    """Strided (frames, frameSize) view following Kaldi's frame count rules."""
    m = int(sampleFreq * frameSizeMs / 1000.0)
    k = int(sampleFreq * frameShiftMs / 1000.0)
    x = np.asarray(samples)
    n = x.shape[-1]
    if snipEdges:
        n = ((n - m) // k) * k + m
        x = x[..., :n]
    return np.lib.stride_tricks.sliding_window_view(x, m, axis=-1)[..., ::k, :]
