"""
CMVN and BatchNorm with the reference's constructor surface
(/root/reference/kaldi_tflite/lib/layers/normalization/{cmvn,batchnorm}.py).
"""

import numpy as np
import torch

from .. import _native as N
from .. import _tensor as T
from .base import Layer


class CMVN(Layer):
    """cmvn.py:38-250 -- centred sliding-window mean (and variance) normalisation."""

    def __init__(self, center=True, norm_vars=False, window=600, min_window=100, padding="SAME",
                 name=None, **kwargs):
        super().__init__(name=name, trainable=False, **kwargs)
        self.center = center
        self.normVar = norm_vars
        self.N = window
        self.minN = min_window
        if not self.center:
            raise NotImplementedError("CMVN with center=False not supported yet")
        if self.N <= 0 or self.minN <= 0:
            raise ValueError("`window` and `min_window` must be > 0")
        self.padding = padding.upper()
        if self.padding not in ["SAME", "VALID"]:
            raise ValueError(f"`padding` should be either 'SAME' or 'VALID', got '{padding}'")

    def numOutputFrames(self, T_in):
        """Frames kept for an utterance of T_in frames (cmvn.py:230-237, python slice semantics)."""
        if self.padding == "SAME":
            return T_in
        return len(range(T_in)[self.N // 2: T_in - (self.N - 1) // 2])

    def compute_output_shape(self, input_shape):
        if self.padding == "SAME":
            return input_shape
        shape = list(input_shape)
        if shape[-2] is not None:
            shape[-2] = self.numOutputFrames(shape[-2])
        return shape

    def get_config(self):
        config = super().get_config()
        config.update({"center": self.center, "norm_vars": self.normVar, "window": self.N,
                       "min_window": self.minN, "padding": self.padding})
        return config

    def forward_ragged(self, x2d, offsets, out_offsets=None, out_rows=None, max_frames=None):
        """x2d (rows, D); `max_frames` is any upper bound on the longest utterance (default: rows)."""
        rows, D = x2d.shape
        max_frames = rows if max_frames is None else max_frames
        valid = self.padding == "VALID"
        if valid and out_offsets is None:
            lens = (offsets[1:] - offsets[:-1]).cpu().tolist()
            oo = np.zeros(len(lens) + 1, dtype=np.int64)
            oo[1:] = np.cumsum([self.numOutputFrames(int(l)) for l in lens])
            out_offsets = torch.from_numpy(oo).to(x2d.device)
            out_rows = int(oo[-1])
        if not valid:
            out_rows = rows
        out = torch.empty((out_rows, D), device=x2d.device, dtype=torch.float32)
        if out_rows > 0:
            N.check(N.lib().ktf_cmvn_forward(T.ptr(x2d), D, T.ptr(offsets), offsets.numel() - 1, rows,
                                             max_frames, self.N, int(self.normVar), int(valid),
                                             T.ptr(out_offsets), T.ptr(out), T.stream_ptr()))
        return out, (out_offsets if valid else offsets)

    def call(self, inputs):
        x = T.as_device(inputs)
        if x.dim() != 3:
            raise ValueError(f"expected input of shape (batch, frames, feats), got {tuple(x.shape)}")
        B, Tn, D = x.shape
        offsets = T.uniform_offsets(B, Tn)
        To = self.numOutputFrames(Tn)
        out_offsets = T.uniform_offsets(B, To) if self.padding == "VALID" else None
        out, _ = self.forward_ragged(x.reshape(B * Tn, D), offsets, out_offsets, B * To, max_frames=Tn)
        return T.like_input(out.reshape(B, To, D), inputs)


class BatchNorm(Layer):
    """batchnorm.py:44-134 -- inference-mode Kaldi BatchNorm: y = gamma (x - mean) / sqrt(var + eps)."""

    def __init__(self, axis=-1, momentum=0.99, target_rms=1.0, epsilon=0.001, mean_initializer=None,
                 variance_initializer=None, name=None, **kwargs):
        super().__init__(name=name, trainable=False, **kwargs)
        if axis != -1:
            raise NotImplementedError("BatchNorm is only supported over the last axis")
        self.axis = axis
        self.momentum = momentum
        self.targetRMS = target_rms
        self.epsilon = epsilon
        self.gamma = None
        self.moving_mean = None
        self.moving_variance = None
        self._dev = None

    def build(self, input_shape):
        dim = input_shape[-1]
        if self.gamma is None:
            self.gamma = np.full((dim,), self.targetRMS, dtype=np.float32)
            self.moving_mean = np.zeros((dim,), dtype=np.float32)
            self.moving_variance = np.ones((dim,), dtype=np.float32)
        elif self.gamma.shape[0] != dim:
            raise ValueError(f"BatchNorm weights are for dim {self.gamma.shape[0]}, input has {dim}")
        super().build(input_shape)

    def get_config(self):
        config = super().get_config()
        config.update({"axis": self.axis, "momentum": self.momentum, "epsilon": self.epsilon,
                       "target_rms": self.targetRMS})
        return config

    def get_weights(self):
        return [self.gamma, self.moving_mean, self.moving_variance]

    def set_weights(self, weights, fmt="kaldi"):
        if fmt not in ["kaldi", "tensorflow"]:
            raise ValueError(f"expected 'fmt' to be either 'kaldi' or 'tensorflow', got {fmt}")
        if len(weights) != 3:
            raise ValueError(f"expected a weight list of length 3, got {len(weights)}")
        g, mean, var = weights
        mean = np.asarray(mean, dtype=np.float32).reshape(-1)
        var = np.asarray(var, dtype=np.float32).reshape(-1)
        if fmt == "kaldi":
            g = np.float32(g) * np.ones_like(mean)           # batchnorm.py:131-132
        g = np.asarray(g, dtype=np.float32).reshape(-1)
        if not (g.shape == mean.shape == var.shape):
            raise ValueError("gamma, mean and variance must have the same shape")
        if self.built and self._build_shape[-1] != mean.shape[0]:
            raise ValueError(f"expected weights of dim {self._build_shape[-1]}, got {mean.shape[0]}")
        self.gamma, self.moving_mean, self.moving_variance = g, mean, var
        self._dev = None
        self._weights_version += 1

    def scale_offset(self):
        """Inference BN folded to y = x * scale + offset (batchnorm.py:81-88, center=False)."""
        scale = (self.gamma / np.sqrt(self.moving_variance + np.float32(self.epsilon))).astype(np.float32)
        offset = (-self.moving_mean * scale).astype(np.float32)
        return scale, offset

    def call(self, inputs, training=False):
        x = T.as_device(inputs)
        self._maybe_build(x.shape)
        if self._dev is None:
            s, o = self.scale_offset()
            self._dev = (T.as_device(s), T.as_device(o))
        D = x.shape[-1]
        y = torch.empty_like(x)
        N.check(N.lib().ktf_scale_offset_forward(T.ptr(x), x.numel() // D, D, T.ptr(self._dev[0]),
                                                 T.ptr(self._dev[1]), T.ptr(y), T.stream_ptr()))
        return T.like_input(y, inputs)
