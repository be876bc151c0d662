"""
StatsPooling with the reference's constructor surface
(/root/reference/kaldi_tflite/lib/layers/stats/stats_pooling.py:34-316).
"""

import torch

from .. import _native as N
from .. import _tensor as T
from .base import Layer


class StatsPooling(Layer):

    def __init__(self, left_context, right_context, input_period=1, output_period=1, include_std=True,
                 padding="SAME", epsilon=1e-10, reduce_time_axis=False, name=None, **kwargs):
        super().__init__(name=name, trainable=False, **kwargs)
        self.leftContext = left_context
        self.rightContext = right_context
        self.inputPeriod = input_period
        self.outputPeriod = output_period
        self.includeStd = include_std
        self.reduce = reduce_time_axis
        if self.leftContext > 0 or self.rightContext < 0:
            raise ValueError("'left_context' must be <= 0 and 'right_context' must be >= 0")
        if self.inputPeriod <= 0 or self.outputPeriod <= 0:
            raise ValueError("'input_period' and 'output_period' must be > 0")
        if self.outputPeriod % self.inputPeriod != 0 and not self.reduce:
            raise ValueError("'output_period' must be a multiple of 'input_period'")
        self.padding = padding.upper()
        if self.padding not in ["VALID", "SAME"]:
            raise ValueError("padding should be either 'VALID' or 'SAME'")
        self.epsilon = epsilon
        self.maxWindowWidth = right_context - left_context + 1

    def get_config(self):
        config = super().get_config()
        config.update({"left_context": self.leftContext, "right_context": self.rightContext,
                       "input_period": self.inputPeriod, "output_period": self.outputPeriod,
                       "include_std": self.includeStd, "padding": self.padding,
                       "epsilon": self.epsilon, "reduce_time_axis": self.reduce})
        return config

    def _eval_steps(self, Tn):
        """(t_start, num_eval) -- stats_pooling.py:157-177."""
        if self.padding == "SAME":
            start, end = 0, Tn
        else:
            start, end = 0, Tn
            if self.leftContext < 0:
                start = -self.leftContext
            if self.rightContext > 0 and self.maxWindowWidth < Tn:
                end = Tn - self.rightContext
            end = end + 1
        return start, len(range(start, end, self.outputPeriod))

    def _windowed(self, Tn):
        return self.padding == "SAME" or Tn > self.maxWindowWidth

    def compute_output_shape(self, input_shape):
        batch, Tn, dim = input_shape
        od = dim * 2 if self.includeStd else dim
        if self.reduce:
            return (batch, 1, od)
        if self.padding == "SAME":
            return (batch, Tn, od)
        if Tn is None:
            return (batch, None, od)
        return (batch, self._eval_steps(Tn)[1] if self._windowed(Tn) else 1, od)

    def reduce_ragged(self, x2d, offsets):
        """(rows, dim) -> (B, dim or 2 dim): stats_pooling.py:211-240 per utterance."""
        rows, D = x2d.shape
        B = offsets.numel() - 1
        od = 2 * D if self.includeStd else D
        out = torch.empty((B, od), device=x2d.device, dtype=torch.float32)
        N.check(N.lib().ktf_stats_reduce(T.ptr(x2d), T.ptr(offsets), B, D, self.inputPeriod,
                                         int(self.includeStd), float(self.epsilon), T.ptr(out),
                                         T.stream_ptr()))
        return out

    def finalize_sums(self, sums, offsets):
        """(B, 2, dim) sums from the fused TDNN epilogue -> (B, dim or 2 dim)."""
        B, _, D = sums.shape
        od = 2 * D if self.includeStd else D
        out = torch.empty((B, od), device=sums.device, dtype=torch.float32)
        N.check(N.lib().ktf_stats_finalize(T.ptr(sums), T.ptr(offsets), B, D, int(self.includeStd),
                                           float(self.epsilon), 1, T.ptr(out), T.stream_ptr()))
        return out

    def call(self, inputs):
        x = T.as_device(inputs)
        if x.dim() != 3:
            raise ValueError(f"expected input of shape (batch, timesteps, feats), got {tuple(x.shape)}")
        B, Tn, D = x.shape
        if self.reduce or not self._windowed(Tn):
            out = self.reduce_ragged(x.reshape(B * Tn, D), T.uniform_offsets(B, Tn)).reshape(B, 1, -1)
            return T.like_input(out, inputs)
        t_start, num_eval = self._eval_steps(Tn)
        right_excl = min(self.rightContext + 1, Tn)              # stats_pooling.py:184-189
        repeat = self.outputPeriod if self.padding == "SAME" else 1
        od = 2 * D if self.includeStd else D
        out = torch.empty((B, num_eval * repeat, od), device=x.device, dtype=torch.float32)
        N.check(N.lib().ktf_stats_windows(T.ptr(x), B, Tn, D, self.leftContext, right_excl,
                                          self.inputPeriod, t_start, num_eval, self.outputPeriod, repeat,
                                          int(self.includeStd), float(self.epsilon), T.ptr(out),
                                          T.stream_ptr()))
        return T.like_input(out, inputs)
