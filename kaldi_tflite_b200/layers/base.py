"""Minimal stand-in for the tf.keras Layer protocol the reference's layers follow
(ctor kwargs, __call__, build, get_config / from_config, compute_output_shape,
set_weights / get_weights)."""

import itertools

_counters = {}


def _auto_name(cls_name):
    n = _counters.get(cls_name, 0)
    _counters[cls_name] = n + 1
    snake = "".join("_" + c.lower() if c.isupper() and i else c.lower() for i, c in enumerate(cls_name))
    return snake if n == 0 else f"{snake}_{n}"


class Layer:
    def __init__(self, name=None, trainable=False, **kwargs):
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        self.name = name if name is not None else _auto_name(type(self).__name__)
        self.trainable = trainable
        self.built = False
        self._build_shape = None
        self._weights_version = 0   # bumped by set_weights(); fused plans that cached the old weights rebuild

    # keras semantics: build() runs once, the first time the layer sees an input shape.
    def build(self, input_shape):
        self.built = True
        self._build_shape = tuple(input_shape)

    def _maybe_build(self, input_shape):
        if not self.built:
            self.build(tuple(input_shape))
            self.built = True

    def call(self, inputs):
        raise NotImplementedError

    def __call__(self, inputs, *args, **kwargs):
        return self.call(inputs, *args, **kwargs)

    def get_config(self):
        return {"name": self.name, "trainable": self.trainable}

    @classmethod
    def from_config(cls, config):
        return cls(**config)

    def compute_output_shape(self, input_shape):
        return input_shape

    def get_weights(self):
        return []

    def set_weights(self, weights, fmt="kaldi"):
        if len(weights) != 0:
            raise ValueError(f"layer '{self.name}' has no weights, got {len(weights)}")
