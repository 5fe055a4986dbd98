"""Keras layer semantics (tensorflow==2.3.1, requirements.txt:4) restated on torch-CPU for the golden generators.

TEST INFRASTRUCTURE ONLY (see tensorflow/__init__.py). What is restated, and from where:
* Attribute tracking and `trainable_weights` order (keras/engine/base_layer.py `Layer.__setattr__`,
  `Layer.trainable_weights`, `_gather_children_attribute`; utils `filter_empty_layer_containers`): a layer's own
  variables come first (creation order inside `build`), then its tracked children depth-first in
  ATTRIBUTE-ASSIGNMENT order, a list attribute expanding in place; duplicates are dropped.
* Variables are created lazily at the first call (`build(input_shape)`), so a model has no weights before it
  has been called once.
* Dense: `x @ kernel + bias`, then the activation; kernel `glorot_uniform`, bias zeros.
* LayerNormalization(epsilon=1e-6): non-fused path (the fused kernel needs epsilon >= 1.001e-5):
  `mean, var = moments(x, -1)`; `inv = rsqrt(var + eps) * gamma`; `y = x * inv + (beta - mean * inv)`.
* TimeDistributed(Dense) == Dense on the last axis; Flatten is row-major; Concatenate(axis).
"""
import math

import numpy as np
import torch

import tensorflow as tf

_INIT_RNG = np.random.default_rng(20201)


def _activation(name):
    if name is None or name == "linear":
        return lambda x: x
    return {"relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}[name]


class Layer:
    def __init__(self, name=None, **_kw):
        object.__setattr__(self, "_children", [])
        object.__setattr__(self, "_own_weights", [])
        self.built = False
        self.trainable = True
        self.name = name or type(self).__name__.lower()

    # ---- tracking (base_layer.Layer.__setattr__ -> _maybe_create_attribute('_layers')) ----
    def __setattr__(self, key, value):
        tracked = getattr(self, "_children", None)
        if tracked is not None:
            if isinstance(value, Layer):
                if not any(value is c for c in tracked):
                    tracked.append(value)
            elif isinstance(value, (list, tuple)) and value and all(isinstance(v, Layer) for v in value):
                for v in value:
                    if not any(v is c for c in tracked):
                        tracked.append(v)
        object.__setattr__(self, key, value)

    def add_weight(self, name, shape, initializer):
        v = tf.Variable(initializer(shape), name=f"{self.name}/{name}:0")
        self._own_weights.append(v)
        return v

    @property
    def trainable_weights(self):
        out, seen = [], set()
        for w in list(self._own_weights) + [w for c in self._children for w in c.trainable_weights]:
            if id(w) not in seen:
                seen.add(id(w))
                out.append(w)
        return out

    weights = trainable_weights
    trainable_variables = trainable_weights

    def build(self, input_shape):
        pass

    def __call__(self, *args, **kwargs):
        if not self.built:
            first = args[0] if args else next(iter(kwargs.values()))
            self.build(tuple(first.shape) if hasattr(first, "shape") else None)
            self.built = True
        return self.call(*args, **kwargs)


def _glorot_uniform(shape):
    lim = math.sqrt(6.0 / (shape[0] + shape[1]))
    return _INIT_RNG.uniform(-lim, lim, size=shape).astype(np.float32)


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform", **kw):
        super().__init__(**kw)
        self.units = int(units)
        self.activation = _activation(activation)
        self.use_bias = use_bias
        self.kernel_initializer = kernel_initializer

    def build(self, input_shape):
        init = self.kernel_initializer
        fn = _glorot_uniform if isinstance(init, str) else init
        self.kernel = self.add_weight("kernel", (int(input_shape[-1]), self.units), fn)
        if self.use_bias:
            self.bias = self.add_weight("bias", (self.units,), lambda s: np.zeros(s, np.float32))

    def call(self, inputs):
        y = torch.matmul(tf._tt(inputs), self.kernel.t)
        if self.use_bias:
            y = y + self.bias.t
        return self.activation(y)


class TimeDistributed(Layer):
    """keras/layers/wrappers.py: applies `layer` to every temporal slice; for Dense that is Dense on the last axis."""

    def __init__(self, layer, **kw):
        super().__init__(**kw)
        self.layer = layer

    def call(self, inputs):
        return self.layer(inputs)


class LayerNormalization(Layer):
    def __init__(self, axis=-1, epsilon=1e-3, **kw):
        super().__init__(**kw)
        assert axis == -1
        self.epsilon = epsilon

    def build(self, input_shape):
        d = int(input_shape[-1])
        self.gamma = self.add_weight("gamma", (d,), lambda s: np.ones(s, np.float32))
        self.beta = self.add_weight("beta", (d,), lambda s: np.zeros(s, np.float32))

    def call(self, inputs):
        x = tf._tt(inputs)
        mean = x.mean(dim=-1, keepdim=True)
        var = ((x - mean.detach()) ** 2).mean(dim=-1, keepdim=True)   # nn.moments: squared_difference(x, stop_gradient(mean))
        inv = torch.rsqrt(var + self.epsilon) * self.gamma.t           # nn.batch_normalization
        return x * inv + (self.beta.t - mean * inv)


class Concatenate(Layer):
    def __init__(self, axis=-1, **kw):
        super().__init__(**kw)
        self.axis = axis

    def call(self, inputs):
        return torch.cat([tf._tt(x) for x in inputs], dim=self.axis).as_subclass(tf.TTensor)


class Flatten(Layer):
    def call(self, inputs):
        x = tf._tt(inputs)
        return x.reshape(x.shape[0], -1)


class Softmax(Layer):
    def __init__(self, axis=-1, **kw):
        super().__init__(**kw)
        self.axis = axis

    def call(self, inputs):
        return torch.softmax(tf._tt(inputs), dim=self.axis)
