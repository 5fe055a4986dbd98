"""tf.keras surface used by agents/** (TEST INFRASTRUCTURE ONLY; see layers.py for the restated semantics)."""
import numpy as np
import torch

import tensorflow as tf
from . import layers  # noqa: F401
from .layers import Layer


class Model(Layer):
    """keras.Model: a Layer whose weights can be saved; `trainable_weights` = children in attribute order.
    save_weights / load_weights write the variables in `trainable_weights` order to `<location>.npz` (the real
    class writes a TF checkpoint; only the round trip inside the reference process matters to the goldens)."""

    def save_weights(self, location):
        np.savez(location + ".npz", *[w.numpy() for w in self.trainable_weights])

    def load_weights(self, location):
        with np.load(location + ".npz") as z:
            for w, k in zip(self.trainable_weights, z.files):
                w.assign(z[k])


class Sequential(Model):
    def __init__(self, layers=None, **kw):  # noqa: A002
        super().__init__(**kw)
        self.layers = list(layers or [])

    def call(self, inputs):
        x = inputs
        for layer in self.layers:
            x = layer(x)
        return x


class initializers:
    class RandomUniform:
        def __init__(self, minval=-0.05, maxval=0.05, seed=None):
            self.minval, self.maxval = minval, maxval
            self._rng = np.random.default_rng(seed if seed is not None else 777)

        def __call__(self, shape, dtype=None):
            return self._rng.uniform(self.minval, self.maxval, size=shape).astype(np.float32)


class losses:
    """keras/losses.py. Reduction.NONE returns per-sample losses; the default reduces with a mean."""

    class Reduction:
        NONE = "none"
        AUTO = "auto"
        SUM_OVER_BATCH_SIZE = "sum_over_batch_size"

    @staticmethod
    def mean_squared_error(y_true, y_pred):
        """K.mean(squared_difference(y_pred, y_true), axis=-1)."""
        y_pred = tf._tt(y_pred)
        y_true = tf._tt(y_true).to(y_pred.dtype)
        return ((y_pred - y_true) ** 2).mean(dim=-1)

    class _Loss:
        def __init__(self, from_logits=False, reduction="auto", **_kw):
            self.from_logits, self.reduction = from_logits, reduction

        def __call__(self, y_true, y_pred):
            per = self.fn(y_true, tf._tt(y_pred))
            return per if self.reduction == "none" else per.mean()

    class MeanSquaredError(_Loss):
        def fn(self, y_true, y_pred):
            return losses.mean_squared_error(y_true, y_pred)

    class SparseCategoricalCrossentropy(_Loss):
        """nn.sparse_softmax_cross_entropy_with_logits: -log_softmax(logits)[label]."""

        def fn(self, y_true, y_pred):
            assert self.from_logits
            lab = torch.as_tensor(np.asarray(y_true)).long().reshape(-1, 1)
            return -torch.log_softmax(y_pred, dim=-1).gather(1, lab).squeeze(1)

    class CategoricalCrossentropy(_Loss):
        """nn.softmax_cross_entropy_with_logits: -sum(labels * log_softmax(logits))."""

        def fn(self, y_true, y_pred):
            assert self.from_logits
            lab = tf._tt(y_true).to(y_pred.dtype)
            return -(lab * torch.log_softmax(y_pred, dim=-1)).sum(dim=-1)


class optimizers:
    class Adam:
        """keras/optimizer_v2/adam.py (non-amsgrad, dense): lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);
        m += (g - m)(1 - b1); v += (g^2 - v)(1 - b2); var -= lr_t * m / (sqrt(v) + eps), eps = 1e-7.
        `clipnorm` is tf.clip_by_norm per variable. Whether a bare `apply_gradients` applies it depends on the
        TensorFlow version (minimize()-only in 2.3, apply_gradients from 2.4); the class
        attribute CLIP_IN_APPLY_GRADIENTS selects which behaviour the golden generator records."""
        CLIP_IN_APPLY_GRADIENTS = True

        def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, clipnorm=None, **_kw):
            self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
            self.clipnorm = clipnorm
            self.iterations = 0
            self._slots = {}

        def apply_gradients(self, grads_and_vars):
            gv = [(g, v) for g, v in grads_and_vars if g is not None]      # _filter_grads
            self.iterations += 1
            t = self.iterations
            b1, b2 = np.float32(self.beta_1), np.float32(self.beta_2)
            lr_t = np.float32(self.learning_rate * np.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t))
            for g, var in gv:
                g = np.asarray(g, dtype=np.float32)
                if self.clipnorm is not None and self.CLIP_IN_APPLY_GRADIENTS:
                    norm = np.sqrt(np.sum(g * g, dtype=np.float32))
                    if norm > self.clipnorm:                                  # clip_by_norm: g * clip / max(norm, clip)
                        g = g * np.float32(self.clipnorm) / norm
                m, v = self._slots.setdefault(id(var), [np.zeros_like(g), np.zeros_like(g)])
                m += (g - m) * (np.float32(1) - b1)
                v += (g * g - v) * (np.float32(1) - b2)
                var.assign(var.numpy() - lr_t * m / (np.sqrt(v) + np.float32(self.epsilon)))
