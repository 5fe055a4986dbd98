"""Stand-in for the single tfp call in the reference env (`sample_action`): Categorical(probs).sample()."""
import numpy as np


class _Categorical:
    _rng = np.random.default_rng(99)

    def __init__(self, probs=None):
        self.probs = np.asarray(probs, dtype=np.float64)

    def sample(self):
        p = self.probs / self.probs.sum(axis=-1, keepdims=True)
        c = np.cumsum(p, axis=-1)
        u = self._rng.random(p.shape[:-1] + (1,))
        import tensorflow as tf
        return tf.convert_to_tensor(np.minimum((u > c).sum(axis=-1), p.shape[-1] - 1).astype(np.int32))


class distributions:
    Categorical = _Categorical
