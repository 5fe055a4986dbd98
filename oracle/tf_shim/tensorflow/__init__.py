"""NumPy-backed stand-in for the handful of TensorFlow ops the reference *environment* code calls
(tf.fill/range/cast/equal/less/maximum/reduce_min/reduce_sum/reshape/tile/transpose/random.*).

TEST INFRASTRUCTURE ONLY. TensorFlow is not installable in the build container, so this shim lets
`tests/golden/make_env_golden.py` import and execute the UNMODIFIED reference environment / reward code from
/root/reference and record golden vectors. It is never imported by the product package, smoke() or bench.py.
Semantics restated: elementwise ops and reductions follow NumPy (identical to TF for these dtypes);
`x / int` on an int32 "tensor" is true division to float64 (tf truediv), matching env.py:207-215.
"""
import numpy as np

float32 = "float32"
int32 = "int32"


class Tensor(np.ndarray):
    """ndarray subclass with .numpy(); int / int -> float64 like tf truediv."""

    def numpy(self):
        return np.asarray(self)


def _t(x, dtype=None):
    return np.asarray(x, dtype=dtype).view(Tensor)


def convert_to_tensor(x, dtype=None):
    return _t(x, dtype)


def constant(x, dtype=None):
    return _t(x, dtype)


def fill(dims, value):
    if np.isscalar(dims):
        dims = (int(dims),)
    return _t(np.full(tuple(int(d) for d in np.atleast_1d(dims)), value))


def range(*args, dtype="int32"):  # noqa: A001
    return _t(np.arange(*args, dtype=dtype))


def zeros(shape, dtype="float32"):
    return _t(np.zeros(shape, dtype=dtype))


def cast(x, dtype):
    return _t(np.asarray(x).astype(dtype))


def equal(a, b):
    return _t(np.equal(np.asarray(a), np.asarray(b)))


def less(a, b):
    return _t(np.less(np.asarray(a), np.asarray(b)))


def greater(a, b):
    return _t(np.greater(np.asarray(a), np.asarray(b)))


def maximum(a, b):
    return _t(np.maximum(np.asarray(a), np.asarray(b)))


def reduce_min(x, axis=None):
    return _t(np.min(np.asarray(x), axis=axis))


def reduce_sum(x, axis=None):
    return _t(np.sum(np.asarray(x), axis=axis))


def reduce_mean(x, axis=None):
    return _t(np.mean(np.asarray(x), axis=axis))


def reshape(x, shape):
    return _t(np.reshape(np.asarray(x), shape))


def transpose(x, perm=None):
    return _t(np.transpose(np.asarray(x), perm))


def tile(x, multiples):
    return _t(np.tile(np.asarray(x), multiples))


def concat(xs, axis=0):
    return _t(np.concatenate([np.asarray(x) for x in xs], axis=axis))


def sort(x, axis=-1, direction="ASCENDING"):
    s = np.sort(np.asarray(x), axis=axis)
    return _t(s if direction == "ASCENDING" else np.flip(s, axis=axis))


def stop_gradient(x):
    return x


class _Random:
    def __init__(self):
        self._rng = np.random.default_rng(1234)

    def set_seed(self, seed):
        self._rng = np.random.default_rng(seed)

    def uniform(self, shape, minval=0, maxval=None, dtype="float32", seed=None):
        if "int" in str(dtype):
            return _t(self._rng.integers(minval, maxval, size=shape, dtype=np.int64).astype(np.int32))
        return _t(self._rng.uniform(minval, 1.0 if maxval is None else maxval, size=shape).astype(dtype))

    def shuffle(self, x, seed=None):
        x = np.asarray(x)
        return _t(x[self._rng.permutation(x.shape[0])])


random = _Random()


class _NN:
    @staticmethod
    def softmax(x, axis=-1):
        x = np.asarray(x, dtype=np.float32)
        e = np.exp(x - x.max(axis=axis, keepdims=True))
        return _t(e / e.sum(axis=axis, keepdims=True))


nn = _NN()
