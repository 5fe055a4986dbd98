"""Stand-in for the TensorFlow symbols the reference's hot-path code calls.

TEST INFRASTRUCTURE ONLY. TensorFlow is not installable in the build container, so this shim lets the golden
generators under tests/golden/ import and execute the UNMODIFIED reference code from /root/reference and record
golden vectors. It is never imported by the product package, smoke() or bench.py.

Two backends behind one module:
* NumPy for the ~15 ops the reference *environment* code calls (tf.fill/range/cast/equal/less/maximum/
  reduce_min/reduce_sum/reshape/tile/transpose/random.*): elementwise ops and reductions follow NumPy
  (identical to TF for these dtypes); `x / int` on an int32 "tensor" is true division to float64 (tf truediv),
  matching env.py:207-215.
* torch-CPU fp32 (autograd) for the model / agent code (agents/models/transformer/**, agents/agent.py,
  agents/trainer.py): tf.matmul/nn.softmax/nn.tanh/argmax/gather_nd/..., tf.GradientTape, and the Keras
  classes in `tensorflow.keras` (Layer/Model attribute tracking, Dense, LayerNormalization, Adam, losses).
  An op runs on torch as soon as one of its operands is a torch tensor; otherwise it stays on NumPy.
"""
import numpy as np
import torch

float32 = "float32"
int32 = "int32"
_TORCH_DT = {"float32": torch.float32, "int32": torch.int32, "int64": torch.int64, "float64": torch.float64}


class Tensor(np.ndarray):
    """ndarray subclass with .numpy(); int / int -> float64 like tf truediv."""

    def numpy(self):
        return np.asarray(self)


class TTensor(torch.Tensor):
    """torch tensor with the `.numpy()` of an eager tf.Tensor and NumPy operands accepted in arithmetic
    (the reference mixes env arrays and model outputs: `logits -= mask * 1e9`, `returns - values`)."""

    def numpy(self):
        return torch.Tensor.numpy(self.detach().as_subclass(torch.Tensor)).copy()

    def __array__(self, dtype=None, copy=None):
        a = torch.Tensor.numpy(self.detach().as_subclass(torch.Tensor))
        return a if dtype is None else a.astype(dtype)

    def __format__(self, spec):
        return format(self.detach().as_subclass(torch.Tensor).item(), spec) if self.dim() == 0 else repr(self)

    def mean(self, *args, axis=None, dtype=None, out=None, **kw):
        """np.mean(tensor) (trainer.py:160,162) dispatches here with NumPy's keyword names."""
        if args or kw:
            return torch.Tensor.mean(self, *args, **kw)
        return torch.Tensor.mean(self) if axis is None else torch.Tensor.mean(self, dim=axis)


def _lift(o):
    if isinstance(o, Variable):
        return o.t
    if isinstance(o, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(o)))
    if isinstance(o, np.generic):
        return o.item()
    return o


def _binary(name):
    base = getattr(torch.Tensor, name)

    def op(self, other):
        return base(self, _lift(other))
    op.__name__ = name
    return op


for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__",
           "__iadd__", "__isub__", "__imul__"):
    setattr(TTensor, _n, _binary(_n))


class Variable:
    """tf.Variable: a named fp32 leaf the tape can differentiate against."""

    def __init__(self, value, name="Variable", trainable=True):
        self.t = torch.as_tensor(np.asarray(value, dtype=np.float32)).clone().requires_grad_(True).as_subclass(TTensor)
        self.name = name
        self.trainable = trainable

    @property
    def shape(self):
        return tuple(self.t.shape)

    def numpy(self):
        return self.t.numpy()

    def assign(self, value):
        with torch.no_grad():
            self.t.as_subclass(torch.Tensor).copy_(torch.as_tensor(np.asarray(value, dtype=np.float32)))
        return self

    def __array__(self, dtype=None, copy=None):
        return self.t.__array__(dtype)


def _is_torch(*xs):
    return any(isinstance(x, (torch.Tensor, Variable)) for x in xs)


def _tt(x, dtype=None):
    """-> TTensor (fp32 unless the source is integer)."""
    if isinstance(x, Variable):
        x = x.t
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
        if x.dtype == torch.float64:
            x = x.to(torch.float32)
    if dtype is not None:
        x = x.to(_TORCH_DT[str(dtype)] if not isinstance(dtype, torch.dtype) else dtype)
    return x.as_subclass(TTensor)


def _t(x, dtype=None):
    return np.asarray(x, dtype=dtype).view(Tensor)


def convert_to_tensor(x, dtype=None):
    return _tt(x, dtype) if _is_torch(x) else _t(x, dtype)


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


def shape(x):
    return tuple(x.shape)


def cast(x, dtype):
    if _is_torch(x):
        return _tt(x, dtype)
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
    if _is_torch(x):
        x = _tt(x)
        return x.mean() if axis is None else x.mean(dim=axis)
    return _t(np.mean(np.asarray(x), axis=axis))


def reshape(x, shape):  # noqa: A002
    if _is_torch(x):
        return _tt(x).reshape(tuple(int(s) for s in shape))
    return _t(np.reshape(np.asarray(x), shape))


def transpose(x, perm=None):
    if _is_torch(x):
        x = _tt(x)
        return x.permute(*perm) if perm is not None else x.permute(*reversed(builtins_range(x.dim())))
    return _t(np.transpose(np.asarray(x), perm))


def tile(x, multiples):
    return _t(np.tile(np.asarray(x), multiples))


def concat(xs, axis=0):
    xs = list(xs)
    if _is_torch(*xs):
        return torch.cat([_tt(x) for x in xs], dim=axis).as_subclass(TTensor)
    return _t(np.concatenate([np.asarray(x) for x in xs], axis=axis))


def stack(xs, axis=0):
    return _t(np.stack([np.asarray(x) for x in xs], axis=axis))


def expand_dims(x, axis):
    if _is_torch(x):
        return _tt(x).unsqueeze(axis)
    return _t(np.expand_dims(np.asarray(x), axis))


def squeeze(x, axis=None):
    if _is_torch(x):
        return _tt(x).squeeze() if axis is None else _tt(x).squeeze(axis)
    return _t(np.squeeze(np.asarray(x), axis=axis))


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _tt(a), _tt(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return torch.matmul(a, b)


def argmax(x, axis=None, output_type="int64"):
    """tf.argmax: lowest index wins ties. Integer result -> NumPy tensor (no gradient flows through it)."""
    return _t(np.argmax(np.asarray(x), axis=axis).astype(str(output_type)))


def gather_nd(params, indices):
    idx = np.asarray(indices)
    out = np.asarray(params)[tuple(idx[..., i] for i in builtins_range(idx.shape[-1]))]
    return _t(out)


def sort(x, axis=-1, direction="ASCENDING"):
    s = np.sort(np.asarray(x), axis=axis)
    return _t(s if direction == "ASCENDING" else np.flip(s, axis=axis))


def stop_gradient(x):
    if isinstance(x, torch.Tensor):
        return x.detach()
    return x


def function(fn=None, **_kw):
    """@tf.function: graph compilation only, no numeric effect."""
    if fn is None:
        return lambda f: f
    return fn


class GradientTape:
    """Eager tape: torch autograd already recorded the graph; gradient() differentiates it."""

    def __init__(self, persistent=False):
        self.persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def gradient(self, target, sources):
        leaves = [s.t for s in sources]
        grads = torch.autograd.grad(target, leaves, allow_unused=True, retain_graph=True)
        return [None if g is None else g.as_subclass(TTensor) for g in grads]


import builtins as _builtins  # noqa: E402
builtins_range = _builtins.range


class _Math:
    @staticmethod
    def sqrt(x):
        if _is_torch(x):
            return torch.sqrt(_tt(x))
        return np.sqrt(x)

    @staticmethod
    def less(a, b):
        return less(a, b)


math = _Math()


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
        if _is_torch(x):
            return torch.softmax(_tt(x), dim=axis)
        x = np.asarray(x, dtype=np.float32)
        e = np.exp(x - x.max(axis=axis, keepdims=True))
        return _t(e / e.sum(axis=axis, keepdims=True))

    @staticmethod
    def tanh(x):
        return torch.tanh(_tt(x))

    @staticmethod
    def relu(x):
        return torch.relu(_tt(x))


nn = _NN()

from . import keras  # noqa: E402,F401
