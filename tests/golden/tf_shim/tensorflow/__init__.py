"""Minimal torch-backed stand-in for the `tensorflow` names the reference's DPPO model code touches.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_ref_golden.py, never by the product, the GPU
tests, smoke() or bench.py).  TensorFlow / Keras cannot be installed in the build image, so this module
lets the reference's OWN Python (model/diffusion/{diffusion,diffusion_vpg,diffusion_ppo,mlp_diffusion,
modules,sampling}.py, model/common/{mlp,critic}.py, imported from /root/reference) execute: the control
flow, clipping order, schedule arithmetic, indexing and loss composition are then the reference's, and
only the primitive ops below are restated from their published TF / Keras 3 semantics on fp32 torch
CPU tensors.  Where TF's gradient convention differs from torch's (tf.maximum ties) TF's is used.
Only what that code path needs is implemented; everything else raises AttributeError.
"""
import builtins as _b
import math as _math
import sys as _sys
import types as _types

import numpy as _np
import torch as _torch

float32 = _torch.float32
float64 = _torch.float64
int32 = _torch.int32
int64 = _torch.int64
bool = _torch.bool  # noqa: A001
newaxis = None
Tensor = _torch.Tensor

def _dims(shape):
    """TF accepts tuples, lists, TensorShape, int tensors and scalars as shapes."""
    if isinstance(shape, _torch.Tensor):
        return [int(v) for v in shape.reshape(-1).tolist()]
    if isinstance(shape, (int, _np.integer)):
        return [int(shape)]
    return [int(v) for v in shape]


def _t(x, dtype=None):
    if isinstance(x, _torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, _np.ndarray):
        return _torch.from_numpy(_np.ascontiguousarray(x)).to(dtype) if dtype is not None else _torch.from_numpy(
            _np.ascontiguousarray(x))
    if dtype is None:
        dtype = _torch.float32 if isinstance(x, float) or (
            isinstance(x, (list, tuple)) and any(isinstance(v, float) for v in x)) else None
    return _torch.tensor(x, dtype=dtype)


# ---- decorators / plumbing ------------------------------------------------------------------------
def function(fn=None, **_kw):
    """tf.function: graph tracing has no numerical effect; run eagerly."""
    if fn is None:
        return lambda f: f
    return fn


def identity(x):
    return x


def convert_to_tensor(x, dtype=None):
    return _t(x, dtype)


def constant(x, dtype=None):
    return _t(x, dtype)


def cast(x, dtype):
    return _t(x).to(dtype)


def shape(x):
    return _torch.tensor(list(x.shape), dtype=_torch.int64)


# ---- creation -------------------------------------------------------------------------------------
def zeros(shape_, dtype=float32):
    return _torch.zeros(_dims(shape_), dtype=dtype)


def zeros_like(x):
    return _torch.zeros_like(x)


def ones_like(x):
    return _torch.ones_like(x)


def fill(dims, value):
    if isinstance(value, _torch.Tensor):
        value = value.item()
    return _torch.full(_dims(dims), value)


def range(*args, dtype=None, **kw):  # noqa: A001
    if "start" in kw or "limit" in kw:
        args = (kw.get("start", 0), kw["limit"], kw.get("delta", 1))
    out = _torch.arange(*args)
    return out if dtype is None else out.to(dtype)


def linspace(a, b, n):
    return _torch.linspace(a, b, int(n), dtype=_torch.float32)


# ---- shape ops ------------------------------------------------------------------------------------
def reshape(x, shape_):
    return _torch.reshape(x, _dims(shape_))


def concat(values, axis):
    return _torch.cat(list(values), dim=axis)


def stack(values, axis=0):
    return _torch.stack(list(values), dim=axis)


def tile(x, multiples):
    return x.repeat(*_dims(multiples))


def expand_dims(x, axis):
    return _torch.unsqueeze(x, axis)


def squeeze(x, axis=None):
    return _torch.squeeze(x) if axis is None else _torch.squeeze(x, axis)


def transpose(x, perm=None):
    return x.t() if perm is None else x.permute(*perm)


def reverse(x, axis):
    return _torch.flip(x, dims=list(axis))


def gather(params, indices, axis=0):
    idx = _t(indices).long()
    return _torch.index_select(params, axis, idx.reshape(-1)).reshape(
        list(params.shape[:axis % params.ndim]) + list(idx.shape) + list(params.shape[axis % params.ndim + 1:]))


def where(cond, x, y):
    return _torch.where(_t(cond), x, y)


def split(value, num_or_size_splits, axis=0):
    """tf.split with an integer: that many equal parts (the reference always divides evenly)."""
    n = int(num_or_size_splits)
    assert value.shape[axis] % n == 0
    return list(_torch.chunk(value, n, dim=axis))


def clip_by_norm(t, clip_norm):
    l2sum = (t * t).sum()
    l2norm = _torch.sqrt(_torch.where(l2sum > 0, l2sum, _torch.ones_like(l2sum)))
    return t * clip_norm / _torch.maximum(l2norm, _torch.tensor(float(clip_norm)))


class GradientTape:
    """tf.GradientTape over torch autograd.  The generator runs with autograd globally OFF (TF records nothing outside a tape);
    entering a tape switches it on, `gradient` differentiates the scalar loss wrt the given variables."""

    def __enter__(self):
        self._prev = _torch.is_grad_enabled()
        _torch.set_grad_enabled(True)
        return self

    def __exit__(self, *exc):
        _torch.set_grad_enabled(self._prev)
        return False

    def watch(self, x):
        pass

    def gradient(self, target, sources):
        return list(_torch.autograd.grad(target, list(sources), allow_unused=True))


def unravel_index(indices, dims):
    """tf.unravel_index: row-major; returns the coordinate arrays stacked on axis 0 (unpackable like a tuple)."""
    idx = _t(indices).long()
    out, rem = [], idx
    for d in reversed(_dims(dims)):
        out.append(rem % d)
        rem = rem // d
    return _torch.stack(out[::-1], dim=0)


def gather_nd(params, indices):
    """tf.gather_nd for index tensors [N, r]: params[i0, i1, ...] per row."""
    idx = _t(indices).long()
    return params[tuple(idx[:, j] for j in _b.range(idx.shape[1]))]


# ---- math -----------------------------------------------------------------------------------------
def sqrt(x):
    return _torch.sqrt(_t(x))


def exp(x):
    return _torch.exp(_t(x))


def square(x):
    return x * x


def abs(x):  # noqa: A001
    return _torch.abs(x)


def sin(x):
    return _torch.sin(x)


def cos(x):
    return _torch.cos(x)


def pow(x, y):  # noqa: A001
    return _torch.pow(_t(x, _torch.float32) if not isinstance(x, _torch.Tensor) else x, y)


def maximum(x, y):
    """tf.maximum; its registered gradient routes a tie to x (xmask = x >= y)."""
    return _torch.where(x >= y, x, y)


def clip_by_value(x, clip_value_min, clip_value_max):
    """tf.clip_by_value = minimum(maximum(x, lo), hi); gradient passes where lo <= x <= hi."""
    return _torch.clamp(x, min=clip_value_min, max=clip_value_max)


def reduce_mean(x, axis=None):
    return x.mean() if axis is None else x.mean(dim=axis)


def reduce_sum(x, axis=None):
    return x.sum() if axis is None else x.sum(dim=axis)


def reduce_prod(x, axis=None):
    return _t(list(x) if not isinstance(x, _torch.Tensor) else x).prod()


class _Math(_types.ModuleType):
    @staticmethod
    def cumprod(x, axis=0):
        return _torch.cumprod(x, dim=axis)

    @staticmethod
    def log(x):
        return _torch.log(_t(x))

    @staticmethod
    def reduce_std(x, axis=None):
        """tf.math.reduce_std: population standard deviation."""
        return x.std(unbiased=False) if axis is None else x.std(dim=axis, unbiased=False)


math = _Math("tensorflow.math")


# ---- random: draws come from an injectable source so the caller can record them ----------------------
class _Random(_types.ModuleType):
    source = None      # callable(kind, shape, **kw) -> tensor; installed by the golden generator (tf.random.source = ...)

    def normal(self, shape_, mean=0.0, stddev=1.0, dtype=float32):
        return self.source("normal", _dims(shape_)) * stddev + mean

    def uniform(self, shape_, minval=0, maxval=None, dtype=float32):
        return self.source("uniform", _dims(shape_), minval=minval, maxval=maxval, dtype=dtype)

    def shuffle(self, value):
        """tf.random.shuffle along axis 0: the permutation comes from the injected source."""
        perm = self.source("shuffle", [int(value.shape[0])])
        return value[perm.long()]


random = _Random("tensorflow.random")


# ---- Keras ----------------------------------------------------------------------------------------
def _walk_variables(obj, seen, out):
    """Variables in Keras' tracking order: attributes in assignment order, lists and sub-layers recursively."""
    if id(obj) in seen:
        return
    seen.add(id(obj))
    if isinstance(obj, _torch.Tensor):
        if getattr(obj, "_is_variable", False):
            out.append(obj)
        return
    if isinstance(obj, (list, tuple)):
        for v in obj:
            _walk_variables(v, seen, out)
        return
    if isinstance(obj, Layer):
        for v in obj.__dict__.values():
            _walk_variables(v, seen, out)


class Layer:
    def __init__(self, *a, **kw):
        pass

    def __call__(self, *a, **kw):
        return self.call(*a, **kw)

    @property
    def variables(self):
        out = []
        _walk_variables(self, set(), out)
        return out

    @property
    def trainable_variables(self):
        return [v for v in self.variables if getattr(v, "trainable", True)]

    def get_weights(self):
        return [v.detach().numpy().copy() for v in self.variables]

    def set_weights(self, weights):
        vs = self.variables
        assert len(vs) == len(weights), (len(vs), len(weights))
        for v, w in zip(vs, weights):
            w = _t(_np.asarray(w), _torch.float32)
            assert tuple(v.shape) == tuple(w.shape), (tuple(v.shape), tuple(w.shape))
            v.data.copy_(w)

    def load_weights(self, path):
        """Stands in for Keras' checkpoint reader: an .npz holding arr_0.. in `variables` order."""
        z = _np.load(path)
        self.set_weights([z[f"arr_{i}"] for i in _b.range(len(z.files))])


class Model(Layer):
    pass


def _variable(t, trainable=True):
    t = t.detach().clone().requires_grad_(True)
    t._is_variable = True
    t.trainable = trainable
    return t


def Variable(initial_value, trainable=True, dtype=None, **_kw):
    return _variable(_t(initial_value, dtype), trainable)


def _relu(x):
    return _torch.clamp(x, min=0)          # keras.activations.relu: max(x, 0), gradient 0 at x == 0


def _mish(x):
    return x * _torch.tanh(_torch.nn.functional.softplus(x))   # keras.activations.mish


_ACT = {None: None, "linear": None, "relu": _relu, "mish": _mish, "tanh": _torch.tanh}


class Dense(Layer):
    """keras.layers.Dense: y = act(x @ kernel[in, out] + bias); built at first call (glorot_uniform, zeros)."""

    def __init__(self, units, activation=None, input_shape=None, **_kw):
        self.units = int(units)
        self.activation = _ACT[activation] if (activation is None or isinstance(activation, str)) else activation
        self.kernel = None
        self.bias = None

    def call(self, x):
        if self.kernel is None:
            fan_in = int(x.shape[-1])
            lim = _math.sqrt(6.0 / (fan_in + self.units))
            self.kernel = _variable((_torch.rand(fan_in, self.units) * 2 - 1) * lim)
            self.bias = _variable(_torch.zeros(self.units))
        y = _torch.matmul(x, self.kernel) + self.bias
        return y if self.activation is None else self.activation(y)


class Sequential(Model):
    def __init__(self, layers=None):
        self.layers = list(layers or [])

    def call(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


class _Unsupported(Layer):
    def __init__(self, *a, **kw):
        raise NotImplementedError(f"{type(self).__name__} is outside the DPPO MLP path this shim covers")


def _ns(name, **members):
    m = _types.ModuleType(name)
    m.__dict__.update(members)
    return m


class _Activation(Layer):
    def __init__(self, name):
        self.fn = _ACT[name]

    def call(self, x):
        return x if self.fn is None else self.fn(x)


def _unsupported(name):
    return type(name, (_Unsupported,), {})


def _no_fn(name):
    def f(*a, **kw):
        raise NotImplementedError(f"{name} is outside the DPPO MLP path this shim covers")
    return f


keras = _ns(
    "tensorflow.keras",
    Model=Model,
    Sequential=Sequential,
    layers=_ns(
        "tensorflow.keras.layers", Layer=Layer, Dense=Dense, Activation=_Activation,
        ReLU=lambda: _Activation("relu"),
        LayerNormalization=_unsupported("LayerNormalization"), Dropout=_unsupported("Dropout"),
        GroupNormalization=_unsupported("GroupNormalization"), Conv1D=_unsupported("Conv1D"),
        Conv1DTranspose=_unsupported("Conv1DTranspose"), Conv2D=_unsupported("Conv2D"),
        Softplus=_unsupported("Softplus"), GELU=_unsupported("GELU"), ELU=_unsupported("ELU"),
    ),
    activations=_ns(
        "tensorflow.keras.activations", relu=_relu, mish=_mish, tanh=_torch.tanh, linear=lambda x: x,
        elu=_no_fn("elu"), gelu=_no_fn("gelu"), softplus=_no_fn("softplus"),
    ),
    losses=_ns("tensorflow.keras.losses", MSE=lambda a, b: ((a - b) ** 2).mean(dim=-1)),
    optimizers=_ns("tensorflow.keras.optimizers",
                   schedules=_ns("tensorflow.keras.optimizers.schedules", LearningRateSchedule=type("LearningRateSchedule", (), {}))),
)

for _m in (math, random, keras, keras.layers, keras.activations, keras.losses):
    _sys.modules[_m.__name__] = _m
