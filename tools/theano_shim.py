"""An eager, NumPy-backed stand-in for the parts of Theano 0.8.2 that `sloika/layers.py`, `sloika/conv.py`,
`sloika/activation.py`, `sloika/config.py` and `sloika/module_tools.py` use.  TEST INFRASTRUCTURE ONLY.

Purpose: Theano is not installable in this image, so the reference's forward pass cannot be run as shipped.
`install()` registers this module under the names `theano`, `theano.tensor`, `theano.tensor.nnet`,
`theano.tensor.signal.pool`, ... in `sys.modules`; after that the reference's own source files import
UNMODIFIED from /root/reference and `network.run(x)` executes their code line by line on concrete float32
arrays (there is no graph: every `T.xxx` call computes at once).  Gate order, the `reshape((-1, 2, size))`
of `Gru.step`, `filter_flip=False`, `pad_first`, `bf1t` / `tbf`, the `scan` recursion and the Lstm column
interleave therefore come from the reference; only the documented semantics of the primitives below come from
here.  Used by `tools/make_golden_forward.py` to write `tests/golden/forward_cases.npz`.

Primitive semantics restated (Theano 0.8.2 documentation / source, un-vendored):
  shared(value)            container with get_value / set_value; takes part in arithmetic as its value
  tensor.tensordot         numpy.tensordot (BLAS, float32 accumulate order differs from Theano's BLAS by rounding)
  tensor.nnet.sigmoid      scalar op `x < -88 ? 0 : x > 15 ? 1 : 1 / (1 + exp(-x))` for float32 (sigm.py c_code)
  tensor.nnet.conv2d       input [b, c, rows, cols], filters [o, c, frows, fcols], border_mode 'valid',
                           subsample = output step, filter_flip=False -> cross-correlation
  scan(fn, sequences, outputs_info)   out_t = fn(seq[t], out_{t-1}), out_{-1} = outputs_info; returns (stack, {})
  map(fn, sequences)       stack of fn(seq[t])
  x.flatten(ndim)          keep the first ndim-1 axes, collapse the rest
  tensor.shape_padleft / shape_padright / shape_padaxis / repeat / concatenate / switch / ...: as documented
"""
import sys
import types

import numpy as np

floatX = 'float32'


class Tensor(np.ndarray):
    """ndarray with Theano's `flatten(ndim)` and `dimshuffle`; everything else is NumPy."""

    def flatten(self, ndim=1):
        arr = np.asarray(self)
        if ndim == 1:
            return _wrap(arr.reshape(-1))
        return _wrap(arr.reshape(arr.shape[:ndim - 1] + (-1,)))

    def dimshuffle(self, *pattern):
        if len(pattern) == 1 and isinstance(pattern[0], (tuple, list)):
            pattern = tuple(pattern[0])
        arr = np.asarray(self)
        keep = [p for p in pattern if p != 'x']
        arr = arr.transpose(keep)
        index = tuple(None if p == 'x' else slice(None) for p in pattern)
        return _wrap(arr[index])


def _wrap(a):
    return np.asarray(a).view(Tensor)


def _val(a):
    if isinstance(a, SharedVariable):
        return a.value
    return a


class SharedVariable(object):
    """`theano.shared`: a mutable container that behaves as its current value inside expressions."""
    __array_priority__ = 100.0

    def __init__(self, value, name=None):
        value = np.array(value)
        if value.dtype.kind == 'f':
            # the reference builds its initial values as float32_array / np.sqrt(python int): float32 under the
            # NumPy 1.x value-based casting it was written for, float64 under NumPy 2 -- keep floatX
            value = value.astype(floatX)
        self._value = value
        self.name = name

    @property
    def value(self):
        if '_value' in self.__dict__:
            return self._value
        return self.container.storage[0]          # unpickled Theano shared variable (models/pretrained.pkl)

    def get_value(self, borrow=False):
        return self.value if borrow else self.value.copy()

    def set_value(self, value, borrow=False):
        self._value = np.array(value, dtype=self.value.dtype)

    def __array__(self, dtype=None, copy=None):
        return self.value if dtype is None else self.value.astype(dtype)

    @property
    def shape(self):
        return self.value.shape

    @property
    def dtype(self):
        return self.value.dtype

    def __getitem__(self, idx):
        return _wrap(self.value[idx])

    def __add__(self, o):
        return _wrap(self.value + _val(o))

    def __radd__(self, o):
        return _wrap(_val(o) + self.value)

    def __sub__(self, o):
        return _wrap(self.value - _val(o))

    def __rsub__(self, o):
        return _wrap(_val(o) - self.value)

    def __mul__(self, o):
        return _wrap(self.value * _val(o))

    def __rmul__(self, o):
        return _wrap(_val(o) * self.value)

    def __truediv__(self, o):
        return _wrap(self.value / _val(o))

    def __neg__(self):
        return _wrap(-self.value)

    def transpose(self, *axes):
        return _wrap(self.value.transpose(*axes))

    def reshape(self, *shape):
        return _wrap(self.value.reshape(*shape))


def shared(value, name=None, **kwargs):
    return SharedVariable(value, name)


# ---- theano.tensor ----------------------------------------------------------------------------------
def _f(fn):
    def wrapped(*args, **kwargs):
        return _wrap(fn(*[_val(a) for a in args], **kwargs))
    wrapped.__name__ = fn.__name__
    return wrapped


def _as32(x):
    """Theano keeps floatX through every op; NumPy upcasts float32 with Python/NumPy float64 scalars only in
    a few places (np.float64 scalars).  Results of float32 inputs stay float32 here."""
    return x


def tensordot(a, b, axes=2):
    return _wrap(np.tensordot(_val(a), _val(b), axes=axes))


def shape(x):
    return np.shape(_val(x))


def zeros(shp, dtype=None):
    return _wrap(np.zeros(tuple(int(s) for s in shp), dtype=dtype or floatX))


def ones(shp, dtype=None):
    return _wrap(np.ones(tuple(int(s) for s in shp), dtype=dtype or floatX))


def constant(x, dtype=None):
    return _wrap(np.asarray(x, dtype=dtype or floatX))


def concatenate(tensors, axis=0):
    return _wrap(np.concatenate([np.asarray(_val(t)) for t in tensors], axis=axis))


def repeat(x, repeats, axis=None):
    return _wrap(np.repeat(_val(x), repeats, axis=axis))


def shape_padleft(x, n_ones=1):
    x = np.asarray(_val(x))
    return _wrap(x.reshape((1,) * n_ones + x.shape))


def shape_padright(x, n_ones=1):
    x = np.asarray(_val(x))
    return _wrap(x.reshape(x.shape + (1,) * n_ones))


def shape_padaxis(x, axis):
    return _wrap(np.expand_dims(np.asarray(_val(x)), axis))


def switch(cond, a, b):
    a, b = _val(a), _val(b)
    with np.errstate(over='ignore', invalid='ignore'):
        return _wrap(np.where(_val(cond), a, b))


def _reduce(fn):
    def wrapped(x, axis=None, keepdims=False):
        if isinstance(axis, list):
            axis = tuple(axis)
        return _wrap(fn(_val(x), axis=axis, keepdims=keepdims))
    return wrapped


def expm1(x):
    with np.errstate(over='ignore'):
        return _wrap(np.expm1(_val(x)))


def exp(x):
    with np.errstate(over='ignore'):
        return _wrap(np.exp(_val(x)))


def sigmoid(x):
    """Theano 0.8.2 `tensor/nnet/sigm.py` ScalarSigmoid.c_code for float32:
    `x < -88.0f ? 0.0f : x > 15.0f ? 1.0f : 1.0f / (1.0f + exp(-x))`; float64 uses (-709, 19)."""
    x = np.asarray(_val(x))
    lo, hi = (-88.0, 15.0) if x.dtype == np.float32 else (-709.0, 19.0)
    one = x.dtype.type(1)
    with np.errstate(over='ignore'):
        y = one / (one + np.exp(-np.clip(x, lo, hi)))
    y = np.where(x < lo, x.dtype.type(0), y)
    return _wrap(np.where(x > hi, one, y).astype(x.dtype))


def relu(x, alpha=0):
    x = np.asarray(_val(x))
    return _wrap(np.where(x > 0, x, alpha * x).astype(x.dtype))


def softmax(x):
    x = np.asarray(_val(x))
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return _wrap(e / e.sum(axis=-1, keepdims=True))


def conv2d(input, filters, input_shape=None, filter_shape=None, border_mode='valid', subsample=(1, 1),
           filter_flip=True, **kwargs):
    """`theano.tensor.nnet.conv2d`: input [batch, in channels, rows, cols], filters [out channels, in channels,
    filter rows, filter cols]; 'valid' positions only, every subsample-th of them; with filter_flip the filter
    is mirrored in both axes (true convolution), without it this is a cross-correlation."""
    assert border_mode == 'valid'
    x = np.asarray(_val(input))
    w = np.asarray(_val(filters))
    if filter_flip:
        w = w[:, :, ::-1, ::-1]
    nb, nc, rows, cols = x.shape
    no, nc2, fr, fc = w.shape
    assert nc == nc2
    orow = (rows - fr) // subsample[0] + 1
    ocol = (cols - fc) // subsample[1] + 1
    out = np.zeros((nb, no, max(orow, 0), max(ocol, 0)), dtype=x.dtype)
    for i in range(fr):
        for j in range(fc):
            # patch[b, c, r, t] = x[b, c, r * s0 + i, t * s1 + j]
            patch = x[:, :, i:i + (orow - 1) * subsample[0] + 1:subsample[0],
                      j:j + (ocol - 1) * subsample[1] + 1:subsample[1]]
            out += np.einsum('bcrt,oc->bort', patch, w[:, :, i, j]).astype(x.dtype)
    return _wrap(out)


def scan(fn, sequences=None, outputs_info=None, non_sequences=None, **kwargs):
    seq = np.asarray(_val(sequences))
    state = _val(outputs_info)
    outs = []
    for t in range(seq.shape[0]):
        state = np.asarray(fn(_wrap(seq[t]), _wrap(state)))
        outs.append(state)
    if outs:
        res = np.stack(outs, axis=0)
    else:
        res = np.zeros((0,) + np.shape(state), dtype=seq.dtype)
    return _wrap(res), {}


def map_(fn, sequences=None, **kwargs):
    seq = np.asarray(_val(sequences))
    return _wrap(np.stack([np.asarray(fn(_wrap(seq[t]))) for t in range(seq.shape[0])], axis=0)), {}


class _Config(object):
    floatX = floatX


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    return mod


def install():
    """Register the shim as `theano` (and sub-modules) in sys.modules.  Idempotent."""
    if 'theano' in sys.modules and getattr(sys.modules['theano'], '__sloika_shim__', False):
        return sys.modules['theano']
    nnet = _module('theano.tensor.nnet', sigmoid=sigmoid, relu=relu, softmax=softmax, conv2d=conv2d)
    pool = _module('theano.tensor.signal.pool')
    signal = _module('theano.tensor.signal', pool=pool)
    sharedvar = _module('theano.tensor.sharedvar', TensorSharedVariable=SharedVariable)
    tensor = _module(
        'theano.tensor', nnet=nnet, signal=signal, sharedvar=sharedvar,
        tensordot=tensordot, shape=shape, zeros=zeros, ones=ones, constant=constant, concatenate=concatenate,
        repeat=repeat, shape_padleft=shape_padleft, shape_padright=shape_padright, shape_padaxis=shape_padaxis,
        switch=switch, expm1=expm1, exp=exp,
        tanh=_f(np.tanh), log=_f(np.log), log1p=_f(np.log1p), sqrt=_f(np.sqrt), sqr=_f(np.square),
        square=_f(np.square), abs_=_f(np.abs), sin=_f(np.sin), clip=_f(np.clip),
        max=_reduce(np.max), sum=_reduce(np.sum), mean=_reduce(np.mean), var=_reduce(np.var),
        erf=None)
    theano = _module('theano', tensor=tensor, config=_Config(), shared=shared, scan=scan, map=map_,
                     __sloika_shim__=True)
    for mod in (theano, tensor, nnet, signal, pool, sharedvar):
        sys.modules[mod.__name__] = mod
    return theano
