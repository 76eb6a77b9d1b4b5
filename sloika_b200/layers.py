"""The `sloika.layers` operator surface of the raw basecall path, re-hosted on B200.

Reference: `sloika/layers.py` -- `Layer` ABC (:32-72), `RNN` (:75-88), `FeedForward` (:114-158),
`Softmax` (:268-314), `Convolution` (:354-419), `Gru` (:952-1021), `Reverse` (:1420-1450),
`Parallel` (:1453-1487), `Serial` (:1524-1560), `birnn` (:1622-1629).

Same class names, constructor signatures, `params()/json()/set_params()/insize/size/name` and the
same `[time, batch, feature]` row-major convention (:13).  What differs is what `run`/`compile`
mean: the reference builds a Theano graph; here a layer owns float32 parameters (host master copy,
mirrored into device buffers on first use) and `run` enqueues hand-written sm_100a kernels through
the C-ABI in `include/sloika_b200.h` on an `engine.Act` (a device activation).  There is no CPU
implementation of `run`: without the CUDA library and a GPU it raises.

Parametrised layers are described declaratively (`_SPEC`: parameter name -> stored shape and the
shape used by `json(params=True)` / `set_params`), so the JSON and `set_params` contracts of the
reference (:139-155, :291-307, :396-415, :985-1008) are produced by one code path.

`Lstm` and `Window` serve the events route.  The other layer classes of the reference
(research RNN zoo, `MaxPool`, ...) are used by no shipped model and are deliberately absent.
"""
import abc
from collections import OrderedDict

import numpy as np

from sloika_b200 import activation, conv
from sloika_b200.config import sloika_dtype


def zeros(size):
    return np.zeros(size, dtype=sloika_dtype)


class Param(object):
    """Stand-in for a Theano shared variable: a float32 host array plus lazily made device mirrors.

    `get_value/set_value` follow the shared-variable API that reference callers use
    (`layers.py:26-30`, `misc/model_convert.py`).
    """

    def __init__(self, value, name=None):
        self.name = name
        self._host = np.ascontiguousarray(value, dtype=sloika_dtype)
        self._dev = {}

    def get_value(self, borrow=False):
        return self._host if borrow else self._host.copy()

    def set_value(self, value):
        self._host = np.ascontiguousarray(value, dtype=sloika_dtype)
        self._dev = {}
        self._absmax = None

    def absmax(self):
        """max |value| (cached): lets the engine decide whether the fp16-split tensor-core kernels apply."""
        if getattr(self, '_absmax', None) is None:
            self._absmax = float(np.abs(self._host).max()) if self._host.size else 0.0
        return self._absmax

    def device(self, dev):
        """float32 buffer on torch device `dev`, uploaded once per device."""
        import torch
        key = str(dev)
        if key not in self._dev:
            self._dev[key] = torch.from_numpy(self._host).to(dev)
        return self._dev[key]

    def __getstate__(self):
        return {'name': self.name, '_host': self._host}

    def __setstate__(self, state):
        self.name = state['name']
        self._host = state['_host']
        self._dev = {}


class Layer(metaclass=abc.ABCMeta):
    """Abstract layer (`layers.py:32-72`)."""

    def compile(self):
        """Return `calc_post`, the callable the basecall workers use (`layers.py:34-36`)."""
        from sloika_b200.engine import CompiledNetwork
        return CompiledNetwork(self)

    @property
    def insize(self):
        return self._insize

    @property
    def size(self):
        return self._size

    @property
    def name(self):
        return self._name

    @abc.abstractmethod
    def params(self):
        """list of parameters (objects with get_value/set_value)"""

    @abc.abstractmethod
    def json(self, params=False):
        """JSON-able description of the layer"""

    @abc.abstractmethod
    def set_params(self, values):
        """set parameters from a dict of arrays"""

    @abc.abstractmethod
    def run(self, inMat):
        """enqueue the layer on a device activation and return the output activation"""


class _Parametrised(Layer):
    """Shared machinery for layers that own weights.

    Subclasses define `_TYPE` (the JSON type string), `_head()` (ordered non-parameter JSON
    fields) and `_spec()` -> ordered `{param: (stored_shape, json_shape)}`; the bias, when the
    layer has one, is always called `b`.
    """
    _TYPE = None

    def _ordered(self):
        return list(self._spec().keys())

    def params(self):
        return [getattr(self, k) for k in self._ordered() if k != 'b' or self.has_bias]

    def json(self, params=False):
        res = OrderedDict([('type', self._TYPE)] + self._head())
        if params:
            res['params'] = OrderedDict(
                (k, getattr(self, k).get_value(borrow=True).reshape(jshape).tolist())
                for k, (_, jshape) in self._spec().items())
        return res

    def set_params(self, values):
        for k, (stored, jshape) in self._spec().items():
            if k == 'b' and not self.has_bias:
                continue
            v = np.asarray(values[k])
            if k == 'b' and len(jshape) == 1:
                assert v.shape[0] == jshape[0], "bias has wrong length"
            else:
                assert v.shape == tuple(jshape), \
                    "parameter {} has shape {}, expected {}".format(k, v.shape, tuple(jshape))
            getattr(self, k).set_value(v.reshape(stored))


class RNN(_Parametrised):
    """Recurrent layers scan over time from a zero state (`layers.py:75-88`)."""


class _Affine(_Parametrised):
    """`W:[size,insize]`, `b:[size]` scaled as in `layers.py:129-130` / `:282-283`."""

    def _make_affine(self, insize, size, init, has_bias, name):
        self.has_bias = has_bias
        self.b = Param(has_bias * init(size), 'b')
        self.W = Param(init((size, insize)) / np.sqrt(size + insize), 'W')
        self._insize, self._size, self._name = insize, size, name

    def _spec(self):
        shape_w, shape_b = (self.size, self.insize), (self.size,)
        return OrderedDict([('W', (shape_w, shape_w)), ('b', (shape_b, shape_b))])


class FeedForward(_Affine):
    """out = fun(inMat W' + b)   (`layers.py:114-158`)."""
    _TYPE = "feed-forward"

    def __init__(self, insize, size, init=zeros, has_bias=False,
                 fun=activation.tanh, name="Feed-forward"):
        self._make_affine(insize, size, init, has_bias, name)
        self.fun = fun

    def _head(self):
        return [('activation', self.fun.__name__), ('size', self.size),
                ('insize', self.insize), ('bias', self.has_bias)]

    def run(self, inMat):
        from sloika_b200 import engine
        return engine.run_feedforward(self, inMat)


class Softmax(_Affine):
    """Row-normalised exp(inMat W' + b) with the row maximum subtracted (`layers.py:268-314`)."""
    _TYPE = "softmax_old"

    def __init__(self, insize, size, init=zeros, has_bias=False, name="Softmax"):
        self._make_affine(insize, size, init, has_bias, name)

    def _head(self):
        return [('size', self.size), ('insize', self.insize), ('bias', self.has_bias)]

    def run(self, inMat):
        from sloika_b200 import engine
        return engine.run_softmax(self, inMat)


class Convolution(_Parametrised):
    """1D convolution over the time dimension (`layers.py:354-419`, `conv.py:90-111`).

    Input `[time, batch, insize]`, output `[ceil((time + padding) / stride), batch, size]`;
    cross-correlation (`filter_flip=False`) over the zero-padded time axis, then bias and `fun`.
    """
    _TYPE = "convolution"

    def __init__(self, insize, size, winlen, stride=1, init=zeros,
                 has_bias=False, fun=activation.tanh, padding_mode='same',
                 name="Convolution"):
        self._insize, self._size, self._name = insize, size, name
        self.winlen = winlen
        self.stride = stride
        self.fun = fun
        self.has_bias = has_bias
        self.padding_mode = padding_mode
        self.padding = conv.calculate_padding(padding_mode, winlen)
        # fan-in/fan-out scaling of `layers.py:387-391`
        scale = np.sqrt(insize * winlen + (size * winlen) / float(stride))
        self.W = Param(init((size, insize, winlen)) / scale, 'W')
        self.b = Param(has_bias * init(size), 'b')

    def _head(self):
        return [("insize", self.insize), ("size", self.size), ("winlen", self.winlen),
                ("stride", self.stride), ("padding_mode", self.padding_mode),
                ("padding", self.padding), ("activation", self.fun.__name__)]

    def _spec(self):
        shape_w, shape_b = (self.size, self.insize, self.winlen), (self.size,)
        return OrderedDict([('W', (shape_w, shape_w)), ('b', (shape_b, shape_b))])

    def run(self, inMat):
        from sloika_b200 import engine
        return engine.run_convolution(self, inMat)


class Gru(RNN):
    """ Gated Recurrent Unit (`layers.py:952-1021`).

    Parameters `iW:[3*size,insize]`, `sW:[2*size,size]`, `sW2:[size,size]`, `b:[3*size]`, row blocks
    ordered [z ; r ; candidate] (scalings :974-977).  One step (:1010-1021):
        vI = x_t iW' + b ; vS = h sW' ; z = gate(vI_z + vS_z) ; r = gate(vI_r + vS_r)
        hbar = fun(vI_c + (r*h) sW2') ; h' = z*h + (1-z)*hbar
    """
    _TYPE = "GRU"

    def __init__(self, insize, size, init=zeros, has_bias=False,
                 fun=activation.tanh, gatefun=activation.sigmoid, name='GRU'):
        self._size, self._insize, self._name = size, insize, name
        self.has_bias = has_bias
        self.fun = fun
        self.gatefun = gatefun
        self.b = Param(has_bias * init(3 * size), 'b')
        self.iW = Param(init((3 * size, insize)) / np.sqrt(insize + size), 'iW')
        self.sW = Param(init((2 * size, size)) / np.sqrt(size + size), 'sW')
        self.sW2 = Param(init((size, size)) / np.sqrt(size + size), 'sW2')

    def _head(self):
        return [('activation', self.fun.__name__), ('gate', self.gatefun.__name__),
                ('size', self.size), ('insize', self.insize), ('bias', self.has_bias)]

    def _spec(self):
        n, m = self.size, self.insize
        return OrderedDict([('iW', ((3 * n, m), (3, n, m))),
                            ('sW', ((2 * n, n), (2, n, n))),
                            ('sW2', ((n, n), (n, n))),
                            ('b', ((3 * n,), (3, n)))])

    def run(self, inMat):
        from sloika_b200 import engine
        return engine.run_gru(self, inMat)


_FORGET_BIAS = 2.0            # layers.py:16


class Lstm(RNN):
    """LSTM with peepholes, "consistent with Currennt" (`layers.py:599-697`).

    Parameters `iW:[4*size,insize]`, `sW:[4*size,size]`, `b:[4*size]`, `p:[3,size]` (scalings and the forget-gate
    bias of :637-640).  The step (:677-691) views the 4*size pre-activations as `(size, 4)`: stored row `4*j + g`
    is gate g of unit j (0 update input, 1 update gate, 2 forget gate, 3 output gate):
        state' = state * gate(s2 + state*p1) + fun(s0) * gate(s1 + state*p0);  out' = fun(state') * gate(s3 + state'*p2)
    `json` / `set_params` keep the reference's `(4, size, ...)` shapes, including its transposed bias in
    `set_params` (:666-668).
    """
    _TYPE = "LSTM"

    def __init__(self, insize, size, init=zeros, has_bias=False, has_peep=False,
                 fun=activation.tanh, gatefun=activation.sigmoid, name="LSTM"):
        self._size, self._insize, self._name = size, insize, name
        self.has_bias = has_bias
        self.has_peep = has_peep
        self.fun = fun
        self.gatefun = gatefun
        self.b = Param(has_bias * (init(4 * size) + np.repeat([0, 0, _FORGET_BIAS, 0], size).astype(sloika_dtype)), 'b')
        self.p = Param(has_peep * init((3, size)) / np.sqrt(size), 'p')
        self.iW = Param(init((4 * size, insize)) / np.sqrt(insize + size), 'iW')
        self.sW = Param(init((4 * size, size)) / np.sqrt(size + size), 'sW')

    def params(self):
        return [self.iW, self.sW] + ([self.b] if self.has_bias else []) + ([self.p] if self.has_peep else [])

    def _head(self):
        return [('activation', self.fun.__name__), ('gate', self.gatefun.__name__), ('size', self.size),
                ('insize', self.insize), ('bias', self.has_bias), ('peep', self.has_peep)]

    def _spec(self):
        n, m = self.size, self.insize
        return OrderedDict([('iW', ((4 * n, m), (4, n, m))), ('sW', ((4 * n, n), (4, n, n))),
                            ('b', ((4 * n,), (4, n))), ('p', ((3, n), (3, n)))])

    def set_params(self, values):
        n, m = self.size, self.insize
        if self.has_bias:
            assert values['b'].shape == (4, n)
            self.b.set_value(np.asarray(values['b']).transpose().reshape(-1))
        if self.has_peep:
            assert values['p'].shape == (3, n)
            self.p.set_value(values['p'])
        assert values['iW'].shape == (4, n, m)
        self.iW.set_value(np.asarray(values['iW']).reshape((4 * n, m)))
        assert values['sW'].shape == (4, n, n)
        self.sW.set_value(np.asarray(values['sW']).reshape((4 * n, n)))

    def run(self, inMat):
        from sloika_b200 import engine
        return engine.run_lstm(self, inMat)


class Window(Layer):
    """Sliding window over the input (`layers.py:317-351`): `w` shifted copies, zero padded by `w // 2` steps either
    side, concatenated on the feature axis.  (The reference's `json` forgets its `return`; this one returns the
    description it builds.)"""

    def __init__(self, insize, w, name="Window"):
        assert w > 0, "Window size must be positive"
        assert w % 2 == 1, 'Window size should be odd'
        self.w = w
        self._insize = insize
        self._name = name

    @property
    def size(self):
        return self.w * self.insize

    def params(self):
        return []

    def json(self, params=False):
        res = OrderedDict([('type', "window")])
        if params:
            res['params'] = OrderedDict([('w', self.w)])
        return res

    def set_params(self, values):
        return

    def run(self, inMat):
        from sloika_b200 import engine
        return engine.run_window(self, inMat)


class _Container(Layer):
    """Combinators hold no weights of their own: `set_params` is a no-op (`layers.py:1446, 1483, 1553`)."""

    def params(self):
        return [p for layer in self._children() for p in layer.params()]

    def set_params(self, values):
        return


class Reverse(_Container):
    """  Runs a recurrent layer in reverse time (`layers.py:1420-1450`): `layer.run(x[::-1])[::-1]`.

    On the device nothing is flipped: `Act.flipped()` toggles a flag and the recurrence kernel walks
    each sequence from its own last step down to 0, which also makes ragged whole-read batches exact.
    """

    def __init__(self, layer, name='Reverse'):
        self.layer = layer
        self._name = name

    def _children(self):
        return [self.layer]

    @property
    def insize(self):
        return self.layer.insize

    @property
    def size(self):
        return self.layer.size

    def json(self, params=False):
        return OrderedDict([('type', "reverse"), ('sublayer', self.layer.json(params))])

    def run(self, inMat):
        return self.layer.run(inMat.flipped()).flipped()


class Parallel(_Container):
    """ Same input to every sub-layer, outputs concatenated on the feature axis
    (`layers.py:1453-1487`).  Sub-layers write straight into their column slice of the output.
    """

    def __init__(self, layers, name='Parallel'):
        assert len(layers) > 0, "A Parallel layer cannot be empty"
        self.layers = layers
        self._name = name
        assert all(x.insize == self.insize for x in self.layers), "Parallel layer has inconsistent sizes"

    def _children(self):
        return self.layers

    @property
    def insize(self):
        return self.layers[0].insize

    @property
    def size(self):
        return sum(x.size for x in self.layers)

    def json(self, params=False):
        return OrderedDict([('type', "parallel"),
                            ('sublayers', [layer.json(params) for layer in self.layers])])

    def run(self, inMat):
        from sloika_b200 import engine
        return engine.run_parallel(self, inMat)


class Serial(_Container):
    """ Output of each layer feeds the next (`layers.py:1524-1560`)."""

    def __init__(self, layers, name='Serial'):
        assert len(layers) > 0, "A Serial layer cannot be empty"
        self.layers = layers
        self._name = name
        assert all(x.size == y.insize for x, y in zip(layers, layers[1:])), \
            "Serial layer has inconistent sizes"

    def _children(self):
        return self.layers

    @property
    def insize(self):
        return self.layers[0].insize

    @property
    def size(self):
        return self.layers[-1].size

    def json(self, params=False):
        return OrderedDict([('type', "serial"),
                            ('sublayers', [layer.json(params) for layer in self.layers])])

    def run(self, inMat):
        act = inMat
        for layer in self.layers:
            act = layer.run(act)
        return act


def birnn(forward, backward, name='BiRNN'):
    """  Bidirectional RNN from two RNNs (`layers.py:1622-1629`)."""
    return Parallel([forward, Reverse(backward)], name=name)
