"""The raw-signal model architectures named by the benchmark configs, built from the B200 layers.

Same structure, sizes, activations and initialisation scaling as the reference's model scripts:
  raw_rgrgr      `models/raw_0.98_rgrgr.py:21-35`   conv(1->96, w11, s5, elu) + 5 alternating GRU(96) + softmax
  raw_rGr        `models/raw_1.00_rGr.py:13-23`     conv(1->128, w11, s2, tanh) + rGRU 110, GRU 142, rGRU 110
  bigger_raw_gru `models/bigger_raw_gru.py:16-38`   conv 32 + birnn(GRU 96) + FF 128 + birnn(GRU 96) + FF 128
  pretrained_like  the architecture of `models/pretrained.pkl`, random weights
The reference scripts themselves also run unchanged through `model_io.network_from_script`.
Weights are `sd * truncnorm(-2, 2)` draws (`module_tools.py:9-13`); seed NumPy for reproducibility.
"""
from functools import partial

from sloika_b200 import activation as act
from sloika_b200.layers import Convolution, FeedForward, Gru, Lstm, Reverse, Serial, Softmax, Window, birnn
from sloika_b200.module_tools import truncated_normal
from sloika_b200.variables import DEFAULT_NBASE, nstate


def raw_rgrgr(klen=5, sd=0.5, nbase=DEFAULT_NBASE, nfeature=1, winlen=11, stride=5, size=96):
    init = partial(truncated_normal, sd=sd)
    gru = lambda: Gru(size, size, init=init, has_bias=True, fun=act.tanh)
    return Serial([Convolution(nfeature, size, winlen, stride, init=init, has_bias=True, fun=act.elu),
                   Reverse(gru()), gru(), Reverse(gru()), gru(), Reverse(gru()),
                   Softmax(size, nstate(klen, nbase=nbase), init=init, has_bias=True)])


def raw_rGr(klen=5, sd=0.5, nbase=DEFAULT_NBASE, nfeature=1, winlen=11, stride=2, sizes=(128, 110, 142, 110)):
    n, k, l, m = sizes
    init = partial(truncated_normal, sd=sd)
    return Serial([Convolution(nfeature, n, winlen, stride, init=init, has_bias=True, fun=act.tanh),
                   Reverse(Gru(n, k, init=init, has_bias=True, fun=act.tanh)),
                   Gru(k, l, init=init, has_bias=True, fun=act.tanh),
                   Reverse(Gru(l, m, init=init, has_bias=True, fun=act.tanh)),
                   Softmax(m, nstate(klen, nbase=nbase), init=init, has_bias=True)])


def bigger_raw_gru(klen=5, sd=0.5, nbase=DEFAULT_NBASE, nfeature=1, winlen=11, stride=2, size=(32, 96, 128)):
    """The reference leaves `layer2` at its all-zero default init (`bigger_raw_gru.py:27`), which is
    degenerate for parity checks; here every layer gets random weights."""
    init = partial(truncated_normal, sd=sd)
    gru = lambda i, o: Gru(i, o, init=init, has_bias=True, fun=act.tanh)
    return Serial([Convolution(nfeature, size[0], winlen, stride, init=init, has_bias=True, fun=act.tanh),
                   birnn(gru(size[0], size[1]), gru(size[0], size[1])),
                   FeedForward(2 * size[1], size[2], init=init, has_bias=True, fun=act.tanh),
                   birnn(gru(size[2], size[1]), gru(size[2], size[1])),
                   FeedForward(2 * size[1], size[2], init=init, has_bias=True, fun=act.tanh),
                   Softmax(size[2], nstate(klen, nbase=nbase), init=init, has_bias=True)])


def pretrained_like(klen=5, sd=0.5, nbase=DEFAULT_NBASE):
    init = partial(truncated_normal, sd=sd)
    return Serial([Convolution(1, 128, 11, 5, init=init, has_bias=True, fun=act.elu),
                   Reverse(Gru(128, 112, init=init, has_bias=True)),
                   Gru(112, 144, init=init, has_bias=True),
                   Reverse(Gru(144, 112, init=init, has_bias=True)),
                   Softmax(112, nstate(klen, nbase=nbase), init=init, has_bias=True)])


def from_weights(arch, weights):
    """Rebuild a model from `Layer.json()` (no params) plus a flat `{dotted.name: ndarray}` dict as
    written by `model_io.weights_of` (used for the committed pretrained fixture)."""
    def build(desc, prefix):
        kind = desc['type']
        if kind == 'serial':
            return Serial([build(d, '{}{}.'.format(prefix, i)) for i, d in enumerate(desc['sublayers'])])
        if kind == 'parallel':
            from sloika_b200.layers import Parallel
            return Parallel([build(d, '{}{}.'.format(prefix, i)) for i, d in enumerate(desc['sublayers'])])
        if kind == 'reverse':
            return Reverse(build(desc['sublayer'], prefix + '0.'))
        if kind == 'window':
            return Window(desc['insize'], desc['w'])
        fun = getattr(act, desc['activation']) if 'activation' in desc else None
        if kind == 'LSTM':
            layer = Lstm(desc['insize'], desc['size'], has_bias=True, has_peep=True, fun=fun, gatefun=getattr(act, desc['gate']))
            for key in ('iW', 'sW', 'b', 'p'):                  # stored layout, no reshaping (see Lstm.set_params)
                getattr(layer, key).set_value(weights[prefix + key])
            return layer
        if kind == 'convolution':
            layer = Convolution(desc['insize'], desc['size'], desc['winlen'], desc['stride'],
                                has_bias=True, fun=fun, padding_mode=desc['padding_mode'])
            layer.set_params({'W': weights[prefix + 'W'], 'b': weights[prefix + 'b']})
        elif kind == 'GRU':
            n, m = desc['size'], desc['insize']
            layer = Gru(m, n, has_bias=True, fun=fun, gatefun=getattr(act, desc['gate']))
            layer.set_params({'iW': weights[prefix + 'iW'].reshape(3, n, m), 'sW': weights[prefix + 'sW'].reshape(2, n, n),
                              'sW2': weights[prefix + 'sW2'], 'b': weights[prefix + 'b'].reshape(3, n)})
        elif kind == 'feed-forward':
            layer = FeedForward(desc['insize'], desc['size'], has_bias=True, fun=fun)
            layer.set_params({'W': weights[prefix + 'W'], 'b': weights[prefix + 'b']})
        elif kind == 'softmax_old':
            layer = Softmax(desc['insize'], desc['size'], has_bias=True)
            layer.set_params({'W': weights[prefix + 'W'], 'b': weights[prefix + 'b']})
        else:
            raise NotImplementedError(kind)
        return layer
    return build(arch, '')
