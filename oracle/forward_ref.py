"""NumPy restatement of the reference forward pass.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Works from the reference's own JSON export of a model (`Layer.json(params=True)`,
`sloika/layers.py:139-148, 291-300, 396-408, 985-997, 1441-1443, 1477-1479, 1547-1549`), so it shares
no code with the product's layer classes.  Tensors are `[time, batch, feature]` (`layers.py:13`).
`dtype=np.float32` reproduces the reference arithmetic type (`bin/basecall_network:5-7` pins
floatX=float32); `dtype=np.float64` is the error-budget twin.

Parity: PINNED.  `tools/make_golden_forward.py` executes the reference's own unmodified `sloika/layers.py`,
`conv.py`, `activation.py` and model scripts (Theano's primitives supplied eagerly by `tools/theano_shim.py`)
and stores their outputs in `tests/golden/forward_cases.npz` / `reads_forward.npz`; `tests/test_oracle.py`
asserts that every function below reproduces them to 1e-6 (2e-5 over whole reads).
"""
import numpy as np


# ---- activations: sloika/activation.py --------------------------------------------------------
def _linear(x):                     # activation.py:8
    return x


def _tanh(x):                       # activation.py:52 (T.tanh)
    return np.tanh(x)


def _sigmoid(x):
    """activation.py:56 (`T.nnet.sigmoid`).  Theano's scalar op (un-vendored, Theano 0.8.2
    `tensor/nnet/sigm.py` ScalarSigmoid.c_code, quoted from memory) evaluates
    `x < lo ? 0 : x > hi ? 1 : 1/(1+exp(-x))` with (lo, hi) = (-88, 15) for float32 and (-709, 19)
    for float64.  The clamp changes results by < 3.1e-7."""
    lo, hi = (-88.0, 15.0) if x.dtype == np.float32 else (-709.0, 19.0)
    one = x.dtype.type(1)
    with np.errstate(over='ignore'):
        y = one / (one + np.exp(-np.clip(x, lo, hi)))
    y = np.where(x < lo, x.dtype.type(0), y)
    return np.where(x > hi, one, y).astype(x.dtype)


def _elu(x):                        # activation.py:38-42: switch(x > 0, x, expm1(x))
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0))).astype(x.dtype)


ACTIVATIONS = {'linear': _linear, 'tanh': _tanh, 'sigmoid': _sigmoid, 'elu': _elu}


def _arr(p, dtype):
    return np.asarray(p, dtype=dtype)


# ---- layers -----------------------------------------------------------------------------------
def convolution(desc, x, dtype):
    """layers.py:417-419 + conv.py:66-111: zero-pad time by `padding`, valid cross-correlation
    (`filter_flip=False`) with step `stride`, add bias, activation."""
    W = _arr(desc['params']['W'], dtype)                 # [C, I, w]
    b = _arr(desc['params']['b'], dtype)
    stride, winlen = desc['stride'], desc['winlen']
    p0, p1 = desc['padding']
    T, B, I = x.shape
    xpad = np.concatenate([np.zeros((p0, B, I), dtype), x, np.zeros((p1, B, I), dtype)], axis=0)
    nout = (T + p0 + p1 - winlen) // stride + 1
    if nout <= 0:
        return np.zeros((0, B, W.shape[0]), dtype)
    # windows[t, k, b, i] = xpad[t*stride + k, b, i]
    idx = (np.arange(nout) * stride)[:, None] + np.arange(winlen)[None, :]
    windows = xpad[idx]                                  # [nout, w, B, I]
    y = np.einsum('tkbi,oik->tbo', windows, W, optimize=True).astype(dtype) + b
    return ACTIVATIONS[desc['activation']](y.astype(dtype))


def feedforward(desc, x, dtype):
    """layers.py:157-158: fun(tensordot(x, W, axes=(2, 1)) + b)."""
    W = _arr(desc['params']['W'], dtype)
    b = _arr(desc['params']['b'], dtype)
    return ACTIVATIONS[desc['activation']]((x @ W.T + b).astype(dtype))


def softmax(desc, x, dtype):
    """layers.py:309-314: t = xW'+b; m = max; e = exp(t - m); e / sum(e)."""
    W = _arr(desc['params']['W'], dtype)
    b = _arr(desc['params']['b'], dtype)
    tmp = (x @ W.T + b).astype(dtype)
    m = np.max(tmp, axis=2, keepdims=True)
    out = np.exp(tmp - m)
    rowsum = np.sum(out, axis=2, keepdims=True)
    return (out / rowsum).astype(dtype)


def gru(desc, x, dtype, lengths=None):
    """layers.py:1010-1021 stepped by RNN.run (:85-88) from h0 = 0.

    json params are `iW:(3,H,I) sW:(2,H,H) sW2:(H,H) b:(3,H)` (:990-996), row blocks [z; r; c].
    """
    H, I = desc['size'], desc['insize']
    iW = _arr(desc['params']['iW'], dtype).reshape(3 * H, I)
    sW = _arr(desc['params']['sW'], dtype).reshape(2 * H, H)
    sW2 = _arr(desc['params']['sW2'], dtype).reshape(H, H)
    b = _arr(desc['params']['b'], dtype).reshape(3 * H)
    fun = ACTIVATIONS[desc['activation']]
    gate = ACTIVATIONS[desc['gate']]
    T, B, _ = x.shape
    one = dtype(1)
    h = np.zeros((B, H), dtype)
    out = np.empty((T, B, H), dtype)
    vI_all = (x @ iW.T + b).astype(dtype)        # same per-step expression, hoisted
    for t in range(T):
        vI = vI_all[t]
        vS = h @ sW.T
        vT = vI[:, :2 * H] + vS
        z = gate(vT[:, :H])
        r = gate(vT[:, H:])
        y = (r * h) @ sW2.T
        hbar = fun(vI[:, 2 * H:] + y)
        h = (z * h + (one - z) * hbar).astype(dtype)
        out[t] = h
    return out


def lstm(desc, x, dtype):
    """layers.py:677-697: LSTM with peepholes stepped from (out, state) = 0.  Stored parameters: `iW [4H, I]`,
    `sW [4H, H]`, `b [4H]`, `p [3, H]`; the step reshapes the 4H pre-activations as `(H, 4)`, i.e. row
    `4*j + g` belongs to unit j and gate g (0 = update input, 1 = update gate, 2 = forget gate, 3 = output gate)."""
    H = desc['size']
    iW = _arr(desc['params']['iW'], dtype).reshape(4 * H, -1)
    sW = _arr(desc['params']['sW'], dtype).reshape(4 * H, H)
    b = _arr(desc['params']['b'], dtype).reshape(4 * H)
    p = _arr(desc['params']['p'], dtype).reshape(3, H)
    fun = ACTIVATIONS[desc['activation']]
    gate = ACTIVATIONS[desc['gate']]
    T, B, _ = x.shape
    out = np.zeros((B, H), dtype)
    state = np.zeros((B, H), dtype)
    res = np.empty((T, B, H), dtype)
    vW_all = (x @ iW.T).astype(dtype)
    for t in range(T):
        sumW = ((vW_all[t] + out @ sW.T) + b).reshape(B, H, 4)
        new_state = state * gate(sumW[:, :, 2] + state * p[1])
        new_state = new_state + fun(sumW[:, :, 0]) * gate(sumW[:, :, 1] + state * p[0])
        out = (fun(new_state) * gate(sumW[:, :, 3] + new_state * p[2])).astype(dtype)
        state = new_state.astype(dtype)
        res[t] = out
    return res


def window(desc, x, dtype):
    """layers.py:346-351: zero-pad w//2 steps either side, concatenate the w shifted copies on the feature
    axis: out[t, b, k*F + f] = xpad[t + k, b, f]."""
    w = desc['w'] if 'w' in desc else desc['params']['w']
    T, B, F = x.shape
    pad = np.zeros((w // 2, B, F), dtype)
    xpad = np.concatenate([pad, x, pad], axis=0)
    return np.concatenate([xpad[k:k + T] for k in range(w)], axis=2)


def with_params(arch, weights, prefix=''):
    """Join an architecture (`json(params=False)`) with a flat `{dotted.name: raw array}` dict (the layout of
    `sloika_b200.model_io.weights_of` and of the golden fixtures) into the description `run` takes."""
    kind = arch['type']
    if kind in ('serial', 'parallel'):
        return dict(arch, sublayers=[with_params(a, weights, '{}{}.'.format(prefix, i))
                                     for i, a in enumerate(arch['sublayers'])])
    if kind == 'reverse':
        return dict(arch, sublayer=with_params(arch['sublayer'], weights, prefix + '0.'))
    names = [k[len(prefix):] for k in weights if k.startswith(prefix) and '.' not in k[len(prefix):]]
    return dict(arch, params={n: weights[prefix + n] for n in names})


def run(desc, x, dtype=np.float32):
    """Evaluate a JSON model description on `x` `[T, B, F]`."""
    dtype = np.dtype(dtype).type
    x = np.asarray(x, dtype=dtype)
    kind = desc['type']
    if kind == 'serial':                           # layers.py:1556-1560
        for sub in desc['sublayers']:
            x = run(sub, x, dtype)
        return x
    if kind == 'parallel':                         # layers.py:1486-1487
        return np.concatenate([run(sub, x, dtype) for sub in desc['sublayers']], axis=2)
    if kind == 'reverse':                          # layers.py:1449-1450
        return run(desc['sublayer'], x[::-1], dtype)[::-1]
    if kind == 'convolution':
        return convolution(desc, x, dtype)
    if kind == 'feed-forward':
        return feedforward(desc, x, dtype)
    if kind == 'softmax_old':
        return softmax(desc, x, dtype)
    if kind == 'GRU':
        return gru(desc, x, dtype)
    if kind == 'LSTM':
        return lstm(desc, x, dtype)
    if kind == 'window':
        return window(desc, x, dtype)
    raise NotImplementedError("oracle has no restatement for layer type {!r}".format(kind))


def run_ragged(desc, signals, dtype=np.float32):
    """Whole-read semantics of the reference basecaller (`basecall.py:117-119`): every read is its
    own batch-of-one sequence.  `signals` is a list of 1-D arrays; returns a list of `[T'_i, S]`."""
    return [run(desc, np.asarray(s, dtype=dtype)[:, None, None], dtype)[:, 0, :] for s in signals]
