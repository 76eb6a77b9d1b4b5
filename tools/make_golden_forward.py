#!/usr/bin/env python3
"""Golden vectors for the forward pass, produced by EXECUTING THE REFERENCE'S OWN SOURCE.

Run in the build container, where /root/reference is mounted:

    python tools/make_golden_forward.py

`tools/theano_shim.py` is registered as `theano`; then `sloika/layers.py`, `sloika/conv.py`,
`sloika/activation.py`, `sloika/module_tools.py` and the model scripts under `models/` are imported
unmodified from /root/reference and `network.run(x)` is called on seeded float32 inputs.  What runs is the
reference's code: `Gru.step` (`layers.py:1010-1021`) under `RNN.run` (:85-88), `conv.conv_1d` (`conv.py:90-111`)
with `pad_first` / `bf1t` / `tbf`, `Reverse.run` (:1449-1450), `Parallel.run` (:1486-1487), `Lstm.step`
(:677-691), `Window.run` (:346-351), `Softmax.run` (:309-314).  `models/pretrained.pkl` is unpickled with the
shim's shared-variable class and run on the 8 bundled reads; the reference's `decode.py` and `bio.py` then
turn the posteriors into base calls.

Written to tests/golden/:
  forward_cases.npz / .json   per case: architecture (`json(params=False)`), raw parameter values, input,
                              output of the reference
  reads_forward.npz           per bundled read: posterior rows (head, tail, every 389th), row maxima
  reads_basecalls.json        (rewritten) basecalls of the 8 reads = reference forward (under the shim) ->
                              reference decode.viterbi -> reference bio.kmers_to_sequence

and checks on the way that `oracle/forward_ref.py` agrees with every output to 1e-6 (the same assertion is
repeated by tests/test_oracle.py from the committed fixtures).
"""
import json
import os
import pickle
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
warnings.simplefilter('ignore', SyntaxWarning)
GOLD = os.path.join(ROOT, 'tests', 'golden')

import theano_shim                                    # noqa: E402
theano_shim.install()
sys.path.insert(0, REF)
import sloika.module_tools as smt                     # noqa: E402  (the reference's, under the shim)
import sloika.layers as ref_layers                    # noqa: E402
assert ref_layers.__file__.startswith(REF)


def arch_of(layer):
    """`json(params=False)` of the reference, with the one hole patched: `Window.json` has no return."""
    if isinstance(layer, ref_layers.Serial):
        return {'type': 'serial', 'sublayers': [arch_of(l) for l in layer.layers]}
    if isinstance(layer, ref_layers.Parallel):
        return {'type': 'parallel', 'sublayers': [arch_of(l) for l in layer.layers]}
    if isinstance(layer, ref_layers.Reverse):
        return {'type': 'reverse', 'sublayer': arch_of(layer.layer)}
    if isinstance(layer, ref_layers.Window):
        return {'type': 'window', 'w': int(layer.w), 'insize': int(layer.insize)}
    desc = json.loads(json.dumps(layer.json(params=False), default=lambda o: int(o)))
    return desc


def weights_of(layer, prefix=''):
    """Raw values of the shared variables, keyed like `sloika_b200.model_io.weights_of`."""
    out = {}
    if isinstance(layer, (ref_layers.Serial, ref_layers.Parallel)):
        for i, child in enumerate(layer.layers):
            out.update(weights_of(child, '{}{}.'.format(prefix, i)))
    elif isinstance(layer, ref_layers.Reverse):
        out.update(weights_of(layer.layer, prefix + '0.'))
    else:
        for key, val in layer.__dict__.items():
            if isinstance(val, theano_shim.SharedVariable):
                out[prefix + key] = np.asarray(val.get_value(), dtype=np.float32)
    return out


def load_script(name):
    path = os.path.join(REF, 'models', name)
    scope = {'__name__': '__sloika_model__', '__file__': path}
    with open(path) as fh:
        exec(compile(fh.read(), path, 'exec'), scope)
    return scope['network']


def randomise_zero_weights(layer, rng):
    """bigger_raw_gru.py:27 / baseline_*.py leave one FeedForward at the all-zero default init; give it values
    so that the case tests something (through the reference's own set_value)."""
    def visit(l):
        if isinstance(l, (ref_layers.Serial, ref_layers.Parallel)):
            for c in l.layers:
                visit(c)
        elif isinstance(l, ref_layers.Reverse):
            visit(l.layer)
        else:
            for key, val in l.__dict__.items():
                if isinstance(val, theano_shim.SharedVariable) and val.value.size and not np.any(val.value):
                    val.set_value((0.2 * rng.standard_normal(val.value.shape)).astype(np.float32))
    visit(layer)


def main():
    from oracle import forward_ref
    rng = np.random.default_rng(20261018)
    np.random.seed(0xbeef)                             # scipy truncnorm draws inside module_tools.truncated_normal
    init = smt.partial(smt.truncated_normal, sd=0.5)
    L = ref_layers
    cases = []

    def add(name, net, x, tol=1e-6):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.asarray(net.run(x.view(theano_shim.Tensor)))
        assert y.dtype == np.float32, (name, y.dtype)
        arch, weights = arch_of(net), weights_of(net)
        mine = forward_ref.run(forward_ref.with_params(arch, weights), x)
        err = float(np.abs(mine - y).max()) if y.size else 0.0
        assert mine.shape == y.shape and err <= tol, (name, mine.shape, y.shape, err)
        cases.append((name, arch, weights, x, y, err))
        print("  {:34s} x {} -> y {}   oracle max|diff| {:.2e}".format(name, x.shape, y.shape, err))

    # ---- Convolution: the raw front ends, every padding mode, strides that do not divide T, even windows ----
    add('conv_1_96_w11_s5_elu', L.Convolution(1, 96, 11, 5, init=init, has_bias=True, fun=smt.elu),
        rng.standard_normal((203, 3, 1)))
    add('conv_1_128_w11_s2_tanh', L.Convolution(1, 128, 11, 2, init=init, has_bias=True, fun=smt.tanh),
        rng.standard_normal((101, 2, 1)))
    add('conv_12_32_w11_s5', L.Convolution(12, 32, 11, 5, init=init, has_bias=True),
        rng.standard_normal((64, 3, 12)))
    add('conv_nobias_linear', L.Convolution(3, 8, 5, 1, init=init, has_bias=False, fun=smt.linear),
        rng.standard_normal((17, 2, 3)))
    for mode in ('same', 'half', 'valid', 'full', 'same_left'):
        for winlen, stride in ((4, 3), (7, 2)):
            add('conv_pad_{}_w{}_s{}'.format(mode, winlen, stride),
                L.Convolution(2, 6, winlen, stride, init=init, has_bias=True, fun=smt.elu, padding_mode=mode),
                rng.standard_normal((29, 2, 2)))
    add('conv_pad_int3', L.Convolution(2, 5, 6, 2, init=init, has_bias=True, padding_mode=3),
        rng.standard_normal((20, 2, 2)))
    add('conv_T_shorter_than_window', L.Convolution(1, 4, 11, 5, init=init, has_bias=True, fun=smt.elu),
        rng.standard_normal((3, 2, 1)))

    # ---- Gru: the sizes of the shipped raw models, both directions, birnn ----
    for insize, size in ((96, 96), (128, 110), (110, 142), (128, 112), (112, 144), (32, 96), (7, 5)):
        add('gru_{}_{}'.format(insize, size), L.Gru(insize, size, init=init, has_bias=True),
            rng.standard_normal((37, 3, insize)))
    add('gru_nobias', L.Gru(6, 9, init=init, has_bias=False), rng.standard_normal((12, 2, 6)))
    add('gru_rev_96', L.Reverse(L.Gru(96, 96, init=init, has_bias=True)), rng.standard_normal((41, 2, 96)))
    add('gru_birnn_32_96', L.birnn(L.Gru(32, 96, init=init, has_bias=True), L.Gru(32, 96, init=init, has_bias=True)),
        rng.standard_normal((33, 2, 32)))
    add('gru_sat', L.Gru(4, 8, init=smt.partial(smt.truncated_normal, sd=40.0), has_bias=True),
        8.0 * rng.standard_normal((25, 2, 4)), tol=1e-5)          # saturated gates: sigmoid clamp region

    # ---- the other operators of the surface ----
    add('feedforward_192_128_tanh', L.FeedForward(192, 128, init=init, has_bias=True, fun=smt.tanh),
        rng.standard_normal((9, 3, 192)))
    add('softmax_96_1025', L.Softmax(96, 1025, init=init, has_bias=True), rng.standard_normal((7, 2, 96)))
    add('window_4_w3', L.Window(4, 3), rng.standard_normal((11, 2, 4)))
    add('window_2_w5', L.Window(2, 5), rng.standard_normal((3, 1, 2)))
    for peep in (True, False):
        add('lstm_12_64_peep{}'.format(int(peep)), L.Lstm(12, 64, init=init, has_bias=True, has_peep=peep),
            rng.standard_normal((29, 3, 12)))
    add('lstm_rev_5_7', L.Reverse(L.Lstm(5, 7, init=init, has_bias=True, has_peep=True)),
        rng.standard_normal((13, 2, 5)))
    add('lstm_nobias', L.Lstm(5, 16, init=init, has_bias=False, has_peep=False), rng.standard_normal((9, 2, 5)))

    # ---- the shipped model scripts, default arguments of bin/train_network.py (klen 5, sd 0.5) ----
    for script, kwargs, T, F in (('raw_0.98_rgrgr.py', {}, 250, 1), ('raw_1.00_rGr.py', {}, 120, 1),
                                 ('bigger_raw_gru.py', {}, 120, 1), ('baseline_raw_gru.py', {}, 120, 1),
                                 ('baseline_gru.py', {}, 60, 4), ('baseline_lstm.py', {}, 60, 4),
                                 ('tiny_gru.py', {}, 60, 4)):
        # klen 5 (1025 states) for the headline architecture, klen 3 (65 states) elsewhere to keep the fixture small
        net = load_script(script)(klen=5 if 'rgrgr' in script else 3, sd=0.5, **kwargs)
        randomise_zero_weights(net, rng)
        add('model_' + script[:-3], net, rng.standard_normal((T, 2, F)), tol=2e-6)

    flat, meta = {}, []
    for name, arch, weights, x, y, err in cases:
        flat[name + '/x'] = x
        flat[name + '/y'] = y
        for k, v in weights.items():
            flat[name + '/w/' + k] = v
        meta.append({'name': name, 'arch': arch})
    np.savez_compressed(os.path.join(GOLD, 'forward_cases.npz'), **flat)
    with open(os.path.join(GOLD, 'forward_cases.json'), 'w') as fh:
        json.dump(meta, fh)
    print("forward: {} cases from the reference's own layers.py / conv.py (Theano primitives: tools/theano_shim.py)"
          .format(len(cases)))

    reads(forward_ref)


def reads(forward_ref):
    """models/pretrained.pkl, unpickled into the reference's own classes, on the 8 bundled reads."""
    from sloika import decode as ref_decode, bio as ref_bio
    from oracle import host_ref
    from sloika_b200.fast5 import Fast5

    import types
    # the pickle also names Theano classes that carry no arithmetic: inert holders
    holder = lambda n, m: type(n, (object,), {'__module__': m, '__setstate__': lambda s, st: s.__dict__.update(st)})
    for modname, names in (('theano.tensor.type', ['TensorType']), ('theano.gof.utils', ['scratchpad']),
                           ('theano.gof.link', ['Container']), ('theano.gof', [])):
        mod = sys.modules.get(modname) or types.ModuleType(modname)
        for n in names:
            setattr(mod, n, holder(n, modname))
        sys.modules[modname] = mod
    with open(os.path.join(REF, 'models', 'pretrained.pkl'), 'rb') as fh:
        net = pickle.load(fh)
    assert type(net).__module__ == 'sloika.layers'
    arch, weights = arch_of(net), weights_of(net)
    committed = np.load(os.path.join(GOLD, 'pretrained_weights.npz'))
    assert sorted(committed.files) == sorted(weights) and all(np.array_equal(committed[k], weights[k]) for k in weights)
    desc = forward_ref.with_params(arch, weights)

    with open(os.path.join(GOLD, 'reads_basecalls.json')) as fh:
        previous = {r['name']: r for r in json.load(fh)}
    kmers = ref_bio.all_kmers(5)
    out, records = {}, []
    for i in range(1, 9):
        name = 'read{}'.format(i)
        f5 = Fast5(os.path.join(REF, 'data', 'reads', name + '.fast5'))
        x = host_ref.prepare_signal(f5.get_read(raw=True))
        post = np.asarray(net.run(x.view(theano_shim.Tensor)))                  # the reference's run()
        mine = forward_ref.run(desc, x)
        err = float(np.abs(mine - post).max())
        prepared = ref_decode.prepare_post(post, min_prob=1e-5, drop_bad=False)
        score, path = ref_decode.viterbi(prepared, 5, skip_pen=0.0, nbase=4)
        seq = ref_bio.kmers_to_sequence([kmers[s] for s in path], always_move=True)
        header = ">{} score {:.0f}, {} {} to {} bases".format(f5.filename_short, score, x.shape[0], 'samples', len(seq))
        rec = dict(previous[name])
        same = rec['seq'] == seq and rec['path'] == [int(p) for p in path]
        rec.update(nsamples=int(x.shape[0]), nsteps=int(post.shape[0]), score=float(score), header=header, seq=seq,
                   path=[int(p) for p in path], mean_max_post=float(post.max(2).mean()),
                   forward='reference layers.py under tools/theano_shim.py')
        records.append(rec)
        sel = np.unique(np.concatenate([np.arange(16), np.arange(post.shape[0] - 16, post.shape[0]),
                                        np.arange(0, post.shape[0], 389)]))
        out[name + '_rows'] = sel.astype(np.int32)
        out[name + '_post'] = post[sel, 0]
        out[name + '_rowmax'] = post[:, 0].max(1)
        print("  {}: {} steps, score {:.1f}, {} bases; oracle max|diff| {:.2e}; same call as the oracle-forward "
              "golden: {}".format(name, post.shape[0], score, len(seq), err, same))
        assert err < 2e-5, (name, err)
    np.savez_compressed(os.path.join(GOLD, 'reads_forward.npz'), **out)
    with open(os.path.join(GOLD, 'reads_basecalls.json'), 'w') as fh:
        json.dump(records, fh)


if __name__ == '__main__':
    main()
    for fn in ('forward_cases.npz', 'forward_cases.json', 'reads_forward.npz', 'reads_basecalls.json'):
        print("  {:28s} {:>9d} B".format(fn, os.path.getsize(os.path.join(GOLD, fn))))
