"""CPU: host-side logic of the drop-in -- operator surface, model loader, fast5 reader, pre/post
processing, and that the C-ABI library loads and exports every symbol include/sloika_b200.h declares
(no compute calls without a GPU)."""
import io
import json
import os
import pickle
import re
import sys
import types

import numpy as np
import pytest

from conftest import REFERENCE, ROOT, needs_reference, scaled_signal
from sloika_b200 import activation, basecall, batch, bio, cabi, conv, layers, maths, model_io, util, zoo
from sloika_b200 import module_tools as smt


# ---------------------------------------------------------------- C ABI
def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'sloika_b200.h')).read()
    declared = set(re.findall(r'\b(sloika_[a-z0-9_]+)\s*\(', header))
    assert declared == set(cabi.EXPORTS), declared ^ set(cabi.EXPORTS)
    lib = cabi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sloika_b200_abi_version() == cabi.ABI_VERSION
    assert b'argument' in lib.sloika_b200_strerror(-1)
    assert lib.sloika_gru_workspace_bytes(800, 1024, 96) == 800 * 1024 * 288 * 4
    assert lib.sloika_viterbi_workspace_bytes(800, 1024, 4, 5) == 800 * 1024 * (512 + 4)   # uint16 per quad of states + a float per event
    assert lib.sloika_viterbi_workspace_bytes(10, 2, 4, 3) == 10 * 2 * 64               # generic kernel: a byte per state


def test_cabi_rejects_bad_arguments_without_touching_the_device():
    lib = cabi.load()
    assert lib.sloika_conv1d_fwd(None, None, None, None, 0, None, 1, 1, 1, 1, 1, 1, 0, 0, 0, None) == -1
    assert lib.sloika_linear_fwd(None, 0, None, None, None, 0, 1, 1, 1, 0, None) == -1
    assert lib.sloika_viterbi_fwd(None, 0, 0, None, 1, 1, 4, 5, 0.0, 1e-5, 0, None, 0, None, None, None, None) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    net = zoo.raw_rgrgr()
    with pytest.raises(cabi.SloikaB200Error):
        net.compile()(np.zeros((100, 1, 1), dtype=np.float32))


# ---------------------------------------------------------------- operator surface (test_layers.py LayerTest mixin)
def _layer_cases():
    np.random.seed(0)
    init = smt.partial(smt.truncated_normal, sd=0.5)
    return [
        (layers.Gru(12, 64, init=init, has_bias=True), ['iW', 'sW', 'sW2', 'b']),
        (layers.Gru(12, 64, init=init), ['iW', 'sW', 'sW2']),
        (layers.Convolution(12, 32, 11, 5, init=init, has_bias=True), ['W', 'b']),
        (layers.FeedForward(12, 8, init=init, has_bias=True), ['W', 'b']),
        (layers.Softmax(12, 9, init=init, has_bias=True), ['W', 'b']),
    ]


@pytest.mark.parametrize('idx', range(5))
def test_layer_api(idx):
    layer, names = _layer_cases()[idx]
    assert [p.name for p in layer.params()] == names
    assert isinstance(layer.insize, int) and isinstance(layer.size, int) and isinstance(layer.name, str)
    json.dumps(layer.json())                       # json() dumps (test_layers.py:318-323)
    desc = json.loads(json.dumps(layer.json(params=True)))
    values = {k: np.array(v, dtype=np.float32) for k, v in desc['params'].items()}
    before = [p.get_value() for p in layer.params()]
    for p in layer.params():
        p.set_value(np.zeros_like(p.get_value()))
    layer.set_params(values)                       # set_params round trip (test_layers.py:325-340)
    for old, p in zip(before, layer.params()):
        np.testing.assert_array_equal(old, p.get_value())


def test_gru_json_shapes_and_bad_set_params():
    g = layers.Gru(5, 7, has_bias=True)
    p = g.json(params=True)['params']
    assert np.array(p['iW']).shape == (3, 7, 5) and np.array(p['sW']).shape == (2, 7, 7)
    assert np.array(p['sW2']).shape == (7, 7) and np.array(p['b']).shape == (3, 7)
    with pytest.raises(AssertionError):
        g.set_params({'iW': np.zeros((21, 5)), 'sW': np.zeros((2, 7, 7)), 'sW2': np.zeros((7, 7)), 'b': np.zeros((3, 7))})


def test_containers():
    a, b = layers.FeedForward(4, 3), layers.FeedForward(3, 2)
    s = layers.Serial([a, b])
    assert (s.insize, s.size) == (4, 2) and len(s.params()) == 2
    with pytest.raises(AssertionError):
        layers.Serial([b, b])
    par = layers.Parallel([layers.FeedForward(4, 3), layers.FeedForward(4, 5)])
    assert (par.insize, par.size) == (4, 8)
    with pytest.raises(AssertionError):
        layers.Parallel([a, b])
    bi = layers.birnn(layers.Gru(4, 6), layers.Gru(4, 6))
    assert bi.json()['sublayers'][1]['type'] == 'reverse' and bi.size == 12
    r = layers.Reverse(layers.Gru(4, 6, has_bias=True))
    assert (r.insize, r.size) == (4, 6) and len(r.params()) == 4
    assert s.json()['type'] == 'serial' and par.json()['type'] == 'parallel'


def test_padding_modes():
    assert conv.calculate_padding('same', 11) == (5, 5)
    assert conv.calculate_padding('same', 4) == (1, 2)
    assert conv.calculate_padding('same_left', 4) == (2, 1)
    assert conv.calculate_padding('half', 5) == (2, 2)
    assert conv.calculate_padding('valid', 5) == (0, 0)
    assert conv.calculate_padding('full', 5) == (4, 4)
    assert conv.calculate_padding(3, 5) == (3, 3)
    assert conv.calculate_padding((1, 2), 5) == (1, 2)
    with pytest.raises(AssertionError):
        conv.calculate_padding('bogus', 5)
    for T in (0, 1, 4, 5, 6, 4000, 114190):
        assert conv.output_length(T, 11, 5, (5, 5)) == -(-T // 5)          # ceil(T/stride) for 'same'


def test_zoo_architectures():
    np.random.seed(1)
    net = zoo.raw_rgrgr()
    kinds = [d['type'] for d in net.json()['sublayers']]
    assert kinds == ['convolution', 'reverse', 'GRU', 'reverse', 'GRU', 'reverse', 'softmax_old']
    assert net.size == 1025 and net.layers[0].stride == 5 and net.layers[0].fun is activation.elu
    rgr = zoo.raw_rGr()
    assert [l.size for l in rgr.layers] == [128, 110, 142, 110, 1025] and rgr.layers[0].stride == 2
    big = zoo.bigger_raw_gru()
    assert [l.size for l in big.layers] == [32, 192, 128, 192, 128, 1025]
    # initialisation scaling of layers.py:974-977: |iW| <= 2*sd/sqrt(I+H)
    g = net.layers[2]
    assert np.abs(g.iW.get_value()).max() <= 2 * 0.5 / np.sqrt(96 + 96) + 1e-6


# ---------------------------------------------------------------- model loader
def _fake_theano_pickle(cuda=False, protocol=3):
    """Pickle of a sloika.layers tree holding Theano shared-variable look-alikes, made by planting
    throw-away modules with the right names (what a real model file contains)."""
    made = {}

    def module(name):
        mod = types.ModuleType(name)
        sys.modules[name] = mod
        made[name] = mod
        return mod

    def klass(mod, name):
        cls = type(name, (object,), {'__module__': mod.__name__})
        setattr(mod, name, cls)
        return cls

    try:
        for name in ('theano', 'theano.tensor', 'theano.gof', 'theano.sandbox', 'theano.sandbox.cuda', 'sloika'):
            module(name)
        shared_mod = module('theano.sandbox.cuda.var' if cuda else 'theano.tensor.sharedvar')
        Shared = klass(shared_mod, 'CudaNdarraySharedVariable' if cuda else 'TensorSharedVariable')
        Container = klass(module('theano.gof.link'), 'Container')
        TensorType = klass(module('theano.tensor.type'), 'TensorType')
        lay = module('sloika.layers')
        actm = module('sloika.activation')
        for fn in ('tanh', 'sigmoid', 'elu'):
            f = types.FunctionType((lambda x: x).__code__, {}, fn)
            f.__module__, f.__qualname__ = 'sloika.activation', fn
            setattr(actm, fn, f)
        Serial, Conv, Gru, Rev, Soft = [klass(lay, n) for n in ('Serial', 'Convolution', 'Gru', 'Reverse', 'Softmax')]

        rng = np.random.default_rng(3)

        def shared(shape, name):
            var, cont, typ = Shared(), Container(), TensorType()
            typ.__dict__.update(dtype='float32', broadcastable=(False,) * len(shape))
            cont.__dict__.update(storage=[rng.standard_normal(shape).astype(np.float32)], readonly=False,
                                 name=name, strict=False, type=typ, allow_downcast=None)
            var.__dict__.update(auto_name='auto_1', owner=None, name=name, index=None, type=typ, container=cont)
            return var

        conv_l = Conv()
        conv_l.__dict__.update(_insize=1, _size=8, _name='Convolution', winlen=11, stride=5, fun=actm.elu,
                               has_bias=True, padding_mode='same', padding=(5, 5),
                               W=shared((8, 1, 11), 'W'), b=shared((8,), 'b'))
        gru_l = Gru()
        gru_l.__dict__.update(_size=6, _insize=8, _name='GRU', has_bias=True, fun=actm.tanh, gatefun=actm.sigmoid,
                              b=shared((18,), 'b'), iW=shared((18, 8), 'iW'), sW=shared((12, 6), 'sW'),
                              sW2=shared((6, 6), 'sW2'))
        rev_l = Rev()
        rev_l.__dict__.update(layer=gru_l, _name='Reverse')
        soft_l = Soft()
        soft_l.__dict__.update(has_bias=True, b=shared((1025,), 'b'), W=shared((1025, 6), 'W'), _insize=6,
                               _size=np.int64(1025), _name='Softmax')
        top = Serial()
        top.__dict__.update(layers=[conv_l, rev_l, soft_l], _name='Serial')
        data = pickle.dumps(top, protocol=protocol)
        expect = {'0.W': conv_l.W.container.storage[0], '1.0.sW2': gru_l.sW2.container.storage[0],
                  '2.b': soft_l.b.container.storage[0]}
        return data, expect
    finally:
        for name in made:
            sys.modules.pop(name, None)


@pytest.mark.parametrize('cuda', [False, True])
def test_loads_theano_style_pickle(cuda):
    data, expect = _fake_theano_pickle(cuda=cuda)
    assert 'theano' not in sys.modules
    net = model_io.loads(data)
    assert isinstance(net, layers.Serial) and isinstance(net.layers[1], layers.Reverse)
    assert isinstance(net.layers[1].layer, layers.Gru) and net.layers[1].layer.gatefun is activation.sigmoid
    assert net.layers[0].fun is activation.elu and net.layers[0].padding == (5, 5)
    assert type(net.layers[2].size) is int and net.size == 1025
    got = model_io.weights_of(net)
    for key, val in expect.items():
        np.testing.assert_array_equal(got[key], val)
        assert got[key].dtype == np.float32 and got[key].flags['C_CONTIGUOUS']
    # the B200 form round-trips through pickle (helpers.compile_model output)
    again = model_io.loads(pickle.dumps(net))
    np.testing.assert_array_equal(model_io.weights_of(again)['1.0.sW2'], expect['1.0.sW2'])


def test_rejects_compiled_function_and_unknown_layers():
    class Fake(object):
        pass
    for mod_name, cls_name in (('theano.compile.function_module', '_constructor_Function'), ('sloika.layers', 'Mut1')):
        data = b'\x80\x03c' + mod_name.encode() + b'\n' + cls_name.encode() + b'\n)\x81.'
        with pytest.raises(model_io.ModelFormatError):
            model_io.loads(data)


def test_compile_model_bad_file(tmp_path):
    from sloika_b200 import helpers
    bad = tmp_path / 'bad.pkl'
    bad.write_bytes(b'not a pickle')
    with pytest.raises(ValueError):
        helpers.compile_model(str(bad))


@needs_reference
def test_loads_real_pretrained_and_matches_fixture(pretrained):
    net = model_io.load_model(os.path.join(REFERENCE, 'models', 'pretrained.pkl'))
    a, b = model_io.weights_of(net), model_io.weights_of(pretrained)
    assert a.keys() == b.keys() and sum(v.size for v in a.values()) == 395713
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    assert net.json() == pretrained.json()


@needs_reference
def test_reference_model_scripts_run_unchanged():
    np.random.seed(2)
    net = model_io.network_from_script(os.path.join(REFERENCE, 'models', 'raw_0.98_rgrgr.py'), klen=5, sd=0.5)
    assert net.json() == zoo.raw_rgrgr().json()
    net = model_io.network_from_script(os.path.join(REFERENCE, 'models', 'raw_1.00_rGr.py'), klen=5, sd=0.5)
    assert net.json() == zoo.raw_rGr().json()
    net = model_io.network_from_script(os.path.join(REFERENCE, 'models', 'bigger_raw_gru.py'), klen=5, sd=0.5)
    assert net.json() == zoo.bigger_raw_gru().json()


# ---------------------------------------------------------------- fast5
@needs_reference
def test_fast5_reader_matches_pinned_lengths_and_fixture(reads_daq):
    from sloika_b200.fast5 import Fast5, iterate_fast5
    pinned = {'read1': 114400, 'read2': 69443, 'read3': 51129, 'read6': 55885}   # test_fast5.py:99-110
    files = list(iterate_fast5(os.path.join(REFERENCE, 'data', 'reads'), paths=True))
    assert len(files) == 8
    assert len(list(iterate_fast5(os.path.join(REFERENCE, 'data', 'reads'), paths=True, limit=3))) == 3
    for fn in files:
        with Fast5(fn) as f5:
            name = f5.filename_short
            daq = f5.get_read(raw=True, scale=False)
            np.testing.assert_array_equal(daq, reads_daq[name])
            if name in pinned:
                assert len(daq) == pinned[name]
            np.testing.assert_array_equal(f5.get_read(raw=True), scaled_signal(reads_daq, name))


def test_fast5_errors(tmp_path):
    from sloika_b200.fast5 import Fast5, Fast5Error
    bad = tmp_path / 'x.fast5'
    bad.write_bytes(b'garbage' * 10)
    with pytest.raises(Fast5Error):
        Fast5(str(bad))
    basecall.calc_post = None
    assert basecall.raw_worker(str(bad), (200, 10), 0, 5, True, True, 1e-5) is None   # basecall.py:107-109


# ---------------------------------------------------------------- pre / post processing
def test_signal_preprocessing_matches_oracle(reads_daq, golden_dir):
    from oracle import host_ref
    data = np.load(os.path.join(golden_dir, 'maths_cases.npz'))
    assert maths.med_mad(data['medmad_x']) == tuple(data['medmad'])
    np.testing.assert_array_equal(maths.mad(data['medmad_x'][:4000].reshape(40, 100), axis=1), data['mad_axis1'])
    for name in ('read7', 'read5'):
        sig = scaled_signal(reads_daq, name)
        got = basecall.prepare_signal(sig, (200, 10), 0)
        np.testing.assert_array_equal(got, host_ref.prepare_signal(sig)[:, 0, 0])
    short = np.random.default_rng(1).standard_normal(300) * np.repeat([1.0, 2.0, 3.0], 100)
    assert basecall.prepare_signal(short, (200, 100), 0) is None          # basecall.py:113-115
    assert util.trim_array(np.arange(10), 2, 0).tolist() == list(range(2, 10))
    assert util.trim_array(np.arange(10), 2, 3).tolist() == list(range(2, 7))
    x = np.random.default_rng(0).standard_normal(1234)
    np.testing.assert_array_equal(batch.trim_open_pore(x, 0.3), host_ref.trim_open_pore(x, 0.3))


def test_sequence_assembly_matches_reference_vectors(golden_dir, read_basecalls, capsys):
    with open(os.path.join(golden_dir, 'bio_cases.json')) as fh:
        cases = json.load(fh)
    kmers = bio.all_kmers(5)
    assert kmers[0] == 'AAAAA' and kmers[1] == 'AAAAC' and kmers[1023] == 'TTTTT'
    assert bio.all_kmers(2, b'AC') == [b'AA', b'AC', b'CA', b'CC']
    for path, (always_move, seq) in zip(cases['paths'], cases['seqs']):
        assert bio.kmers_to_sequence([kmers[i] for i in path], always_move=always_move) == seq
        assert bio.states_to_sequence(path, 5, 'ACGT', always_move=always_move) == seq
    gold = read_basecalls['read3']
    printer = basecall.SeqPrinter(5, datatype='samples', transducer=True, alphabet='ACGT')
    n = printer.write('read3', gold['score'], gold['path'], gold['nsamples'])
    out = capsys.readouterr().out
    assert out == gold['header'] + '\n' + gold['seq'] + '\n' and n == len(gold['seq'])


def test_prepare_post_matches_oracle():
    from oracle import decode_ref
    from sloika_b200 import decode
    post = np.random.default_rng(0).random((7, 1, 65)).astype(np.float32)
    np.testing.assert_array_equal(decode.prepare_post(post, 1e-5), decode_ref.prepare_post(post, 1e-5))
    np.testing.assert_array_equal(decode.prepare_post(post, 1e-5, drop_bad=True),
                                  decode_ref.prepare_post(post, 1e-5, drop_bad=True))


def test_reader_pool_reads_like_serial(tmp_path, reads_daq):
    """`--jobs`: a forked pool parses fast5 files in parallel and hands back exactly what the serial reader does,
    exceptions included (as values, so that the caller reports them in order)."""
    from h5write import write_fast5
    files = []
    for i in (1, 4, 6):
        name = 'read{}'.format(i)
        offset, rng_, digi = reads_daq[name + '_scaling']
        fn = str(tmp_path / (name + '.fast5'))
        write_fast5(fn, reads_daq[name], float(offset), float(rng_), float(digi), read_number=i)
        files.append(fn)
    files.append(str(tmp_path / 'missing.fast5'))
    serial = [basecall._read_raw_safe(f) for f in files]
    pool = basecall.make_reader_pool(2)
    try:
        par = basecall.read_files(files, pool)
        later = basecall.read_files(files[:2], pool, wait=False).get()
    finally:
        pool.close()
    assert len(later) == 2 and np.array_equal(later[1][0], serial[1][0])
    assert [type(x) for x in basecall.read_files(files)] == [type(x) for x in serial]
    assert basecall.make_reader_pool(1) is None
    for a, b in zip(serial[:3], par[:3]):
        assert a[1] == b[1] and np.array_equal(a[0], b[0])
    assert isinstance(serial[3], Exception) and isinstance(par[3], Exception)


def test_seqprinter_uses_device_assembled_sequence_only_when_conventions_match(capsys):
    """`CalledPath.sequence` (assembled on the device by raw_batch) is printed as is when it was built with the
    printer's own (kmer_len, alphabet, transducer) conventions; otherwise the k-mers are spelled on the host."""
    path = basecall.CalledPath([0, 1, 5, 5, 21])
    host_seq = bio.states_to_sequence(path, 5, 'ACGT', always_move=True)
    path.sequence, path.assembly = 'N' * 7, (5, 'ACGT', True)               # a marker no assembly would produce
    printer = basecall.SeqPrinter(5, datatype='samples', transducer=True, alphabet='ACGT')
    assert printer.write('r', -3.2, path, 100) == 7
    other = basecall.SeqPrinter(5, datatype='samples', transducer=False, alphabet='ACGT')
    n = other.write('r', -3.2, path, 100)
    out = capsys.readouterr().out.splitlines()
    assert out[0] == '>r score -3, 100 samples to 7 bases' and out[1] == 'NNNNNNN'
    assert out[3] == bio.states_to_sequence(path, 5, 'ACGT', always_move=False) and n == len(out[3])
    assert host_seq != 'NNNNNNN'
    from sloika_b200 import transducer
    assert transducer.argmax(3.0, 9.0, 1.0) == (1, 9.0)                      # transducer.py:9-11


def test_unpickler_refuses_globals_a_model_does_not_need():
    """A model pickle cannot name arbitrary callables (find_class used to fall through to pickle's own)."""
    import pickle

    class Evil(object):
        def __reduce__(self):
            import os
            return (os.system, ('true',))
    with pytest.raises(model_io.ModelFormatError):
        model_io.loads(pickle.dumps(Evil(), protocol=3))


def test_sub_batches_under_a_padded_sample_budget():
    """Reads sorted longest first are grouped so that (reads x longest read) stays under the budget; an outlier
    longer than the budget is called on its own (one ultra-long read used to inflate a 64-read batch)."""
    from sloika_b200.basecall import plan_sub_batches
    lengths = [1000000, 60000, 59000, 58000, 100, 90]
    groups = plan_sub_batches(lengths, 200000)
    assert groups[0] == [0]
    assert sorted(p for g in groups for p in g) == list(range(6))
    for g in groups[1:]:
        assert len(g) * max(lengths[p] for p in g) <= 200000
    assert plan_sub_batches([], 10) == [] and plan_sub_batches([5], 1) == [[0]]
