"""CPU: the oracle restatements against the golden vectors produced from the reference itself
(tools/make_golden.py).  These pin the checker that the GPU parity tests rely on."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, scaled_signal
from oracle import cbind, decode_ref, forward_ref, host_ref


def _post_for(case, data):
    name = case['name']
    if name + '/post' in data:
        return data[name + '/post']
    return decode_ref.prepare_post(data[name + '/raw'], min_prob=1e-5)


def test_reference_known_answers(decode_cases):
    """test/unit/test_decode.py:233-241 literal expectations."""
    meta, data = decode_cases
    score, path = decode_ref.viterbi(data['kat_post3/post'], 3)
    assert path == [49, 7, 63, 63]
    assert score == pytest.approx(-11.130084569094556, abs=1e-7)
    score, path = decode_ref.viterbi(data['kat_post3/post'], 3, skip_pen=3.0)
    assert path == [49, 7, 31, 63, 63]
    assert score == pytest.approx(-11.936803444063674, abs=1e-7)


def test_numpy_oracle_matches_reference_vectors(decode_cases):
    meta, data = decode_cases
    for case in meta:
        post = _post_for(case, data)
        score, path = decode_ref.viterbi(post, case['klen'], skip_pen=case['skip_pen'], log=case['log'],
                                         nbase=case['nbase'])
        assert path == data[case['name'] + '/path'].tolist(), case['name']
        assert score == data[case['name'] + '/score'], case['name']


def test_c_oracle_matches_reference_vectors(decode_cases):
    meta, data = decode_cases
    for case in meta:
        post = _post_for(case, data)
        if post.dtype != np.float32:
            continue        # the C restatement is float32 (the arithmetic type of the real path)
        lp = post if case['log'] else decode_ref.log_post(post)
        score, paths = cbind.viterbi_batch(lp[:, None, :], None, klen=case['klen'], nbase=case['nbase'],
                                           skip_pen=case['skip_pen'])
        assert paths[0] == data[case['name'] + '/path'].tolist(), case['name']
        assert score[0] == data[case['name'] + '/score'], case['name']


def test_c_oracle_ragged_batch():
    rng = np.random.default_rng(5)
    T, B, S = 90, 5, 1025
    logits = 3 * rng.standard_normal((T, B, S))
    logits[:, :, 0] += 6
    post = np.exp(logits - logits.max(2, keepdims=True))
    post = (post / post.sum(2, keepdims=True)).astype(np.float32)
    lp = decode_ref.log_post(1e-5 + (1.0 - 1e-5) * post).astype(np.float32)
    lengths = np.array([90, 1, 2, 45, 89], dtype=np.int32)
    score, paths = cbind.viterbi_batch(lp, lengths, skip_pen=0.0)
    for b in range(B):
        s, p = decode_ref.viterbi(lp[:lengths[b], b], 5, skip_pen=0.0, log=True)
        assert p == paths[b] and s == score[b]


def test_host_vectors(golden_dir):
    data = np.load(os.path.join(golden_dir, 'maths_cases.npz'))
    med, mad = host_ref.med_mad(data['medmad_x'])
    assert (med, mad) == tuple(data['medmad'])
    assert np.array_equal(host_ref.med_mad(data['medmad_x'][:4000].reshape(40, 100), axis=1)[1], data['mad_axis1'])
    with open(os.path.join(golden_dir, 'bio_cases.json')) as fh:
        bio = json.load(fh)
    kmers = host_ref.all_kmers(5)
    for path, (always_move, seq) in zip(bio['paths'], bio['seqs']):
        assert host_ref.kmers_to_sequence([kmers[i] for i in path], always_move=always_move) == seq


def test_feedforward_formula():
    """test/unit/test_layers.py:58-69: FeedForward == x.W' + b (linear) / tanh of it."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((10, 20, 12)).astype(np.float32)
    W = rng.standard_normal((8, 12)).astype(np.float32)
    b = rng.standard_normal(8).astype(np.float32)
    for fun, f in (('linear', lambda v: v), ('tanh', np.tanh)):
        desc = {'type': 'feed-forward', 'activation': fun, 'params': {'W': W.tolist(), 'b': b.tolist()}}
        got = forward_ref.run(desc, x)
        np.testing.assert_almost_equal(got, f(x.dot(W.T) + b), decimal=5)


def test_softmax_rows_sum_to_one_and_structure():
    """test/unit/test_layers.py:71-125: softmax rows sum to 1; Reverse/Parallel/Serial structure."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((7, 3, 5)).astype(np.float32)
    W = rng.standard_normal((9, 5)).astype(np.float32)
    sm = {'type': 'softmax_old', 'params': {'W': W.tolist(), 'b': np.zeros(9).tolist()}}
    out = forward_ref.run(sm, x)
    np.testing.assert_allclose(out.sum(2), 1.0, rtol=1e-5)
    ff = {'type': 'feed-forward', 'activation': 'tanh', 'params': {'W': W.tolist(), 'b': np.zeros(9).tolist()}}
    rev = forward_ref.run({'type': 'reverse', 'sublayer': ff}, x)
    np.testing.assert_array_equal(rev, forward_ref.run(ff, x))          # pointwise layer: flip cancels
    par = forward_ref.run({'type': 'parallel', 'sublayers': [ff, ff]}, x)
    assert par.shape == (7, 3, 18)
    np.testing.assert_array_equal(par[:, :, :9], par[:, :, 9:])


def test_oracle_reproduces_bundled_read_basecall(pretrained, reads_daq, read_basecalls, golden_dir):
    """Forward (float32) + decode + assembly of a bundled read == the golden record made with the
    reference's decode.py / bio.py.  read5 (32 890 samples) keeps the CPU suite short."""
    name = 'read5'
    gold = read_basecalls[name]
    x = host_ref.prepare_signal(scaled_signal(reads_daq, name))
    assert x.shape[0] == gold['nsamples']
    post = forward_ref.run(pretrained.json(params=True), x)
    slices = np.load(os.path.join(golden_dir, 'reads_post_slices.npz'))
    np.testing.assert_allclose(post[:8, 0], slices[name + '_head'], atol=2e-6)
    np.testing.assert_allclose(post[-8:, 0], slices[name + '_tail'], atol=2e-6)
    np.testing.assert_allclose(post[:, 0].max(1), slices[name + '_rowmax'], atol=2e-6)
    score, path = decode_ref.decode_post(post, 5, 1e-5, skip=0.0)
    assert path == gold['path']
    rec = host_ref.fasta_record(name, score, path, x.shape[0])
    assert rec == gold['header'] + '\n' + gold['seq'] + '\n'
    assert gold['embedded_basecall_identity'] > 0.8     # anchor: agrees with the ONT call in the fast5


def test_float64_twin_bounds_float32_error(pretrained, reads_daq):
    """Error budget: float32 oracle vs float64 twin on a short real read stays far below 1e-4."""
    x = host_ref.prepare_signal(scaled_signal(reads_daq, 'read7'))[:3000]
    desc = pretrained.json(params=True)
    p32 = forward_ref.run(desc, x, np.float32)
    p64 = forward_ref.run(desc, x, np.float64)
    assert np.abs(p32 - p64).max() < 2e-5


# ------------------------------------------------------------------ remap decode
def _remap_cases():
    g = np.load(os.path.join(GOLDEN, 'remap_cases.npz'))
    with open(os.path.join(GOLDEN, 'remap_cases.json')) as fh:
        meta = json.load(fh)
    return g, meta


def _same_score(a, b):
    return np.float32(a) == np.float32(b) or (np.isnan(a) and np.isnan(b))


def test_remap_oracles_match_reference_goldens():
    """oracle/remap_ref.py and oracle/remap_ref.c reproduce, bit for bit, what the reference's own
    transducer.map_to_sequence / viterbi_helpers.slip_update returned (tools/make_golden_remap.py), including the
    known-answer recipe of test/unit/test_viterbi.py and the NaN behaviour of slip=None."""
    from oracle import cbind, remap_ref
    g, meta = _remap_cases()
    kinds = set()
    for m in meta:
        k = m['key']
        kinds.add(m['kind'])
        if m['kind'] == 'slip_update':
            for fn in (remap_ref.slip_update, cbind.slip_update):
                fs, fp = fn(g[k + '_x'], m['slip'])
                assert np.array_equal(fs, g[k + '_score']) and np.array_equal(fp, g[k + '_pos']), k
            continue
        p0 = g[k + '_prior0'] if k + '_prior0' in g else None
        p1 = g[k + '_prior1'] if k + '_prior1' in g else None
        s, p = remap_ref.map_to_sequence(g[k + '_trans'], g[k + '_seq'], m['slip'], p0, p1, m['log'])
        assert np.array_equal(p, g[k + '_path']) and _same_score(s, g[k + '_score']), k
        lt = g[k + '_trans'] if m['log'] else np.log(g[k + '_trans'])
        s, p = cbind.remap_batch(lt[:, None, :], g[k + '_seq'][None], slip=m['slip'],
                                 prior0=None if p0 is None else p0[None], prior1=None if p1 is None else p1[None])
        assert np.array_equal(p[0], g[k + '_path']) and _same_score(s[0], g[k + '_score']), k
    assert kinds == {'slip_update', 'map_to_sequence'}


def test_slip_update_reference_unit_test_recipe():
    """test/unit/test_viterbi.py:13-33 restated: the helper equals the plain-Python recurrence on the seeded input."""
    from oracle import cbind
    np.random.seed(0xdeadbeef)
    x = np.random.normal(size=10).astype(np.float32)
    slip = 5.0
    y1s, y1i = cbind.slip_update(x, slip)
    y2s = np.zeros(len(x), dtype=np.float32)
    y2i = np.zeros(len(x), dtype=np.int64)
    y2s[0] = y2s[1] = -1e38
    y2s[2] = x[0] - slip
    for j in range(3, len(x)):
        if y2s[j - 1] >= x[j - 2]:
            y2s[j], y2i[j] = y2s[j - 1], y2i[j - 1]
        else:
            y2s[j], y2i[j] = x[j - 2], j - 2
        y2s[j] -= slip
    np.testing.assert_almost_equal(y1s, y2s)
    np.testing.assert_equal(y1i, y2i)


# ---- forward pass: pinned to outputs of the reference's own source (tools/make_golden_forward.py) ----
def test_forward_oracle_matches_reference_outputs(forward_cases):
    """`forward_ref.run` == `network.run(x)` of the unmodified sloika/layers.py + conv.py, every case."""
    from conftest import case_weights
    meta, data = forward_cases
    names = [c['name'] for c in meta]
    for must in ('conv_1_96_w11_s5_elu', 'conv_12_32_w11_s5', 'gru_96_96', 'gru_128_110', 'gru_110_142', 'gru_128_112',
                 'gru_112_144', 'gru_rev_96', 'gru_birnn_32_96', 'model_raw_0.98_rgrgr', 'model_raw_1.00_rGr',
                 'model_bigger_raw_gru', 'lstm_12_64_peep1', 'window_4_w3'):
        assert must in names
    for case in meta:
        name = case['name']
        desc = forward_ref.with_params(case['arch'], case_weights(data, name))
        got = forward_ref.run(desc, data[name + '/x'])
        ref = data[name + '/y']
        assert got.shape == ref.shape and got.dtype == np.float32, name
        tol = 1e-5 if name == 'gru_sat' else 2e-6 if name.startswith('model_') else 1e-6
        assert np.abs(got - ref).max() <= tol, (name, float(np.abs(got - ref).max()))


def test_forward_oracle_whole_read_matches_reference(pretrained, reads_daq):
    """models/pretrained.pkl on a bundled read: oracle posteriors vs the reference's (stored rows), 2e-5."""
    fwd = np.load(os.path.join(GOLDEN, 'reads_forward.npz'))
    desc = pretrained.json(params=True)
    for name in ('read7', 'read5'):
        x = host_ref.prepare_signal(scaled_signal(reads_daq, name))
        post = forward_ref.run(desc, x)[:, 0]
        assert np.abs(post[fwd[name + '_rows']] - fwd[name + '_post']).max() < 2e-5
        assert np.abs(post.max(1) - fwd[name + '_rowmax']).max() < 2e-5


# ---- non-transducer decoder: pinned to outputs of the reference's own olddecode.py (tools/make_golden_olddecode.py) ----
def test_olddecode_oracle_matches_reference_outputs():
    from oracle import olddecode_ref
    data = np.load(os.path.join(GOLDEN, 'olddecode_cases.npz'))
    with open(os.path.join(GOLDEN, 'olddecode_cases.json')) as fh:
        meta = json.load(fh)
    assert len(meta) >= 12
    for case in meta:
        name = case['name']
        post = data[name + '/post']
        if case['mode'] == 'profile':
            prior = None if case['prior'] is None else np.array(case['prior'])
            est = olddecode_ref.estimate_transitions(np.exp(post) if case['log'] else post, trans=prior)
            assert np.array_equal(est, data[name + '/est']), name
            score, seq = olddecode_ref.decode_profile(post, trans=data[name + '/ltrans'], log=case['log'], slip=case['slip'])
        else:
            score, seq = olddecode_ref.decode_profile(post, trans=None, log=case['log'], slip=case['slip'])
        assert score == data[name + '/score'] and np.array_equal(seq, data[name + '/seq']), name
