"""GPU: the events route -- non-transducer decoder, event features, events worker, forward-only
scoring -- against the goldens made by the reference's own olddecode.py and against the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import forward_ref, olddecode_ref
from sloika_b200 import basecall, features, layers, olddecode, validate
from sloika_b200 import module_tools as smt

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cases():
    data = np.load(os.path.join(GOLDEN, 'olddecode_cases.npz'))
    with open(os.path.join(GOLDEN, 'olddecode_cases.json')) as fh:
        return json.load(fh), data


def test_decode_profile_matches_reference(cases):
    """Given the same float32 log-posteriors and float64 weights the kernel's scores and state sequences are those of
    the reference's decode_profile, bit for bit (log taken by NumPy so that both sides start from identical values)."""
    meta, data = cases
    for case in meta:
        name = case['name']
        lpost = olddecode_ref.log_post(data[name + '/post'], case['log'])
        trans = data[name + '/ltrans'] if case['mode'] == 'profile' else None
        score, seq = olddecode.decode_profile(lpost, trans=trans, log=True, slip=case['slip'])
        assert np.array_equal(seq, data[name + '/seq']), name
        assert score == data[name + '/score'], name
        # probabilities in: the log is taken on the device (logf vs NumPy's log: <= 1 ulp) -- same path, score to 1e-4
        if not case['log']:
            score2, seq2 = olddecode.decode_profile(data[name + '/post'], trans=trans, log=False, slip=case['slip'])
            assert np.array_equal(seq2, data[name + '/seq']) and abs(score2 - score) < 1e-4 * max(1.0, abs(score)), name


def test_estimate_transitions_matches_reference(cases):
    meta, data = cases
    for case in meta:
        if case['mode'] != 'profile' or case['log']:
            continue
        prior = None if case['prior'] is None else np.array(case['prior'])
        est = olddecode.estimate_transitions(data[case['name'] + '/post'], trans=prior)
        np.testing.assert_allclose(est, data[case['name'] + '/est'], rtol=2e-6, atol=1e-12)


def test_decode_post_non_transducer():
    """basecall.decode_post for a non-transducer model with a bad state (basecall.py:43-50): bad events dropped,
    transition weights estimated, old decoder -- same state sequence as the reference path restated in the oracle."""
    rng = np.random.default_rng(3)
    T, K = 70, 1024
    logits = 6.0 * rng.standard_normal((T, 1, K + 1))
    logits[rng.random(T) < 0.15, 0, 0] += 12.0                     # some events call the bad state
    e = np.exp(logits - logits.max(2, keepdims=True))
    post = (e / e.sum(2, keepdims=True)).astype(np.float32)
    score, call = basecall.decode_post(post, 5, False, True, 1e-5)
    p2 = post[:, 0]
    p2 = p2[p2.argmax(1) > 0, 1:]
    p2 = p2 / p2.sum(1, keepdims=True)
    p2 = (1e-5 + (1.0 - 1e-5) * p2).astype(np.float32)
    est = olddecode_ref.estimate_transitions(p2)
    s_ref, q_ref = olddecode_ref.decode_profile(p2, trans=np.log(1e-10 + est))
    assert len(call) == p2.shape[0] and np.array_equal(call, q_ref)
    assert abs(score - s_ref) < 1e-3


def test_event_features_and_worker(monkeypatch, capsys):
    """features.from_events (features.py:6-37) and basecall.events_worker (basecall.py:54-85) on a synthetic event
    table, Window + birnn(Lstm) model (models/baseline_lstm.py architecture), transducer decode."""
    rng = np.random.default_rng(5)
    n = 260
    ev = np.zeros(n, dtype=[('start', '<f8'), ('length', '<f8'), ('mean', '<f8'), ('stdv', '<f8')])
    ev['length'] = rng.integers(3, 40, n)
    ev['start'] = np.cumsum(ev['length']) - ev['length']
    ev['mean'] = 90 + 12 * rng.standard_normal(n)
    ev['stdv'] = 1.5 + 0.3 * rng.random(n)
    feat = features.from_events(ev, tag='')
    assert feat.shape == (n, 4) and feat.dtype == np.float32
    np.testing.assert_allclose(feat.mean(0), 0, atol=1e-5)
    np.testing.assert_allclose(feat.std(0), 1, atol=1e-4)
    expect3 = np.fabs(np.ediff1d(ev['mean'], to_end=0))
    np.testing.assert_allclose(feat[:, 3], (expect3 - expect3.mean()) / expect3.std(), rtol=1e-4, atol=1e-5)

    np.random.seed(6)
    init = smt.partial(smt.truncated_normal, sd=0.5)
    lstm = lambda i, o: layers.Lstm(i, o, init=init, has_bias=True, has_peep=True)
    net = layers.Serial([layers.Window(4, 3), layers.birnn(lstm(12, 32), lstm(12, 32)),
                         layers.FeedForward(64, 32, init=init, has_bias=True),
                         layers.Softmax(32, 1025, init=init, has_bias=True)])

    class FakeFast5(object):
        filename_short = 'synthetic_read'

        def __init__(self, fn):
            if 'missing' in fn:
                raise IOError("no such file")

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

        def get_section_events(self, section, analysis='Segmentation'):
            assert section == 'template' and analysis == 'Segmentation'
            return ev
    import sloika_b200.fast5
    monkeypatch.setattr(sloika_b200.fast5, 'Fast5', FakeFast5)
    basecall.calc_post = net.compile()
    try:
        res = basecall.events_worker('read.fast5', 'template', 'Segmentation', (50, 1), 5, True, True, 1e-5, skip=0.0)
        assert basecall.events_worker('missing.fast5', 'template', 'Segmentation', (50, 1), 5, True, True, 1e-5) is None
        assert basecall.events_worker('read.fast5', 'template', 'Segmentation', (200, 100), 5, True, True, 1e-5) is None
    finally:
        basecall.calc_post = None
    err = capsys.readouterr().err
    assert 'Error getting events' in err and 'Read too short' in err
    name, score, call, nev = res
    assert name == 'synthetic_read' and nev == n - 51
    x = features.from_events(ev[50:-1], tag='')[:, None, :]
    from oracle import decode_ref
    post = forward_ref.run(net.json(params=True), x)
    s_ref, p_ref = decode_ref.decode_post(post, 5, 1e-5, skip=0.0)
    assert list(call) == list(p_ref) and abs(score - s_ref) < 1e-2


def test_forward_only_scoring():
    """validate.wrap_network (validate_network.py:45-54): mean cross-entropy and correct calls of a batch."""
    np.random.seed(9)
    init = smt.partial(smt.truncated_normal, sd=0.5)
    net = layers.Serial([layers.Window(4, 3), layers.FeedForward(12, 24, init=init, has_bias=True),
                         layers.Softmax(24, 65, init=init, has_bias=True)])
    rng = np.random.default_rng(2)
    events = rng.standard_normal((50, 7, 4)).astype(np.float32)
    labels = rng.integers(0, 65, (50, 7)).astype(np.int32)
    post = forward_ref.run(net.json(params=True), events)
    labels[:25] = post[:25].argmax(2)                               # half of them right
    loss, ncorr = validate.wrap_network(net.compile())(events, labels)
    ref_loss = float(np.mean(-np.log(np.take_along_axis(post, labels[:, :, None].astype(np.int64), 2))))
    assert abs(loss - ref_loss) < 1e-4 * max(1.0, ref_loss)
    assert ncorr == int((post.argmax(2) == labels).sum())
    lab = np.array([[1, 0, 0, 2, 0], [0, 3, 0, 0, 4]])
    assert validate.remove_blanks(lab).tolist() == [[1, 1, 1, 2, 2], [0, 3, 3, 3, 4]]
