"""Remap decode on the device (csrc/remap.cu through the C ABI) against the reference's golden outputs and the
C oracle.  SURVEY.md section 8 row f1: transducer.map_to_sequence + viterbi_helpers.slip_update."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import cbind
from sloika_b200 import transducer, viterbi_helpers

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _cases():
    g = np.load(os.path.join(GOLDEN, 'remap_cases.npz'))
    with open(os.path.join(GOLDEN, 'remap_cases.json')) as fh:
        return g, json.load(fh)


def _same_score(a, b):
    return np.float32(a) == np.float32(b) or (np.isnan(a) and np.isnan(b))


def test_slip_update_matches_reference():
    g, meta = _cases()
    n = 0
    for m in meta:
        if m['kind'] != 'slip_update':
            continue
        k = m['key']
        fs, fp = viterbi_helpers.slip_update(g[k + '_x'], m['slip'])
        assert fs.dtype == np.float32 and fp.dtype == np.int64
        assert np.array_equal(fs, g[k + '_score']) and np.array_equal(fp, g[k + '_pos']), k
        n += 1
    assert n >= 12
    with pytest.raises(ValueError):
        viterbi_helpers.slip_update(np.zeros(5, dtype=np.float64), 1.0)      # Cython's buffer dtype check


def test_map_to_sequence_matches_reference_goldens():
    """Same call as the reference's (`transducer.map_to_sequence(post, seq, slip=, prior_initial=, prior_final=,
    log=)`); paths must be identical, scores identical for log input (log=False adds device-vs-NumPy logf)."""
    g, meta = _cases()
    n = 0
    for m in meta:
        if m['kind'] != 'map_to_sequence':
            continue
        k = m['key']
        p0 = g[k + '_prior0'] if k + '_prior0' in g else None
        p1 = g[k + '_prior1'] if k + '_prior1' in g else None
        score, path = transducer.map_to_sequence(g[k + '_trans'], list(g[k + '_seq']), slip=m['slip'],
                                                 prior_initial=p0, prior_final=p1, log=m['log'])
        assert path.dtype == np.int32 and len(path) == m['nev']
        assert np.array_equal(path, g[k + '_path']), k
        if m['log']:
            assert _same_score(score, g[k + '_score']), (k, score, g[k + '_score'])
        else:
            np.testing.assert_allclose(score, g[k + '_score'], rtol=2e-6)
        n += 1
    assert n == 10


@pytest.mark.parametrize('T,B,S,P,slip', [(200, 37, 65, 120, 5.0), (800, 64, 1025, 400, 5.0), (300, 16, 65, 1500, 0.5),
                                          (150, 9, 17, 40, 0.0)])
def test_map_to_sequence_batch_bit_exact_vs_c_oracle(T, B, S, P, slip):
    """Ragged batches (event counts and sequence lengths differ per read), priors on: paths and scores of the
    device kernel equal the C restatement bit for bit on the same float32 log-transducer."""
    rng = np.random.default_rng(T + B + P)
    lt = np.log(np.maximum(rng.dirichlet(np.full(S, 0.05), size=(T, B)), 1e-30)).astype(np.float32)
    lt = np.maximum(lt, -60.0).astype(np.float32)
    npos = rng.integers(max(3, P // 3), P + 1, size=B).astype(np.int32)
    npos[0], npos[-1] = P, 3
    nev = rng.integers(max(2, T // 2), T + 1, size=B).astype(np.int32)
    nev[0], nev[1] = T, 1
    seqs = [rng.integers(1, S, size=n).astype(np.int32) for n in npos]
    pri0 = [np.log(rng.dirichlet(np.ones(n))) for n in npos]
    pri1 = [np.log(rng.dirichlet(np.ones(n))) for n in npos]
    score, paths = transducer.map_to_sequence_batch(lt, seqs, nev=nev, slip=slip, prior_initial=pri0, prior_final=pri1)
    seq_pad = np.zeros((B, P), dtype=np.int32)
    p0_pad, p1_pad = np.zeros((B, P)), np.zeros((B, P))
    for b in range(B):
        seq_pad[b, :npos[b]], p0_pad[b, :npos[b]], p1_pad[b, :npos[b]] = seqs[b], pri0[b], pri1[b]
    ref_score, ref_paths = cbind.remap_batch(lt, seq_pad, nev=nev, npos=npos, slip=slip, prior0=p0_pad, prior1=p1_pad)
    assert np.array_equal(score, ref_score)
    for b in range(B):
        assert np.array_equal(paths[b], ref_paths[b, :nev[b]]), b
        assert paths[b].max() < npos[b]


def test_map_to_sequence_argument_checks():
    lt = np.zeros((4, 1, 17), dtype=np.float32)
    with pytest.raises(AssertionError):
        transducer.map_to_sequence(lt[:, 0], [1, 2, 3], slip=-1.0)          # transducer.py:26
    with pytest.raises(AssertionError):
        transducer.map_to_sequence(lt[:, 0], [1, 2], slip=1.0)              # fewer than 3 positions
    with pytest.raises(AssertionError):
        transducer.map_to_sequence(lt[:, 0], [1, 2, 40], slip=1.0)          # column outside the transducer
