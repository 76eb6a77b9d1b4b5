#!/usr/bin/env python3
"""Golden vectors for the non-transducer decoder, produced by the reference's unmodified `sloika/olddecode.py`
(NumPy only; imported from /root/reference in the build container):

    python tools/make_golden_olddecode.py      ->  tests/golden/olddecode_cases.npz / .json

Cases: k = 3 (64 states) and k = 5 (1024 states) posteriors, flat and peaky; `decode_profile` with the per-event
weights of `estimate_transitions` (the call of `basecall.decode_post`, `basecall.py:47-50`), with a fixed prior,
`decode_simple`, a slip probability, log input, one- and two-event reads, and posteriors with exact ties.
`oracle/olddecode_ref.py` is checked against every one on the way.
"""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
warnings.simplefilter('ignore', SyntaxWarning)
GOLD = os.path.join(ROOT, 'tests', 'golden')


def main():
    from sloika import olddecode as ref
    from oracle import olddecode_ref as mine
    assert ref.__file__.startswith(REF)
    rng = np.random.default_rng(20261019)
    cases, meta = {}, []

    def posts(T, K, scale):
        logits = scale * rng.standard_normal((T, K))
        e = np.exp(logits - logits.max(1, keepdims=True))
        return (e / e.sum(1, keepdims=True)).astype(np.float32)

    def add(name, post, mode, prior=None, slip=0.0, log=False):
        eta = 1e-10
        if mode == 'profile':                       # basecall.decode_post, non-transducer branch
            est = ref.estimate_transitions(post if not log else np.exp(post), trans=prior)
            est_mine = mine.estimate_transitions(post if not log else np.exp(post), trans=prior)
            assert np.array_equal(est, est_mine), name
            ltrans = np.log(eta + est)
            score, seq = ref.decode_profile(post, trans=ltrans, log=log, slip=slip)
            s2, q2 = mine.decode_profile(post, trans=ltrans, log=log, slip=slip)
            cases[name + '/est'] = est
            cases[name + '/ltrans'] = ltrans
        else:
            ltrans = None
            score, seq = ref.decode_simple(post, log=log, slip=slip)
            s2, q2 = mine.decode_profile(post, trans=None, log=log, slip=slip)
        assert score == s2 and np.array_equal(seq, q2), name
        cases[name + '/post'] = post
        cases[name + '/score'] = np.asarray(score, dtype=np.float64)
        cases[name + '/seq'] = np.asarray(seq, dtype=np.int64)
        meta.append(dict(name=name, mode=mode, prior=None if prior is None else list(prior), slip=slip, log=log))
        print("  {:24s} T {:4d} K {:5d}  score {:.6f}".format(name, post.shape[0], post.shape[1], float(score)))

    add('k3_flat', posts(40, 64, 1.0), 'profile')
    add('k3_peaky', posts(60, 64, 6.0), 'profile')
    add('k3_simple', posts(30, 64, 3.0), 'simple')
    add('k3_slip', posts(30, 64, 3.0), 'profile', slip=1e-3)
    add('k3_prior', posts(30, 64, 3.0), 'profile', prior=np.array([0.1, 0.8, 0.1]))
    add('k5_peaky', posts(80, 1024, 8.0), 'profile')
    add('k5_flat', posts(50, 1024, 2.0), 'profile')
    add('k5_simple_slip', posts(40, 1024, 5.0), 'simple', slip=0.01)
    add('k5_log', np.log(posts(30, 1024, 5.0) + np.float32(1e-10)), 'simple', log=True)
    add('k5_T1', posts(1, 1024, 5.0), 'profile')
    add('k5_T2', posts(2, 1024, 5.0), 'profile')
    ties = np.full((25, 64), 1.0 / 64, dtype=np.float32)
    ties[np.arange(25), rng.integers(0, 64, 25)] = 0.5
    ties /= ties.sum(1, keepdims=True)
    add('k3_ties', ties, 'profile')
    np.savez_compressed(os.path.join(GOLD, 'olddecode_cases.npz'), **cases)
    with open(os.path.join(GOLD, 'olddecode_cases.json'), 'w') as fh:
        json.dump(meta, fh, indent=1)
    print("olddecode: {} cases from the reference's own olddecode.py; oracle restatement identical".format(len(meta)))


if __name__ == '__main__':
    main()
