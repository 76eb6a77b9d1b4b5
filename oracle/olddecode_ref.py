"""NumPy restatement of the reference's non-transducer decoder (`sloika/olddecode.py:13-118`).  TEST INFRASTRUCTURE ONLY.

Parity: PINNED -- `tools/make_golden_olddecode.py` runs the reference's unmodified `olddecode.py` (NumPy only) and
stores its outputs in `tests/golden/olddecode_cases.npz`; `tests/test_oracle.py` asserts that the functions below
reproduce them bit for bit.

Arithmetic type.  `decode_profile` takes float32 log-posteriors but adds float64 transition weights
(`score = pscore + ev_trans[0]`, :41; `np.log(eta + trans)` is float64).  Under the NumPy >= 2 promotion rules of this
image that makes the whole recursion float64 (a float64 NumPy scalar is not demoted to the array's float32 as it was
by NumPy 1.x value-based casting); the goldens were produced here, so the restatement states float64 explicitly.
"""
import numpy as np

_ETA = 1e-10
_NSTEP = 4
_NSKIP = 16
_STEP_FACTOR = np.log(_NSTEP)
_SKIP_FACTOR = np.log(_NSKIP)


def log_post(post, log=False):
    """olddecode.py:22-25: lpost = log(eta + post) in the input's own dtype (float32 for real posteriors)."""
    lpost = np.array(post, copy=True)
    if not log:
        np.add(_ETA, lpost, lpost)
        np.log(lpost, lpost)
    return lpost


def decode_profile(post, trans=None, log=False, slip=0.0):
    """olddecode.py:13-73.  Returns (score float64, state sequence int64[T])."""
    T, nstate = post.shape
    lpost = log_post(post, log)
    if trans is None:
        tr = np.zeros((T, 3))
    else:
        tr = np.array(trans, dtype=np.float64, copy=True)
        tr[:, 1] -= _STEP_FACTOR
        tr[:, 2] -= _SKIP_FACTOR
    log_slip = np.log(_ETA + slip)
    pscore = lpost[0].astype(np.float64)
    tb = np.zeros((T, nstate), dtype=np.int64)
    idx = np.arange(nstate)
    for ev in range(1, T):
        t0, t1, t2 = tr[ev - 1]
        score = pscore + t0                                   # stay
        iscore = idx.copy()
        new = np.amax(pscore) + log_slip                      # slip: from the best state, first maximum
        inew = np.argmax(pscore)
        iscore = np.where(score > new, iscore, inew)
        score = np.fmax(score, new)
        p4 = pscore.reshape((_NSTEP, -1))                     # step
        nrem = p4.shape[1]
        new = np.repeat(np.amax(p4, axis=0), _NSTEP) + t1
        inew = np.repeat(nrem * np.argmax(p4, axis=0) + np.arange(nrem), _NSTEP)
        iscore = np.where(score > new, iscore, inew)
        score = np.fmax(score, new)
        p16 = pscore.reshape((_NSKIP, -1))                    # skip
        nrem = p16.shape[1]
        new = np.repeat(np.amax(p16, axis=0), _NSKIP) + t2
        inew = np.repeat(nrem * np.argmax(p16, axis=0) + np.arange(nrem), _NSKIP)
        iscore = np.where(score > new, iscore, inew)
        score = np.fmax(score, new)
        tb[ev - 1] = iscore
        pscore = score + lpost[ev]
    seq = np.zeros(T, dtype=np.int64)
    seq[-1] = np.argmax(pscore)
    for ev in range(T, 1, -1):
        seq[ev - 2] = tb[ev - 2][seq[ev - 1]]
    return np.amax(pscore), seq


def estimate_transitions(post, trans=None):
    """olddecode.py:94-118."""
    assert trans is None or len(trans) == 3, 'Incorrect number of transitions'
    res = np.zeros((len(post), 3))
    res[:] = _ETA
    for ev in range(1, len(post)):
        stay = np.sum(post[ev - 1] * post[ev])
        p = post[ev].reshape((-1, _NSTEP))
        step = np.sum(post[ev - 1] * np.tile(np.sum(p, axis=1), _NSTEP)) / _NSTEP
        p = post[ev].reshape((-1, _NSKIP))
        skip = np.sum(post[ev - 1] * np.tile(np.sum(p, axis=1), _NSKIP)) / _NSKIP
        res[ev - 1] = [stay, step, skip]
    if trans is None:
        trans = np.sum(res, axis=0)
        trans /= np.sum(trans)
    res *= trans
    res /= np.sum(res, axis=1).reshape((-1, 1))
    return res
