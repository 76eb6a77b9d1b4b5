"""NumPy restatement of the reference's remap decode.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows `sloika/transducer.py:14-73` (map_to_sequence) and `sloika/viterbi_helpers.pyx:12-35` (slip_update).
PINNED: `tests/test_oracle.py` checks both functions bit for bit against `tests/golden/remap_cases.npz`, which
`tools/make_golden_remap.py` produced by running the reference's own transducer.py on its own compiled
viterbi_helpers (the known-answer recipe of `test/unit/test_viterbi.py:10-33` is one of the stored cases).

Arithmetic notes that matter for bit-exactness (all observed on the reference, not assumed):
  * scores are float32 (`sloika_dtype`), the slip penalty is `np.float32(slip)`;
  * `np.float32(None)` is NaN, and the reference tests `slip is not None` AFTER that conversion
    (transducer.py:29, :55), so `slip=None` does not disable slips: it runs them with a NaN penalty and the
    score comes out NaN.  Restated as is (two golden cases cover it);
  * priors are float64 (`util.geometric_prior`); `pscore += prior` adds in float64 and rounds to float32;
  * `log=False` takes `np.log` of the float32 input in float32.
"""
import numpy as np

_STAY = 0


def slip_update(x, slip):
    """viterbi_helpers.pyx:12-35.  x float32 [n >= 3]; returns (from_score float32 [n], from_pos int64 [n])."""
    x = np.asarray(x, dtype=np.float32)
    slip = np.float32(slip)
    n = len(x)
    score = np.zeros(n, dtype=np.float32)
    pos = np.zeros(n, dtype=np.int64)
    score[0] = score[1] = np.float32(-1e38)
    score[2] = x[0] - slip
    for j in range(3, n):
        if score[j - 1] >= x[j - 2]:                       # tie keeps the older source
            best, src = score[j - 1], pos[j - 1]
        else:
            best, src = x[j - 2], j - 2
        score[j] = np.float32(best) - slip
        pos[j] = src
    return score, pos


def map_to_sequence(trans, sequence, slip=None, prior_initial=None, prior_final=None, log=True):
    """transducer.py:14-73.  Returns (score float32, path int32 [nev])."""
    assert slip is None or slip >= 0.0, 'Slip penalty should be non-negative'
    with np.errstate(invalid='ignore'):
        pen = np.float32(np.nan) if slip is None else np.float32(slip)
        trans = np.asarray(trans)
        seq = np.asarray(sequence, dtype=np.int64)
        nev, npos = len(trans), len(seq)
        lt = trans if log else np.log(trans)
        back = np.zeros((nev, npos), dtype=np.int32)
        prev = np.zeros(npos, dtype=np.float32)
        if prior_initial is not None:
            prev += prior_initial                              # float64 add, float32 store
        prev += np.fmax(lt[0][seq], lt[0][_STAY])
        here = np.arange(npos)
        for i in range(1, nev):
            row = lt[i]
            emit = row[seq]
            cur = prev + row[_STAY]                            # stay
            src = here.copy()
            step = prev[:-1] + emit[1:]                        # step from the previous position
            take = step > cur[1:]                              # tie -> stay
            cur[1:][take] = step[take]
            src[1:][take] = here[:-1][take]
            fs, fp = slip_update(prev, pen)                    # slip from any position <= j - 2
            fs = fs + emit
            keep = fs <= cur                                   # tie -> no slip; NaN -> slip
            src = np.where(keep, src, fp)
            cur = np.where(keep, cur, fs)
            back[i] = src
            prev = cur.astype(np.float32)
        if prior_final is not None:
            prev += prior_final
        path = np.empty(nev, dtype=np.int32)
        path[0] = np.argmax(prev)                              # first maximum; NaN counts as the maximum
        score = prev[path[0]]
        for i in range(1, nev):
            path[i] = back[nev - i][path[i - 1]]
        return score, path[::-1].copy()
