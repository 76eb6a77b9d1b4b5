"""NumPy restatement of the transducer Viterbi decode.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows `sloika/decode.py:21-36` (`prepare_post`) and `:39-93` (`viterbi`), called from
`sloika/basecall.py:26-51` (`decode_post`).  PINNED: `tools/make_golden.py` compares it, bit for
bit (scores and paths), with the reference functions imported from /root/reference on the
known-answer matrices of `test/unit/test_decode.py:233-256` and on seeded random posteriors;
`tests/test_oracle.py` replays those golden vectors.

State j in [0, K) is k-mer j (column j+1 of the posterior); column 0 is "stay".  With
p = previous scores (`decode.py:60-82`):
    step[j] = max_{a<nb}    p[a*K/nb    + j//nb   ]          argmax = first maximum (lowest a)
    skip[j] = max_{a<nb^2}  p[a*K/nb^2  + j//nb^2 ] - skip_pen
    move[j] = lpost[j+1] + max(step, skip)          from = step if step > skip else skip  (tie -> skip)
    stay[j] = p[j] + lpost[0]
    v[j]    = max(move, stay)                       traceback = from if move > stay else -1 (tie -> stay)
Arithmetic stays in the dtype of `post` (float32 on the real path: `skip_pen` is a python float and
does not upcast).
"""
import numpy as np

_ETA = 1e-10


def prepare_post(post, min_prob=1e-5, drop_bad=False):
    """decode.py:21-36: drop the batch axis, optionally drop bad-state rows, floor probabilities."""
    post = np.squeeze(post, axis=1)
    if drop_bad:
        keep = np.argmax(post, axis=1) > 0
        post = post[keep, 1:]
        post = post / np.sum(post, axis=1, keepdims=True)
    return min_prob + (1.0 - min_prob) * post


def log_post(post):
    """decode.py:56: lpost = log(post + 1e-10), in the dtype of post."""
    return np.log(post + _ETA)


def viterbi(post, klen, skip_pen=0.0, log=False, nbase=4, return_traceback=False):
    """decode.py:39-93.  Returns (score, path) and, optionally, the int32 traceback matrix."""
    nev, nst = post.shape
    assert klen >= 3, "Kmer not long enough to apply Viterbi with skips"
    K = nbase ** klen
    assert nst == K + 1
    nstep, nskip = nbase, nbase * nbase
    rstep, rskip = K // nstep, K // nskip

    lpost = post if log else log_post(post)
    v = lpost[0, 1:].copy()
    tb = np.empty((nev, K), dtype=np.int32)
    tb[0] = -1
    rem_step = np.arange(rstep)
    rem_skip = np.arange(rskip)
    for i in range(1, nev):
        p = v
        ps = p.reshape(nstep, rstep)
        best_a = np.argmax(ps, axis=0)                       # first maximum
        score_step = np.repeat(ps[best_a, rem_step], nstep)
        from_step = np.repeat(rstep * best_a + rem_step, nstep)
        pk = p.reshape(nskip, rskip)
        best_b = np.argmax(pk, axis=0)
        score_skip = np.repeat(pk[best_b, rem_skip], nskip) - skip_pen
        from_skip = np.repeat(rskip * best_b + rem_skip, nskip)

        move = lpost[i, 1:] + np.maximum(score_step, score_skip)
        src = np.where(score_step > score_skip, from_step, from_skip)
        stay = p + lpost[i, 0]
        tb[i] = np.where(move > stay, src, -1)
        v = np.maximum(move, stay)

    cur = int(np.argmax(v))
    path = [cur]
    for i in range(nev - 1, 0, -1):
        t = int(tb[i, cur])
        if t >= 0:
            path.append(t)
            cur = t
    path.reverse()
    score = v[int(np.argmax(v))] if nev > 0 else None
    if return_traceback:
        return score, path, tb
    return score, path


def decode_post(post, kmer_len, min_prob, skip=5.0, nbase=4):
    """The transducer branch of `basecall.decode_post` (basecall.py:43-46)."""
    assert post.shape[2] == nbase ** kmer_len + 1
    return viterbi(prepare_post(post, min_prob=min_prob, drop_bad=False), kmer_len,
                   skip_pen=skip, nbase=nbase)
