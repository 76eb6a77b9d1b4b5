"""ctypes binding of oracle/_build/liboracle.so (the C Viterbi restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE])
    return os.path.join(_HERE, '_build', 'liboracle.so')


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, '_build', 'liboracle.so')
        srcs = [os.path.join(_HERE, f) for f in ('viterbi_ref.c', 'remap_ref.c')]
        if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(f) for f in srcs):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.sloika_oracle_viterbi_batch.restype = ctypes.c_int
    return _LIB


def viterbi_batch(lpost, lengths=None, klen=5, nbase=4, skip_pen=0.0, nthreads=None):
    """lpost [T, B, S] float32 log-posteriors -> (scores[B], list of paths)."""
    lpost = np.ascontiguousarray(lpost, dtype=np.float32)
    T, B, S = lpost.shape
    assert S == nbase ** klen + 1
    paths = np.zeros((B, T), dtype=np.int32)
    plen = np.zeros(B, dtype=np.int32)
    score = np.zeros(B, dtype=np.float32)
    if lengths is not None:
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
    rc = lib().sloika_oracle_viterbi_batch(
        lpost.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(T), ctypes.c_long(B),
        lengths.ctypes.data_as(ctypes.c_void_p) if lengths is not None else None,
        ctypes.c_int(nbase), ctypes.c_int(klen), ctypes.c_float(skip_pen),
        ctypes.c_int(nthreads or os.cpu_count() or 1),
        paths.ctypes.data_as(ctypes.c_void_p), plen.ctypes.data_as(ctypes.c_void_p),
        score.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError("oracle viterbi failed")
    return score, [paths[b, :plen[b]].tolist() for b in range(B)]


def slip_update(x, slip):
    """viterbi_helpers.slip_update through the C restatement: (from_score float32, from_pos int64)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    fs = np.zeros(len(x), dtype=np.float32)
    fp = np.zeros(len(x), dtype=np.int64)
    lib().sloika_oracle_slip_update(x.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(len(x)), ctypes.c_float(slip),
                                    fs.ctypes.data_as(ctypes.c_void_p), fp.ctypes.data_as(ctypes.c_void_p))
    return fs, fp


def remap_batch(lt, seq, nev=None, npos=None, slip=None, prior0=None, prior1=None, nthreads=None):
    """lt [T, B, S] float32 LOG transducer, seq [B, P] int32 state columns -> (scores [B], paths [B, T] int32)."""
    lt = np.ascontiguousarray(lt, dtype=np.float32)
    seq = np.ascontiguousarray(seq, dtype=np.int32)
    T, B, S = lt.shape
    P = seq.shape[1]
    opt = lambda a, dt: None if a is None else np.ascontiguousarray(a, dtype=dt)
    nev, npos, prior0, prior1 = opt(nev, np.int32), opt(npos, np.int32), opt(prior0, np.float64), opt(prior1, np.float64)
    ptr = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    paths = np.zeros((B, T), dtype=np.int32)
    score = np.zeros(B, dtype=np.float32)
    fn = lib().sloika_oracle_remap_batch
    fn.restype = ctypes.c_int
    rc = fn(ptr(lt), ctypes.c_long(T), ctypes.c_long(B), ctypes.c_long(S), ptr(nev), ptr(seq), ctypes.c_long(P), ptr(npos),
            ctypes.c_float(float('nan') if slip is None else slip), ptr(prior0), ptr(prior1),
            ctypes.c_int(nthreads or os.cpu_count() or 1), ptr(paths), ptr(score))
    if rc != 0:
        raise RuntimeError("oracle remap failed")
    return score, paths
