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
        src = os.path.join(_HERE, 'viterbi_ref.c')
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
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
