"""ctypes binding of `libsloika_b200.so` (the C ABI declared in `include/sloika_b200.h`).

The library is built in-tree by `__graft_entry__.build()` / `make -C sloika_b200/csrc`.  There is no
fallback: if it is missing, `load()` raises, and every device operator of this package goes through
it.  ctypes releases the GIL for the duration of each call; calls only enqueue work on a stream.
"""
import ctypes
import os

ABI_VERSION = 1
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libsloika_b200.so')

_p = ctypes.c_void_p
_i = ctypes.c_int
_l = ctypes.c_long
_d = ctypes.c_double
_z = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/sloika_b200.h one to one
EXPORTS = {
    'sloika_b200_abi_version': (_i, []),
    'sloika_b200_strerror': (ctypes.c_char_p, [_i]),
    'sloika_b200_device_info': (_i, [_p, _p, _p]),
    'sloika_conv1d_fwd': (_i, [_p, _p, _p, _p, _l, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    'sloika_conv1d_fwd_ex': (_i, [_p, _p, _p, _p, _l, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    'sloika_linear_fwd_gated': (_i, [_p, _l, _p, _p, _p, _l, _l, _i, _i, _i, _p, ctypes.c_float, _p]),
    'sloika_linear_fwd': (_i, [_p, _l, _p, _p, _p, _l, _l, _i, _i, _i, _p]),
    'sloika_linear_fwd_ex': (_i, [_p, _l, _p, _p, _p, _l, _l, _i, _i, _i, _i, _p]),
    'sloika_softmax_fwd': (_i, [_p, _l, _p, _p, _p, _l, _l, _i, _i, _p]),
    'sloika_softmax_slices': (_i, [_i, _i, _i]),
    'sloika_softmax_logits_fwd': (_i, [_p, _l, _p, _p, _p, _l, _p, _l, _i, _i, _i, _i, _p]),
    'sloika_softmax_logits_blocked_fwd': (_i, [_p, _p, _p, _p, _l, _p, _l, _i, _i, _i, _p]),
    'sloika_softmax_normalise_fwd': (_i, [_p, _l, _p, _i, _l, _i, _p]),
    'sloika_gru_workspace_bytes': (_z, [_i, _i, _i]),
    'sloika_gru_fwd': (_i, [_p, _l, _p, _p, _p, _p, _p, _l, _p, _z, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    'sloika_gru_recurrence_fwd': (_i, [_p, _l, _p, _p, _p, _l, _p, _i, _i, _i, _i, _i, _i, _p]),
    'sloika_b200_set_gemm_sm_budget': (_i, [_i]),
    'sloika_lstm_recurrence_fwd': (_i, [_p, _l, _p, _p, _p, _l, _p, _i, _i, _i, _i, _i, _i, _p]),
    'sloika_olddecode_workspace_bytes': (_z, [_i, _i, _i]),
    'sloika_olddecode_fwd': (_i, [_p, _l, _l, _p, _l, _p, _i, _i, _i, ctypes.c_double, _i, _p, _z, _p, _p, _p]),
    'sloika_transitions_fwd': (_i, [_p, _l, _i, _i, _p, _p]),
    'sloika_score_fwd': (_i, [_p, _l, _p, _l, _i, _p, _p, _p]),
    'sloika_window_fwd': (_i, [_p, _l, _p, _l, _p, _i, _i, _i, _i, _p]),
    'sloika_gru_fused_workspace_bytes': (_z, [_i, _i]),
    'sloika_gru_fused_fwd': (_i, [_p, _l, _p, _p, _p, _p, _p, _l, _p, _z, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    'sloika_gru_fwd_gated': (_i, [_p, _l, _p, _p, _p, _p, _p, _l, _p, _l, _p, _z, _p, _i, _i, _i, _i, _i, _i, _i, _l, _p,
                                  ctypes.c_float, _p]),
    'sloika_blocked_bytes': (_z, [_i, _i, _i]),
    'sloika_block_layout_fwd': (_i, [_p, _p, _l, _i, _i, _i, _i, _p]),
    'sloika_gru_seq_fwd_gated': (_i, [_p, _l, _i, _p, _p, _p, _p, _p, _p, _p, _l, _p, _l, _p, _i, _i, _i, _i, _i, _i, _i, _l, _p,
                                      ctypes.c_float, _p]),
    'sloika_gru_seq_fwd': (_i, [_p, _l, _p, _p, _p, _p, _p, _l, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    'sloika_gru_recurrence_fwd_ex': (_i, [_p, _l, _p, _p, _p, _l, _p, _i, _i, _i, _i, _i, _i, _l, _p]),
    'sloika_viterbi_workspace_bytes': (_z, [_i, _i, _i, _i]),
    'sloika_viterbi_fwd': (_i, [_p, _l, _l, _p, _i, _i, _i, _i, _d, _d, _i, _p, _z, _p, _p, _p, _p]),
    'sloika_viterbi_logits_fwd': (_i, [_p, _l, _l, _p, _i, _p, _i, _i, _i, _i, _d, _d, _p, _z, _p, _p, _p, _p]),
    'sloika_remap_workspace_bytes': (_z, [_i, _i, _i]),
    'sloika_remap_fwd': (_i, [_p, _l, _l, _p, _i, _i, _i, _p, _l, _p, _i, _d, _i, _p, _p, _l, _i, _p, _z, _p, _p, _p]),
    'sloika_slip_update_fwd': (_i, [_p, _i, ctypes.c_float, _p, _p, _p]),
    'sloika_prepare_workspace_bytes': (_z, [_l, _i, _i]),
    'sloika_prepare_signal_fwd': (_i, [_p, _p, _i, _l, _i, _i, _d, _i, _p, _z, _p, _l, _i, _p, _p]),
    'sloika_path_to_bases_fwd': (_i, [_p, _l, _p, _i, _i, _i, _i, _p, _p, _l, _p, _p]),
}

SLOIKA_VIT_POST, SLOIKA_VIT_LOG = 0, 1

_lib = None


class SloikaB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare every prototype.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SloikaB200Error(
            "{} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the B200 path)".format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.sloika_b200_abi_version() != ABI_VERSION:
        raise SloikaB200Error("ABI version mismatch: library {} vs binding {}".format(
            lib.sloika_b200_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().sloika_b200_strerror(code).decode('ascii', 'replace')
        raise SloikaB200Error("{} failed: {} (code {})".format(what, msg, code))


def ptr(tensor):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if tensor is None else ctypes.c_void_p(tensor.data_ptr())


def stream_ptr(device=None):
    """Current torch CUDA stream of `device` as a void* for the ABI."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
