"""Drop-in for `sloika.transducer` (`sloika/transducer.py`): Viterbi mapping of a transducer posterior onto a known
sequence -- the remap decode used when training data is prepared (`sloika/tools/chunkify_raw.py:262-274`,
`sloika/batch.py:141-155`).  The dynamic programme runs on the device, one CTA per read (`csrc/remap.cu`);
`map_to_sequence_batch` maps many reads in one launch, `map_to_sequence` keeps the reference's one-read signature.
"""
import numpy as np

from sloika_b200 import cabi
from sloika_b200.config import sloika_dtype

_NEG_LARGE = -50000.0
_STAY = 0


def argmax(*args):
    res = max(enumerate(args), key=lambda x: x[1])
    return res


def _device_of(t):
    import torch
    if isinstance(t, torch.Tensor) and t.is_cuda:
        return t.device
    return torch.device('cuda', torch.cuda.current_device())


def map_to_sequence_batch(trans, sequences, nev=None, slip=None, prior_initial=None, prior_final=None, log=True,
                          return_device=False, timing=None):
    """Map a batch of reads.

    :param trans: `[T, B, nstate]` float32 transducer posteriors (ndarray or torch CUDA tensor, any row strides);
        column 0 is the stay state
    :param sequences: list of B integer sequences (state columns, i.e. k-mer state + 1), or an int32 `[B, P]`
        array together with the lengths in `sequences.npos` semantics (pass a list for ragged input)
    :param nev: events per read (None = T)
    :param slip: slip penalty >= 0, or None -- which, as in the reference, runs slips with a NaN penalty
    :param prior_initial, prior_final: lists / `[B, P]` arrays of float64 priors, or None
    :param log: `trans` is already log-scaled
    :param timing: optional dict; receives `kernel_ms` (CUDA events around the launch, synchronises)

    :returns: (scores float32 [B], list of B int32 path arrays) or the device tensors (scores, paths [B, T])
    """
    import torch
    assert slip is None or slip >= 0.0, 'Slip penalty should be non-negative'
    lib = cabi.load()
    dev = _device_of(trans)
    if not isinstance(trans, torch.Tensor):
        tr = np.asarray(trans)
        if tr.dtype != np.float32:
            tr = tr.astype(sloika_dtype)
        trans = torch.from_numpy(np.ascontiguousarray(tr)).to(dev)
    assert trans.dim() == 3 and trans.dtype == torch.float32, "trans must be float32 [time, batch, state]"
    T, B, S = trans.shape
    if trans.stride(2) != 1 and S > 1:
        trans = trans.contiguous()
    ld_t = trans.stride(0) if T > 1 else B * S
    ld_b = trans.stride(1) if B > 1 else S

    def ragged(rows, dtype, fill):
        rows = [np.asarray(r, dtype=dtype) for r in rows]
        assert len(rows) == B, "one entry per read expected"
        width = max(len(r) for r in rows)
        out = np.full((B, width), fill, dtype=dtype)
        for b, r in enumerate(rows):
            out[b, :len(r)] = r
        return out, np.array([len(r) for r in rows], dtype=np.int32)

    seq, npos = ragged(sequences, np.int32, 0)
    P = seq.shape[1]
    assert npos.min() >= 3, "sequences need at least 3 positions (the reference's slip_update indexes element 2)"
    assert seq.min() >= 0 and seq.max() < S, "sequence entries are columns of trans"
    seq_d = torch.from_numpy(seq).to(dev)
    npos_d = torch.from_numpy(npos).to(dev)
    nev_d = None if nev is None else torch.as_tensor(np.asarray(nev, dtype=np.int32), device=dev)

    def prior(p):
        if p is None:
            return None
        arr, lens = ragged(p, np.float64, 0.0)
        assert np.array_equal(lens, npos), "priors must match the sequence lengths"
        return torch.from_numpy(arr).to(dev)

    p0, p1 = prior(prior_initial), prior(prior_final)
    with torch.cuda.device(dev):
        nbytes = lib.sloika_remap_workspace_bytes(T, B, P)
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        paths = torch.zeros((B, T), dtype=torch.int32, device=dev)
        score = torch.empty(B, dtype=torch.float32, device=dev)
        if timing is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        cabi.check(lib.sloika_remap_fwd(
            cabi.ptr(trans), ld_t, ld_b, cabi.ptr(nev_d), T, B, S, cabi.ptr(seq_d), P, cabi.ptr(npos_d), P,
            0.0 if slip is None else float(slip), 0 if slip is None else 1, cabi.ptr(p0), cabi.ptr(p1), P,
            1 if log else 0, cabi.ptr(ws), nbytes, cabi.ptr(paths), cabi.ptr(score), cabi.stream_ptr(dev)),
            'map_to_sequence')
        if timing is not None:
            ev1.record()
            torch.cuda.synchronize(dev)
            timing['kernel_ms'] = ev0.elapsed_time(ev1)
    if return_device:
        return score, paths
    score_h, paths_h = score.cpu().numpy(), paths.cpu().numpy()
    lens = np.full(B, T, dtype=np.int64) if nev is None else np.asarray(nev)
    return score_h, [paths_h[b, :lens[b]].copy() for b in range(B)]


def map_to_sequence(trans, sequence, slip=None, prior_initial=None, prior_final=None, log=True):
    """  Find Viterbi path through sequence for transducer (`transducer.py:14-73`)

    :param trans: A 2D :class:`nd.array` Transducer to be mapped
    :param sequence: A 1D :class:`nd.array` Sequence of bases to be mapped against
    :param slip: slip penalty (in log-space)
    :param prior_initial: A 1D :class:`nd.array` containing prior over initial position
    :param prior_final: A 1D :class:`nd.array` containing prior over final position
    :param log: Transducer is log-scaled

    :returns: Tuple containing score for path and array containing path
    """
    trans = np.asarray(trans)
    assert trans.ndim == 2, "trans must be [event, state]"
    score, paths = map_to_sequence_batch(
        trans[:, None, :], [sequence], slip=slip,
        prior_initial=None if prior_initial is None else [prior_initial],
        prior_final=None if prior_final is None else [prior_final], log=log)
    return score[0], paths[0]
