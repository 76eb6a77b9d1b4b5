"""The reference's non-transducer decoder on the B200 (`sloika/olddecode.py:13-118`), reached from
`basecall.decode_post` for models that are not transducers (`sloika/basecall.py:47-50`).

Same names and signatures; the dynamic programming and the per-event sums run in `csrc/olddecode.cu`
(`sloika_olddecode_fwd`, `sloika_transitions_fwd`), there is no host implementation.  `post` may be a NumPy array or
a torch CUDA tensor `[T, K]`; results come back as the reference returns them: `(score, int64 state per event)`.
"""
import itertools

import numpy as np

from sloika_b200 import cabi

_ETA = 1e-10
_NSTEP = 4
_NSKIP = 16
_STEP_FACTOR = np.log(_NSTEP)
_SKIP_FACTOR = np.log(_NSKIP)


def _device_post(post):
    import torch
    if isinstance(post, np.ndarray):
        if not torch.cuda.is_available():
            raise cabi.SloikaB200Error("no CUDA device: the B200 decode has no CPU fallback")
        post = torch.from_numpy(np.ascontiguousarray(post, dtype=np.float32)).cuda()
    if post.dtype != torch.float32:
        post = post.float()
    assert post.dim() == 2, "post must be [events, states]"
    return post.contiguous()


def decode_profile(post, trans=None, log=False, slip=0.0):
    """  Viterbi-style decoding with per-event transition weights (`olddecode.py:13-73`)

    :param post: posterior probabilities of kmers by event `[T, K]`
    :param trans: per-event log-scaled weights `[T, 3]` (stay, step, skip) or any iterable of such rows; None == no
        transition weights
    :param log: posterior probabilities are in log-space
    :returns: (score, state for every event)
    """
    import torch
    lib = cabi.load()
    post = _device_post(post)
    T, K = post.shape
    dev = post.device
    ltrans = None
    if trans is not None:
        if isinstance(trans, torch.Tensor):
            rows = trans.to(device=dev, dtype=torch.float64).reshape(-1, 3)[:max(T - 1, 0)].clone()
        else:
            if not isinstance(trans, np.ndarray):
                trans = np.array(list(itertools.islice(iter(trans), max(T - 1, 0))), dtype=np.float64).reshape(-1, 3)
            rows = torch.from_numpy(np.array(trans, dtype=np.float64, copy=True)[:max(T - 1, 0)]).to(dev)
        rows[:, 1] -= _STEP_FACTOR                              # olddecode.py:31-32
        rows[:, 2] -= _SKIP_FACTOR
        ltrans = torch.zeros((T, 3), dtype=torch.float64, device=dev)
        ltrans[:rows.shape[0]] = rows
    nbytes = lib.sloika_olddecode_workspace_bytes(T, 1, K)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    seq = torch.empty((1, T), dtype=torch.int32, device=dev)
    score = torch.empty(1, dtype=torch.float64, device=dev)
    cabi.check(lib.sloika_olddecode_fwd(cabi.ptr(post), K, T * K, cabi.ptr(ltrans), T * 3, None, T, 1, K, float(slip),
                                        1 if log else 0, cabi.ptr(ws), nbytes, cabi.ptr(seq), cabi.ptr(score),
                                        cabi.stream_ptr(dev)), 'olddecode')
    return np.float64(score.item()), seq[0].cpu().numpy().astype(np.int64)


def decode_transition(post, trans, log=False, slip=0.0):
    """  Viterbi-style decoding with weighted transitions (`olddecode.py:76-83`)"""
    return decode_profile(post, trans=itertools.repeat(trans), log=log, slip=slip)


def decode_simple(post, log=False, slip=0.0):
    """  Viterbi-style decoding with uniform transitions (`olddecode.py:86-91`)"""
    return decode_profile(post, log=log, slip=slip)


def estimate_transitions(post, trans=None, return_device=False):
    """  Naive estimate of transition behaviour from posteriors (`olddecode.py:94-118`)

    :param post: posterior probabilities of kmers by event
    :param trans: prior belief of transition behaviour (None = use global estimate)
    :returns: float64 `[T, 3]` (NumPy, or the device tensor with `return_device`)
    """
    import torch
    assert trans is None or len(trans) == 3, 'Incorrect number of transitions'
    lib = cabi.load()
    post = _device_post(post)
    T, K = post.shape
    res = torch.full((T, 3), _ETA, dtype=torch.float64, device=post.device)
    cabi.check(lib.sloika_transitions_fwd(cabi.ptr(post), K, T, K, cabi.ptr(res), cabi.stream_ptr(post.device)),
               'estimate_transitions')
    if trans is None:
        prior = res.sum(0)
        prior = prior / prior.sum()
    else:
        prior = torch.as_tensor(np.asarray(trans, dtype=np.float64), device=post.device)
    res = res * prior
    res = res / res.sum(1, keepdim=True)
    return res if return_device else res.cpu().numpy()
