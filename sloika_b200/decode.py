"""Transducer Viterbi decode on the B200 (reference `sloika/decode.py:21-93`).

`prepare_post` and `viterbi` keep the reference signatures so `basecall.decode_post` reads the same;
`viterbi_batch` is the batched device entry (one CTA per read) that the GPU basecall path uses --
it fuses `prepare_post` (the min_prob floor) and the `log(post + 1e-10)` of `decode.py:56` into the
kernel and never materialises the int32 traceback of the reference.

All dynamic programming happens in `sloika_viterbi_fwd` (csrc/viterbi.cu); there is no host
implementation here.  Arithmetic is float32 (the dtype of real posteriors); float64 input, which the
reference would decode in float64, is cast to float32.
"""
import numpy as np

from sloika_b200 import cabi
from sloika_b200 import variables as sv

_ETA = 1e-10


def prepare_post(post, min_prob=1e-5, drop_bad=False):
    """ Sanitised posterior matrix for decoding (`decode.py:21-36`)

    Host/NumPy or torch input `[T, 1, S]`; returns `[T, S]` (`[T', S-1]` renormalised when
    `drop_bad`).  The device basecall path does not call this: `viterbi_batch(min_prob=...)` applies
    the same floor inside the kernel.
    """
    is_np = isinstance(post, np.ndarray)
    post = post.squeeze(1)
    if drop_bad:
        keep = post.argmax(1) > 0
        post = post[keep][:, 1:]
        post = post / post.sum(1, keepdims=True) if is_np else post / post.sum(1, keepdim=True)
    return min_prob + (1.0 - min_prob) * post


def _as_device(post, device=None):
    import torch
    if isinstance(post, np.ndarray):
        if not torch.cuda.is_available():
            raise cabi.SloikaB200Error("no CUDA device: the B200 decode has no CPU fallback")
        dev = device or torch.device('cuda', torch.cuda.current_device())
        return torch.from_numpy(np.ascontiguousarray(post, dtype=np.float32)).to(dev)
    if post.dtype != torch.float32:
        post = post.float()
    return post


def viterbi_batch(post, lengths=None, klen=5, skip_pen=0.0, min_prob=1e-5, nbase=4, log=False,
                  return_device=False):
    """Best paths of a batch of reads.

    :param post: `[T, B, S]` float32 posteriors (torch CUDA tensor, any row strides, or ndarray);
        raw network output -- the min_prob floor of `prepare_post` is applied in the kernel
    :param lengths: events per read (None = T for all)
    :param min_prob: floor; pass 0.0 if `post` already went through `prepare_post`
    :param log: `post` holds log-probabilities (`decode.py:56` `log=True`); min_prob is ignored

    :returns: (scores float32[B], list of B lists of k-mer states) -- or the device tensors
        `(scores, paths[B, T], path_len[B])` when `return_device`
    """
    import torch
    from sloika_b200 import engine
    lib = cabi.load()
    if isinstance(post, engine.LogitsAct):
        return _viterbi_logits(post, klen, skip_pen, min_prob, nbase, return_device)
    if isinstance(post, engine.Act):
        post, lengths = post.data, (post.lengths if lengths is None else lengths)
    post = _as_device(post)
    assert post.dim() == 3, "post must be [time, batch, state]"
    T, B, S = post.shape
    assert klen >= 3, "Kmer not long enough to apply Viterbi with skips"
    assert sv.nstate(klen, transducer=True, nbase=nbase) == S
    if post.stride(2) != 1 and S > 1:
        post = post.contiguous()
    # strides of size-1 dimensions are arbitrary in torch: only pass strides that are stepped over
    ld_t = post.stride(0) if T > 1 else B * S
    ld_b = post.stride(1) if B > 1 else S
    dev = post.device
    lens = None
    if lengths is not None:
        lens = torch.as_tensor(lengths, dtype=torch.int32, device=dev).contiguous()
    with torch.cuda.device(dev):
        nbytes = lib.sloika_viterbi_workspace_bytes(T, B, nbase, klen)
        tb = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        paths = torch.empty((B, max(T, 1)), dtype=torch.int32, device=dev)
        plen = torch.empty(B, dtype=torch.int32, device=dev)
        score = torch.empty(B, dtype=torch.float32, device=dev)
        from sloika_b200.engine import launch
        launch('viterbi', 1, lib.sloika_viterbi_fwd,
               cabi.ptr(post), ld_t, ld_b, cabi.ptr(lens), T, B, nbase, klen,
               float(skip_pen), float(min_prob), cabi.SLOIKA_VIT_LOG if log else cabi.SLOIKA_VIT_POST,
               cabi.ptr(tb), nbytes, cabi.ptr(paths), cabi.ptr(plen), cabi.ptr(score),
               cabi.stream_ptr(dev))
    if return_device:
        return score, paths, plen
    score_h = score.cpu().numpy()
    plen_h = plen.cpu().numpy()
    paths_h = paths.cpu().numpy()
    return score_h, [paths_h[b, :plen_h[b]].tolist() for b in range(B)]


def _viterbi_logits(la, klen, skip_pen, min_prob, nbase, return_device):
    """Decode straight from the final layer's logits + row statistics (`engine.LogitsAct`)."""
    import torch
    from sloika_b200.engine import launch
    lib = cabi.load()
    data = la.data
    T, B, S = data.shape
    assert klen >= 3, "Kmer not long enough to apply Viterbi with skips"
    assert sv.nstate(klen, transducer=True, nbase=nbase) == S
    dev = data.device
    ld_t = data.stride(0) if T > 1 else B * data.stride(1)
    ld_b = data.stride(1)
    with torch.cuda.device(dev):
        nbytes = lib.sloika_viterbi_workspace_bytes(T, B, nbase, klen)
        tb = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        paths = torch.empty((B, max(T, 1)), dtype=torch.int32, device=dev)
        plen = torch.empty(B, dtype=torch.int32, device=dev)
        score = torch.empty(B, dtype=torch.float32, device=dev)
        launch('viterbi', 2, lib.sloika_viterbi_logits_fwd,          # row statistics + decode
               cabi.ptr(data), ld_t, ld_b, cabi.ptr(la.stats), la.n_slices, cabi.ptr(la.lengths), T, B, nbase, klen,
               float(skip_pen), float(min_prob), cabi.ptr(tb), nbytes, cabi.ptr(paths), cabi.ptr(plen),
               cabi.ptr(score), cabi.stream_ptr(dev))
    if return_device:
        return score, paths, plen
    score_h, plen_h, paths_h = score.cpu().numpy(), plen.cpu().numpy(), paths.cpu().numpy()
    return score_h, [paths_h[b, :plen_h[b]].tolist() for b in range(B)]


def viterbi(post, klen, skip_pen=0.0, log=False, nbase=4):
    """  Viterbi decoding of a kmer transducer (`decode.py:39-93`)

    :param post: A 2d array `[events, states]`, already through `prepare_post`
    :param klen: Length of kmer
    :param log: post array is in log space

    :returns: (score, list of k-mer states)
    """
    if isinstance(post, np.ndarray):
        post3 = post[:, None, :]
    else:
        post3 = post.unsqueeze(1)
    score, paths = viterbi_batch(post3, None, klen=klen, skip_pen=skip_pen, min_prob=0.0, nbase=nbase, log=log)
    return score[0], paths[0]


def paths_to_sequences(paths, path_len, kmer_len=5, alphabet='ACGT', always_move=True):
    """Base sequences of a batch of best paths, assembled on the device.

    Same strings as `bio.kmers_to_sequence([kmers[s] for s in path], always_move)` (`bio.py:228-237`), which
    `SeqPrinter.write` computes per read on the host (`basecall.py:141-149`).

    :param paths: int32 `[B, T]` k-mer states (device tensor from `viterbi_batch(..., return_device=True)`, or array)
    :param path_len: int32 `[B]` valid entries per read
    :param always_move: transducer convention -- a k-mer followed by itself is a move, not a stay

    :returns: list of B `str`
    """
    import torch
    lib = cabi.load()
    if isinstance(alphabet, bytes):
        alphabet = alphabet.decode('ascii')
    if not isinstance(paths, torch.Tensor):
        paths = torch.as_tensor(np.ascontiguousarray(paths, dtype=np.int32))
    if not isinstance(path_len, torch.Tensor):
        path_len = torch.as_tensor(np.ascontiguousarray(path_len, dtype=np.int32))
    dev = paths.device if paths.is_cuda else torch.device('cuda', torch.cuda.current_device())
    paths = paths.to(dev, dtype=torch.int32).contiguous()
    path_len = path_len.to(dev, dtype=torch.int32).contiguous()
    B, T = paths.shape
    if B == 0:
        return []
    ld_out = max(1, kmer_len * T)
    with torch.cuda.device(dev):
        alpha = torch.tensor(list(alphabet.encode('ascii')), dtype=torch.uint8, device=dev)
        out = torch.empty((B, ld_out), dtype=torch.uint8, device=dev)
        out_len = torch.empty(B, dtype=torch.int32, device=dev)
        from sloika_b200.engine import launch
        launch('path_to_bases', 1, lib.sloika_path_to_bases_fwd,
               cabi.ptr(paths), paths.stride(0), cabi.ptr(path_len), B, kmer_len, len(alphabet),
               1 if always_move else 0, cabi.ptr(alpha), cabi.ptr(out), ld_out, cabi.ptr(out_len),
               cabi.stream_ptr(dev))
    lens = out_len.cpu().numpy()
    width = int(lens.max()) if B else 0
    chars = out[:, :max(width, 1)].cpu().numpy()
    return [chars[b, :lens[b]].tobytes().decode('ascii') for b in range(B)]
