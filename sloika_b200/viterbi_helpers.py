"""Drop-in for `sloika.viterbi_helpers` (the reference's Cython module, `sloika/viterbi_helpers.pyx`)."""
import numpy as np

from sloika_b200 import cabi


def slip_update(x, slip):
    """Score and source of a geometric slip into every position (`viterbi_helpers.pyx:12-35`).

    :param x: 1D float32 array (length >= 3; the reference indexes element 2 unconditionally)
    :param slip: slip penalty (log space)

    :returns: (from_score float32 [n], from_pos int64 [n]) -- computed on the device
    """
    import torch
    x = np.asarray(x)
    if x.dtype != np.float32 or x.ndim != 1:
        raise ValueError("Buffer dtype mismatch, expected 'DTYPE_t' (float32) 1D array")   # what Cython raises
    if len(x) < 3:
        raise IndexError("slip_update needs at least 3 positions")
    lib = cabi.load()
    dev = torch.device('cuda', torch.cuda.current_device())
    xd = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    fs = torch.empty(len(x), dtype=torch.float32, device=dev)
    fp = torch.empty(len(x), dtype=torch.int64, device=dev)
    cabi.check(lib.sloika_slip_update_fwd(cabi.ptr(xd), len(x), float(np.float32(slip)), cabi.ptr(fs), cabi.ptr(fp),
                                          cabi.stream_ptr(dev)), 'slip_update')
    return fs.cpu().numpy(), fp.cpu().numpy()
