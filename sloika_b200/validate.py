"""Forward-only scoring of a network on labelled chunks (`bin/validate_network.py:46-54`): mean categorical
cross-entropy and number of correct argmax calls of one batch, computed on the device from the posteriors the
forward pass leaves there (`csrc/validate.cu`)."""
import numpy as np

from sloika_b200 import cabi


def remove_blanks(labels):
    """Blank labels inherit the previous label (`validate_network.py:37-42`; non-transducer models)."""
    for lbl_ch in labels:
        for i in range(1, len(lbl_ch)):
            if lbl_ch[i] == 0:
                lbl_ch[i] = lbl_ch[i - 1]
    return labels


def wrap_network(calc_post):
    """`fv(events [T, B, F] float32, labels [T, B] int32) -> (loss, ncorrect)` as the Theano function of
    `validate_network.py:45-54` returns them."""
    import torch

    def fv(events, labels):
        lib = cabi.load()
        dev = calc_post.device
        x = torch.from_numpy(np.ascontiguousarray(events, dtype=np.float32)).to(dev)
        post = calc_post.forward_device(x).data
        T, B, S = post.shape
        lab = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.int32)).to(dev)
        assert lab.shape == (T, B), "labels must be [time, batch] like the posteriors"
        ld = post.stride(1) if B > 1 else (post.stride(0) if T > 1 else S)
        assert post.stride(2) == 1 and (T == 1 or B == 1 or post.stride(0) == B * ld)
        loss = torch.zeros(1, dtype=torch.float64, device=dev)
        ncorr = torch.zeros(1, dtype=torch.int64, device=dev)
        cabi.check(lib.sloika_score_fwd(cabi.ptr(post), ld, cabi.ptr(lab), T * B, S, cabi.ptr(loss), cabi.ptr(ncorr),
                                        cabi.stream_ptr(dev)), 'score')
        return float(loss.item()) / (T * B), int(ncorr.item())
    return fv
