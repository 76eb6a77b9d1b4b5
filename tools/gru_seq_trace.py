"""Per-step timeline of the sequences-on-lanes GRU launch (csrc/gru_seq.cu built with -DGRU_TC_TRACE): clock64() stamps of
CTA 0 for scan steps 100..163, relative to the issuing warp's "h ready".

    make -C sloika_b200/csrc EXTRA=-DGRU_TC_TRACE && python tools/gru_seq_trace.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
LAYOUT = int(os.environ.get('LAYOUT', '0'))   # sloika_gru_seq_fwd: bit 0 x blocked, bit 1 y blocked (timing only here)

from sloika_b200 import cabi

NAMES = {0: 'compute: r, z complete (d1)', 1: 'compute: r*h published', 2: 'compute: x operand free (dx)',
         3: 'compute: z read, x(s+1) operand published', 4: 'compute: candidate complete (d2)', 5: 'compute: h published',
         6: 'compute: x(s+1) operand stored', 7: 'compute: x(s+2) loads issued', 14: 'compute: z denominators done',
         8: 'issuer: h ready', 9: 'issuer: phase 1 + c projection issued', 10: 'issuer: r*h ready',
         11: 'issuer: phase 2 issued', 12: 'issuer: z | r accumulators free', 13: 'issuer: z | r projection issued'}


def main():
    lib = cabi.load()
    dev = torch.device('cuda:0')
    T, B, H, I = 800, int(os.environ.get('B', '1024')), 96, 96
    g = torch.Generator().manual_seed(1)
    iW = (torch.randn(3 * H, I, generator=g) * 0.2).to(dev); sW = (torch.randn(2 * H, H, generator=g) * 0.1).to(dev)
    sW2 = (torch.randn(H, H, generator=g) * 0.1).to(dev); b = (torch.randn(3 * H, generator=g) * 0.1).to(dev)
    x = torch.tanh(torch.randn(T, B, I, device=dev)); y = torch.empty(T, B, H, device=dev)
    for _ in range(2):
        rc = lib.sloika_gru_seq_fwd(cabi.ptr(x), I, cabi.ptr(iW), cabi.ptr(sW), cabi.ptr(sW2), cabi.ptr(b), cabi.ptr(y), H,
                                    None, T, B, I, H, 0, 1, 2, LAYOUT, cabi.stream_ptr(dev))
        assert rc == 0, rc
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (64 * 16))()
    ctypes.CDLL(cabi.LIB_PATH).sloika_debug_gru_seq_trace(buf)
    t = np.array(buf, dtype=np.int64).reshape(64, 16)
    step = np.diff(t[:, 8])
    print("cycles per step (issuer: h ready): median {:.0f}  min {:.0f}  max {:.0f}".format(np.median(step), step.min(), step.max()))
    rel = (t - t[:, 8:9])[1:-1]
    for k in sorted(NAMES, key=lambda k: np.median(rel[:, k])):
        print("   {:44s} +{:6.0f}".format(NAMES[k], np.median(rel[:, k])))


if __name__ == '__main__':
    main()
