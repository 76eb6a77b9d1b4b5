"""Per-step timeline of the fused GRU launch (csrc/gru_fused.cu built with -DGRU_TC_TRACE): clock64() stamps of
cluster 0 -- recurrence CTA (rank 0) and projection CTA (rank 2) -- for scan steps 100..163.  Clocks of different SMs
have different origins, so each CTA's stamps are shown relative to its own first stamp of the step.

    make -C sloika_b200/csrc EXTRA=-DGRU_TC_TRACE && python tools/gru_fused_trace.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sloika_b200 import cabi

NAMES = {0: 'R issuer: h ready', 1: 'R issuer: phase-1 issued', 2: 'R issuer: ring slot s+2 requested', 3: 'R issuer: r*h ready',
         4: 'R issuer: phase-2 issued', 5: 'R compute: vI slot landed', 6: 'R compute: r done', 7: 'R compute: r*h published',
         8: 'R compute: c done', 9: 'R compute: h published',
         10: 'P issuer: operand ready', 11: 'P issuer: MMAs issued', 12: 'P issuer: x(s+2) requested',
         13: 'P compute: MMAs done', 14: 'P compute: operand s+1 staged', 15: 'P compute: accumulators read',
         16: 'P compute: ring slot free', 17: 'P compute: stores issued', 18: 'P compute: writers joined',
         19: 'P compute: published'}


def main():
    lib = cabi.load()
    dev = torch.device('cuda:0')
    T, B, H, I = 800, int(os.environ.get('B', '1024')), 96, 96
    g = torch.Generator().manual_seed(1)
    iW = (torch.randn(3 * H, I, generator=g) * 0.2).to(dev); sW = (torch.randn(2 * H, H, generator=g) * 0.1).to(dev)
    sW2 = (torch.randn(H, H, generator=g) * 0.1).to(dev); b = (torch.randn(3 * H, generator=g) * 0.1).to(dev)
    x = torch.tanh(torch.randn(T, B, I, device=dev)); y = torch.empty(T, B, H, device=dev)
    nws = lib.sloika_gru_fused_workspace_bytes(B, H)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    for _ in range(2):
        rc = lib.sloika_gru_fused_fwd(cabi.ptr(x), I, cabi.ptr(iW), cabi.ptr(sW), cabi.ptr(sW2), cabi.ptr(b), cabi.ptr(y), H,
                                      cabi.ptr(ws), nws, None, T, B, I, H, 0, 1, 2, cabi.stream_ptr(dev))
        assert rc == 0, rc
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (64 * 24))()
    ctypes.CDLL(cabi.LIB_PATH).sloika_debug_gru_fused_trace(buf)
    t = np.array(buf, dtype=np.int64).reshape(64, 24)
    for first, ids in ((0, range(0, 10)), (13, range(10, 20))):
        step = np.diff(t[:, first])
        print("cycles per step ({}): median {:.0f}  min {:.0f}  max {:.0f}".format(NAMES[first], np.median(step), step.min(), step.max()))
        rel = (t - t[:, first:first + 1])[1:-1]
        for k in sorted(ids, key=lambda k: np.median(rel[:, k])):
            print("   {:36s} +{:6.0f}".format(NAMES[k], np.median(rel[:, k])))


if __name__ == '__main__':
    main()
