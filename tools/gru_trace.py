"""Per-step timeline of the tensor-memory GRU recurrence (csrc/gru_tc.cu built with -DGRU_TC_TRACE):
clock64() stamps of the issuing warp and of one compute warp of CTA 0, steps 100..163.

    make -C sloika_b200/csrc clean && make -C sloika_b200/csrc EXTRA=-DGRU_TC_TRACE && python tools/gru_trace.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sloika_b200 import cabi, engine, layers
from sloika_b200 import module_tools as smt

NAMES = ['issuer: h ready', 'issuer: phase-1 issued', 'issuer: r*h ready', 'issuer: phase-2 issued',
         'compute: r done (d1)', 'compute: r loaded', 'compute: r*h published', 'compute: c done (d2)',
         'compute: c loaded', 'compute: h published', 'compute: r*h computed', 'compute: r*h stored', 'compute: r*h fenced']


def main():
    lib = cabi.load()
    np.random.seed(1)
    g = layers.Gru(96, 96, init=smt.partial(smt.truncated_normal, sd=0.5), has_bias=True)
    x = torch.tanh(torch.randn((800, 1024, 96), device='cuda'))
    for cfg in (os.environ.get('SLOIKA_B200_GRU_TC', '1,8'),):
        engine.run_gru(g, engine.Act(x))
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * (64 * 16))()
        ctypes.CDLL(cabi.LIB_PATH).sloika_debug_gru_trace(buf)
        t = np.array(buf, dtype=np.int64).reshape(64, 16)[:, :13]
        base = t[:, 0:1]
        rel = (t - base)[1:-1]
        step = np.diff(t[:, 0])
        print("config {}: cycles per step (issuer h-ready to h-ready): median {:.0f}".format(cfg, np.median(step)))
        order = np.argsort(np.median(rel, axis=0))
        for k in order:
            print("   {:28s} +{:6.0f}".format(NAMES[k], np.median(rel[:, k])))


if __name__ == '__main__':
    main()
