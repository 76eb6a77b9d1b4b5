"""Time the device pre-processing (csrc/prepare.cu) against the NumPy path on whole-read-sized signals."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sloika_b200 import basecall, engine

DEV = torch.device('cuda:0')
rng = np.random.default_rng(4)
for B, N in ((256, 60000), (1024, 40000), (64, 400000)):
    sigs = [rng.standard_normal(N + int(rng.integers(0, 1000))) * 8 + 95 for _ in range(B)]
    for _ in range(2):
        basecall.prepare_signals_device(sigs, (200, 10), 0, device=DEV)
    engine.TIMER.reset(); engine.TIMER.enabled = True
    t0 = time.time()
    x, lens_d, lens_h = basecall.prepare_signals_device(sigs, (200, 10), 0, device=DEV)
    torch.cuda.synchronize()
    wall = time.time() - t0
    engine.TIMER.enabled = False
    kms = engine.TIMER.totals_ms()['prepare_signal'][0]
    nsamp = sum(len(s) for s in sigs)
    t0 = time.time()
    ref = [basecall.prepare_signal(s, (200, 10), 0) for s in sigs[:16]]
    host = (time.time() - t0) / sum(len(s) for s in sigs[:16])
    same = all(np.array_equal(x[:len(r), b, 0].cpu().numpy(), r) for b, r in enumerate(ref))
    print("%4d reads x ~%6d samples: kernel %.2f ms (%.0f M samples/s), call incl. host packing + H2D %.1f ms; "
          "NumPy %.1f M samples/s per core; first 16 reads identical: %s" % (B, N, kms, nsamp / kms / 1e3, wall * 1e3,
                                                                               1e-6 / host, same))
