"""Time the remap decode (transducer.map_to_sequence, row f1) on chunk-sized problems: B reads of T events mapped
onto P-position sequences, 1025 states; device-timed GPU kernel vs the C oracle on the host cores (bounded sample),
paths and scores compared bit for bit on that sample.

    python tools/remap_bench.py > profiles/r1_remap_bench.txt
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import cbind
from sloika_b200 import transducer

DEV = torch.device('cuda:0')
S = 1025


def main():
    gen = torch.Generator(device=DEV)
    gen.manual_seed(11)
    rng = np.random.default_rng(11)
    ncpu = os.cpu_count() or 1
    print("# remap decode: slip 5.0, geometric-like priors off, log-transducer input; CPU = C oracle on {} threads".format(ncpu))
    print("# {:>7s} {:>6s} {:>6s} {:>10s} {:>12s} {:>10s} {:>12s} {:>8s} {:>5s}".format(
        "events", "reads", "npos", "gpu_ms", "gpu_Gcell/s", "cpu_reads", "cpu_Gcell/s", "ratio", "same"))
    for T, B, P in [(800, 1024, 400), (800, 4096, 400), (2000, 1024, 1000), (4000, 592, 2000), (8000, 148, 4000)]:
        lt = torch.log_softmax(3.0 * torch.randn((T, B, S), generator=gen, device=DEV), dim=-1)
        seqs = [rng.integers(1, S, size=P).astype(np.int32) for _ in range(B)]
        for _ in range(2):
            score, paths = transducer.map_to_sequence_batch(lt, seqs, slip=5.0, return_device=True)
        torch.cuda.synchronize()
        t0 = time.time()
        score, paths = transducer.map_to_sequence_batch(lt, seqs, slip=5.0, return_device=True)
        torch.cuda.synchronize()
        wall_ms = (time.time() - t0) * 1e3                       # includes the host-side packing of the sequences
        # kernel alone, device-timed (CUDA events around the C-ABI call)
        tm = {}
        score, paths = transducer.map_to_sequence_batch(lt, seqs, slip=5.0, return_device=True, timing=tm)
        ms = tm['kernel_ms']
        cells = float(T) * B * P
        n_cpu = int(max(1, min(B, ncpu * max(1, int(4e8 // (T * P))))))
        lt_h = lt[:, :n_cpu].cpu().numpy()
        seq_h = np.stack(seqs[:n_cpu])
        t0 = time.time()
        ref_s, ref_p = cbind.remap_batch(lt_h, seq_h, slip=5.0)
        dt = time.time() - t0
        same = np.array_equal(score[:n_cpu].cpu().numpy(), ref_s) and np.array_equal(paths[:n_cpu].cpu().numpy(), ref_p)
        gpu, cpu = cells / ms / 1e6, float(T) * n_cpu * P / dt / 1e9
        print("  {:7d} {:6d} {:6d} {:10.2f} {:12.2f} {:10d} {:12.3f} {:8.0f} {:>5s}   (wall incl. host packing {:.1f} ms)".format(
            T, B, P, ms, gpu, n_cpu, cpu, gpu / cpu, "yes" if same else "NO", wall_ms))
        sys.stdout.flush()
        del lt, score, paths
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
