"""Decode-only sweep (BASELINE.json configs[4], SURVEY.md section 8 "Config 5"): Viterbi best-path decode over
synthetic posteriors (row softmax of 3*N(0,1) logits, 1025 states) for read lengths 1k .. 100k events.

The floor and the log of decode.prepare_post / decode.py:56 are applied once on the device
(lp = log(1e-5 + (1 - 1e-5) * post + 1e-10), float32) so that both arms decode the very same log-posteriors
(`log=True`) and their paths can be compared bit for bit.
GPU: `decode.viterbi_batch` (one CTA per read), device-timed with CUDA events, input resident in HBM.
CPU: the oracle (C restatement of decode.py:39-93, one thread per read) on a bounded sample of the same reads.

    python tools/decode_sweep.py > profiles/r1_decode_sweep.txt
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import cbind
from sloika_b200 import decode

DEV = torch.device('cuda:0')
S = 1025
# (events per read, reads): sized so that posteriors + traceback stay below ~80 GB
CASES = [(1000, 1024), (3000, 1024), (10000, 1024), (30000, 444), (100000, 148)]


def synth(T, B, gen):
    post = torch.empty((T, B, S), dtype=torch.float32, device=DEV)
    step = max(1, (1 << 28) // (B * S))                    # ~1 GB of logits at a time
    for t0 in range(0, T, step):
        t1 = min(T, t0 + step)
        logits = torch.randn((t1 - t0, B, S), generator=gen, device=DEV) * 3.0
        p = torch.softmax(logits, dim=-1)
        post[t0:t1] = torch.log((1e-5 + (1.0 - 1e-5) * p) + 1e-10)
        del logits, p
    return post


def main():
    gen = torch.Generator(device=DEV)
    gen.manual_seed(5)
    ncpu = os.cpu_count() or 1
    print("# decode-only sweep: {} states, min_prob 1e-5, skip 0.0; GPU = viterbi_batch (device-timed), "
          "CPU = oracle C port on {} threads".format(S, ncpu))
    print("# {:>8s} {:>6s} {:>10s} {:>14s} {:>10s} {:>12s} {:>14s} {:>8s} {:>6s}".format(
        "events", "reads", "gpu_ms", "gpu_Mevents/s", "GB/s", "cpu_reads", "cpu_Mevents/s", "ratio", "same"))
    for T, B in CASES:
        post = synth(T, B, gen)
        for _ in range(2):
            decode.viterbi_batch(post, None, log=True, return_device=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            score, paths, plen = decode.viterbi_batch(post, None, log=True, return_device=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gpu_rate = T * B / ms / 1e3                                     # M events/s
        gbs = (T * B * (S * 4.0 + 1024.0)) / ms / 1e6                   # posteriors read + traceback written
        # CPU on a bounded sample: enough reads for ~10 s at ~15 k events/s/thread
        n_cpu = int(max(1, min(B, ncpu * max(1, 150000 // T))))
        sample = post[:, :n_cpu].cpu().numpy()
        t0 = time.time()
        s_ref, p_ref = cbind.viterbi_batch(sample, None, 5, 4, 0.0)
        dt = time.time() - t0
        cpu_rate = T * n_cpu / dt / 1e6
        plen_h = plen[:n_cpu].cpu().numpy()
        paths_h = paths[:n_cpu].cpu().numpy()
        same = all(paths_h[b, :plen_h[b]].tolist() == list(p_ref[b]) for b in range(n_cpu))
        print("  {:8d} {:6d} {:10.2f} {:14.1f} {:10.0f} {:12d} {:14.3f} {:8.0f} {:>6s}".format(
            T, B, ms, gpu_rate, gbs, n_cpu, cpu_rate, gpu_rate / cpu_rate, "yes" if same else "NO"))
        sys.stdout.flush()
        del post, score, paths, plen
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
