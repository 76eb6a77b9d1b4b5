#!/usr/bin/env python3
"""Benchmark of the raw basecall hot path (forward + Viterbi) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--in-flight F] [--scaling weak|strong]
                    [--workload raw_rgrgr|raw_rGr|bigger_raw_gru|bigger_raw_gru_wide|pretrained_like|decode_sweep]
    python bench.py --impl reference ...        # the CPU arm: oracle port on the host cores
Default = BASELINE.json configs[2] (raw_rgrgr, the config the metric is quoted on); the other workloads are
configs[1] (raw_rGr), configs[3] (bigger_raw_gru and a widened [64, 256, 256] variant) and configs[4] (decode only).

A step = one pass of the hot path over one batch of synthetic raw-signal chunks
(x ~ N(0,1) float32 [4000, 1024, 1] per GPU, seeded; weights truncated-normal sd 0.5, seeded):
conv -> GRU stack (projection + recurrence per layer) -> softmax -> Viterbi incl. backtrace.
metric = raw samples/s basecalled, whole job over all ranks (read sharding, weak scaling: every rank
owns its own 1024 chunks; no data-path collective).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')     # one hardware queue per stream of the pipelined step (sloika_b200/__init__.py)
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "raw_samples_per_s_basecalled_fwd_viterbi"
UNIT = "samples/s"
CHUNK_LEN = 4000
BATCH_PER_GPU = 1024
WEIGHT_SEED = 0xdeadbeef & 0x7fffffff
INPUT_SEED = 20261017


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='raw_rgrgr')
    ap.add_argument('--batch', type=int, default=BATCH_PER_GPU, help='chunks per GPU')
    ap.add_argument('--chunk', type=int, default=CHUNK_LEN, help='raw samples per chunk')
    ap.add_argument('--cpu-sample-chunks', type=int, default=None,
                    help='chunks in the bounded CPU sample (default: 384 once for cpu_baseline, 32 per step for --impl reference)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: --batch chunks per GPU; strong: --batch chunks in total, split over the GPUs')
    ap.add_argument('--in-flight', type=int, default=None,
                    help='batches pipelined on separate CUDA streams (1 = one batch after the other); default: 16 for the '
                         'stride-5 workloads (~6 GB per batch), 4 otherwise')
    return ap.parse_args()


def default_in_flight(workload, steps):
    """Batches pipelined on separate CUDA streams when --in-flight is not given.  Stride-5 workloads (~6.5 GB per batch): up
    to 20, and a count that divides the timed steps so that the last round is a full one (a GRU layer holds 8 SMs per batch
    for ~4 ms: a lone batch at the end of the run would leave the device idle behind it); the stride-2 workloads need 2.5 x
    the memory per batch and stay at 4."""
    if workload in ('raw_rgrgr', 'pretrained_like'):
        rounds = -(-max(steps, 1) // 20)
        return -(-max(steps, 1) // rounds)
    return 4


def build_network(workload):
    from sloika_b200 import zoo
    np.random.seed(WEIGHT_SEED)
    if workload == 'bigger_raw_gru_wide':
        return zoo.bigger_raw_gru(size=(64, 256, 256))       # widened variant (H > 144: step-wise scan)
    return getattr(zoo, workload)()


def stride_of(net):
    return net.layers[0].stride


def flops_per_sample(net):
    """Algorithmic FLOP per raw sample from the layer shapes."""
    from sloika_b200 import layers as L

    def per_step(layer):
        if isinstance(layer, L.Convolution):
            return 2 * layer.insize * layer.winlen * layer.size
        if isinstance(layer, L.Gru):
            return 2 * (3 * layer.size * layer.insize + 3 * layer.size * layer.size)
        if isinstance(layer, (L.FeedForward, L.Softmax)):
            return 2 * layer.insize * layer.size
        if isinstance(layer, L.Reverse):
            return per_step(layer.layer)
        return sum(per_step(sub) for sub in layer.layers)
    return per_step(net) / float(stride_of(net))


# ---------------------------------------------------------------------------------------------
# algorithmic HBM bytes per raw sample of each kernel class (every activation
# written once / read once, fp32, weights resident, vI not materialised, uint8 traceback)
def algorithmic_bytes_per_sample(net):
    from sloika_b200 import layers as L
    s = float(stride_of(net))
    out = {}

    def add(name, nbytes):
        out[name] = out.get(name, 0.0) + nbytes / s

    def walk(layer, first=False):
        if isinstance(layer, L.Convolution):
            out['conv1d'] = 4.0 * layer.insize + 4.0 * layer.size / s
        elif isinstance(layer, L.Gru):
            add('gru_layer', 4.0 * layer.insize + 4.0 * layer.size)
        elif isinstance(layer, L.FeedForward):
            add('feedforward', 4.0 * layer.insize + 4.0 * layer.size)
        elif isinstance(layer, L.Softmax):
            add('softmax', 4.0 * layer.insize + 4.0 * layer.size)
            add('viterbi', 4.0 * layer.size + (layer.size - 1) / 2.0)      # posteriors read + uint16-per-quad traceback
        elif isinstance(layer, L.Reverse):
            walk(layer.layer)
        else:
            for sub in layer.layers:
                walk(sub)
    walk(net)
    return out


# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
              'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
def cpu_reference_step(desc, x, klen=5, min_prob=1e-5, skip=0.0, pool=None):
    """One bounded CPU sample: oracle forward (NumPy/BLAS, all host threads) + the reference-semantics
    NumPy Viterbi per chunk, chunks spread over `pool` processes (mirrors --jobs, iterators.py:343-351)."""
    from oracle import forward_ref
    post = forward_ref.run(desc, x)
    cols = [np.ascontiguousarray(post[:, b:b + 1]) for b in range(post.shape[1])]
    args = [(c, klen, min_prob, skip) for c in cols]
    if pool is None:
        res = [_decode_one(a) for a in args]
    else:
        res = pool.map(_decode_one, args)
    return res


def _decode_one(arg):
    from oracle import decode_ref
    post, klen, min_prob, skip = arg
    return decode_ref.decode_post(post, klen, min_prob, skip=skip)


def time_cpu_reference(net, chunk, nchunks, steps, warmup):
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    desc = net.json(params=True)
    rng = np.random.default_rng(INPUT_SEED)
    x = rng.standard_normal((chunk, nchunks, 1)).astype(np.float32)
    ctx = mp.get_context('fork')
    with ctx.Pool(min(cores, nchunks)) as pool:
        for _ in range(warmup):
            cpu_reference_step(desc, x, pool=pool)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_reference_step(desc, x, pool=pool)
        dt = time.perf_counter() - t0
    return chunk * nchunks * steps / dt, dt / steps, cores


# ---------------------------------------------------------------------------------------------
def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    net = build_network(args.workload)
    nchunks = args.cpu_sample_chunks or 32
    value, sec_per_step, cores = time_cpu_reference(net, args.chunk, nchunks, args.steps, args.warmup)
    sample = "{} chunks x {} samples per step (of {} per GPU), oracle NumPy forward + NumPy Viterbi over {} processes".format(
        nchunks, args.chunk, args.batch, min(cores, nchunks))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "Theano 0.8.2 is not installable in this image; the CPU arm is the reference-semantics port (oracle/)",
    }
    print(json.dumps(line), flush=True)


def workload_name(args):
    if args.scaling == 'strong':
        return "{}: {} chunks x {} raw samples in total (split over the GPUs), fwd+Viterbi".format(
            args.workload, args.batch, args.chunk)
    return "{}: {} chunks x {} raw samples per GPU, fwd+Viterbi".format(args.workload, args.batch, args.chunk)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: decode only, synthetic posteriors (row softmax of 3*N(0,1) logits, 1025 states)
SWEEP_CASES = [(1000, 1024), (3000, 1024), (10000, 1024), (30000, 444), (100000, 148)]   # (events per read, reads)


def run_decode_sweep(args):
    import torch
    import torch.distributed as dist
    from sloika_b200 import cabi, decode
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cabi.load()
    S = 1025
    gen = torch.Generator(device=dev)
    gen.manual_seed(5 + rank)

    def synth(T, B):
        post = torch.empty((T, B, S), dtype=torch.float32, device=dev)
        step = max(1, (1 << 28) // (B * S))
        for t0 in range(0, T, step):
            t1 = min(T, t0 + step)
            p = torch.softmax(torch.randn((t1 - t0, B, S), generator=gen, device=dev) * 3.0, dim=-1)
            post[t0:t1] = torch.log((1e-5 + (1.0 - 1e-5) * p) + 1e-10)
        return post

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    rows, total_events, total_ms = [], 0, 0.0
    cpu = None
    for T, B in SWEEP_CASES:
        post = synth(T, B)
        for _ in range(max(args.warmup, 3)):
            decode.viterbi_batch(post, None, log=True, return_device=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        reps = max(1, min(args.steps, 5))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            score, paths, plen = decode.viterbi_batch(post, None, log=True, return_device=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        rows.append({"events_per_read": T, "reads_per_gpu": B, "ms": ms, "events_per_s": T * B * world / (ms * 1e-3),
                     "GBps_per_gpu": T * B * (S * 4.0 + 512.0) / (ms * 1e-3) / 1e9})
        total_events += T * B * world
        total_ms += ms
        if rank == 0 and world == 1 and not args.no_cpu_baseline and T == 10000:
            # bounded CPU sample: the C restatement of decode.py:39-93 on the host cores, same log-posteriors
            from oracle import cbind
            ncpu = os.cpu_count() or 1
            n_cpu = min(B, 2 * ncpu)
            sample = post[:, :n_cpu].cpu().numpy()
            t0 = time.perf_counter()
            s_ref, p_ref = cbind.viterbi_batch(sample, None, 5, 4, 0.0)
            dt = time.perf_counter() - t0
            plen_h, paths_h = plen[:n_cpu].cpu().numpy(), paths[:n_cpu].cpu().numpy()
            same = all(paths_h[b, :plen_h[b]].tolist() == list(p_ref[b]) for b in range(n_cpu))
            cpu = {"value": T * n_cpu / dt, "unit": "events/s", "cores": ncpu, "kind": "port",
                   "sample": "{} reads x {} events, oracle/viterbi_ref.c over {} threads, {:.1f} s; paths identical to the device's: {}".format(
                       n_cpu, T, ncpu, dt, same)}
        del post, score, paths, plen
        torch.cuda.empty_cache()
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    value = total_events / (total_ms * 1e-3)
    big = rows[2]
    line = {"metric": "events_per_s_viterbi_decoded", "value": value, "unit": "events/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "decode_sweep: Viterbi over synthetic posteriors, reads of 1k/3k/10k/30k/100k events x 1025 states per GPU",
                       "cases": rows, "l2": "posteriors per case exceed L2; no flush needed"},
            "e2e": None, "gpu_launches": len(SWEEP_CASES) * max(1, min(args.steps, 5)) * world,
            "roofline": {"kernel": "viterbi_k1024 (10k x 1024 case)", "bound": "hbm", "achieved": big["GBps_per_gpu"],
                         "peak": hbm_peak, "unit": "GB/s", "frac": big["GBps_per_gpu"] / hbm_peak, "traffic": None,
                         "algorithmic_bytes_per_event": S * 4.0 + 512.0},
            "cpu_baseline": cpu, "clocks": clocks}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from sloika_b200 import basecall, cabi, decode, engine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cabi.load()

    net = build_network(args.workload)
    calc_post = net.compile().to(dev)
    T, B = args.chunk, args.batch
    if args.scaling == 'strong':
        B = max(1, B // world)                              # the batch is split over the ranks
    gen = torch.Generator().manual_seed(INPUT_SEED + rank)
    x_host = torch.randn((T, B, 1), generator=gen, dtype=torch.float32).pin_memory()
    x_dev = x_host.to(dev)
    samples_per_step = T * B
    if args.in_flight is None:
        args.in_flight = default_in_flight(args.workload, args.steps)
    K = max(1, args.in_flight)
    calc_post.prepare()
    main = torch.cuda.current_stream(dev)
    streams = basecall._pipeline_streams(dev, K)       # the streams the public API pipelines on (warm allocator pools)

    def step_device():
        out = calc_post.forward_device(x_dev, fused_decode=True)
        return decode.viterbi_batch(out, None, klen=5, skip_pen=0.0, min_prob=1e-5, return_device=True)

    def run_e2e(steps):
        """`steps` batches through the public host-buffer API (`basecall.basecall_chunk_stream`): every batch is copied
        from pinned host memory to the device and its scores / paths / lengths are copied back, all inside the timed
        region; `in_flight` batches are pipelined, each on its own stream."""
        res = None
        for res in basecall.basecall_chunk_stream((x_host for _ in range(steps)), kmer_len=5, min_prob=1e-5, skip=0.0,
                                                  network=calc_post, in_flight=K):
            pass
        return res

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(steps, in_flight):
        """`steps` batches, `in_flight` of them pipelined: batch i runs on stream i % in_flight.  Device time from an
        event all streams wait on to an event that waits on all streams."""
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        engine.set_batches_in_flight(in_flight)
        barrier()
        start.record(main)
        use = streams[:in_flight]
        for st in use:
            st.wait_event(start)
        res = None
        for i in range(steps):
            with torch.cuda.stream(use[i % in_flight]):
                res = step_device()
        for st in use:
            main.wait_stream(st)
        end.record(main)
        barrier()
        return start.elapsed_time(end), res

    # ---- warm-up (every stream: allocator pools, kernel attributes) ----
    timed(max(args.warmup, 3) * K, K)

    # ---- device-resident timing (value); per-kernel events are recorded on the launching streams ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    engine.TIMER.reset()
    engine.TIMER.enabled = True
    ms_dev, res = timed(args.steps, K)
    engine.TIMER.enabled = False
    launches = engine.TIMER.launches
    kernel_ms = engine.TIMER.totals_ms()
    ms_dev = max_over_ranks(ms_dev)

    # ---- the same step alone on the device (one batch in flight): latency of a batch, per-kernel times undisturbed ----
    iso_steps = min(args.steps, 5)
    timed(2, 1)
    engine.TIMER.reset()
    engine.TIMER.enabled = True
    ms_single, _ = timed(iso_steps, 1)
    engine.TIMER.enabled = False
    kernel_ms_single = engine.TIMER.totals_ms()
    ms_single = max_over_ranks(ms_single) / iso_steps
    engine.set_batches_in_flight(K)

    # ---- end-to-end timing through the host-buffer API (e2e): H2D + D2H inside, wall clock on host ----
    run_e2e(2 * K)
    barrier()
    wall0 = time.perf_counter()
    res_e2e = run_e2e(args.steps)
    barrier()
    wall_e2e = (time.perf_counter() - wall0) * 1e3
    wall_e2e = max_over_ranks(wall_e2e)
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = samples_per_step * world * args.steps / (ms_dev * 1e-3)
    e2e_value = samples_per_step * world * args.steps / (wall_e2e * 1e-3)
    scores, paths, plen = res_e2e
    d2h = scores.nbytes + paths.nbytes + plen.nbytes

    # ---- roofline of the dominant kernel ----
    peaks = {}
    peaks_src = "fallback"
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            peaks = json.load(fh)
        peaks_src = "measured"
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    alg = algorithmic_bytes_per_sample(net)
    # a GRU layer's algorithmic bytes (x read once, h written once) are charged to its recurrence kernel;
    # the projection kernel's vI round trip is extra (non-algorithmic) traffic by SURVEY section 8d
    # GRU layers run either as projection GEMM + recurrence kernel or as one fused launch (csrc/gru_fused.cu); the
    # layer's algorithmic bytes go to whichever kernel ran it, in proportion to the launches
    def gru_share(kms, name):
        calls = {k: kms.get(k, (0.0, 0))[1] for k in ('gru_recurrence', 'gru_fused', 'gru_seq')}
        total = sum(calls.values())
        return alg.get('gru_layer', 0.0) * calls[name] / total if total else 0.0
    alg_by_kernel = {'conv1d': alg.get('conv1d', 0.0), 'gru_recurrence': gru_share(kernel_ms_single, 'gru_recurrence'),
                     'gru_fused': gru_share(kernel_ms_single, 'gru_fused'), 'gru_seq': gru_share(kernel_ms_single, 'gru_seq'),
                     'block_layout': 0.0,
                     'gru_projection': 0.0, 'feedforward': alg.get('feedforward', 0.0),
                     'softmax': alg.get('softmax', 0.0), 'viterbi': alg.get('viterbi', 0.0)}
    def table(kms, steps):
        out = {}
        for name, (ms, calls) in kms.items():
            nbytes = alg_by_kernel.get(name, 0.0) * samples_per_step * steps
            out[name] = {"ms_per_step": ms / steps, "calls_per_step": calls / steps,
                         "algorithmic_GBps": (nbytes / (ms * 1e-3) / 1e9) if ms > 0 else None}
        return out
    breakdown_single = table(kernel_ms_single, iso_steps)  # one batch in flight: undisturbed durations
    alg_by_kernel['gru_recurrence'] = gru_share(kernel_ms, 'gru_recurrence')
    alg_by_kernel['gru_fused'] = gru_share(kernel_ms, 'gru_fused')
    alg_by_kernel['gru_seq'] = gru_share(kernel_ms, 'gru_seq')
    breakdown = table(kernel_ms, args.steps)               # spans inside the pipelined region: they overlap each other
    alg_by_kernel['gru_recurrence'] = gru_share(kernel_ms_single, 'gru_recurrence')
    alg_by_kernel['gru_fused'] = gru_share(kernel_ms_single, 'gru_fused')
    alg_by_kernel['gru_seq'] = gru_share(kernel_ms_single, 'gru_seq')
    # dominant kernel = largest share of the step when a batch has the GPU to itself; its roofline numbers come from
    # that pass (in the pipelined region several launches of the same kernel share the SMs, which stretches every
    # launch without saying anything about the kernel); the pipelined spans are reported beside them
    dominant = max(kernel_ms_single, key=lambda k: kernel_ms_single[k][0])
    dom_ms, dom_calls = kernel_ms_single[dominant]
    dom_bytes_per_launch = alg_by_kernel.get(dominant, 0.0) * samples_per_step * iso_steps / dom_calls
    achieved = dom_bytes_per_launch / (dom_ms / dom_calls * 1e-3) / 1e9
    traffic_table = {}
    try:
        if args.workload == 'raw_rgrgr' and B == BATCH_PER_GPU and T == CHUNK_LEN:
            with open(os.path.join(ROOT, 'profiles', 'r2g_traffic.json')) as fh:
                traffic_table = json.load(fh)
    except Exception:
        traffic_table = {}
    traffic = traffic_table.get(dominant)
    pipe_ms, pipe_calls = kernel_ms.get(dominant, (0.0, 0))
    roofline = {"kernel": dominant, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peaks_src,
                "algorithmic_bytes_per_launch": dom_bytes_per_launch,
                "avg_launch_ms": dom_ms / dom_calls,
                "measured_in": "single_batch pass of this run (CUDA events on the launching stream, one batch in flight)",
                "share_of_kernel_time": dom_ms / sum(v[0] for v in kernel_ms_single.values())}
    # the same figures for the kernel with the largest share INSIDE the timed (pipelined) region, from the spans recorded
    # there: `in_flight` batches share the SMs and HBM, so a launch is stretched by its neighbours -- the fraction says
    # how much of the device's bandwidth one launch uses while it runs, not how good the kernel is
    tr_alg = dict(alg_by_kernel, gru_recurrence=gru_share(kernel_ms, 'gru_recurrence'), gru_fused=gru_share(kernel_ms, 'gru_fused'),
                  gru_seq=gru_share(kernel_ms, 'gru_seq'))
    tr_dom = max(kernel_ms, key=lambda k: kernel_ms[k][0])
    tr_ms, tr_calls = kernel_ms[tr_dom]
    tr_bytes = tr_alg.get(tr_dom, 0.0) * samples_per_step * args.steps / tr_calls
    tr_achieved = tr_bytes / (tr_ms / tr_calls * 1e-3) / 1e9
    roofline["timed_region"] = {"kernel": tr_dom, "achieved": tr_achieved, "frac": tr_achieved / hbm_peak,
                                "algorithmic_bytes_per_launch": tr_bytes, "avg_launch_ms": tr_ms / tr_calls,
                                "traffic": traffic_table.get(tr_dom), "batches_in_flight": K,
                                "share_of_kernel_time": tr_ms / sum(v[0] for v in kernel_ms.values())}
    # and for the step as a whole: all algorithmic bytes of a batch over the device time per batch
    step_bytes = sum(alg.values()) * samples_per_step
    step_gbps = step_bytes / (ms_dev / args.steps * 1e-3) / 1e9
    roofline["step"] = {"algorithmic_bytes": step_bytes, "achieved": step_gbps, "frac": step_gbps / hbm_peak,
                        "traffic": (sum(traffic_table.get(k, 0.0) * c / args.steps for k, (_, c) in kernel_ms.items())
                                    if traffic_table else None)}
    if 'gru_recurrence' in kernel_ms_single:
        # the recurrence is bound by the latency of its dependent steps, not by HBM: say so on the line
        rec_ms, rec_calls = kernel_ms_single['gru_recurrence']
        steps_per_launch = T // stride_of(net)
        pipe_name = 'gru_seq' if 'gru_seq' in kernel_ms else ('gru_fused' if 'gru_fused' in kernel_ms else 'gru_recurrence')
        roofline["latency"] = {"kernel": "gru_recurrence", "us_per_time_step_one_batch": 1e3 * rec_ms / rec_calls / steps_per_launch,
                               "pipelined_kernel": pipe_name,
                               "us_per_time_step_pipelined": 1e3 * kernel_ms[pipe_name][0] / kernel_ms[pipe_name][1] / steps_per_launch,
                               "time_steps_per_launch": steps_per_launch,
                               "ncu": "profiles/r2g_step_kernels_ncu.txt (warps active, issue active, tensor pipe)"}
    total_alg = sum(alg.values())
    paper = min(hbm_peak * 1e9 / total_alg, float(peaks.get('bf16_tflops_sustained', 1400.0)) * 1e12 / flops_per_sample(net))

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        nchunks = args.cpu_sample_chunks or 384
        cpu_value, sec_per_step, cores = time_cpu_reference(net, T, nchunks, steps=1, warmup=0)
        cpu_baseline = {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "{} chunks x {} samples once: oracle NumPy forward (BLAS threads) + NumPy Viterbi over {} processes, {:.1f} s".format(
                            nchunks, T, min(cores, nchunks), sec_per_step)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "chunks_per_gpu": B, "chunk_len": T, "stride": stride_of(net),
                   "batches_in_flight": K, "value_excludes_h2d": True,
                   "sharding": "reads x{} (no collective)".format(world),
                   "l2": "working set per step (posteriors {:.2f} GB) exceeds L2; no flush needed".format(
                       4.0 * net.size * (T // stride_of(net)) * B / 1e9)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * 4),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": wall_e2e / args.steps},
        "gpu_launches": launches * world,          # kernels of this repo enqueued in the timed region, all ranks
        "roofline": roofline,
        "kernels": breakdown,
        "single_batch": {"ms_per_step": ms_single, "value": samples_per_step * world / (ms_single * 1e-3),
                         "kernels": breakdown_single},
        "paper_roofline_samples_per_s": paper,
        "frac_of_paper_roofline": value / world / paper,
        "cpu_baseline": cpu_baseline,
        "clocks": clocks,
        "device_memory_gb": {"reserved_peak": torch.cuda.max_memory_reserved(dev) / 1e9,
                             "allocated_peak": torch.cuda.max_memory_allocated(dev) / 1e9},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    elif args.workload == 'decode_sweep':
        run_decode_sweep(args)
    else:
        run_b200_arm(args)


if __name__ == '__main__':
    main()
