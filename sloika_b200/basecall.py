"""Basecall workers (reference `sloika/basecall.py`).

Drop-in surface: `init_worker(model)`, `decode_post(...)`, `raw_worker(fast5_file_name, ...)`,
`SeqPrinter` keep the reference's names, arguments, return tuples and error behaviour
(`basecall.py:12-23, 26-51, 88-121, 124-163`).  On top of that `basecall_signals` / `raw_batch` run a
*batch* of whole reads through the device in one go (ragged batch, reads stay independent exactly as
with the reference's batch-of-one calls), which is how the GPU is actually fed.
"""
import sys

import numpy as np

from sloika_b200 import bio, cabi, decode, util
from sloika_b200.config import sloika_dtype
from sloika_b200.maths import mad
from sloika_b200.variables import nstate, DEFAULT_ALPHABET

calc_post = None


def init_worker(model, device=None):
    """ Worker init function: load the model once per process into the global `calc_post`
    (`basecall.py:12-23`).  `model` is a Sloika model pickle or a file from `helpers.compile_model`.
    """
    from sloika_b200 import helpers
    global calc_post
    calc_post = helpers.load_calc_post(model, device=device)


def decode_post(post, kmer_len, transducer, bad, min_prob, skip=5.0, trans=None, nbase=4, eta=1e-10):
    """ Decode Viterbi state sequence from a posterior matrix `[T, 1, S]` (`basecall.py:26-51`)

    :returns: score, Viterbi path
    """
    assert post.shape[2] == nstate(kmer_len, transducer=transducer, bad_state=bad, nbase=nbase)
    if not transducer:
        # legacy models (`basecall.py:47-50`): drop bad events, estimate per-event transition weights, old decoder
        from sloika_b200 import olddecode
        assert nbase == 4, "Modified bases not supported by old decoder"
        post = decode._as_device(post)
        post = decode.prepare_post(post, min_prob=min_prob, drop_bad=bad and not transducer)
        est = olddecode.estimate_transitions(post, trans=trans, return_device=True)
        return olddecode.decode_profile(post, trans=(eta + est).log(), log=False)
    assert post.shape[1] == 1, "decode_post decodes one read; use decode.viterbi_batch for batches"
    score, paths = decode.viterbi_batch(post, None, klen=kmer_len, skip_pen=skip, min_prob=min_prob,
                                        nbase=nbase)
    return score[0], paths[0]


def prepare_signal(signal, trim, open_pore_fraction):
    """Trim and normalise one raw signal (`basecall.py:111-118`); None when nothing is left."""
    from sloika_b200 import batch
    signal = batch.trim_open_pore(signal, open_pore_fraction)
    signal = util.trim_array(signal, *trim)
    if signal.size == 0:
        return None
    return ((signal - np.median(signal)) / mad(signal)).astype(sloika_dtype)


_PINNED = {}


def _pinned(tag, nelem, dtype):
    """Grow-only pinned staging buffer (page-locked allocations cost milliseconds; batches come in a loop)."""
    import torch
    buf = _PINNED.get(tag)
    if buf is None or buf.dtype != dtype or buf.numel() < nelem:
        buf = torch.empty(max(int(nelem * 1.25), 1), dtype=dtype).pin_memory()
        _PINNED[tag] = buf
    return buf[:nelem]


def prepare_signals_device(signals, trim=(200, 10), open_pore_fraction=0, window_size=100, device=None):
    """`prepare_signal` for a batch of raw reads on the device (`basecall.py:111-118` per read): open-pore trim, end
    trim and median / MAD normalisation in float64 with exact order statistics, written straight into the padded
    time-major batch the network consumes.  Bit-identical to the NumPy path.

    :param signals: list of 1-D raw signals (any float / int dtype; used as float64 like the reference)
    :returns: (x `[Tmax, B, 1]` float32 CUDA tensor, lengths int32 CUDA tensor `[B]`, lengths as a NumPy array);
        length 0 = nothing left after trimming, -1 = the reference's `trim_open_pore` would raise (no window
        stands out: `is_read` empty)
    """
    import torch
    lib = cabi.load()
    dev = device if device is not None else (calc_post.device if calc_post is not None else
                                             torch.device('cuda', torch.cuda.current_device()))
    B = len(signals)
    lens = np.array([len(s) for s in signals], dtype=np.int64)
    offsets = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    flat = _pinned('raw', int(offsets[-1]), torch.float64)
    flat_np = flat.numpy()
    for b, s in enumerate(signals):
        flat_np[offsets[b]:offsets[b + 1]] = s
    max_len = int(lens.max()) if B else 0
    with torch.cuda.device(dev):
        sig_d = flat.to(dev, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()      # the staging buffer is reused by the next call
        off_d = torch.from_numpy(offsets).to(dev)
        nbytes = lib.sloika_prepare_workspace_bytes(max_len, B, window_size)
        ws = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=dev)
        x = torch.empty((max(max_len, 1), B), dtype=torch.float32, device=dev)
        out_len = torch.empty(B, dtype=torch.int32, device=dev)
        from sloika_b200.engine import launch
        launch('prepare_signal', 1, lib.sloika_prepare_signal_fwd,
               cabi.ptr(sig_d), cabi.ptr(off_d), B, max_len, int(trim[0]), int(trim[1]), float(open_pore_fraction),
               int(window_size), cabi.ptr(ws), nbytes, cabi.ptr(x), B, max(max_len, 1), cabi.ptr(out_len),
               cabi.stream_ptr(dev))
        lens_h = out_len.cpu().numpy()
    T = max(int(lens_h.max()), 1)
    return x[:T].unsqueeze(2), out_len.clamp(min=0), lens_h


class CalledPath(list):
    """A best path (list of k-mer states) that also carries the base sequence assembled on the device
    (`decode.paths_to_sequences`); `SeqPrinter.write` uses it instead of spelling the k-mers on the host."""
    sequence = None
    assembly = None          # (kmer_len, alphabet, always_move) the sequence was assembled with


def basecall_signals(signals, kmer_len=5, min_prob=1e-5, skip=0.0, nbase=4, network=None, assemble=None):
    """Forward + Viterbi for a batch of normalised whole-read signals (list of 1-D float32 arrays).

    Reads are packed into one padded `[Tmax, B, 1]` batch with per-read lengths; every kernel honours
    the lengths, so each read sees exactly the computation it would see alone (`basecall.py:117-119`
    feeds one read per call).  Returns a list of `(score, path)`.

    :param assemble: optional `(alphabet, always_move)`: also assemble the base sequences on the device
        (`bio.kmers_to_sequence`, bio.py:228-237); paths are then `CalledPath` lists with `.sequence` set
    """
    import torch
    net = network if network is not None else calc_post
    if net is None:
        raise RuntimeError("init_worker() has not been called")
    if not signals:
        return []
    dev = net.device
    lens = np.array([len(s) for s in signals], dtype=np.int32)
    host = _pinned('signals', int(lens.max()) * len(signals), torch.float32).view(int(lens.max()), len(signals))
    host.zero_()
    for b, s in enumerate(signals):
        host[:len(s), b] = torch.from_numpy(np.ascontiguousarray(s, dtype=np.float32))
    x = host.to(dev, non_blocking=True).unsqueeze(2)
    torch.cuda.current_stream(dev).synchronize()          # the staging buffer is reused by the next call
    return basecall_prepared(x, torch.from_numpy(lens).to(dev), kmer_len, min_prob, skip, nbase, net, assemble)


def basecall_prepared(x, lengths, kmer_len=5, min_prob=1e-5, skip=0.0, nbase=4, network=None, assemble=None):
    """Forward + Viterbi for a padded device batch `x [Tmax, B, 1]` with per-read `lengths` (int32 CUDA tensor), e.g.
    the output of `prepare_signals_device`.  Same results as `basecall_signals`."""
    net = network if network is not None else calc_post
    if net is None:
        raise RuntimeError("init_worker() has not been called")
    out = net.forward_device(x, lengths, fused_decode=(kmer_len == 5 and nbase == 4))
    if assemble is None:
        score, paths = decode.viterbi_batch(out, None, klen=kmer_len, skip_pen=skip, min_prob=min_prob, nbase=nbase)
        return list(zip(score.tolist(), paths))
    alphabet, always_move = assemble
    if isinstance(alphabet, bytes):
        alphabet = alphabet.decode('ascii')
    score, paths_d, plen_d = decode.viterbi_batch(out, None, klen=kmer_len, skip_pen=skip, min_prob=min_prob,
                                                  nbase=nbase, return_device=True)
    seqs = decode.paths_to_sequences(paths_d, plen_d, kmer_len, alphabet, always_move=always_move)
    paths_h, plen_h = paths_d.cpu().numpy(), plen_d.cpu().numpy()
    calls = []
    for b, sc in enumerate(score.cpu().numpy().tolist()):
        path = CalledPath(paths_h[b, :plen_h[b]].tolist())
        path.sequence, path.assembly = seqs[b], (kmer_len, alphabet, bool(always_move))
        calls.append((sc, path))
    return calls


def basecall_chunks(x_host, kmer_len=5, min_prob=1e-5, skip=0.0, nbase=4, network=None):
    """Forward + Viterbi for a dense batch of equal-length chunks held in HOST memory.

    :param x_host: float32 torch CPU tensor `[T, B]` or `[T, B, 1]` (pinned memory makes the copy
        asynchronous); the synthetic-chunk shape of the throughput configs (4000 x 1024)
    :returns: (scores float32[B], paths int32[B, T'], path_len int32[B]) as NumPy arrays on the host
    """
    import torch
    net = network if network is not None else calc_post
    if net is None:
        raise RuntimeError("init_worker() has not been called")
    if x_host.dim() == 2:
        x_host = x_host.unsqueeze(2)
    x = x_host.to(net.device, non_blocking=True)
    out = net.forward_device(x, fused_decode=(kmer_len == 5 and nbase == 4))
    score, paths, plen = decode.viterbi_batch(out, None, klen=kmer_len, skip_pen=skip, min_prob=min_prob,
                                              nbase=nbase, return_device=True)
    return score.cpu().numpy(), paths.cpu().numpy(), plen.cpu().numpy()


_PIPELINE_STREAMS = {}        # device -> CUDA streams of the batch pipeline, kept so that their allocator pools stay warm


def _pipeline_streams(dev, n):
    import torch
    have = _PIPELINE_STREAMS.setdefault(str(dev), [])
    while len(have) < n:
        have.append(torch.cuda.Stream(dev))
    return have[:n]


def basecall_chunk_stream(batches, kmer_len=5, min_prob=1e-5, skip=0.0, nbase=4, network=None, in_flight=4):
    """`basecall_chunks` over a stream of host batches with `in_flight` batches on the device at once, each on its
    own CUDA stream: its host->device copy, its kernels and the device->host copy of its results are ordered on that
    stream, and the streams overlap one another -- copies under kernels, and the latency-bound recurrence kernels
    (a batch of 1024 chunks occupies a quarter to a half of the SMs, `csrc/gru_tc.cu`) beside the bandwidth-bound
    GEMM / Viterbi kernels of the neighbouring batches.

    :param batches: iterable of float32 torch CPU tensors `[T, B]` / `[T, B, 1]` (pinned memory for real overlap)
    :param in_flight: number of batches pipelined (1 = strictly one after the other)
    :returns: generator of (scores float32[B], paths int32[B, T'], path_len int32[B]) NumPy arrays, in input order
    """
    import collections
    import torch
    from sloika_b200 import engine
    net = network if network is not None else calc_post
    if net is None:
        raise RuntimeError("init_worker() has not been called")
    dev = net.device
    in_flight = max(1, int(in_flight))
    net.prepare()
    streams = _pipeline_streams(dev, in_flight)
    x_dev = [None] * in_flight
    host_bufs = [None] * in_flight         # pinned result buffers of a slot, reused (cudaHostAlloc synchronises the device)
    pending = collections.deque()          # (host result tensors, copy-out event, device tensors kept alive)
    caller = torch.cuda.current_stream(dev)
    ready = torch.cuda.Event()
    ready.record(caller)

    def finish(item):
        host, ev, _keep = item
        ev.synchronize()
        return tuple(t.numpy().copy() for t in host)            # the pinned buffers go back to their slot

    saved = engine.BATCHES_IN_FLIGHT
    engine.set_batches_in_flight(in_flight)
    try:
        for k, xh in enumerate(batches):
            if xh.dim() == 2:
                xh = xh.unsqueeze(2)
            slot = k % in_flight
            if len(pending) == in_flight:
                yield finish(pending.popleft())                 # the batch that last used this stream and buffer
            st = streams[slot]
            with torch.cuda.stream(st):
                if k < in_flight:
                    st.wait_event(ready)                        # work the caller enqueued before us
                if x_dev[slot] is None or x_dev[slot].shape != xh.shape:
                    x_dev[slot] = torch.empty(xh.shape, dtype=torch.float32, device=dev)
                x_dev[slot].copy_(xh, non_blocking=True)
                out = net.forward_device(x_dev[slot], fused_decode=(kmer_len == 5 and nbase == 4))
                res = decode.viterbi_batch(out, None, klen=kmer_len, skip_pen=skip, min_prob=min_prob, nbase=nbase,
                                           return_device=True)
                if host_bufs[slot] is None or any(b.shape != t.shape for b, t in zip(host_bufs[slot], res)):
                    host_bufs[slot] = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in res]
                host = host_bufs[slot]
                for buf, t in zip(host, res):
                    buf.copy_(t, non_blocking=True)
                out_ev = torch.cuda.Event()
                out_ev.record(st)
            pending.append((host, out_ev, res))
        while pending:
            yield finish(pending.popleft())
    finally:
        engine.set_batches_in_flight(saved)
        for st in streams:
            caller.wait_stream(st)


def _read_raw(fast5_file_name):
    from sloika_b200.fast5 import Fast5
    with Fast5(fast5_file_name) as f5:
        return f5.get_read(raw=True), f5.filename_short


def _read_raw_safe(fast5_file_name):
    """(signal, short name) or the exception raised by the reader (pool workers cannot write to our stderr in order)."""
    try:
        return _read_raw(fast5_file_name)
    except Exception as e:                                   # reported by the caller like raw_worker does
        return e


def read_files(fast5_file_names, reader_pool=None, wait=True):
    """(signal, short name) or the raised exception for every file; with a pool and `wait=False` an AsyncResult whose
    `.get()` gives that list (lets the caller read the next batch while the current one is on the device)."""
    if reader_pool is None:
        return [_read_raw_safe(fn) for fn in fast5_file_names]
    res = reader_pool.map_async(_read_raw_safe, fast5_file_names)
    return res.get() if wait else res


def make_reader_pool(jobs):
    """Process pool that parses fast5 files (pure-Python HDF5 reader: ~2 ms per read, the host-side bottleneck once
    the rest runs on the device).  Create it BEFORE the first CUDA call: the workers are forked and never touch CUDA."""
    if jobs is None or jobs <= 1:
        return None
    import multiprocessing
    return multiprocessing.get_context('fork').Pool(jobs)


def raw_worker(fast5_file_name, trim, open_pore_fraction, kmer_len, transducer, bad, min_prob,
               alphabet=DEFAULT_ALPHABET, skip=5.0, trans=None):
    """ Worker function for basecalling one fast5 file from raw data (`basecall.py:88-121`)

    :returns: (read name, score, path of k-mer states, number of samples) or None for unreadable /
        too-short reads (message on stderr, as the reference)
    """
    try:
        signal, sn = _read_raw(fast5_file_name)
    except Exception as e:
        sys.stderr.write("Error getting raw data for file {}\n{!r}\n".format(fast5_file_name, e))
        return None
    inMat = prepare_signal(signal, trim, open_pore_fraction)
    if inMat is None:
        sys.stderr.write("Read too short in file {}\n".format(fast5_file_name))
        return None
    inMat = inMat[:, None, None]
    post = calc_post.forward_device(_to_device(inMat))
    score, call = decode_post(post.data, kmer_len, transducer, bad, min_prob, skip, trans, nbase=len(alphabet))
    return sn, score, call, inMat.shape[0]


def raw_batch(fast5_file_names, trim=(200, 10), open_pore_fraction=0, kmer_len=5, transducer=True,
              bad=True, min_prob=1e-5, alphabet=DEFAULT_ALPHABET, skip=0.0, trans=None, reader_pool=None, loaded=None):
    """`raw_worker` over many files in one device batch; same result tuples, None for bad reads.

    :param reader_pool: optional `make_reader_pool(jobs)` pool that reads the files in parallel (`--jobs`)
    :param loaded: optional result of `read_files(fast5_file_names, ...)` obtained earlier (prefetched by the caller
        while the previous batch was on the device)
    """
    import os
    if not transducer:
        # legacy (non-transducer) raw models: the old decoder works on one read at a time
        return [raw_worker(fn, trim, open_pore_fraction, kmer_len, transducer, bad, min_prob, alphabet=alphabet,
                           skip=skip, trans=trans) for fn in fast5_file_names]
    names, raws, slots = [], [], []
    results = [None] * len(fast5_file_names)
    if loaded is None:
        loaded = read_files(fast5_file_names, reader_pool)
    for i, (fn, item) in enumerate(zip(fast5_file_names, loaded)):
        if isinstance(item, Exception):
            sys.stderr.write("Error getting raw data for file {}\n{!r}\n".format(fn, item))
            continue
        signal, sn = item
        names.append(sn)
        raws.append(signal)
        slots.append(i)
    if not raws:
        return results
    if os.environ.get('SLOIKA_B200_HOST_PREP'):
        # NumPy pre-processing, one read at a time (the reference's own path; kept for A/B checks)
        keep, signals = [], []
        for k, signal in enumerate(raws):
            sig = prepare_signal(signal, trim, open_pore_fraction)
            if sig is None:
                sys.stderr.write("Read too short in file {}\n".format(fast5_file_names[slots[k]]))
                continue
            keep.append(k)
            signals.append(sig)
        calls = basecall_signals(signals, kmer_len=kmer_len, min_prob=min_prob, skip=skip, nbase=len(alphabet),
                                 assemble=(alphabet, transducer))
        for k, sig, (score, path) in zip(keep, signals, calls):
            results[slots[k]] = (names[k], np.float32(score), path, len(sig))
        return results
    # trimming and normalisation on the device.  Every read of a device batch is padded to the longest one, and the
    # logits alone take ~4 KB per event and slot: reads are therefore taken longest first and cut into sub-batches
    # under a padded-sample budget (many short reads together, an ultra-long read on its own), and a sub-batch that
    # still runs out of device memory is split and retried instead of aborting the run.
    order = sorted(range(len(raws)), key=lambda k: -len(raws[k]))
    for group in plan_sub_batches([len(raws[k]) for k in order], padded_sample_budget()):
        _raw_sub_batch([order[g] for g in group], raws, names, slots, fast5_file_names, results, trim, open_pore_fraction,
                       kmer_len, min_prob, skip, alphabet, transducer)
    return results


def padded_sample_budget():
    """Padded raw samples (batch size x longest read) one device batch may hold: SLOIKA_B200_BATCH_SAMPLES, else a third
    of the free device memory at ~1.5 KB per sample (logits + traceback + activations of a stride-5 network)."""
    import os
    env = os.environ.get('SLOIKA_B200_BATCH_SAMPLES')
    if env:
        return max(1, int(env))
    import torch
    free, _total = torch.cuda.mem_get_info(calc_post.device if calc_post is not None else None)
    return max(1 << 20, int(free / 3 / 1500))


def plan_sub_batches(lengths, budget):
    """Greedy grouping of reads sorted longest first: a group is closed when adding the next read would push
    (group size x longest read of the group) over `budget`.  Returns lists of positions; a read longer than the
    budget gets a group of its own."""
    groups, cur, longest = [], [], 0
    for pos, n in enumerate(lengths):
        if cur and (len(cur) + 1) * max(longest, n) > budget:
            groups.append(cur)
            cur, longest = [], 0
        cur.append(pos)
        longest = max(longest, n)
    if cur:
        groups.append(cur)
    return groups


def _raw_sub_batch(ks, raws, names, slots, fast5_file_names, results, trim, open_pore_fraction, kmer_len, min_prob, skip,
                   alphabet, transducer):
    import torch
    try:
        x, lens_d, lens_h = prepare_signals_device([raws[k] for k in ks], trim, open_pore_fraction)
        if (lens_h < 0).any():
            raise ValueError("zero-size array to reduction operation minimum which has no identity")   # trim_open_pore
        calls = basecall_prepared(x, lens_d, kmer_len=kmer_len, min_prob=min_prob, skip=skip, nbase=len(alphabet),
                                  assemble=(alphabet, transducer))
    except torch.cuda.OutOfMemoryError:
        x = lens_d = calls = None
        torch.cuda.empty_cache()
        if len(ks) == 1:
            sys.stderr.write("Read does not fit in device memory, file {}\n".format(fast5_file_names[slots[ks[0]]]))
            return
        half = len(ks) // 2
        for part in (ks[:half], ks[half:]):
            _raw_sub_batch(part, raws, names, slots, fast5_file_names, results, trim, open_pore_fraction, kmer_len, min_prob,
                           skip, alphabet, transducer)
        return
    for j, (k, (score, path)) in enumerate(zip(ks, calls)):
        if lens_h[j] == 0:
            sys.stderr.write("Read too short in file {}\n".format(fast5_file_names[slots[k]]))
            continue
        results[slots[k]] = (names[k], np.float32(score), path, int(lens_h[j]))


def events_worker(fast5_file_name, section, segmentation, trim, kmer_len, transducer,
                  bad, min_prob, alphabet=DEFAULT_ALPHABET, skip=5.0, trans=None):
    """ Worker function for basecalling one fast5 file from events (`basecall.py:54-85`)

    :param section: part of read to basecall, 'template' or 'complement'
    :param segmentation: location of segmentation analysis for extracting target read section
    :param trim: (int, int) events to remove from read beginning and end
    :returns: (read name, score, state path, number of events) or None (message on stderr, as the reference)
    """
    from sloika_b200 import features
    from sloika_b200.fast5 import Fast5
    try:
        with Fast5(fast5_file_name) as f5:
            ev = f5.get_section_events(section, analysis=segmentation)
            sn = f5.filename_short
    except Exception as e:
        sys.stderr.write("Error getting events for section {!r} in file {}\n{!r}\n".format(section, fast5_file_name, e))
        return None
    ev = util.trim_array(ev, *trim)
    if ev.size == 0:
        sys.stderr.write("Read too short in file {}\n".format(fast5_file_name))
        return None
    inMat = features.from_events(ev, tag='')[:, None, :]
    post = calc_post.forward_device(_to_device(inMat))
    score, call = decode_post(post.data, kmer_len, transducer, bad, min_prob, skip, trans, nbase=len(alphabet))
    return sn, score, call, inMat.shape[0]


def _to_device(inMat):
    import torch
    return torch.from_numpy(np.ascontiguousarray(inMat)).to(calc_post.device)


class SeqPrinter(object):
    """ Formats fasta records and writes them to stdout or a file (`basecall.py:124-163`)

    :param kmer_len: length of kmer to use for converting states to kmers
    :param datatype: collective noun for the model input, e.g. "events" or "samples"
    :param transducer: if True, a kmer followed by itself is a move, not a stay
    :param fname: name of output file or None to use sys.stdout
    """

    def __init__(self, kmer_len, datatype="events", transducer=False, fname=None, alphabet=DEFAULT_ALPHABET):
        if isinstance(alphabet, bytes):
            alphabet = alphabet.decode('ascii')
        self.kmer_len = kmer_len
        self.alphabet = alphabet
        self.kmers = bio.all_kmers(kmer_len, alphabet=alphabet)
        self.transducer = transducer
        self.datatype = datatype
        self.close_fh = fname is not None
        self.fh = open(fname, 'w') if self.close_fh else sys.stdout

    def __del__(self):
        if getattr(self, 'close_fh', False):
            self.fh.close()

    def write(self, read_name, score, call, nev):
        if getattr(call, 'assembly', None) == (self.kmer_len, self.alphabet, bool(self.transducer)):
            seq = call.sequence                  # assembled on the device with the same conventions
        else:
            seq = bio.states_to_sequence(call, self.kmer_len, self.alphabet, always_move=self.transducer)
        self.fh.write(">{} score {:.0f}, {} {} to {} bases\n".format(read_name, score,
                                                                     nev, self.datatype, len(seq)))
        self.fh.write(seq + '\n')
        return len(seq)
