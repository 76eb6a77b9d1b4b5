"""Execution of a layer tree on one B200: device activations and the per-layer kernel launches.

This replaces what Theano does for the reference (`Layer.compile`, `sloika/layers.py:34-36`:
`th.function([x], self.run(x))`): `CompiledNetwork` is the `calc_post` callable of
`sloika/basecall.py:12-23, 119`, and the `run_*` functions are the bodies of the layers' `run`
methods.  Every operator goes through the C ABI (`cabi`), on the current torch stream; torch is
used for device memory, streams and nothing else.
"""
import os

import numpy as np

from sloika_b200 import cabi
from sloika_b200.activation import code_of
from sloika_b200.config import sloika_dtype
from sloika_b200.conv import output_length


class Act(object):
    """A device activation `[T, B, F]` (float32), possibly a column slice of a wider buffer.

    data     torch CUDA tensor of shape [T, B, F]; rows (t, b) are `ld` floats apart
    lengths  optional int32 CUDA tensor [B]: valid steps per sequence (ragged whole-read batch)
    reverse  time direction flag toggled by `Reverse` (nothing is ever flipped in memory)
    bounded  True when the producing layer guarantees |x| <= 1 (tanh / sigmoid / GRU / softmax outputs): the
             next GEMM may then use the fp16-split tensor-core kernel (SLOIKA_GEMM_TC_F16)
    absmax   optional 1-element CUDA tensor holding max |x| as measured by the producing kernel (unbounded
             activations such as elu): lets the next GEMM choose its operand format on the device
    """

    def __init__(self, data, lengths=None, reverse=False, bounded=False, absmax=None):
        self._blocked = None
        self._set_data(data)
        self.lengths = lengths
        self.reverse = reverse
        self.bounded = bounded
        self.absmax = absmax

    def _set_data(self, data):
        assert data.dim() == 3
        T, B, F = data.shape
        # strides of size-1 dimensions are meaningless in torch; derive the row distance from a
        # dimension that is really stepped over
        if B > 1:
            ld = data.stride(1)
            assert T == 1 or data.stride(0) == B * ld, "time-major [T, B, F] layout expected"
        elif T > 1:
            ld = data.stride(0)
        else:
            ld = F
        assert (F == 1 or data.stride(2) == 1) and ld >= F, "feature axis must be contiguous"
        self._data = data
        self._ld = ld
        self._shape = (T, B, F)

    @classmethod
    def from_blocked(cls, blocked, shape, lengths=None, reverse=False, bounded=False, absmax=None):
        """An activation that so far exists only in the BLOCKED layout of the sequences-on-lanes GRU kernel
        (include/sloika_b200.h, `sloika_gru_seq_fwd`): the next such layer reads it as it is, anybody else asks for
        `.data` and gets a row-major copy made on the spot (`sloika_block_layout_fwd`)."""
        self = cls.__new__(cls)
        self._blocked, self._data, self._ld, self._shape = blocked, None, None, tuple(int(v) for v in shape)
        self.lengths, self.reverse, self.bounded, self.absmax = lengths, reverse, bounded, absmax
        return self

    @property
    def blocked(self):
        return self._blocked

    @property
    def data(self):
        if self._data is None:
            T, B, F = self._shape
            y = _padded_rows(T, B, F, self._blocked.device)
            launch('block_layout', 1, cabi.load().sloika_block_layout_fwd, cabi.ptr(self._blocked), cabi.ptr(y),
                   _row_stride(y), T, B, F, 0, cabi.stream_ptr(y.device))
            self._set_data(y)
        return self._data

    @property
    def T(self):
        return self._shape[0]

    @property
    def B(self):
        return self._shape[1]

    @property
    def F(self):
        return self._shape[2]

    @property
    def ld(self):
        if self._data is None:
            self.data
        return self._ld

    @property
    def device(self):
        return self._blocked.device if self._data is None else self._data.device

    def flipped(self):
        if self._data is None:
            return Act.from_blocked(self._blocked, self._shape, self.lengths, not self.reverse, self.bounded, self.absmax)
        other = Act(self._data, self.lengths, not self.reverse, self.bounded, self.absmax)
        other._blocked = self._blocked
        return other

    def like(self, data, lengths='same', bounded=False, absmax=None):
        return Act(data, self.lengths if lengths == 'same' else lengths, self.reverse, bounded, absmax)


class KernelTimer(object):
    """Optional per-kernel device timing: CUDA events recorded on the launching stream around each
    C-ABI call (used by bench.py for the roofline object; off by default)."""

    def __init__(self):
        self.spans = {}       # kernel name -> [(start_event, end_event)]
        self.launches = 0     # kernels enqueued since reset (counted even when timing is off)
        self.enabled = False

    def reset(self):
        self.spans = {}
        self.launches = 0

    def totals_ms(self):
        """name -> (total ms, number of calls); call after synchronising."""
        return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in self.spans.items()}


TIMER = KernelTimer()

# Number of batches the caller keeps in flight on separate CUDA streams (`basecall.basecall_chunk_stream`, bench.py).
# Only a scheduling hint for the recurrence kernel (how many sequences share a CTA); results do not depend on it.
BATCHES_IN_FLIGHT = 1
_CONCURRENT_BRANCHES = 1      # > 1 while the branches of a Parallel layer are being enqueued on side streams


def set_batches_in_flight(k, gemm_sms=None):
    """Tell the kernels how many batches the caller pipelines on separate streams (1 = none).

    Two scheduling knobs follow from it, neither changes results: the recurrence kernel packs more sequences per
    CTA (fewer, busier SMs per batch), and the tensor-core GEMMs are held to the SMs the other batches' recurrence
    kernels leave free (`gemm_sms`, default 148 - 32 per further batch, never under a quarter of the device;
    SLOIKA_B200_GEMM_SMS overrides), so that their CTAs start at once instead of queueing behind 1 ms kernels."""
    global BATCHES_IN_FLIGHT
    k = max(1, int(k))
    BATCHES_IN_FLIGHT = k
    if gemm_sms is None:
        env = os.environ.get('SLOIKA_B200_GEMM_SMS')
        if env:
            gemm_sms = int(env)
        elif k >= SEQ_MIN_IN_FLIGHT:
            gemm_sms = 92            # the GRU layers run 128 sequences per CTA (8 SMs per 1024-sequence batch): most SMs stay free
        else:
            gemm_sms = 0 if k == 1 else max(37, 148 - 32 * (k - 1))
    cabi.check(cabi.load().sloika_b200_set_gemm_sm_budget(int(gemm_sms)), 'set_gemm_sm_budget')
    return k


def launch(name, nkernels, fn, *args):
    """Enqueue one C-ABI call, bracketed by events when profiling is on."""
    import torch
    TIMER.launches += nkernels
    if TIMER.enabled:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        code = fn(*args)
        end.record()
        TIMER.spans.setdefault(name, []).append((start, end))
    else:
        code = fn(*args)
    cabi.check(code, name)


def _row_stride(t):
    """Distance in floats between consecutive (time, batch) rows of a [T, B, F] tensor."""
    T, B, F = t.shape
    if B > 1:
        return t.stride(1)
    return t.stride(0) if T > 1 else F


def _empty(shape, device):
    import torch
    return torch.empty(shape, dtype=torch.float32, device=device)


def _out_buffer(act, T, F, out):
    """Output activation: a fresh buffer whose rows are padded to 16 bytes (so the next layer's TMA-fed GEMM
    and 128-bit accesses apply to odd widths such as 110 or 142), or the caller's column slice (Parallel)."""
    if out is not None:
        assert out.shape == (T, act.B, F)
        return out
    return _padded_rows(T, act.B, F, act.device)


GEMM_AUTO, GEMM_SIMT, GEMM_TC, GEMM_TC_F16 = 0, 1, 2, 3      # include/sloika_b200.h
_F16_WEIGHT_LIMIT = 1.0e3
_F16_INPUT_LIMIT = 1.0e4          # measured max |x| below which an unbounded activation may use the fp16 split


def _bounded_fun(fun):
    return code_of(fun) in (1, 2)               # tanh, sigmoid


def _gemm_algo(act, *params):
    """fp16-split tensor-core GEMM when the input is provably in [-1, 1] and the weights are far inside the
    fp16 range; otherwise let the library choose (tf32 split / SIMT).  SLOIKA_B200_NO_F16=1 disables it."""
    if act.bounded and not os.environ.get('SLOIKA_B200_NO_F16') and \
            all(p.absmax() < _F16_WEIGHT_LIMIT for p in params):
        return GEMM_TC_F16
    return GEMM_AUTO


def _linear(name, lib, act, W, b, y, ldy, N, fun_code, dev):
    """y = fun(x W' + b) through sloika_linear_fwd_ex; the fp16-split request degrades to AUTO when the
    tensor-core kernel cannot take the shape."""
    tc_ok = act.ld % 4 == 0 and act.data.data_ptr() % 16 == 0 and act.T * act.B >= 128 and act.F <= 512
    if not act.bounded and act.absmax is not None and tc_ok and W.absmax() < _F16_WEIGHT_LIMIT \
            and not os.environ.get('SLOIKA_B200_NO_F16'):
        # range known only on the device: enqueue both tensor-core forms, the kernel-side gate runs exactly one
        launch(name, 2, lib.sloika_linear_fwd_gated,
               cabi.ptr(act.data), act.ld, cabi.ptr(W.device(dev)), cabi.ptr(b.device(dev)), cabi.ptr(y), ldy,
               act.T * act.B, act.F, N, fun_code, cabi.ptr(act.absmax), _F16_INPUT_LIMIT, cabi.stream_ptr(dev))
        return
    algo = _gemm_algo(act, W)
    args = lambda a: (cabi.ptr(act.data), act.ld, cabi.ptr(W.device(dev)), cabi.ptr(b.device(dev)), cabi.ptr(y), ldy,
                      act.T * act.B, act.F, N, fun_code, a, cabi.stream_ptr(dev))
    if algo == GEMM_TC_F16 and tc_ok:
        launch(name, 1, lib.sloika_linear_fwd_ex, *args(GEMM_TC_F16))
    else:
        launch(name, 1, lib.sloika_linear_fwd_ex, *args(GEMM_AUTO))


def run_convolution(layer, act, out=None):
    lib = cabi.load()
    assert act.F == layer.insize, "Convolution input has {} features, expected {}".format(act.F, layer.insize)
    if act.ld != act.F:
        raise ValueError("Convolution needs a dense input")
    if act.reverse:
        # Reverse only flips a direction flag that the recurrent kernels honour; conv(x[::-1])[::-1] differs from
        # conv(x) for even windows, asymmetric padding or stride > 1, and no shipped model reverses a convolution
        raise NotImplementedError("Reverse(Convolution) is not implemented on the device")
    Tout = output_length(act.T, layer.winlen, layer.stride, layer.padding)
    dev = act.device
    bounded = _bounded_fun(layer.fun)
    absmax = None
    if not bounded:                              # elu / linear outputs: let the kernel report their range
        import torch
        absmax = torch.zeros(1, dtype=torch.float32, device=dev)
    lengths = None
    if act.lengths is not None:
        # per-read output length: each read is padded/convolved on its own (conv.py:66-111)
        span = act.lengths + (layer.padding[0] + layer.padding[1] - layer.winlen)
        lengths = (span.clamp(min=-layer.stride) // layer.stride + 1).clamp(min=0).to(act.lengths.dtype)
    if out is None and layer.insize == 1 and layer.winlen == 11 and layer.size % 4 == 0 and layer.size <= 96 and Tout > 0 \
            and layer.stride <= 24 and act.B <= 32 * 65535 and _seq_mode(act.B):
        # the GRU stack behind it runs with sequences on the lanes: write the blocked layout straight away
        yb = _blocked_empty(Tout, act.B, layer.size, dev)
        launch('conv1d', 1, lib.sloika_conv1d_fwd_ex, cabi.ptr(act.data), cabi.ptr(layer.W.device(dev)),
               cabi.ptr(layer.b.device(dev)), cabi.ptr(yb), -1, cabi.ptr(act.lengths), act.T, act.B, layer.insize,
               layer.size, layer.winlen, layer.stride, layer.padding[0], layer.padding[1], code_of(layer.fun),
               cabi.ptr(absmax), cabi.stream_ptr(dev))
        return Act.from_blocked(yb, (Tout, act.B, layer.size), lengths, act.reverse, bounded, absmax)
    y = _out_buffer(act, Tout, layer.size, out)
    launch('conv1d', 1, lib.sloika_conv1d_fwd_ex,
           cabi.ptr(act.data), cabi.ptr(layer.W.device(dev)), cabi.ptr(layer.b.device(dev)), cabi.ptr(y),
           _row_stride(y), cabi.ptr(act.lengths), act.T, act.B, layer.insize, layer.size, layer.winlen,
           layer.stride, layer.padding[0], layer.padding[1], code_of(layer.fun), cabi.ptr(absmax),
           cabi.stream_ptr(dev))
    return act.like(y, lengths, bounded=bounded, absmax=absmax)


def run_feedforward(layer, act, out=None):
    lib = cabi.load()
    assert act.F == layer.insize
    y = _out_buffer(act, act.T, layer.size, out)
    _linear('feedforward', lib, act, layer.W, layer.b, y, _row_stride(y), layer.size, code_of(layer.fun), act.device)
    return act.like(y, bounded=_bounded_fun(layer.fun))


def _padded_rows(T, B, F, device):
    """[T, B, F] view of a buffer whose rows are padded to a multiple of 8 floats: 16-byte aligned rows let the kernels
    use 128-bit accesses and TMA, and 32-byte (one L2 sector) aligned rows keep the 128-byte row segments of the GEMM's
    TMA-store boxes from straddling sectors (the 1025-column logits GEMM is 8 % faster at a pitch of 1032 floats than
    at 1028, tools/gemm_bench.py with PITCH=8)."""
    pitch = (F + 7) // 8 * 8
    return _empty((T, B, pitch), device)[:, :, :F]


class LogitsAct(object):
    """Un-normalised output of the final Softmax layer in the decoder's layout: `data[t, b, :]` holds the
    k-mer state logits first and the stay/blank logit LAST, `stats[t*B + b, s]` the (max, sum exp) pair
    of column slice s.  Produced by `run_softmax_logits`, consumed by `decode.viterbi_batch`; the
    posterior matrix itself is never materialised on this path."""

    def __init__(self, data, stats, n_slices, lengths):
        self.data, self.stats, self.n_slices, self.lengths = data, stats, n_slices, lengths


def _softmax_tc_ok(lib, layer, act, algo):
    return act.ld % 4 == 0 and act.data.data_ptr() % 16 == 0 and act.T * act.B >= 128 and \
        lib.sloika_softmax_slices(layer.insize, layer.size, algo) > 0


def run_softmax_logits(layer, act):
    """Final Softmax for the fused basecall path: logits + row statistics (see LogitsAct); None when the
    tensor-core kernel cannot take this shape (caller falls back to `run_softmax`)."""
    import torch
    lib = cabi.load()
    assert act.F == layer.insize
    dev = act.device
    if act.blocked is not None and act._data is None and act.B % 128 == 0 and act.bounded and \
            layer.W.absmax() < _F16_WEIGHT_LIMIT and not os.environ.get('SLOIKA_B200_NO_F16') and \
            lib.sloika_softmax_slices(layer.insize, layer.size, GEMM_TC_F16) > 0:
        # the last GRU layer left its output in the blocked layout: the GEMM's TMA reads it as it is
        nsl = lib.sloika_softmax_slices(layer.insize, layer.size, GEMM_TC_F16)
        y = _padded_rows(act.T, act.B, layer.size, dev)
        stats = torch.empty((act.T * act.B, nsl, 2), dtype=torch.float32, device=dev)
        launch('softmax', 1, lib.sloika_softmax_logits_blocked_fwd,
               cabi.ptr(act.blocked), cabi.ptr(layer.W.device(dev)), cabi.ptr(layer.b.device(dev)),
               cabi.ptr(y), _row_stride(y), cabi.ptr(stats), act.T * act.B, layer.insize, layer.size, 1, cabi.stream_ptr(dev))
        act.blocked.record_stream(torch.cuda.current_stream(dev))
        return LogitsAct(y, stats, nsl, act.lengths)
    algo = _gemm_algo(act, layer.W)
    if not _softmax_tc_ok(lib, layer, act, algo):
        return None
    nsl = lib.sloika_softmax_slices(layer.insize, layer.size, algo)
    y = _padded_rows(act.T, act.B, layer.size, dev)
    stats = torch.empty((act.T * act.B, nsl, 2), dtype=torch.float32, device=dev)
    launch('softmax', 1, lib.sloika_softmax_logits_fwd,
           cabi.ptr(act.data), act.ld, cabi.ptr(layer.W.device(dev)), cabi.ptr(layer.b.device(dev)),
           cabi.ptr(y), _row_stride(y), cabi.ptr(stats), act.T * act.B, layer.insize, layer.size, 1, algo,
           cabi.stream_ptr(dev))
    return LogitsAct(y, stats, nsl, act.lengths)


def run_softmax(layer, act, out=None):
    import torch
    lib = cabi.load()
    assert act.F == layer.insize
    dev = act.device
    y = out if out is not None else _padded_rows(act.T, act.B, layer.size, dev)
    ldy = _row_stride(y)
    algo = _gemm_algo(act, layer.W)
    if _softmax_tc_ok(lib, layer, act, algo) and ldy % 4 == 0 and y.data_ptr() % 16 == 0:
        # tensor-core logits + per-slice row statistics, then one normalising pass (read + write)
        nsl = lib.sloika_softmax_slices(layer.insize, layer.size, algo)
        stats = torch.empty((act.T * act.B, nsl, 2), dtype=torch.float32, device=dev)
        launch('softmax', 1, lib.sloika_softmax_logits_fwd,
               cabi.ptr(act.data), act.ld, cabi.ptr(layer.W.device(dev)), cabi.ptr(layer.b.device(dev)),
               cabi.ptr(y), ldy, cabi.ptr(stats), act.T * act.B, layer.insize, layer.size, 0, algo,
               cabi.stream_ptr(dev))
        launch('softmax_normalise', 1, lib.sloika_softmax_normalise_fwd,
               cabi.ptr(y), ldy, cabi.ptr(stats), nsl, act.T * act.B, layer.size, cabi.stream_ptr(dev))
    else:
        launch('softmax', 2, lib.sloika_softmax_fwd,
               cabi.ptr(act.data), act.ld, cabi.ptr(layer.W.device(dev)), cabi.ptr(layer.b.device(dev)),
               cabi.ptr(y), ldy, act.T * act.B, layer.insize, layer.size, cabi.stream_ptr(dev))
    return act.like(y, bounded=True)


def _fused_gru_ok(layer, act):
    """The projection-in-the-recurrence launch (csrc/gru_fused.cu) applies: throughput mode (enough sequences in
    flight on the device that 32 per recurrence CTA is the right packing, as in `sloika_gru_recurrence_fwd_ex`),
    sizes that fit the tensor memory / shared memory of its projection CTA, and an input inside the fp16 range:
    returns 'fused' when that is known here (bounded activations), 'gated' when only the device knows (the producing
    kernel left max |x| in `act.absmax`: both forms are enqueued, `sloika_gru_fwd_gated`), else None.
    SLOIKA_B200_FUSED_GRU=1 forces it for any batch, =0 turns it off."""
    env = os.environ.get('SLOIKA_B200_FUSED_GRU', '')
    if env == '0':
        return None
    busy = act.B * max(1, BATCHES_IN_FLIGHT) * _CONCURRENT_BRANCHES > 16 * 148
    aligned = act.blocked is not None or (act.ld % 4 == 0 and act.data.data_ptr() % 16 == 0)
    shape_ok = (busy or env not in ('', '0')) and layer.size <= 96 and layer.insize <= 96 and aligned \
        and code_of(layer.fun) == 1 and code_of(layer.gatefun) == 2 and layer.iW.absmax() < _F16_WEIGHT_LIMIT
    if not shape_ok:
        return None
    # with many batches in flight the sequences-on-lanes launch (csrc/gru_seq.cu: 128 sequences per CTA, ~60 % of the SM
    # time per sequence, but 3.5 x the latency per layer) is the better form; SLOIKA_B200_GRU_SEQ=1 / 0 forces / forbids it
    seq = _seq_mode(act.B)
    if act.bounded:
        return 'seq' if seq else 'fused'
    if act.absmax is not None and act.T * act.B >= 128 and not os.environ.get('SLOIKA_B200_NO_F16'):
        return 'seq_gated' if seq else 'gated'
    return None


SEQ_MIN_IN_FLIGHT = 6


def _seq_mode(B):
    """GRU layers run as the sequences-on-lanes launch (and the layers around them speak its blocked layout): enough
    batches in flight, and batches of at least 512 sequences (small ones -- strong scaling splits a batch over the GPUs --
    would put a single CTA to work for 4 ms per layer: they keep the fused form).  SLOIKA_B200_GRU_SEQ=1 / 0 forces /
    forbids it; SLOIKA_B200_FUSED_GRU=0 (no throughput forms at all) forbids it as well."""
    env = os.environ.get('SLOIKA_B200_GRU_SEQ', '')
    if env == '0' or os.environ.get('SLOIKA_B200_FUSED_GRU', '') == '0':
        return False
    return env == '1' or (B >= 512 and max(1, BATCHES_IN_FLIGHT) * _CONCURRENT_BRANCHES >= SEQ_MIN_IN_FLIGHT)


def _blocked_empty(T, B, F, dev):
    import torch
    n = cabi.load().sloika_blocked_bytes(T, B, F) // 4
    return torch.empty(max(n, 1), dtype=torch.float32, device=dev)


def _run_gru_seq(layer, act, form, out):
    """The layer as the sequences-on-lanes launch: blocked activations in and out (`Act.from_blocked`)."""
    import torch
    lib = cabi.load()
    dev = act.device
    T, B, I, H = act.T, act.B, layer.insize, layer.size
    yb = _blocked_empty(T, B, H, dev)
    stream = torch.cuda.current_stream(dev)
    common = (cabi.ptr(layer.iW.device(dev)), cabi.ptr(layer.sW.device(dev)), cabi.ptr(layer.sW2.device(dev)),
              cabi.ptr(layer.b.device(dev)))
    if form == 'seq_gated':
        y = _padded_rows(T, B, H, dev)                       # written only if the input leaves the fp16 range
        vI = _padded_rows(T, B, 3 * H, dev)
        if act.blocked is not None and act._data is None:    # the convolution wrote the blocked layout
            scratch = _padded_rows(T, B, I, dev)             # its row-major copy, made only if the range check fails
            xptr, ldx, xblk, sptr = cabi.ptr(act.blocked), _row_stride(scratch), 1, cabi.ptr(scratch)
        else:
            scratch = _blocked_empty(T, B, I, dev)
            xptr, ldx, xblk, sptr = cabi.ptr(act.data), act.ld, 0, cabi.ptr(scratch)
        launch('gru_seq', 5, lib.sloika_gru_seq_fwd_gated, xptr, ldx, xblk, *common, cabi.ptr(yb), sptr,
               cabi.ptr(y), _row_stride(y), cabi.ptr(vI), _row_stride(vI), cabi.ptr(act.lengths), T, B, I, H,
               1 if act.reverse else 0, code_of(layer.fun), code_of(layer.gatefun),
               B * max(1, BATCHES_IN_FLIGHT) * _CONCURRENT_BRANCHES, cabi.ptr(act.absmax), _F16_INPUT_LIMIT, cabi.stream_ptr(dev))
        for t in (scratch, y, vI):
            t.record_stream(stream)
    else:
        xb = act.blocked
        if xb is None:
            xb = _blocked_empty(T, B, I, dev)
            launch('block_layout', 1, lib.sloika_block_layout_fwd, cabi.ptr(act.data), cabi.ptr(xb), act.ld, T, B, I, 1,
                   cabi.stream_ptr(dev))
            xb.record_stream(stream)
        launch('gru_seq', 1, lib.sloika_gru_seq_fwd, cabi.ptr(xb), 0, *common, cabi.ptr(yb), 0, cabi.ptr(act.lengths), T, B, I, H,
               1 if act.reverse else 0, code_of(layer.fun), code_of(layer.gatefun), 3, cabi.stream_ptr(dev))
    res = Act.from_blocked(yb, (T, B, H), act.lengths, act.reverse, bounded=True)
    if out is not None:                                      # a Parallel branch: its column slice of the shared buffer
        launch('block_layout', 1, lib.sloika_block_layout_fwd, cabi.ptr(yb), cabi.ptr(out), _row_stride(out), T, B, H, 0,
               cabi.stream_ptr(dev))
        yb.record_stream(stream)
        return act.like(out, bounded=True)
    return res


def run_gru(layer, act, out=None):
    lib = cabi.load()
    assert act.F == layer.insize
    form = _fused_gru_ok(layer, act)
    if form in ('seq', 'seq_gated'):
        return _run_gru_seq(layer, act, form, out)
    y = _out_buffer(act, act.T, layer.size, out)
    dev = act.device
    if form == 'gated':
        import torch
        nbytes = lib.sloika_gru_fused_workspace_bytes(act.B, layer.size)
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        vI = _padded_rows(act.T, act.B, 3 * layer.size, dev)     # written only if the input leaves the fp16 range
        launch('gru_fused', 3, lib.sloika_gru_fwd_gated,
               cabi.ptr(act.data), act.ld, cabi.ptr(layer.iW.device(dev)), cabi.ptr(layer.sW.device(dev)),
               cabi.ptr(layer.sW2.device(dev)), cabi.ptr(layer.b.device(dev)), cabi.ptr(y), _row_stride(y), cabi.ptr(vI),
               _row_stride(vI), cabi.ptr(ws), nbytes, cabi.ptr(act.lengths), act.T, act.B, layer.insize, layer.size,
               1 if act.reverse else 0, code_of(layer.fun), code_of(layer.gatefun),
               act.B * max(1, BATCHES_IN_FLIGHT) * _CONCURRENT_BRANCHES, cabi.ptr(act.absmax), _F16_INPUT_LIMIT,
               cabi.stream_ptr(dev))
        stream = torch.cuda.current_stream(dev)
        ws.record_stream(stream)
        vI.record_stream(stream)
        return act.like(y, bounded=True)
    if form == 'fused':
        import torch
        nbytes = lib.sloika_gru_fused_workspace_bytes(act.B, layer.size)
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        launch('gru_fused', 1, lib.sloika_gru_fused_fwd,
               cabi.ptr(act.data), act.ld, cabi.ptr(layer.iW.device(dev)), cabi.ptr(layer.sW.device(dev)),
               cabi.ptr(layer.sW2.device(dev)), cabi.ptr(layer.b.device(dev)), cabi.ptr(y), _row_stride(y), cabi.ptr(ws),
               nbytes, cabi.ptr(act.lengths), act.T, act.B, layer.insize, layer.size, 1 if act.reverse else 0,
               code_of(layer.fun), code_of(layer.gatefun), cabi.stream_ptr(dev))
        ws.record_stream(torch.cuda.current_stream(dev))
        return act.like(y, bounded=True)
    # sloika_gru_fwd == input projection for all steps + recurrence; issued as its two halves so
    # that each kernel can be timed on its own
    H = layer.size
    vI = _padded_rows(act.T, act.B, 3 * H, dev)          # 16-byte row pitch also for odd H
    st = cabi.stream_ptr(dev)
    _linear('gru_projection', lib, act, layer.iW, layer.b, vI, _row_stride(vI), 3 * H, 0, dev)
    launch('gru_recurrence', 1, lib.sloika_gru_recurrence_fwd_ex,
           cabi.ptr(vI), _row_stride(vI), cabi.ptr(layer.sW.device(dev)), cabi.ptr(layer.sW2.device(dev)), cabi.ptr(y),
           _row_stride(y), cabi.ptr(act.lengths), act.T, act.B, H, 1 if act.reverse else 0,
           code_of(layer.fun), code_of(layer.gatefun), act.B * max(1, BATCHES_IN_FLIGHT) * _CONCURRENT_BRANCHES, st)
    # h is a convex combination of act(.) values when the gates are sigmoids
    return act.like(y, bounded=_bounded_fun(layer.fun) and code_of(layer.gatefun) == 2)


def run_lstm(layer, act, out=None):
    """Lstm.run (`layers.py:693-697`): input projection for all steps (tensor-core GEMM) + the recurrence kernel."""
    lib = cabi.load()
    assert act.F == layer.insize
    y = _out_buffer(act, act.T, layer.size, out)
    dev = act.device
    H = layer.size
    vW = _padded_rows(act.T, act.B, 4 * H, dev)
    _linear('lstm_projection', lib, act, layer.iW, layer.b, vW, _row_stride(vW), 4 * H, 0, dev)
    launch('lstm_recurrence', 1, lib.sloika_lstm_recurrence_fwd,
           cabi.ptr(vW), _row_stride(vW), cabi.ptr(layer.sW.device(dev)), cabi.ptr(layer.p.device(dev)), cabi.ptr(y),
           _row_stride(y), cabi.ptr(act.lengths), act.T, act.B, H, 1 if act.reverse else 0,
           code_of(layer.fun), code_of(layer.gatefun), cabi.stream_ptr(dev))
    # out = fun(state) * gate(.): bounded when fun is and the gate is a sigmoid
    return act.like(y, bounded=_bounded_fun(layer.fun) and code_of(layer.gatefun) == 2)


def run_window(layer, act, out=None):
    """Window.run (`layers.py:346-351`)."""
    lib = cabi.load()
    assert act.F == layer.insize
    if act.reverse:
        # a window is symmetric in time (w odd, equal padding): windowing the reversed input equals reversing the
        # windowed input with the feature blocks in reverse order -- not supported in place, so say so
        raise NotImplementedError("Reverse(Window) is not implemented on the device")
    y = _out_buffer(act, act.T, layer.size, out)
    launch('window', 1, lib.sloika_window_fwd, cabi.ptr(act.data), act.ld, cabi.ptr(y), _row_stride(y),
           cabi.ptr(act.lengths), act.T, act.B, act.F, layer.w, cabi.stream_ptr(act.device))
    return act.like(y, bounded=act.bounded, absmax=act.absmax)


_SLICE_WRITERS = {}


_SIDE_STREAMS = {}            # (device, launching stream) -> side streams for the branches of a Parallel layer


def _side_streams(dev, n):
    import torch
    key = (str(dev), torch.cuda.current_stream(dev).cuda_stream)
    have = _SIDE_STREAMS.setdefault(key, [])
    while len(have) < n:
        have.append(torch.cuda.Stream(dev))
    return have[:n]


def run_parallel(layer, act, out=None):
    """Parallel.run (`layers.py:1486-1487`): sub-layers write their column slice of one buffer.

    The branches are independent, so every branch after the first is enqueued on a side stream that forks from and
    joins the launching stream: the two directions of a birnn (`layers.py:1622-1629`) scan at the same time instead
    of paying the recurrence latency twice per layer (each occupies at most 128 SMs at one CTA per SM)."""
    import torch
    from sloika_b200 import layers as L
    y = _out_buffer(act, act.T, layer.size, out)
    dev = act.device
    nsub = len(layer.layers)
    concurrent = nsub > 1 and not os.environ.get('SLOIKA_B200_SERIAL_BRANCHES')
    main = torch.cuda.current_stream(dev)
    sides = _side_streams(dev, nsub - 1) if concurrent else []
    if concurrent:
        fork = torch.cuda.Event()
        fork.record(main)
    global _CONCURRENT_BRANCHES
    saved_branches = _CONCURRENT_BRANCHES
    if concurrent:
        _CONCURRENT_BRANCHES = saved_branches * nsub       # the recurrence kernels of the branches share the SMs
    col = 0
    result = None
    bounded = True
    for idx, sub in enumerate(layer.layers):
        view = y[:, :, col:col + sub.size]
        inner, flips = sub, 0
        while isinstance(inner, L.Reverse):
            inner, flips = inner.layer, flips + 1
        src = act.flipped() if flips % 2 else act
        stream = sides[idx - 1] if (concurrent and idx > 0) else main
        if stream is not main:
            stream.wait_event(fork)
        with torch.cuda.stream(stream):
            if isinstance(inner, L.Gru):
                res = run_gru(inner, src, out=view)
            elif isinstance(inner, L.Lstm):
                res = run_lstm(inner, src, out=view)
            elif isinstance(inner, L.FeedForward):
                res = run_feedforward(inner, src, out=view)
            elif isinstance(inner, L.Softmax):
                res = run_softmax(inner, src, out=view)
            else:
                res = sub.run(act)                  # nested containers: run, then place
                view.copy_(res.data)
                res = act.like(view, res.lengths, bounded=res.bounded)
                flips = 0
        if stream is not main:
            main.wait_stream(stream)
        if result is None:
            result = res.flipped() if flips % 2 else res
        bounded = bounded and res.bounded
        col += sub.size
    _CONCURRENT_BRANCHES = saved_branches
    return Act(y, result.lengths, act.reverse, bounded)


class CompiledNetwork(object):
    """`calc_post`: the callable `Layer.compile()` returns (`layers.py:34-36`, used at
    `basecall.py:119`).

    NumPy in -> NumPy out (drop-in for the reference: float32 `[T, B, F]` C-contiguous, anything
    else raises TypeError like a Theano function does); torch CUDA tensor in -> torch CUDA tensor
    out (device path, no host copies).  `lengths` enables the ragged whole-read batch.
    """

    def __init__(self, network, device=None):
        self.network = network
        self._device = device

    @property
    def device(self):
        import torch
        if self._device is None:
            if not torch.cuda.is_available():
                raise cabi.SloikaB200Error("no CUDA device: the B200 path has no CPU fallback")
            self._device = torch.device('cuda', torch.cuda.current_device())
        return self._device

    def to(self, device):
        import torch
        self._device = torch.device(device)
        return self

    def prepare(self):
        """Upload every parameter to the device and wait for it: callers that enqueue batches on several CUDA
        streams call this once first, so that no stream reads a weight buffer another stream is still filling."""
        import torch
        from sloika_b200 import layers as L

        def visit(layer):
            if isinstance(layer, L._Container):
                for child in layer._children():
                    visit(child)
            else:
                for p in layer.params():
                    p.device(self.device)
                for key in getattr(layer, '_ordered', lambda: [])():
                    getattr(layer, key).device(self.device)
        visit(self.network)
        torch.cuda.synchronize(self.device)
        return self

    def forward_device(self, x, lengths=None, fused_decode=False):
        """x: float32 CUDA tensor [T, B, F] (dense); lengths: optional int32 CUDA tensor [B].

        With `fused_decode` the final Softmax layer hands un-normalised logits + row statistics to the
        decoder (`LogitsAct`) instead of writing posteriors; only `decode.viterbi_batch` understands
        that form.  Falls back to posteriors when the network or shape does not allow it."""
        import torch
        from sloika_b200 import layers as L
        cabi.load()
        if x.dtype != torch.float32 or x.dim() != 3:
            raise TypeError("calc_post expects a float32 [time, batch, feature] tensor")
        x = x.contiguous()
        if lengths is not None:
            lengths = lengths.to(device=x.device, dtype=torch.int32).contiguous()
        with torch.cuda.device(x.device):
            net = self.network
            if fused_decode and isinstance(net, L.Serial) and isinstance(net.layers[-1], L.Softmax):
                act = Act(x, lengths)
                for layer in net.layers[:-1]:
                    act = layer.run(act)
                out = run_softmax_logits(net.layers[-1], act)
                if out is None:
                    out = net.layers[-1].run(act)
            else:
                out = net.run(Act(x, lengths))
        return out

    def __call__(self, inMat, lengths=None):
        import torch
        if isinstance(inMat, np.ndarray):
            if inMat.dtype != np.dtype(sloika_dtype) or inMat.ndim != 3:
                raise TypeError("calc_post expects a float32 ndarray of shape [time, batch, feature], "
                                "got {} {}".format(inMat.dtype, inMat.shape))
            dev = self.device
            x = torch.from_numpy(np.ascontiguousarray(inMat)).to(dev)
            lens = None if lengths is None else torch.as_tensor(np.asarray(lengths), dtype=torch.int32, device=dev)
            out = self.forward_device(x, lens)
            return out.data.contiguous().cpu().numpy()
        out = self.forward_device(inMat, lengths)
        return out.data
