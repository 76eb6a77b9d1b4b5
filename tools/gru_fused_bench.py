"""Time one GRU layer alone, as projection GEMM + recurrence and as the fused launch (csrc/gru_fused.cu), with one
batch and with K batches on K streams.  Env: T, B, H, I, K."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
LAYOUT = int(os.environ.get('LAYOUT', '0'))   # sloika_gru_seq_fwd: bit 0 x blocked, bit 1 y blocked (timing only here)
from sloika_b200 import cabi
lib = cabi.load()
dev = torch.device('cuda:0')
T, B, H, I, K = (int(os.environ.get(k, d)) for k, d in (('T', '800'), ('B', '1024'), ('H', '96'), ('I', '96'), ('K', '4')))
g = torch.Generator(device='cpu').manual_seed(1)
iW = (torch.randn(3 * H, I, generator=g) * 0.2).to(dev)
sW = (torch.randn(2 * H, H, generator=g) * 0.1).to(dev)
sW2 = (torch.randn(H, H, generator=g) * 0.1).to(dev)
b = (torch.randn(3 * H, generator=g) * 0.1).to(dev)
streams = [torch.cuda.Stream(dev) for _ in range(K)]
xs = [torch.tanh(torch.randn(T, B, I, device=dev)) for _ in range(K)]
ys = [torch.empty(T, B, H, device=dev) for _ in range(K)]
vIs = [torch.empty(T, B, 3 * H, device=dev) for _ in range(K)]
nws = lib.sloika_gru_fused_workspace_bytes(B, H)
wss = [torch.empty(max(nws, 1), dtype=torch.uint8, device=dev) for _ in range(K)]


def unfused(i, st, hint):
    rc = lib.sloika_linear_fwd_ex(cabi.ptr(xs[i]), I, cabi.ptr(iW), cabi.ptr(b), cabi.ptr(vIs[i]), 3 * H, T * B, I, 3 * H, 0, 3, st)
    assert rc == 0, rc
    rc = lib.sloika_gru_recurrence_fwd_ex(cabi.ptr(vIs[i]), 3 * H, cabi.ptr(sW), cabi.ptr(sW2), cabi.ptr(ys[i]), H, None,
                                          T, B, H, 0, 1, 2, hint, st)
    assert rc == 0, rc


def fused(i, st, hint):
    rc = lib.sloika_gru_fused_fwd(cabi.ptr(xs[i]), I, cabi.ptr(iW), cabi.ptr(sW), cabi.ptr(sW2), cabi.ptr(b), cabi.ptr(ys[i]), H,
                                  cabi.ptr(wss[i]), nws, None, T, B, I, H, 0, 1, 2, st)
    assert rc == 0, rc


def seq(i, st, hint):
    rc = lib.sloika_gru_seq_fwd(cabi.ptr(xs[i]), I, cabi.ptr(iW), cabi.ptr(sW), cabi.ptr(sW2), cabi.ptr(b), cabi.ptr(ys[i]), H,
                                None, T, B, I, H, 0, 1, 2, LAYOUT, st)
    assert rc == 0, rc


def timeit(fn, k, reps=3):
    for _ in range(2):
        for i in range(k):
            with torch.cuda.stream(streams[i]):
                fn(i, streams[i].cuda_stream, B * k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams[:k]:
        s.wait_event(e0)
    for _ in range(reps):
        for i in range(k):
            with torch.cuda.stream(streams[i]):
                fn(i, streams[i].cuda_stream, B * k)
    for s in streams[:k]:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for k in (1, K):
    u, f, q = timeit(unfused, k), timeit(fused, k), timeit(seq, k)
    print("T=%d B=%d H=%d I=%d  %d batch(es) in flight: projection+recurrence %.3f ms, fused %.3f ms, sequences-on-lanes %.3f ms (per round of %d)" % (T, B, H, I, k, u, f, q, k))
unfused(0, streams[0].cuda_stream, B); fused(1, streams[1].cuda_stream, B)
torch.cuda.synchronize()
xs[1].copy_(xs[0]); fused(1, streams[1].cuda_stream, B); torch.cuda.synchronize()
print("max |fused - unfused| = %.3g" % (ys[0] - ys[1]).abs().max().item())
seq(2 % K, streams[0].cuda_stream, B) if K > 2 else None
if K > 2:
    xs[2].copy_(xs[0]); seq(2, streams[2].cuda_stream, B); torch.cuda.synchronize()
    print("max |sequences-on-lanes - unfused| = %.3g" % (ys[0] - ys[2]).abs().max().item())
