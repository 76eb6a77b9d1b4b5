"""Time the logits GEMM (M = 819 200, K = 96, N = 1025, fp16-split form with row statistics) fed by a row-major and by a
blocked activation (sloika_softmax_logits_fwd / sloika_softmax_logits_blocked_fwd)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sloika_b200 import cabi

lib = cabi.load()
dev = torch.device('cuda:0')
T, B, K, N = 800, 1024, 96, 1025
M = T * B
x = torch.tanh(torch.randn(T, B, K, device=dev))
xb = torch.empty(lib.sloika_blocked_bytes(T, B, K) // 4, device=dev)
st = cabi.stream_ptr(dev)
assert lib.sloika_block_layout_fwd(cabi.ptr(x), cabi.ptr(xb), K, T, B, K, 1, st) == 0
W = torch.randn(N, K, device=dev) * 0.3
b = torch.randn(N, device=dev)
nsl = lib.sloika_softmax_slices(K, N, 3)
y1 = torch.empty(M, 1032, device=dev)
y2 = torch.empty(M, 1032, device=dev)
s1 = torch.empty(M, nsl, 2, device=dev)
s2 = torch.empty(M, nsl, 2, device=dev)
row = lambda: lib.sloika_softmax_logits_fwd(cabi.ptr(x), K, cabi.ptr(W), cabi.ptr(b), cabi.ptr(y1), 1032, cabi.ptr(s1), M, K, N, 1, 3, st)
blk = lambda: lib.sloika_softmax_logits_blocked_fwd(cabi.ptr(xb), cabi.ptr(W), cabi.ptr(b), cabi.ptr(y2), 1032, cabi.ptr(s2), M, K, N, 1, st)
for name, fn in (('row-major', row), ('blocked', blk)):
    for _ in range(2):
        assert fn() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-10s x: %.3f ms" % (name, e0.elapsed_time(e1) / 5))
print("identical logits:", torch.equal(y1[:, :N], y2[:, :N]), " identical statistics:", torch.equal(s1, s2))
