import sys, os
sys.path.insert(0, '/root/repo')
import torch
from sloika_b200 import cabi
lib = cabi.load(); dev = torch.device('cuda:0')
M, K = 819200, 96
x = torch.tanh(torch.randn(M, K, device=dev)); st = cabi.stream_ptr(dev)
def run(N, pitch, reps=5):
    W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
    y = torch.empty(M, pitch, device=dev)
    f = lambda: lib.sloika_linear_fwd_ex(cabi.ptr(x), K, cabi.ptr(W), cabi.ptr(b), cabi.ptr(y), pitch, M, K, N, 0, 3, st)
    for _ in range(2): assert f() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("N=%4d pitch=%4d: %.3f ms  %.0f GB/s written" % (N, pitch, ms, M * N * 4 / ms / 1e6))
for N, pitch in ((224, 224), (224, 1032), (224, 4128), (256, 256), (256, 1032), (1024, 1024), (1025, 1032), (448, 448), (448, 1032)):
    run(N, pitch)
