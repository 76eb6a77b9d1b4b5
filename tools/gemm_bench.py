"""Time sloika_linear_fwd_ex (tensor-core vs SIMT) on the benchmark's projection / logits shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sloika_b200 import cabi
lib = cabi.load()
dev = torch.device('cuda:0')
PITCH = int(os.environ.get('PITCH', '0'))
def run(M, K, N, algo, reps=5):
    pad = lambda n: (n + 3) // 4 * 4
    padn = (lambda n: (n + PITCH - 1) // PITCH * PITCH) if PITCH else pad      # env PITCH: output row pitch a multiple of this many floats
    x = torch.tanh(torch.randn(M, pad(K), device=dev))[:, :K]; W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
    y = torch.empty(M, padn(N), device=dev)[:, :N]
    st = cabi.stream_ptr(dev)
    for _ in range(2):
        rc = lib.sloika_linear_fwd_ex(cabi.ptr(x), x.stride(0), cabi.ptr(W), cabi.ptr(b), cabi.ptr(y), y.stride(0), M, K, N, 0, algo, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        rc = lib.sloika_linear_fwd_ex(cabi.ptr(x), x.stride(0), cabi.ptr(W), cabi.ptr(b), cabi.ptr(y), y.stride(0), M, K, N, 0, algo, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = (M * K + M * N) * 4 / 1e9
    return rc, ms, gb / ms * 1e3, 2.0 * M * K * N / ms / 1e9
shapes = [(819200, 96, 288), (819200, 96, 1025), (819200, 128, 336), (819200, 144, 336)]
if os.environ.get("SHAPES") == "one":
    shapes = [(2048000, 128, 336)]
elif os.environ.get("SHAPES") == "sweep":
    shapes = [(819200, 96, 288), (819200, 96, 1025), (2048000, 128, 330), (2048000, 110, 426), (2048000, 142, 330),
              (2048000, 110, 1025), (819200, 128, 336), (819200, 112, 432), (819200, 144, 336), (819200, 112, 1025),
              (2048000, 32, 288), (2048000, 192, 128), (2048000, 128, 1025)]
elif os.environ.get("SHAPES") == "bench":
    shapes = [(819200, 96, 288), (819200, 96, 1025)]
elif os.environ.get("SHAPES") == "rGr":
    shapes = [(2048000, 128, 330), (2048000, 110, 426), (2048000, 142, 330), (2048000, 128, 336), (2048000, 112, 432), (2048000, 110, 1025)]
for (M, K, N) in shapes:
    for algo in ((3,) if os.environ.get('ALGOS') == 'f16' else (3, 2, 1)):
        rc, ms, gbs, tf = run(M, K, N, algo)
        if os.environ.get('TERSE'):
            print("%d %d %d %s %.3f" % (M, K, N, os.environ.get('SLOIKA_B200_GEMM_STAGES', '-'), ms)); continue
        print("M=%d K=%d N=%d algo=%s rc=%d: %.3f ms  %.0f GB/s  %.1f TFLOP/s  dbg=%s" % (M, K, N, {1: 'simt', 2: 'tc', 3: 'tc_f16'}[algo], rc, ms, gbs, tf, os.environ.get('SLOIKA_B200_GEMM_DBG', '0')))
