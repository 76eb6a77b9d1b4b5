"""Achievable HBM bandwidth for pure writes, pure reads and copies (torch fill_ / sum / copy_ on 4 GiB), to put the
write-dominated kernels (softmax logits, GRU projection) in context."""
import torch
dev = torch.device('cuda:0')
n = 1 << 30                      # floats: 4 GiB
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timeit(lambda: a.fill_(1.0)); print("write  (fill_ 4 GiB): %.3f ms  %.0f GB/s" % (ms, 4 * n / ms / 1e6))
ms = timeit(lambda: a.sum());      print("read   (sum   4 GiB): %.3f ms  %.0f GB/s" % (ms, 4 * n / ms / 1e6))
ms = timeit(lambda: b.copy_(a));   print("copy   (4 GiB -> 4 GiB): %.3f ms  %.0f GB/s (read + write)" % (ms, 8 * n / ms / 1e6))
