"""Time the GRU recurrence kernel alone (default: the benchmark shape T'=800, B=1024, H=96; env T, B, H)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sloika_b200 import cabi
lib = cabi.load()
dev = torch.device('cuda:0')
T, B, H = int(os.environ.get('T', '800')), int(os.environ.get('B', '1024')), int(os.environ.get('H', '96'))
vI = torch.randn(T, B, 3 * H, device=dev)
sW = torch.randn(2 * H, H, device=dev) * 0.1
sW2 = torch.randn(H, H, device=dev) * 0.1
y = torch.empty(T, B, H, device=dev)
st = cabi.stream_ptr(dev)
def run():
    return lib.sloika_gru_recurrence_fwd(cabi.ptr(vI), 3 * H, cabi.ptr(sW), cabi.ptr(sW2), cabi.ptr(y), H, None, T, B, H, 0, 1, 2, st)
for _ in range(2): rc = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): rc = run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("H=%d rc=%d %.3f ms  %.2f us/step  dbg=%s threads=%s" % (H, rc, ms, ms * 1e3 / T, os.environ.get('SLOIKA_B200_GRU_DBG', '0'), os.environ.get('SLOIKA_B200_GRU_THREADS', 'default')))
