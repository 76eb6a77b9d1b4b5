"""One launch of each of the row-f kernels at a representative size (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sloika_b200 import basecall, decode, transducer
DEV = torch.device('cuda:0')
rng = np.random.default_rng(1)
gen = torch.Generator(device=DEV); gen.manual_seed(1)
T, B, P, S = 800, 1024, 400, 1025
lt = torch.log_softmax(3.0 * torch.randn((T, B, S), generator=gen, device=DEV), dim=-1)
seqs = [rng.integers(1, S, size=P).astype(np.int32) for _ in range(B)]
for _ in range(2):
    transducer.map_to_sequence_batch(lt, seqs, slip=5.0, return_device=True)
sigs = [rng.standard_normal(40000) * 8 + 95 for _ in range(256)]
for _ in range(2):
    basecall.prepare_signals_device(sigs, (200, 10), 0, device=DEV)
paths = torch.from_numpy(rng.integers(0, 1024, size=(1024, 800)).astype(np.int32)).to(DEV)
plen = torch.full((1024,), 800, dtype=torch.int32, device=DEV)
for _ in range(2):
    decode.paths_to_sequences(paths, plen, 5, 'ACGT', True)
torch.cuda.synchronize()
