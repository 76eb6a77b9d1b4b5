"""Posterior error of the device path on a real read (bundled read7, pretrained weights) against the float64 oracle,
for the operand formats of the first GRU projection (fp16 split chosen on the device vs tf32 split), next to the
float32 NumPy oracle's own distance from float64.  Puts the 1e-4 max-abs parity bound in context."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np

from oracle import forward_ref
from sloika_b200 import basecall, zoo

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def main():
    with open(os.path.join(GOLDEN, 'pretrained_arch.json')) as fh:
        arch = json.load(fh)
    weights = dict(np.load(os.path.join(GOLDEN, 'pretrained_weights.npz')))
    net = zoo.from_weights(arch, weights)
    daq = np.load(os.path.join(GOLDEN, 'reads_daq.npz'))
    calc = net.compile()
    for name in ('read7', 'read5'):
        offset, rng, digi = daq[name + '_scaling']
        sig = basecall.prepare_signal((daq[name] + offset) * (rng / digi), (200, 10), 0)
        x = sig[:, None, None]
        ref64 = forward_ref.run(net.json(params=True), x, np.float64)
        ref32 = forward_ref.run(net.json(params=True), x, np.float32)
        out = {}
        for label, env in (('gated fp16/tf32', {}), ('tf32 only', {'SLOIKA_B200_NO_F16': '1'}),
                           ('mma.sync recurrence', {'SLOIKA_B200_GRU': 'v4'}),
                           ('mma.sync + tf32', {'SLOIKA_B200_GRU': 'v4', 'SLOIKA_B200_NO_F16': '1'})):
            os.environ.update(env)
            out[label] = calc(x)
            for k in env:
                os.environ.pop(k, None)
        print("%s: %d steps" % (name, ref64.shape[0]))
        lp64 = np.log(ref64.max(2))

        def bias(post):
            return float(np.mean(np.log(post.astype(np.float64).max(2)) - lp64))
        print("   float32 NumPy oracle vs float64 : max abs %.3e   mean log-posterior bias of the best state %+.2e" % (
            np.abs(ref32 - ref64).max(), bias(ref32)))
        for label, post in out.items():
            print("   device (%-19s) vs float64 : max abs %.3e   vs float32 oracle: %.3e   bias %+.2e" % (
                label, np.abs(post - ref64).max(), np.abs(post - ref32).max(), bias(post)))


if __name__ == '__main__':
    main()
