"""sloika_b200: the raw-signal basecall hot path of nanoporetech/sloika on B200 (sm_100a).

Modules mirror the reference's names for this path (`layers`, `activation`, `conv`, `decode`,
`basecall`, `helpers`, `bio`, `maths`, `variables`, `module_tools`) so a caller swaps
`import sloika.X` for `import sloika_b200.X`; `sloika_b200.install_as_sloika()` does that swap
in `sys.modules` for unmodified model scripts and pickles.
"""
__version__ = '0.1.0'

import os as _os

# The throughput path keeps up to 16 batches in flight, each on its own CUDA stream.  The driver maps streams onto
# CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8): streams that share a queue serialise each other's kernels
# (measured: 12 batches in flight 4.9 ms per batch with 8 queues, 4.2 ms with 32).  Read when the CUDA context is
# created, so it has to be in the environment before the first CUDA call of the process.
_os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')


def install_as_sloika():
    """Alias this package as `sloika` so `models/*.py` (`import sloika.module_tools as smt`) run as is."""
    import importlib
    import sys
    sys.modules.setdefault('sloika', sys.modules[__name__])
    for sub in ('config', 'variables', 'activation', 'conv', 'layers', 'module_tools', 'decode',
                'bio', 'maths', 'basecall', 'helpers', 'batch', 'util', 'fast5', 'transducer', 'viterbi_helpers'):
        mod = importlib.import_module('sloika_b200.' + sub)
        sys.modules.setdefault('sloika.' + sub, mod)
