"""Raw-signal trimming used by the basecall path (reference `sloika/batch.py:194-220`).

Only `trim_open_pore` is on the path; the rest of the reference module is chunkify / training-data
preparation (out of scope).
"""
import numpy as np

from sloika_b200 import maths

TRIM_OPEN_PORE_LOCAL_VAR_METHODS = frozenset(['mad', 'std'])


def trim_open_pore(signal, max_op_fraction=0.3, var_method='mad', window_size=100):
    """Locate the read inside `signal` by thresholding the local variation (`batch.py:194-220`).

    The signal is cut into windows of `window_size`; windows whose variation exceeds the
    `100*max_op_fraction` percentile are "probably read", and the slice from the first to the last
    such window is returned.
    """
    assert var_method in TRIM_OPEN_PORE_LOCAL_VAR_METHODS, "var_method not understood: {}".format(var_method)
    nwin = len(signal) // window_size
    windows = signal[:nwin * window_size].reshape((nwin, window_size))
    local_var = windows.std(1) if var_method == 'std' else maths.mad(windows, axis=1)
    is_read = np.flatnonzero(local_var > np.percentile(local_var, 100 * max_op_fraction))
    return signal[is_read.min() * window_size:(is_read.max() + 1) * window_size]
