"""Median / MAD normalisation of the raw signal (reference `sloika/maths.py:4-45`).

Host-side float64 NumPy, like the reference: two medians per read, a few hundred microseconds --
not a kernel.
"""
import numpy as np

_MAD_FACTOR = 1.4826    # consistency with the s.d. of a normal distribution


def med_mad(data, factor=None, axis=None, keepdims=False):
    """(median, scaled median-absolute-deviation) of `data` (`maths.py:4-26`)."""
    scale = _MAD_FACTOR if factor is None else factor
    centre = np.median(data, axis=axis, keepdims=True)
    spread = scale * np.median(np.abs(data - centre), axis=axis, keepdims=True)
    if axis is None:
        return centre.ravel()[0], spread.ravel()[0]
    if keepdims:
        return centre, spread
    return centre.squeeze(axis), spread.squeeze(axis)


def mad(data, factor=None, axis=None, keepdims=False):
    """Scaled MAD alone (`maths.py:29-45`)."""
    return med_mad(data, factor=factor, axis=axis, keepdims=keepdims)[1]


def studentise(x, axis=None):
    """(x - mean) / sd with sd = 1 where it vanishes (`maths.py:48-60`)."""
    m = np.mean(x, axis=axis, keepdims=True)
    s = np.std(x, axis=axis, keepdims=True)
    return (x - m) / np.where(s > 0.0, s, 1.0)
