"""Event features for the events route (`sloika/features.py:6-37`): host code, as in the reference (a handful of
column operations on a per-read event table before the network runs)."""
import numpy as np

from sloika_b200.config import sloika_dtype
from sloika_b200.maths import studentise


def from_events(ev, tag='scaled_', normalise=True, nanonet=False):
    """  Create a matrix of features from events

    :param ev: A structured array with fields 'mean', 'stdv' and 'length' (prefixed by `tag` for the first two)
    :param tag: Prefix of which fields to read
    :param normalise: Perform normalisation (Studentisation) of features.
    :param nanonet: Use Nanonet-like features

    :returns: A :class:`ndarray` with studentised features `[events, 4]`: mean, stdv, length, |delta mean|
    """
    nev = len(ev)
    features = np.zeros((nev, 4), dtype=sloika_dtype)
    features[:, 0] = ev[tag + 'mean']
    features[:, 1] = ev[tag + 'stdv']
    features[:, 2] = ev['length']
    features[:, 3] = np.fabs(np.ediff1d(ev[tag + 'mean'], to_end=0))       # zero padded delta mean
    if normalise:
        features = studentise(features, axis=0)
    if nanonet:
        features[:, 3] = np.ediff1d(ev[tag + 'mean'], to_end=0)             # delta mean uncentred
        features[:, 3] /= np.std(features[:, 3])
    return np.ascontiguousarray(features, dtype=sloika_dtype)
