"""NumPy/stdlib restatement of the host-side pre/post-processing.  TEST INFRASTRUCTURE ONLY.

Signal: `sloika/batch.py:194-220` (`trim_open_pore`), `sloika/util.py:94-99` (`trim_array`),
`sloika/maths.py:4-45` (`med_mad`, `mad`), normalisation at `sloika/basecall.py:117-118`.
Sequence: `sloika/bio.py:12-24` (`all_kmers`), `:160-179` (`max_overlap`), `:206-225`
(`reduce_kmers`), `:228-237` (`kmers_to_sequence`), FASTA record at `sloika/basecall.py:157-163`.
PINNED against the reference modules (`maths.py`, `bio.py` import cleanly) by
`tools/make_golden.py`; `batch.py`/`util.py` import h5py/Bio at module top and cannot be imported,
so `trim_open_pore`/`trim_array` are restated from source only.
"""
import itertools

import numpy as np


def med_mad(data, factor=1.4826, axis=None):
    dmed = np.median(data, axis=axis, keepdims=True)
    dmad = factor * np.median(np.abs(data - dmed), axis=axis, keepdims=True)
    if axis is None:
        return dmed.ravel()[0], dmad.ravel()[0]
    return dmed.squeeze(axis), dmad.squeeze(axis)


def trim_open_pore(signal, max_op_fraction=0.3, window_size=100):
    """batch.py:194-220 with var_method='mad': keep first..last window whose MAD exceeds the
    `100*max_op_fraction` percentile of the window MADs."""
    ml = len(signal) // window_size
    chunks = signal[:ml * window_size].reshape(ml, window_size)
    _, local_var = med_mad(chunks, axis=1)
    keep = np.nonzero(local_var > np.percentile(local_var, 100 * max_op_fraction))[0]
    return signal[keep.min() * window_size:(keep.max() + 1) * window_size]


def trim_array(x, from_start, from_end):
    assert from_start >= 0 and from_end >= 0
    return x[from_start:(len(x) - from_end if from_end else None)]


def normalise(signal):
    """basecall.py:117-118: (signal - median) / mad, cast to float32, shape [T, 1, 1]."""
    med, mad = med_mad(signal)
    return ((signal - med) / mad)[:, None, None].astype(np.float32)


def prepare_signal(signal, trim=(200, 10), open_pore_fraction=0.0):
    """basecall.py:111-118 up to the network input; None for reads that trim to nothing."""
    signal = trim_open_pore(signal, open_pore_fraction)
    signal = trim_array(signal, *trim)
    if signal.size == 0:
        return None
    return normalise(signal)


def all_kmers(length, alphabet='ACGT'):
    return [''.join(x) for x in itertools.product(alphabet, repeat=length)]


def kmers_to_sequence(kmers, always_move=False):
    """bio.py:228-237: overlap consecutive k-mers by the smallest shift that matches (identical
    k-mers count as a stay only when `always_move` is False), append the new bases."""
    kmers = list(kmers)
    seq = kmers[0]
    for k1, k2 in zip(kmers, kmers[1:]):
        klen = len(k1)
        if (not always_move) and k1 == k2:
            continue
        move = klen
        for i in range(1, klen):
            if k1[i:] == k2[:-i]:
                move = i
                break
        seq += k2 if move >= klen else k2[-move:]
    return seq


def fasta_record(read_name, score, call, nev, kmer_len=5, alphabet='ACGT', datatype='samples',
                 transducer=True):
    """basecall.py:157-163 (`SeqPrinter.write`)."""
    kmers = all_kmers(kmer_len, alphabet)
    seq = kmers_to_sequence([kmers[i] for i in call], always_move=transducer)
    head = ">{} score {:.0f}, {} {} to {} bases\n".format(read_name, score, nev, datatype, len(seq))
    return head + seq + '\n'
