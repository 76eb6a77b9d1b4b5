"""Padding rules of the 1-D convolution (reference `sloika/conv.py:10-63`).

Only the host-side geometry lives here; the convolution itself is the CUDA kernel behind
`sloika_conv1d_fwd` (zero padding is done by predication inside the kernel, never by a padded copy,
unlike `conv.py:66-77`).
"""

PADDING_MODES = frozenset(['same', 'half', 'valid', 'full', 'same_left'])


def calculate_padding(mode, winlen):
    """(start, end) zero padding for a window of `winlen` (`conv.py:10-63`).

    'same' -> ((w-1)//2, w//2); 'half' -> (w//2, w//2); 'valid' -> (0, 0); 'full' -> (w-1, w-1);
    'same_left' -> (w//2, (w-1)//2); an int p -> (p, p); a pair of ints is used as is.
    """
    assert winlen > 0, "winlen must be positive"
    if isinstance(mode, int):
        return (mode, mode)
    if isinstance(mode, (tuple, list)) and len(mode) == 2 and all(isinstance(m, int) for m in mode):
        return tuple(mode)
    assert mode in PADDING_MODES, 'Padding mode "{}" not supported'.format(mode)
    lo, hi = (winlen - 1) // 2, winlen // 2
    table = {'same': (lo, hi), 'half': (hi, hi), 'valid': (0, 0),
             'full': (winlen - 1, winlen - 1), 'same_left': (hi, lo)}
    return table[mode]


def output_length(ntime, winlen, stride, padding):
    """Number of output steps of a valid strided correlation over the padded input
    (what `T.nnet.conv2d(..., subsample=(1, stride))` yields at `conv.py:107-108`)."""
    span = ntime + padding[0] + padding[1] - winlen
    return 0 if span < 0 else span // stride + 1
