"""Small helpers of the basecall path (reference `sloika/util.py:94-99`, `:22-40`)."""


def trim_array(x, from_start, from_end):
    """Drop `from_start` leading and `from_end` trailing entries (`util.py:94-99`)."""
    assert from_start >= 0
    assert from_end >= 0
    stop = None if from_end == 0 else -from_end
    return x[from_start:stop]


def get_kwargs(args, names):
    """Pick `names` out of an argparse namespace as a dict (used by `bin/basecall_network.py:100`)."""
    return {name: getattr(args, name) for name in names}
