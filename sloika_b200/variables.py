"""k-mer / state counting constants (reference: `sloika/variables.py:1-26`)."""
DEFAULT_ALPHABET = b'ACGT'
DEFAULT_NBASE = len(DEFAULT_ALPHABET)


def nkmer(kmer, nbase=DEFAULT_NBASE):
    """Number of k-mers of length `kmer` over `nbase` letters (`variables.py:5-13`)."""
    return nbase ** kmer


def nstate(kmer, transducer=True, bad_state=True, nbase=DEFAULT_NBASE):
    """Number of network output states (`variables.py:16-26`): k-mers plus one stay/bad state."""
    return nkmer(kmer, nbase=nbase) + (transducer or bad_state)
