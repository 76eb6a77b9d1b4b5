"""k-mer path -> base sequence (reference `sloika/bio.py:12-24, 160-179, 206-237`).

Host post-processing of the decoded path, as in the reference (`SeqPrinter.write`,
`sloika/basecall.py:157-163`).  States are integers on this path, so the overlap test is done on
integers (k-mer j = sum base_i * nbase^(k-1-i)) instead of on strings; the string API of the
reference is kept on top of it.
"""
import itertools


def all_kmers(length, alphabet='ACGT'):
    """All k-mers in alphabet order (`bio.py:12-24`); bytes alphabet -> bytes k-mers."""
    if isinstance(alphabet, bytes):
        letters = alphabet.decode('utf-8')
        return [''.join(x).encode('utf-8') for x in itertools.product(letters, repeat=length)]
    return [''.join(x) for x in itertools.product(alphabet, repeat=length)]


def max_overlap(kmers, allow_identical=True):
    """Smallest shift aligning each k-mer with its successor (`bio.py:160-179`)."""
    kmers = list(kmers)
    moves = []
    for prev, nxt in zip(kmers, kmers[1:]):
        klen = len(prev)
        if allow_identical and prev == nxt:
            moves.append(0)
            continue
        shift = next((i for i in range(1, klen) if prev[i:] == nxt[:-i]), klen)
        moves.append(shift)
    return moves


def reduce_kmers(kmers, moves):
    """Stitch k-mers together given their moves (`bio.py:206-225`)."""
    kmers = list(kmers)
    pieces = [kmers[0]]
    for kmer, move in zip(kmers[1:], moves):
        if move == 0:
            continue
        pieces.append(kmer if move >= len(kmer) else kmer[-move:])
    return type(kmers[0])().join(pieces)


def kmers_to_sequence(kmers, always_move=False):
    """Sequence from overlapping k-mers (`bio.py:228-237`)."""
    return reduce_kmers(kmers, max_overlap(kmers, not always_move))


def states_to_sequence(states, kmer_len, alphabet='ACGT', always_move=True):
    """Same result as `kmers_to_sequence([kmers[i] for i in states], always_move)` computed on the
    integer states: shift m matches iff  prev mod nbase^(k-m) == next div nbase^m."""
    nbase = len(alphabet)
    states = [int(s) for s in states]
    if not states:
        return ''
    pows = [nbase ** i for i in range(kmer_len + 1)]

    def spell(value, nletters):
        return ''.join(alphabet[(value // pows[nletters - 1 - i]) % nbase] for i in range(nletters))

    out = [spell(states[0], kmer_len)]
    for prev, nxt in zip(states, states[1:]):
        if not always_move and prev == nxt:
            continue
        move = kmer_len
        for m in range(1, kmer_len):
            if prev % pows[kmer_len - m] == nxt // pows[m]:
                move = m
                break
        out.append(spell(nxt % pows[move], move))
    return ''.join(out)
