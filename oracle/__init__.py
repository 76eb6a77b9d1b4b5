"""CPU oracle for the Sloika raw basecall hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under `sloika_b200/` imports this package.  It may be imported only by `tests/`,
`__graft_entry__.smoke()`, `tools/` (golden generation, sweeps that print a CPU column) and the CPU-baseline /
`--impl reference` legs of `bench.py`, and there only as the checker or the reported CPU baseline -- never as the product.

Parity status (see DESIGN.md "Oracle"):
  * decode (`decode_ref`), k-mer assembly and signal normalisation (`host_ref`): PINNED -- checked
    against the reference's own `sloika/decode.py`, `sloika/bio.py`, `sloika/maths.py` imported from
    /root/reference in the build container (`tools/make_golden.py`) and against the known-answer
    tests of `test/unit/test_decode.py:233-256`, `test_maths.py`, `test_bio.py`.
  * forward pass (`forward_ref`): FeedForward / Softmax / Serial / Parallel / Reverse semantics are
    pinned by the NumPy formulas in `test/unit/test_layers.py:58-125`; **Gru and Convolution are
    "parity unpinned"**: Theano 0.8.2 is not installable here and the reference holds no numeric
    known-answer for them.  They are anchored indirectly: the oracle's basecalls of the bundled
    reads with `models/pretrained.pkl` agree with the basecalls embedded in those fast5 files
    (identity reported by `tools/make_golden.py`).
  * remap decode (`remap_ref.py`, `remap_ref.c`: transducer.map_to_sequence + viterbi_helpers.slip_update): PINNED --
    bit for bit against outputs of the reference's unmodified `sloika/transducer.py` running on its own
    `viterbi_helpers.pyx` (compiled with Cython outside the repo by `tools/make_golden_remap.py`), including the
    recipe of `test/unit/test_viterbi.py` and the NaN behaviour of `slip=None`.
"""
