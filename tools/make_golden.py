#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the REFERENCE itself.

Run in the build container, where /root/reference is mounted:

    python tools/make_golden.py

What it pins (and prints):
  * decode: `sloika/decode.py` (prepare_post, viterbi) is imported unmodified from /root/reference
    and run on (a) the known-answer matrix `post3` of `test/unit/test_decode.py:22-191` with the
    expected results of :233-241, (b) the modified-base case of :244-256, (c) seeded random
    posteriors (stay-heavy and flat, klen 3/4/5, nbase 4/5, several skip penalties).  The oracle
    restatement (`oracle/decode_ref.py`, `oracle/viterbi_ref.c`) must agree bit for bit.
  * host pre/post-processing: `sloika/maths.py` (med_mad) and `sloika/bio.py` (kmers_to_sequence)
    imported unmodified; outputs stored.
  * bundled reads: raw DAQ signals of data/reads/read{1..8}.fast5 (read with sloika_b200.fast5;
    lengths pinned by test/unit/test_fast5.py:99-110), the weights of models/pretrained.pkl, and
    the basecalls obtained by  oracle forward (float32)  ->  REFERENCE decode.viterbi  ->  REFERENCE
    bio.kmers_to_sequence, i.e. `bin/basecall_network.py raw` with CLI defaults
    (--trim 200 10 --open_pore_fraction 0 --min_prob 1e-5 --skip 0 --kmer_len 5).
    Theano is not installable, so the forward pass is the oracle restatement ("parity unpinned" for
    Gru/Convolution); as an external anchor the basecalls are compared with the ONT basecalls
    embedded in the fast5 files and the identities are stored.

Nothing here is read at test time from /root/reference: the fixtures are committed.
"""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, 'test', 'unit'))
warnings.simplefilter('ignore', SyntaxWarning)
GOLD = os.path.join(ROOT, 'tests', 'golden')


def softmax_rows(logits):
    e = np.exp(logits - logits.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


def edit_distance(a, b):
    a = np.frombuffer(a.encode(), dtype=np.uint8)
    b = np.frombuffer(b.encode(), dtype=np.uint8)
    prev = np.arange(len(b) + 1)
    idx = np.arange(len(b) + 1)
    for i in range(1, len(a) + 1):
        cur = np.minimum(prev[:-1] + (b != a[i - 1]), prev[1:] + 1)
        cur = np.concatenate([[i], cur])
        prev = np.minimum.accumulate(cur - idx) + idx
    return int(prev[-1])


def decode_fixtures():
    from sloika import decode as ref_decode
    from oracle import decode_ref, cbind
    import test_decode as ref_test

    cases = {}
    meta = []

    def add(name, post, klen, nbase, skip_pen, log=False, store_post=True):
        score, path = ref_decode.viterbi(post, klen, skip_pen=skip_pen, log=log, nbase=nbase)
        o_score, o_path = decode_ref.viterbi(post, klen, skip_pen=skip_pen, log=log, nbase=nbase)
        assert o_path == [int(p) for p in path] and o_score == score, name
        if post.dtype == np.float32:
            lp = post if log else decode_ref.log_post(post)
            c_score, c_paths = cbind.viterbi_batch(lp[:, None, :], None, klen=klen, nbase=nbase, skip_pen=skip_pen)
            assert c_paths[0] == o_path and c_score[0] == score, name + ' (C oracle)'
        if store_post:
            cases[name + '/post'] = post
        cases[name + '/path'] = np.asarray(path, dtype=np.int32)
        cases[name + '/score'] = np.asarray(score)
        meta.append(dict(name=name, klen=klen, nbase=nbase, skip_pen=skip_pen, log=log))

    ref_test.TestDecode.setUpClass()
    post3 = np.asarray(ref_test.TestDecode.post3)
    add('kat_post3', post3, 3, 4, 0.0)
    add('kat_post3_skip3', post3, 3, 4, 3.0)
    score, path = ref_decode.viterbi(post3, 3)
    assert abs(score - (-11.130084569094556)) < 1e-7 and path == [49, 7, 63, 63]
    score, path = ref_decode.viterbi(post3, 3, skip_pen=3.0)
    assert abs(score - (-11.936803444063674)) < 1e-7 and path == [49, 7, 31, 63, 63]

    ref_test.TestDecodeModifiedBases.setUpClass()
    modpost = np.asarray(ref_test.TestDecodeModifiedBases.post)
    add('kat_modbase', modpost, 3, 5, 5.0)
    assert [int(p) for p in cases['kat_modbase/path']] == [x - 1 for x in ref_test.TestDecodeModifiedBases.seq if x]

    rng = np.random.default_rng(20261017)
    for name, T, klen, nbase, skip_pen, stay_boost, scale in [
            ('rand_k5_flat', 40, 5, 4, 0.0, 0.0, 3.0),
            ('rand_k5_stay', 120, 5, 4, 0.0, 7.0, 3.0),
            ('rand_k5_stay_skip5', 100, 5, 4, 5.0, 6.0, 3.0),
            ('rand_k5_peaky', 80, 5, 4, 0.0, 12.0, 8.0),
            ('rand_k3_nb4', 80, 3, 4, 1.5, 4.0, 3.0),
            ('rand_k4_nb4', 80, 4, 4, 0.0, 5.0, 3.0),
            ('rand_k3_nb5', 70, 3, 5, 2.0, 4.0, 3.0),
            ('rand_k5_T1', 1, 5, 4, 0.0, 0.0, 3.0),
            ('rand_k5_T2', 2, 5, 4, 0.0, 3.0, 3.0)]:
        S = nbase ** klen + 1
        logits = scale * rng.standard_normal((T, 1, S))
        logits[:, :, 0] += stay_boost
        post = softmax_rows(logits).astype(np.float32)
        prepared = ref_decode.prepare_post(post, min_prob=1e-5)
        o_prepared = decode_ref.prepare_post(post, min_prob=1e-5)
        assert prepared.dtype == np.float32 and np.array_equal(prepared, o_prepared)
        cases[name + '/raw'] = post       # tests rebuild the prepared matrix with the pinned prepare_post
        add(name, prepared, klen, nbase, skip_pen, store_post=False)
    # floor ties: most states sit exactly at the min_prob floor -> exercises the tie rules
    T, S = 120, 1025
    post = np.zeros((T, 1, S), dtype=np.float32)
    hot = rng.integers(0, S, size=T)
    post[np.arange(T), 0, hot] = 1.0
    prepared = ref_decode.prepare_post(post, min_prob=1e-5)
    cases['ties_k5/raw'] = post
    add('ties_k5', prepared, 5, 4, 0.0, store_post=False)
    # log=True path
    logits = 3.0 * rng.standard_normal((60, 1025))
    logits[:, 0] += 6.0
    add('rand_k5_log', np.log(softmax_rows(logits).astype(np.float32) + np.float32(1e-10)), 5, 4, 0.0, log=True)

    np.savez_compressed(os.path.join(GOLD, 'decode_cases.npz'), **cases)
    with open(os.path.join(GOLD, 'decode_cases.json'), 'w') as fh:
        json.dump(meta, fh, indent=1)
    print("decode: {} cases, oracle (NumPy + C) == reference decode.viterbi bit for bit".format(len(meta)))


def host_fixtures():
    from sloika import bio as ref_bio, maths as ref_maths
    from oracle import host_ref
    rng = np.random.default_rng(7)
    out = {}
    x = rng.standard_normal(4001) * 11 + 90
    med, mad = ref_maths.med_mad(x)
    assert (med, mad) == host_ref.med_mad(x)
    out['medmad_x'] = x
    out['medmad'] = np.array([med, mad])
    xm = x[:4000].reshape(40, 100)
    out['mad_axis1'] = ref_maths.mad(xm, axis=1)
    assert np.array_equal(out['mad_axis1'], host_ref.med_mad(xm, axis=1)[1])
    kmers = ref_bio.all_kmers(5)
    paths, seqs = [], []
    for n in range(40):
        path = [int(rng.integers(1024))]
        for _ in range(int(rng.integers(0, 80))):
            c, s = rng.random(), path[-1]
            if c < 0.25:
                path.append(s)
            elif c < 0.7:
                path.append((s * 4 + int(rng.integers(4))) % 1024)
            elif c < 0.9:
                path.append((s * 16 + int(rng.integers(16))) % 1024)
            else:
                path.append(int(rng.integers(1024)))
        for always_move in (True, False):
            seq = ref_bio.kmers_to_sequence([kmers[i] for i in path], always_move=always_move)
            assert seq == host_ref.kmers_to_sequence([kmers[i] for i in path], always_move=always_move)
            paths.append(path)
            seqs.append([always_move, seq])
    with open(os.path.join(GOLD, 'bio_cases.json'), 'w') as fh:
        json.dump({'paths': paths, 'seqs': seqs}, fh)
    np.savez_compressed(os.path.join(GOLD, 'maths_cases.npz'), **out)
    print("host: med_mad + {} kmers_to_sequence cases pinned to reference maths.py / bio.py".format(len(seqs)))


def read_fixtures():
    from sloika import decode as ref_decode, bio as ref_bio
    from sloika_b200 import model_io
    from sloika_b200.fast5 import Fast5
    from oracle import forward_ref, host_ref

    model = model_io.load_model(os.path.join(REF, 'models', 'pretrained.pkl'))
    desc = model.json(params=True)
    weights = model_io.weights_of(model)
    np.savez_compressed(os.path.join(GOLD, 'pretrained_weights.npz'), **weights)
    with open(os.path.join(GOLD, 'pretrained_arch.json'), 'w') as fh:
        json.dump(model.json(params=False), fh, indent=1)

    pinned = {'read1': 114400, 'read2': 69443, 'read3': 51129, 'read6': 55885}   # test_fast5.py:99-110
    kmers = ref_bio.all_kmers(5)
    signals, records, slices = {}, [], {}
    for i in range(1, 9):
        name = 'read{}'.format(i)
        f5 = Fast5(os.path.join(REF, 'data', 'reads', name + '.fast5'))
        daq = f5.get_read(raw=True, scale=False)
        if name in pinned:
            assert len(daq) == pinned[name], (name, len(daq))
        meta = f5.channel_meta
        signals[name] = daq
        signals[name + '_scaling'] = np.array([float(meta['offset']), float(meta['range']),
                                               float(meta['digitisation'])])
        x = host_ref.prepare_signal(f5.get_read(raw=True))
        post = forward_ref.run(desc, x)
        prepared = ref_decode.prepare_post(post, min_prob=1e-5, drop_bad=False)
        score, path = ref_decode.viterbi(prepared, 5, skip_pen=0.0, nbase=4)
        seq = ref_bio.kmers_to_sequence([kmers[s] for s in path], always_move=True)
        header = ">{} score {:.0f}, {} {} to {} bases".format(f5.filename_short, score, x.shape[0], 'samples', len(seq))
        rec = dict(name=name, nsamples=int(x.shape[0]), nsteps=int(post.shape[0]), score=float(score),
                   header=header, seq=seq, path=[int(p) for p in path],
                   mean_max_post=float(post.max(2).mean()))
        fq_path = '/Analyses/Basecall_1D_000/BaseCalled_template/Fastq'
        if fq_path in f5._h5:
            fq = f5._h5.dataset(fq_path)[()].decode().split('\n')[1]
            ed = edit_distance(seq, fq)
            rec['embedded_basecall_identity'] = 1.0 - ed / max(len(seq), len(fq))
        records.append(rec)
        slices[name + '_head'] = post[:8, 0]
        slices[name + '_tail'] = post[-8:, 0]
        slices[name + '_rowmax'] = post[:, 0].max(1)
        print("  {}: {} samples -> {} steps, score {:.1f}, {} bases, identity vs embedded call {}".format(
            name, rec['nsamples'], rec['nsteps'], score, len(seq),
            '{:.3f}'.format(rec['embedded_basecall_identity']) if 'embedded_basecall_identity' in rec else 'n/a'))
    np.savez_compressed(os.path.join(GOLD, 'reads_daq.npz'), **signals)
    np.savez_compressed(os.path.join(GOLD, 'reads_post_slices.npz'), **slices)
    with open(os.path.join(GOLD, 'reads_basecalls.json'), 'w') as fh:
        json.dump(records, fh)
    print("reads: 8 bundled reads basecalled (oracle forward fp32 + reference decode/bio)")


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    decode_fixtures()
    host_fixtures()
    read_fixtures()
    print("fixture sizes:")
    for fn in sorted(os.listdir(GOLD)):
        print("  {:32s} {:>9d} B".format(fn, os.path.getsize(os.path.join(GOLD, fn))))
