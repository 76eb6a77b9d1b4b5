"""Small invocations of every kernel family, meant to be run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sloika_b200 import basecall, decode, engine, layers, olddecode, transducer, validate, zoo
from sloika_b200 import module_tools as smt

DEV = torch.device('cuda:0')


def main():
    np.random.seed(1)
    torch.manual_seed(1)
    # conv + GRU (narrow, reverse and forward) + f16/tf32 GEMMs + softmax + fused Viterbi, ragged batch
    net = zoo.raw_rgrgr().compile()
    x = torch.randn((640, 9, 1), device=DEV)
    lens = torch.tensor([640, 1, 7, 640, 333, 12, 300, 500, 5], dtype=torch.int32, device=DEV)
    for lengths in (None, lens):
        out = net.forward_device(x, lengths, fused_decode=True)
        score, paths, plen = decode.viterbi_batch(out, None, min_prob=1e-5, return_device=True)
        decode.paths_to_sequences(paths, plen, 5, 'ACGT', True)
        post = net.forward_device(x, lengths)
        decode.viterbi_batch(post, None, min_prob=1e-5)
    # wide GRU kernels (H = 110 / 142) and the elu -> tf32 GEMM path
    for build in (zoo.raw_rGr, zoo.pretrained_like):
        net = build().compile()
        out = net.forward_device(torch.randn((400, 5, 1), device=DEV), None, fused_decode=True)
        decode.viterbi_batch(out, None, min_prob=1e-5)
    # every shape of the tensor-memory recurrence (groups per CTA, compute warps per group, sequences per group)
    g = layers.Gru(40, 96, init=smt.partial(smt.truncated_normal, sd=0.5), has_bias=True)
    xg = torch.tanh(torch.randn((50, 37, 40), device=DEV))
    for shape in ('1,4', '1,8', '1,16', '2,4', '2,8', '4,4', '1,8,16', '2,8,16'):
        os.environ['SLOIKA_B200_GRU_TC'] = shape
        engine.run_gru(g, engine.Act(xg, None))
        engine.run_gru(g, engine.Act(xg, torch.randint(1, 51, (37,), dtype=torch.int32, device=DEV), reverse=True))
    os.environ.pop('SLOIKA_B200_GRU_TC')
    # the 3-CTA cluster launch with the projection inside (csrc/gru_fused.cu): partial last cluster, ragged, reversed
    os.environ['SLOIKA_B200_FUSED_GRU'] = '1'
    for I, H, B in ((40, 96, 37), (96, 96, 70), (32, 50, 5)):
        gf = layers.Gru(I, H, init=smt.partial(smt.truncated_normal, sd=0.5), has_bias=True)
        xf = torch.tanh(torch.randn((50, B, I), device=DEV))
        engine.run_gru(gf, engine.Act(xf, None, bounded=True))
        engine.run_gru(gf, engine.Act(xf, torch.randint(1, 51, (B,), dtype=torch.int32, device=DEV), reverse=True, bounded=True))
    # the sequences-on-lanes launch (csrc/gru_seq.cu) with its layout conversions: a conv -> GRU x 3 stack (gated first layer,
    # blocked activations between the layers), partial last block of 128, ragged, reversed
    os.environ['SLOIKA_B200_GRU_SEQ'] = '1'
    init = smt.partial(smt.truncated_normal, sd=0.5)
    seq_net = layers.Serial([layers.Convolution(1, 96, 11, 5, init=init, has_bias=True, fun=smt.elu),
                             layers.Reverse(layers.Gru(96, 96, init=init, has_bias=True)),
                             layers.Gru(96, 50, init=init, has_bias=True),
                             layers.Reverse(layers.Gru(50, 80, init=init, has_bias=True))])
    xs = torch.randn((300, 150, 1), device=DEV)
    seq_net.run(engine.Act(xs, torch.randint(20, 301, (150,), dtype=torch.int32, device=DEV))).data
    seq_net.run(engine.Act(xs * 3.0e4, None)).data                      # outside the fp16 range: the other verdict of the gate
    os.environ.pop('SLOIKA_B200_GRU_SEQ')
    os.environ.pop('SLOIKA_B200_FUSED_GRU')
    # 144 < H <= 256: the 4-CTA cluster kernel (distributed shared memory), ragged and reversed
    for H in (160, 256):
        gw = layers.Gru(24, H, init=smt.partial(smt.truncated_normal, sd=0.5), has_bias=True)
        xw = torch.tanh(torch.randn((20, 11, 24), device=DEV))
        engine.run_gru(gw, engine.Act(xw, torch.randint(1, 21, (11,), dtype=torch.int32, device=DEV), reverse=(H == 256)))
    # events route: Window + birnn(Lstm), old decoder, transition estimates, scoring
    init = smt.partial(smt.truncated_normal, sd=0.5)
    lstm = lambda i, o: layers.Lstm(i, o, init=init, has_bias=True, has_peep=True)
    ev_net = layers.Serial([layers.Window(4, 3), layers.birnn(lstm(12, 64), lstm(12, 64)),
                            layers.FeedForward(128, 64, init=init, has_bias=True),
                            layers.Softmax(64, 1024, init=init, has_bias=True)])
    post = ev_net.compile().forward_device(torch.randn((70, 1, 4), device=DEV)).data
    est = olddecode.estimate_transitions(post[:, 0], return_device=True)
    olddecode.decode_profile(post[:, 0], trans=(1e-10 + est).log())
    validate.wrap_network(ev_net.compile())(np.random.standard_normal((30, 3, 4)).astype(np.float32),
                                             np.random.randint(0, 1024, (30, 3)).astype(np.int32))
    # remap decode, both launch geometries
    rng = np.random.default_rng(2)
    for T, B, S, P in ((60, 5, 65, 40), (40, 3, 65, 1100)):
        lt = torch.log_softmax(torch.randn((T, B, S), device=DEV), dim=-1)
        seqs = [rng.integers(1, S, size=int(n)) for n in rng.integers(3, P + 1, size=B)]
        transducer.map_to_sequence_batch(lt, seqs, slip=5.0)
    # pre-processing: ragged reads, ties, a read that trims to nothing, a percentile threshold
    sigs = [rng.standard_normal(n) * 5 + 90 for n in (1000, 2550, 399, 250)] + [np.round(rng.standard_normal(1500) * 3) + 100]
    for frac in (0, 0.3):
        basecall.prepare_signals_device(sigs, (200, 10), frac, device=DEV)
    torch.cuda.synchronize()
    print("sanitize smoke done")


if __name__ == '__main__':
    main()
