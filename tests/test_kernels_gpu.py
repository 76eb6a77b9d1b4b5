"""GPU: every kernel behind the C ABI against the CPU oracle on the same seeded inputs.

Tolerances: forward kernels are float32 with a different summation order than BLAS -> 2e-5 on
pre-softmax activations in [-1, 1] (elu outputs: relative), 1e-4 max-abs on posteriors (the
north-star bound); Viterbi is float32 add/max only -> scores and paths are compared EXACTLY when both
sides get identical log-posteriors.
"""
import os

import numpy as np
import pytest
import torch

from oracle import cbind, decode_ref, forward_ref
from sloika_b200 import activation as act
from sloika_b200 import cabi, decode, engine, layers, zoo
from sloika_b200 import module_tools as smt

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _init():
    return smt.partial(smt.truncated_normal, sd=0.5)


def _run(layer, x, lengths=None, reverse=False):
    a = engine.Act(torch.from_numpy(x).to(DEV), None if lengths is None else torch.as_tensor(lengths, dtype=torch.int32, device=DEV), reverse)
    out = layer.run(a)
    torch.cuda.synchronize()
    return out.data.cpu().numpy(), out


def _oracle(layer, x):
    return forward_ref.run(layer.json(params=True), x)


# ------------------------------------------------------------------ convolution
@pytest.mark.parametrize('C,stride,fun,T,B', [(96, 5, act.elu, 403, 37), (128, 2, act.tanh, 250, 5),
                                               (32, 2, act.tanh, 64, 33), (128, 5, act.elu, 1, 1),
                                               (96, 5, act.elu, 4000, 16)])
def test_conv_raw(C, stride, fun, T, B):
    np.random.seed(C + T)
    layer = layers.Convolution(1, C, 11, stride, init=_init(), has_bias=True, fun=fun)
    x = (3 * np.random.standard_normal((T, B, 1))).astype(np.float32)
    got, _ = _run(layer, x)
    ref = _oracle(layer, x)
    assert got.shape == ref.shape == (-(-T // stride), B, C)
    np.testing.assert_allclose(got, ref, atol=2e-5, rtol=2e-5)


def test_conv_generic_signature():
    """The reference's own layer test shape: Convolution(12, 32, 11, 5) on [100, 20, 12] (test_layers.py:454-459)."""
    np.random.seed(3)
    for mode in ('same', 'valid', 'full', (2, 7)):
        layer = layers.Convolution(12, 32, 11, 5, init=_init(), has_bias=True, padding_mode=mode)
        x = np.random.standard_normal((100, 20, 12)).astype(np.float32)
        got, _ = _run(layer, x)
        ref = _oracle(layer, x)
        assert got.shape == ref.shape
        np.testing.assert_allclose(got, ref, atol=2e-5, rtol=2e-5)
    # even window, no bias, Cout not a multiple of 4
    layer = layers.Convolution(1, 30, 4, 3, init=_init(), has_bias=False, fun=act.linear)
    x = np.random.standard_normal((50, 3, 1)).astype(np.float32)
    np.testing.assert_allclose(_run(layer, x)[0], _oracle(layer, x), atol=2e-5, rtol=2e-5)


def test_conv_ragged_equals_per_read():
    np.random.seed(4)
    layer = layers.Convolution(1, 96, 11, 5, init=_init(), has_bias=True, fun=act.elu)
    lengths = [503, 1, 250, 499, 7]
    x = np.zeros((503, 5, 1), dtype=np.float32)
    for b, n in enumerate(lengths):
        x[:n, b, 0] = np.random.standard_normal(n)
    x_dirty = x.copy()
    for b, n in enumerate(lengths):
        x_dirty[n:, b, 0] = 99.0                     # garbage past the end must not be read
    got, out = _run(layer, x_dirty, lengths)
    assert out.lengths.cpu().tolist() == [-(-n // 5) for n in lengths]
    for b, n in enumerate(lengths):
        ref = _oracle(layer, x[:n, b:b + 1])
        np.testing.assert_allclose(got[:ref.shape[0], b], ref[:, 0], atol=2e-5, rtol=2e-5)


# ------------------------------------------------------------------ feed-forward / softmax
@pytest.mark.parametrize('I,O,fun', [(192, 128, act.tanh), (12, 8, act.linear), (110, 37, act.sigmoid), (5, 130, act.elu)])
def test_feedforward(I, O, fun):
    np.random.seed(I)
    layer = layers.FeedForward(I, O, init=_init(), has_bias=True, fun=fun)
    x = np.random.standard_normal((33, 7, I)).astype(np.float32)
    np.testing.assert_allclose(_run(layer, x)[0], _oracle(layer, x), atol=2e-5, rtol=2e-5)


@pytest.mark.parametrize('I', [96, 110, 112, 128])
def test_softmax(I):
    np.random.seed(I)
    layer = layers.Softmax(I, 1025, init=_init(), has_bias=True)
    layer.W.set_value(layer.W.get_value() * 6)       # trained-model-like magnitudes -> peaky rows
    x = np.tanh(np.random.standard_normal((40, 9, I))).astype(np.float32)
    got = _run(layer, x)[0]
    ref = _oracle(layer, x)
    assert np.abs(got - ref).max() < 2e-5
    np.testing.assert_allclose(got.sum(2), 1.0, rtol=1e-5)


# ------------------------------------------------------------------ GRU
@pytest.mark.parametrize('I,H,T,B,reverse', [(96, 96, 120, 19, False), (96, 96, 120, 8, True),
                                             (128, 112, 80, 5, True), (112, 144, 80, 9, False),
                                             (128, 110, 60, 3, True), (110, 142, 60, 4, False),
                                             (32, 96, 50, 1, False), (12, 64, 10, 20, False),
                                             (7, 5, 30, 2, True), (40, 128, 40, 11, False),
                                             (3, 33, 25, 17, True), (16, 80, 25, 8, False), (16, 16, 25, 8, True)])
def test_gru(I, H, T, B, reverse):
    np.random.seed(I * 1000 + H)
    g = layers.Gru(I, H, init=_init(), has_bias=True)
    # trained-model-like magnitudes (|w| up to ~4): gates saturate, errors would compound if wrong
    g.sW.set_value(g.sW.get_value() * 6)
    g.sW2.set_value(g.sW2.get_value() * 6)
    layer = layers.Reverse(g) if reverse else g
    x = np.tanh(np.random.standard_normal((T, B, I))).astype(np.float32)
    got = _run(layer, x)[0]
    ref = _oracle(layer, x)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 5e-5


@pytest.mark.parametrize('env', [{'SLOIKA_B200_GRU': 'v1'}, {'SLOIKA_B200_GRU': 'v3'}, {'SLOIKA_B200_GRU': 'v4'},
                                 {'SLOIKA_B200_GRU_TC': '1,4'}, {'SLOIKA_B200_GRU_TC': '1,8'},
                                 {'SLOIKA_B200_GRU_TC': '1,16'}, {'SLOIKA_B200_GRU_TC': '2,4'},
                                 {'SLOIKA_B200_GRU_TC': '2,8'}, {'SLOIKA_B200_GRU_TC': '4,4'},
                                 {'SLOIKA_B200_GRU_TC': '1,8,16'}, {'SLOIKA_B200_GRU_TC': '2,8,16'}])
@pytest.mark.parametrize('I,H,T,B,reverse', [(96, 96, 90, 37, False), (40, 110, 50, 21, True), (20, 128, 40, 9, False),
                                             (24, 48, 30, 50, True)])
def test_gru_every_kernel_generation(env, I, H, T, B, reverse, monkeypatch):
    """Each selectable recurrence kernel -- FFMA2 (v1), 3xTF32 mma.sync (v3), fp16x3 mma.sync (v4) and every
    (groups per CTA, compute warps per group) shape of the tcgen05 / tensor-memory kernel -- against the oracle, with a
    ragged batch whose size is not a multiple of any CTA tile."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    np.random.seed(H + B)
    g = layers.Gru(I, H, init=_init(), has_bias=True)
    g.sW.set_value(g.sW.get_value() * 5)
    g.sW2.set_value(g.sW2.get_value() * 5)
    layer = layers.Reverse(g) if reverse else g
    lengths = [int(v) for v in np.random.randint(1, T + 1, size=B)]
    lengths[0] = T
    x = np.tanh(np.random.standard_normal((T, B, I))).astype(np.float32)
    got, _ = _run(layer, x, lengths)
    for b in list(range(0, B, 5)) + [B - 1]:
        n = lengths[b]
        ref = _oracle(layer, x[:n, b:b + 1])
        assert np.abs(got[:n, b] - ref[:, 0]).max() < 5e-5, (env, b)
        assert np.all(got[n:, b] == 0)


def test_gru_other_activation_pair_takes_the_generic_kernel():
    """A GRU whose activations are not tanh / sigmoid (any `fun` / `gatefun` the constructor accepts, layers.py:965)
    runs on the FFMA2 kernel of gru.cu."""
    np.random.seed(4)
    g = layers.Gru(12, 40, init=_init(), has_bias=True, fun=act.sigmoid, gatefun=act.tanh)
    x = np.tanh(np.random.standard_normal((30, 6, 12))).astype(np.float32)
    got = _run(g, x)[0]
    ref = _oracle(g, x)
    assert np.abs(got - ref).max() < 5e-5


def test_gru_ragged_equals_per_read():
    np.random.seed(11)
    for reverse in (False, True):
        g = layers.Gru(24, 96, init=_init(), has_bias=True)
        layer = layers.Reverse(g) if reverse else g
        lengths = [70, 1, 33, 69, 70, 2, 50, 64, 8, 41]
        x = np.tanh(np.random.standard_normal((70, len(lengths), 24))).astype(np.float32)
        got, _ = _run(layer, x, lengths)
        for b, n in enumerate(lengths):
            ref = _oracle(layer, x[:n, b:b + 1])
            assert np.abs(got[:n, b] - ref[:, 0]).max() < 5e-5, (reverse, b)
            assert np.all(got[n:, b] == 0)


def test_birnn_parallel_writes_in_place():
    np.random.seed(12)
    init = _init()
    net = layers.Serial([layers.birnn(layers.Gru(20, 48, init=init, has_bias=True), layers.Gru(20, 48, init=init, has_bias=True)),
                         layers.FeedForward(96, 40, init=init, has_bias=True),
                         layers.Parallel([layers.FeedForward(40, 8, init=init, has_bias=True),
                                          layers.Reverse(layers.Gru(40, 16, init=init, has_bias=True))])])
    x = np.tanh(np.random.standard_normal((45, 6, 20))).astype(np.float32)
    got = _run(net, x)[0]
    ref = _oracle(net, x)
    assert got.shape == ref.shape == (45, 6, 24)
    assert np.abs(got - ref).max() < 5e-5


@pytest.mark.parametrize('seq', ['0', '1'])
def test_birnn_through_throughput_forms(seq, monkeypatch):
    """Parallel branches with the GRU layers forced onto the fused (seq = 0) / sequences-on-lanes (seq = 1) launch: each
    branch writes its column slice of the shared buffer (the blocked result is converted into the slice), bounded input."""
    monkeypatch.setenv('SLOIKA_B200_FUSED_GRU', '1')
    monkeypatch.setenv('SLOIKA_B200_GRU_SEQ', seq)
    np.random.seed(13)
    init = _init()
    net = layers.Serial([layers.birnn(layers.Gru(32, 96, init=init, has_bias=True), layers.Gru(32, 96, init=init, has_bias=True)),
                         layers.FeedForward(192, 64, init=init, has_bias=True),
                         layers.birnn(layers.Gru(64, 48, init=init, has_bias=True), layers.Gru(64, 80, init=init, has_bias=True))])
    x = np.tanh(np.random.standard_normal((60, 140, 32))).astype(np.float32)
    lengths = np.random.randint(1, 61, size=140).astype(np.int32)
    lengths[0] = 60
    a = engine.Act(torch.from_numpy(x).to(DEV), torch.from_numpy(lengths).to(DEV), bounded=True)
    engine.TIMER.reset()
    engine.TIMER.enabled = True
    try:
        got = net.run(a).data.cpu().numpy()
        ran = set(engine.TIMER.totals_ms())
    finally:
        engine.TIMER.enabled = False
    assert ('gru_seq' if seq == '1' else 'gru_fused') in ran
    for b in (0, 1, 70, 139):
        ref = _oracle(net, x[:lengths[b], b:b + 1])
        assert np.abs(got[:lengths[b], b] - ref[:, 0]).max() < 5e-5, b
        assert np.all(got[lengths[b]:, b] == 0)


# ------------------------------------------------------------------ whole networks
@pytest.mark.parametrize('name,T,B', [('raw_rgrgr', 1000, 6), ('raw_rGr', 400, 5), ('bigger_raw_gru', 300, 4),
                                      ('pretrained_like', 900, 3)])
def test_network_posteriors(name, T, B):
    np.random.seed(len(name))
    net = getattr(zoo, name)()
    x = np.random.standard_normal((T, B, 1)).astype(np.float32)
    post = net.compile()(x)
    ref = _oracle(net, x)
    assert post.shape == ref.shape and post.dtype == np.float32
    assert np.abs(post - ref).max() < 1e-4            # north-star bound
    with pytest.raises(TypeError):
        net.compile()(x.astype(np.float64))           # Theano function rejects the wrong dtype too


# ------------------------------------------------------------------ Viterbi
def _kernel_vs_golden(case, data, **kw):
    name = case['name']
    if name + '/post' in data:
        post = data[name + '/post']
    else:
        post = decode_ref.prepare_post(data[name + '/raw'], min_prob=1e-5)
    return post


def test_viterbi_reference_known_answers(decode_cases):
    """test/unit/test_decode.py:233-256 through the device kernel (float64 inputs are decoded in
    float32: paths identical, score to 1e-5)."""
    meta, data = decode_cases
    score, path = decode.viterbi(data['kat_post3/post'], 3)
    assert path == [49, 7, 63, 63] and abs(score - (-11.130084569094556)) < 1e-5
    score, path = decode.viterbi(data['kat_post3/post'], 3, skip_pen=3.0)
    assert path == [49, 7, 31, 63, 63] and abs(score - (-11.936803444063674)) < 1e-5
    score, path = decode.viterbi(data['kat_modbase/post'], 3, skip_pen=5.0, nbase=5)
    assert path == data['kat_modbase/path'].tolist()


def test_viterbi_golden_vectors_bit_exact(decode_cases):
    """Identical float32 log-posteriors in -> identical score bits and path out."""
    meta, data = decode_cases
    for case in meta:
        post = _kernel_vs_golden(case, data)
        if post.dtype != np.float32:
            continue
        lp = post if case['log'] else decode_ref.log_post(post)
        score, path = decode.viterbi(lp, case['klen'], skip_pen=case['skip_pen'], log=True, nbase=case['nbase'])
        assert path == data[case['name'] + '/path'].tolist(), case['name']
        assert score == data[case['name'] + '/score'], case['name']


def test_viterbi_fused_prepare_and_log(decode_cases):
    """Raw posteriors in, min_prob floor + log done in the kernel (device logf vs NumPy log may differ
    by an ulp, so the score is compared to 1e-3 and the path must still match on these cases)."""
    meta, data = decode_cases
    for case in meta:
        name = case['name']
        if name + '/raw' not in data:
            continue
        score, paths = decode.viterbi_batch(data[name + '/raw'], None, klen=case['klen'], skip_pen=case['skip_pen'],
                                            min_prob=1e-5, nbase=case['nbase'])
        assert paths[0] == data[name + '/path'].tolist(), name
        assert abs(score[0] - float(data[name + '/score'])) < 1e-3 * max(1.0, abs(float(data[name + '/score']))), name


def test_viterbi_ragged_batch_bit_exact():
    rng = np.random.default_rng(9)
    T, B, S = 257, 21, 1025
    logits = 3 * rng.standard_normal((T, B, S))
    logits[:, :, 0] += 6
    post = np.exp(logits - logits.max(2, keepdims=True))
    post = (post / post.sum(2, keepdims=True)).astype(np.float32)
    lp = decode_ref.log_post(decode_ref.prepare_post(post.reshape(T * B, 1, S)).reshape(T, B, S)).astype(np.float32)
    lengths = rng.integers(1, T + 1, size=B).astype(np.int32)
    lengths[:4] = [T, 1, 2, 0]
    for skip in (0.0, 2.5):
        score, paths = decode.viterbi_batch(lp, lengths, skip_pen=skip, log=True)
        ok = lengths > 0
        ref_score, ref_paths = cbind.viterbi_batch(lp[:, ok], lengths[ok], skip_pen=skip)
        got_paths = [p for p, k in zip(paths, ok) if k]
        assert got_paths == ref_paths
        assert np.array_equal(score[ok], ref_score)
        assert paths[3] == [] and score[3] == 0.0


def test_viterbi_strided_posteriors_view():
    """The decoder reads the network's output in place, including a column-padded buffer."""
    rng = np.random.default_rng(10)
    T, B, S = 60, 3, 1025
    buf = torch.rand((T, B, 1032), device=DEV)
    view = buf[:, :, :S]
    lp = torch.log_softmax(view * 4, dim=2)
    padded = torch.full((T, B, 1032), -1.0, device=DEV)
    padded[:, :, :S] = lp
    s1, p1 = decode.viterbi_batch(padded[:, :, :S], None, log=True)
    s2, p2 = decode.viterbi_batch(lp.contiguous(), None, log=True)
    assert p1 == p2 and np.array_equal(s1, s2)
    ref_s, ref_p = cbind.viterbi_batch(lp.cpu().numpy(), None)
    assert p1 == ref_p and np.array_equal(s1, ref_s)


def test_error_codes_surface_as_exceptions():
    with pytest.raises(AssertionError):
        decode.viterbi(np.zeros((4, 17), dtype=np.float32), 2)             # klen >= 3 (decode.py:50)
    with pytest.raises(AssertionError):
        decode.viterbi(np.zeros((4, 66), dtype=np.float32), 3)             # nstate mismatch (decode.py:52)
    with pytest.raises(cabi.SloikaB200Error):                              # C-ABI error codes become exceptions
        cabi.check(-2, 'unsupported thing')


# ------------------------------------------------------------------ tcgen05 3xTF32 GEMM vs fp32 SIMT vs float64
def _linear_abi(x, W, b, act_code, algo, ldy=None):
    lib = cabi.load()
    M, K = x.shape
    N = W.shape[0]
    ldy = ldy or N
    xd, Wd = torch.from_numpy(x).to(DEV), torch.from_numpy(W).to(DEV)
    bd = None if b is None else torch.from_numpy(b).to(DEV)
    y = torch.full((M, ldy), -7.0, dtype=torch.float32, device=DEV)
    rc = lib.sloika_linear_fwd_ex(cabi.ptr(xd), K, cabi.ptr(Wd), cabi.ptr(bd), cabi.ptr(y), ldy, M, K, N, act_code, algo,
                                  cabi.stream_ptr(torch.device(DEV)))
    torch.cuda.synchronize()
    return rc, y.cpu().numpy()


@pytest.mark.parametrize('M,K,N,act_code', [(4096, 96, 288, 0), (1000, 96, 288, 1), (128, 96, 96, 0), (777, 128, 336, 0),
                                            (3000, 112, 432, 0), (2500, 144, 336, 2), (2048, 96, 1025, 0),
                                            (640, 32, 96, 3), (513, 40, 50, 0), (300, 8, 16, 0), (20000, 192, 128, 1),
                                            (1500, 256, 300, 0)])
def test_tensor_core_gemm_matches_fp32(M, K, N, act_code):
    rng = np.random.default_rng(M + K + N)
    x = np.tanh(rng.standard_normal((M, K))).astype(np.float32)
    W = (rng.standard_normal((N, K)) * 1.5).astype(np.float32)          # trained-model-like magnitudes
    b = rng.standard_normal(N).astype(np.float32)
    rc, y_tc = _linear_abi(x, W, b, act_code, 2)
    assert rc == 0
    rc, y_simt = _linear_abi(x, W, b, act_code, 1)
    assert rc == 0
    exact = x.astype(np.float64) @ W.astype(np.float64).T + b
    fun = {0: lambda v: v, 1: np.tanh, 2: lambda v: 1 / (1 + np.exp(-v)), 3: lambda v: np.where(v > 0, v, np.expm1(np.minimum(v, 0)))}[act_code]
    exact = fun(exact)
    err_tc = np.abs(y_tc - exact).max()
    err_simt = np.abs(y_simt - exact).max()
    scale = np.abs(exact).max()
    # the 3xTF32 split must be as good as plain fp32 accumulation (both ~1e-6 relative to the row scale)
    assert err_tc <= max(4 * err_simt, 4e-6 * scale), (err_tc, err_simt, scale)


@pytest.mark.parametrize('M,K,N,act_code', [(4096, 96, 288, 0), (1000, 96, 288, 1), (128, 96, 96, 0), (777, 128, 336, 0),
                                            (3000, 110, 426, 0), (2500, 144, 336, 2), (2048, 96, 1025, 0),
                                            (640, 32, 96, 3), (513, 40, 50, 0), (300, 8, 16, 0), (20000, 192, 128, 1),
                                            (1500, 256, 300, 0)])
def test_tensor_core_gemm_fp16_split_matches_fp32(M, K, N, act_code):
    """SLOIKA_GEMM_TC_F16 (fp16 hi/lo operands, for inputs bounded by 1) is as accurate as fp32 accumulation,
    including inputs so small that the lo parts fall into the fp16 subnormal range."""
    rng = np.random.default_rng(M + K + N)
    x = np.tanh(rng.standard_normal((M, K))).astype(np.float32)
    x[::7] *= 1e-4                                                      # rows of tiny activations
    W = (rng.standard_normal((N, K)) * 1.5).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    Kp = (K + 3) // 4 * 4                                               # TMA needs 16-byte rows
    xp = np.zeros((M, Kp), dtype=np.float32)
    xp[:, :K] = x
    lib = cabi.load()
    xd, Wd, bd = torch.from_numpy(xp).to(DEV), torch.from_numpy(W).to(DEV), torch.from_numpy(b).to(DEV)
    out = {}
    for algo in (3, 1):
        y = torch.full((M, N), -7.0, dtype=torch.float32, device=DEV)
        rc = lib.sloika_linear_fwd_ex(cabi.ptr(xd), Kp, cabi.ptr(Wd), cabi.ptr(bd), cabi.ptr(y), N, M, K, N, act_code, algo,
                                      cabi.stream_ptr(torch.device(DEV)))
        assert rc == 0
        torch.cuda.synchronize()
        out[algo] = y.cpu().numpy()
    exact = x.astype(np.float64) @ W.astype(np.float64).T + b
    fun = {0: lambda v: v, 1: np.tanh, 2: lambda v: 1 / (1 + np.exp(-v)), 3: lambda v: np.where(v > 0, v, np.expm1(np.minimum(v, 0)))}[act_code]
    exact = fun(exact)
    err_tc, err_simt = np.abs(out[3] - exact).max(), np.abs(out[1] - exact).max()
    scale = np.abs(exact).max()
    assert err_tc <= max(4 * err_simt, 4e-6 * scale), (err_tc, err_simt, scale)


def test_engine_picks_fp16_split_only_for_bounded_inputs(monkeypatch):
    """tanh / GRU outputs feed the fp16-split GEMM, elu outputs and huge weights do not; results agree either way."""
    np.random.seed(4)
    net = zoo.raw_rgrgr()
    x = torch.randn((600, 16, 1), device=DEV)
    calc = net.compile()
    with_f16 = calc.forward_device(x, None).data.cpu().numpy()
    monkeypatch.setenv('SLOIKA_B200_NO_F16', '1')
    without = calc.forward_device(x, None).data.cpu().numpy()
    monkeypatch.delenv('SLOIKA_B200_NO_F16')
    np.testing.assert_allclose(with_f16, without, atol=2e-5)
    a = engine.Act(torch.zeros((4, 2, 8), device=DEV), bounded=True)
    big = layers.Param(np.full((3, 8), 5e4, dtype=np.float32))
    assert engine._gemm_algo(a, net.layers[-1].W) == engine.GEMM_TC_F16
    assert engine._gemm_algo(a, big) == engine.GEMM_AUTO
    assert engine._gemm_algo(engine.Act(torch.zeros((4, 2, 8), device=DEV)), net.layers[-1].W) == engine.GEMM_AUTO
    elu_out = engine.run_convolution(layers.Convolution(1, 8, 11, 2, has_bias=True, fun=act.elu),
                                     engine.Act(torch.randn((100, 2, 1), device=DEV)))
    assert not elu_out.bounded
    tanh_out = engine.run_convolution(layers.Convolution(1, 8, 11, 2, has_bias=True, fun=act.tanh),
                                      engine.Act(torch.randn((100, 2, 1), device=DEV)))
    assert tanh_out.bounded


def test_tensor_core_gemm_column_slice_and_fallback():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((512, 64)).astype(np.float32)
    W = rng.standard_normal((48, 64)).astype(np.float32)
    rc, y = _linear_abi(x, W, None, 0, 2, ldy=100)                     # writes only its 48 columns of 100
    assert rc == 0
    np.testing.assert_allclose(y[:, :48], x @ W.T, atol=2e-4)
    assert np.all(y[:, 48:] == -7.0)
    xo = rng.standard_normal((512, 110)).astype(np.float32)            # ld not a multiple of 4: no TMA
    Wo = rng.standard_normal((30, 110)).astype(np.float32)
    rc, _ = _linear_abi(xo, Wo, None, 0, 2)
    assert rc == -2                                                    # forced TC refuses
    rc, y = _linear_abi(xo, Wo, None, 0, 0)                            # AUTO falls back to SIMT
    assert rc == 0
    np.testing.assert_allclose(y, xo @ Wo.T, atol=2e-4)


# ------------------------------------------------------------------ fused softmax -> Viterbi (logits + row statistics)
def test_fused_logits_decode_equals_posterior_decode():
    """The fused basecall path (final layer emits logits + per-slice (max, sum exp), the decoder applies
    softmax division, min_prob floor and log itself) must give the same paths as decoding the
    materialised posteriors, for dense and ragged batches."""
    np.random.seed(21)
    net = zoo.raw_rgrgr()
    net.layers[-1].W.set_value(net.layers[-1].W.get_value() * 4)        # peaky posteriors like a trained model
    calc = net.compile()
    x = torch.randn((1500, 9, 1), device=DEV)
    lengths = torch.tensor([1500, 1, 7, 640, 1499, 12, 300, 1000, 5], dtype=torch.int32, device=DEV)
    for lens in (None, lengths):
        std = calc.forward_device(x, lens)
        fused = calc.forward_device(x, lens, fused_decode=True)
        assert isinstance(fused, engine.LogitsAct)
        s1, p1 = decode.viterbi_batch(std, None, min_prob=1e-5)
        s2, p2 = decode.viterbi_batch(fused, None, min_prob=1e-5)
        assert p1 == p2
        np.testing.assert_allclose(s1, s2, rtol=1e-5, atol=1e-3)
        # and both agree with the oracle decoding the same posteriors
        post = std.data.cpu().numpy()
        nev = std.lengths.cpu().numpy() if lens is not None else [post.shape[0]] * post.shape[1]
        for b in (0, 3):
            s_ref, p_ref = decode_ref.decode_post(post[:nev[b], b:b + 1], 5, 1e-5, skip=0.0)
            assert p_ref == p1[b]
    # min_prob = 0 disables the exact floor shortcut
    s3, p3 = decode.viterbi_batch(calc.forward_device(x, None, fused_decode=True), None, min_prob=0.0)
    s4, p4 = decode.viterbi_batch(calc.forward_device(x, None), None, min_prob=0.0)
    assert p3 == p4


def test_softmax_normalise_matches_oracle_with_stay_column():
    np.random.seed(22)
    layer = layers.Softmax(96, 1025, init=_init(), has_bias=True)
    layer.W.set_value(layer.W.get_value() * 8)
    x = np.tanh(np.random.standard_normal((70, 4, 96))).astype(np.float32)
    got = _run(layer, x)[0]
    ref = _oracle(layer, x)
    assert got.shape == ref.shape and np.abs(got - ref).max() < 2e-5
    a = engine.Act(torch.from_numpy(x).to(DEV))
    la = engine.run_softmax_logits(layer, a)
    torch.cuda.synchronize()
    logits = la.data.cpu().numpy()
    W, b = layer.W.get_value().astype(np.float64), layer.b.get_value().astype(np.float64)
    exact = x.astype(np.float64) @ W.T + b
    np.testing.assert_allclose(logits[:, :, :1024], exact[:, :, 1:], atol=2e-4)    # k-mer states first
    np.testing.assert_allclose(logits[:, :, 1024], exact[:, :, 0], atol=2e-4)      # stay last
    st = la.stats.cpu().numpy().reshape(70, 4, la.n_slices, 2)
    m = st[..., 0].max(-1)
    s = (st[..., 1] * np.exp(st[..., 0] - m[..., None])).sum(-1)
    np.testing.assert_allclose(m, exact.max(-1), atol=2e-4)
    np.testing.assert_allclose(s, np.exp(exact - exact.max(-1, keepdims=True)).sum(-1), rtol=1e-4)


# ------------------------------------------------------------------ config 5: long reads, decode only
def test_viterbi_long_reads_bit_exact():
    """Decode-only sweep shape (synthetic posteriors, tens of thousands of events per read): scores and paths
    bit-identical to the C oracle given the same log-posteriors; the traceback lives in HBM (30 MB/read)."""
    rng = np.random.default_rng(33)
    T, B, S = 30000, 3, 1025
    lp = torch.empty((T, B, S), device=DEV)
    gen = torch.Generator(device=DEV).manual_seed(5)
    for b in range(B):                                   # built on the device to keep the test quick
        logits = 3 * torch.randn((T, S), generator=gen, device=DEV)
        logits[:, 0] += 6
        lp[:, b] = torch.log_softmax(logits, dim=1)
    lengths = np.array([T, 12345, 29999], dtype=np.int32)
    score, paths = decode.viterbi_batch(lp, lengths, log=True)
    ref_score, ref_paths = cbind.viterbi_batch(lp.cpu().numpy(), lengths)
    assert paths == ref_paths
    assert np.array_equal(score, ref_score)
    assert all(len(p) > 1000 for p in paths)


def test_batch_not_multiple_of_cta_tile():
    """1023 sequences (the GRU kernels tile 8 per CTA, Viterbi 1 per CTA): last CTA partially filled."""
    np.random.seed(44)
    net = zoo.raw_rgrgr()
    x = torch.randn((300, 13, 1), device=DEV)
    post13 = net.compile().forward_device(x).data.cpu().numpy()
    ref = _oracle(net, x.cpu().numpy())
    assert np.abs(post13 - ref).max() < 1e-4


def test_gated_gemm_chooses_operand_format_on_device():
    """sloika_conv1d_fwd_ex reports max |y|; sloika_linear_fwd_gated runs the fp16-split GEMM below the limit and the
    tf32-split one above it -- where the fp16 form would overflow -- with no host synchronisation in between."""
    lib = cabi.load()
    rng = np.random.default_rng(12)
    dev = torch.device(DEV)
    st = cabi.stream_ptr(dev)
    # conv range report
    conv = layers.Convolution(1, 16, 11, 5, has_bias=True, fun=act.elu)
    x = torch.from_numpy((rng.standard_normal((500, 6, 1)) * 3).astype(np.float32)).to(dev)
    out = engine.run_convolution(conv, engine.Act(x))
    assert not out.bounded and out.absmax is not None
    assert float(out.absmax.item()) == float(out.data.abs().max().item())
    # gated GEMM on small and on huge inputs
    M, K, N = 1024, 96, 288
    W = (rng.standard_normal((N, K)) * 0.5).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    Wd, bd = torch.from_numpy(W).to(dev), torch.from_numpy(b).to(dev)
    for scale in (1.0, 1.0e5):
        xh = (rng.standard_normal((M, K)) * scale).astype(np.float32)
        xd = torch.from_numpy(xh).to(dev)
        amax = xd.abs().max().reshape(1).contiguous()
        y = torch.full((M, N), -7.0, dtype=torch.float32, device=dev)
        rc = lib.sloika_linear_fwd_gated(cabi.ptr(xd), K, cabi.ptr(Wd), cabi.ptr(bd), cabi.ptr(y), N, M, K, N, 0,
                                         cabi.ptr(amax), 1.0e4, st)
        assert rc == 0
        torch.cuda.synchronize()
        exact = xh.astype(np.float64) @ W.astype(np.float64).T + b
        got = y.cpu().numpy()
        assert np.isfinite(got).all()
        assert np.abs(got - exact).max() <= 4e-6 * np.abs(exact).max()
    # the engine uses it for elu -> GRU projection and results agree with the forced tf32 path
    np.random.seed(3)
    net = zoo.pretrained_like().compile()
    xs = torch.randn((900, 8, 1), device=dev)
    a = net.forward_device(xs, None).data.cpu().numpy()
    os.environ['SLOIKA_B200_NO_F16'] = '1'
    try:
        bref = net.forward_device(xs, None).data.cpu().numpy()
    finally:
        del os.environ['SLOIKA_B200_NO_F16']
    np.testing.assert_allclose(a, bref, atol=2e-5)


@pytest.mark.parametrize('I,H,T,B,reverse,ragged', [(96, 96, 200, 16, False, False), (40, 110, 60, 5, True, True),
                                                    (24, 64, 50, 9, False, True), (128, 144, 40, 8, True, False)])
def test_gru_one_call_entry_point(I, H, T, B, reverse, ragged):
    """`sloika_gru_fwd` (projection + recurrence behind one C call, dense vI workspace -- the form a foreign caller
    binds) against the oracle; covers the dense, unaligned vI pitch (3H not a multiple of 4) of the recurrence."""
    lib = cabi.load()
    np.random.seed(I + H + T)
    g = layers.Gru(I, H, init=_init(), has_bias=True)
    g.sW.set_value(g.sW.get_value() * 4)
    x = np.tanh(np.random.standard_normal((T, B, I))).astype(np.float32)
    lengths = np.random.randint(1, T + 1, size=B).astype(np.int32) if ragged else None
    dev = torch.device(DEV)
    xd = torch.from_numpy(x).to(dev)
    y = torch.full((T, B, H), 7.0, dtype=torch.float32, device=dev)
    nbytes = lib.sloika_gru_workspace_bytes(T, B, H)
    assert nbytes == T * B * 3 * H * 4
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    ld = None if lengths is None else torch.from_numpy(lengths).to(dev)
    args = [cabi.ptr(xd), I, cabi.ptr(g.iW.device(dev)), cabi.ptr(g.sW.device(dev)), cabi.ptr(g.sW2.device(dev)),
            cabi.ptr(g.b.device(dev)), cabi.ptr(y), H, cabi.ptr(ws), nbytes, cabi.ptr(ld), T, B, I, H,
            1 if reverse else 0, 1, 2, cabi.stream_ptr(dev)]
    assert lib.sloika_gru_fwd(*args) == 0
    torch.cuda.synchronize()
    got = y.cpu().numpy()
    layer = layers.Reverse(g) if reverse else g
    for b in range(B):
        n = T if lengths is None else int(lengths[b])
        ref = _oracle(layer, x[:n, b:b + 1])
        assert np.abs(got[:n, b] - ref[:, 0]).max() < 5e-5, b
        assert np.all(got[n:, b] == 0)
    args[9] = nbytes - 4                                       # workspace too small
    assert lib.sloika_gru_fwd(*args) == -3


@pytest.mark.parametrize('stepwise', [False, True])
@pytest.mark.parametrize('I,H,T,B,reverse,ragged', [(24, 160, 40, 6, False, False), (32, 200, 30, 130, True, True),
                                                    (16, 256, 25, 3, False, True), (40, 300, 12, 5, True, True)])
def test_gru_wider_than_one_sm(I, H, T, B, reverse, ragged, stepwise, monkeypatch):
    """H > 144 does not fit one SM.  144 < H <= 256 runs on a 4-CTA cluster that keeps the weights in the four CTAs'
    shared memory and exchanges the state through distributed shared memory (gru_h16.cu, `gru_cluster_kernel`); wider
    still (and with SLOIKA_B200_GRU_STEPWISE) the scan runs as per-step GEMMs + gate kernels (gru.cu, `gru_stepwise`).
    Same results as the oracle, ragged and reversed included (B = 130 takes the tensor-core GEMM)."""
    if stepwise:
        monkeypatch.setenv('SLOIKA_B200_GRU_STEPWISE', '1')
    np.random.seed(H + T)
    g = layers.Gru(I, H, init=_init(), has_bias=True)
    g.sW.set_value(g.sW.get_value() * 3)
    layer = layers.Reverse(g) if reverse else g
    x = np.tanh(np.random.standard_normal((T, B, I))).astype(np.float32)
    lengths = list(np.random.randint(1, T + 1, size=B)) if ragged else None
    got, _ = _run(layer, x, lengths)
    for b in range(min(B, 12)):
        n = T if lengths is None else int(lengths[b])
        ref = _oracle(layer, x[:n, b:b + 1])
        assert np.abs(got[:n, b] - ref[:, 0]).max() < 5e-5, b
        assert np.all(got[n:, b] == 0)


# ------------------------------------------------------------------ the reference's own outputs (tools/make_golden_forward.py)
_SUPPORTED = {'serial', 'parallel', 'reverse', 'convolution', 'GRU', 'feed-forward', 'softmax_old', 'LSTM', 'window'}


def _types(arch):
    out = {arch['type']}
    for sub in arch.get('sublayers', []):
        out |= _types(sub)
    if 'sublayer' in arch:
        out |= _types(arch['sublayer'])
    return out


def test_cuda_path_matches_reference_outputs(forward_cases):
    """Every case of tests/golden/forward_cases.npz: the CUDA path against what the reference's unmodified
    layers.py / conv.py computed (not against the oracle).  1e-4 max-abs on posteriors (north-star bound), 2e-5
    on bounded hidden activations, 2e-5 relative on elu outputs."""
    from conftest import case_weights
    meta, data = forward_cases
    ran, bad = 0, []
    for case in meta:
        name = case['name']
        assert _types(case['arch']) <= _SUPPORTED, name
        try:
            net = zoo.from_weights(case['arch'], case_weights(data, name))
        except NotImplementedError as err:
            bad.append((name, repr(err)))
            continue
        x = data[name + '/x']
        ref = data[name + '/y']
        got = net.compile()(x)
        assert got.shape == ref.shape and got.dtype == np.float32, name
        if ref.size == 0:
            continue
        err = np.abs(got - ref)
        if name.startswith('model_') or name.startswith('softmax'):
            ok = err.max() < 1e-4
        else:
            ok = (err <= 2e-5 + 2e-5 * np.abs(ref)).all()
        if not ok:
            bad.append((name, float(err.max())))
        ran += 1
    assert not bad, bad
    assert ran >= 40


# ------------------------------------------------------------------ events route operators (row f4)
@pytest.mark.parametrize('I,H,T,B,reverse,peep', [(12, 64, 60, 7, False, True), (12, 64, 40, 5, True, True),
                                                  (5, 16, 30, 9, False, False), (20, 100, 25, 3, True, True),
                                                  (8, 128, 20, 2, False, True)])
def test_lstm(I, H, T, B, reverse, peep):
    """Lstm (layers.py:599-697) against the oracle (pinned to the reference's own Lstm.step by forward_cases.npz),
    ragged batch: every read sees the computation the reference gives it alone."""
    np.random.seed(I + H)
    g = layers.Lstm(I, H, init=_init(), has_bias=True, has_peep=peep)
    g.sW.set_value(g.sW.get_value() * 4)
    layer = layers.Reverse(g) if reverse else g
    lengths = [int(v) for v in np.random.randint(1, T + 1, size=B)]
    lengths[0] = T
    x = np.random.standard_normal((T, B, I)).astype(np.float32)
    got, _ = _run(layer, x, lengths)
    for b, n in enumerate(lengths):
        ref = _oracle(layer, x[:n, b:b + 1])
        assert np.abs(got[:n, b] - ref[:, 0]).max() < 5e-5, b
        assert np.all(got[n:, b] == 0)


def test_window_and_events_model():
    np.random.seed(8)
    x = np.random.standard_normal((23, 4, 4)).astype(np.float32)
    for w in (1, 3, 5):
        layer = layers.Window(4, w)
        got, _ = _run(layer, x, [23, 1, 7, 22])
        for b, n in enumerate([23, 1, 7, 22]):
            ref = forward_ref.window({'w': w}, x[:n, b:b + 1], np.float32)
            assert np.array_equal(got[:n, b], ref[:, 0]) and np.all(got[n:, b] == 0)
    # models/baseline_lstm.py architecture: window + birnn(Lstm) + FF + birnn(Lstm) + FF + softmax
    init = _init()
    lstm = lambda i, o: layers.Lstm(i, o, init=init, has_bias=True, has_peep=True)
    net = layers.Serial([layers.Window(4, 3), layers.birnn(lstm(12, 64), lstm(12, 64)),
                         layers.FeedForward(128, 64, init=init, has_bias=True),
                         layers.birnn(lstm(64, 64), lstm(64, 64)),
                         layers.FeedForward(128, 64, init=init, has_bias=True),
                         layers.Softmax(64, 1025, init=init, has_bias=True)])
    x = np.random.standard_normal((150, 3, 4)).astype(np.float32)
    post = net.compile()(x)
    ref = _oracle(net, x)
    assert post.shape == ref.shape and np.abs(post - ref).max() < 1e-4


def test_feedforward_wide_input_on_tensor_cores():
    """K up to 512 (the widened bigger_raw_gru joins two 256-unit GRUs into a FeedForward) runs on the tcgen05 GEMM."""
    np.random.seed(31)
    for K, N in ((512, 256), (384, 96), (300, 40)):
        layer = layers.FeedForward(K, N, init=_init(), has_bias=True, fun=act.tanh)
        x = np.tanh(np.random.standard_normal((40, 16, K))).astype(np.float32)
        a = engine.Act(torch.from_numpy(x).to(DEV), bounded=True)
        got = layer.run(a).data.cpu().numpy()
        ref = _oracle(layer, x)
        assert np.abs(got - ref).max() < 2e-5, (K, N)


@pytest.mark.parametrize('I,H,T,B,reverse,ragged', [(96, 96, 60, 70, False, False), (96, 96, 45, 64, True, True),
                                                    (32, 96, 30, 5, False, True), (64, 64, 25, 130, True, False),
                                                    (96, 80, 20, 33, False, True), (40, 50, 20, 17, True, True),
                                                    (20, 32, 15, 200, False, False)])
def test_gru_fused_projection(I, H, T, B, reverse, ragged, monkeypatch):
    """The GRU layer with the projection inside the recurrence launch (csrc/gru_fused.cu: clusters of two recurrence
    CTAs and one projection CTA, vI handed over through an L2-resident ring) against the oracle; batches that do not
    fill a cluster, ragged and reversed included."""
    monkeypatch.setenv('SLOIKA_B200_FUSED_GRU', '1')
    np.random.seed(I + H + B)
    g = layers.Gru(I, H, init=_init(), has_bias=True)
    g.sW.set_value(g.sW.get_value() * 5)
    g.sW2.set_value(g.sW2.get_value() * 5)
    layer = layers.Reverse(g) if reverse else g
    lengths = [int(v) for v in np.random.randint(1, T + 1, size=B)] if ragged else None
    if lengths:
        lengths[0] = T
    x = np.tanh(np.random.standard_normal((T, B, I))).astype(np.float32)
    a = engine.Act(torch.from_numpy(x).to(DEV), None if lengths is None else torch.as_tensor(lengths, dtype=torch.int32, device=DEV),
                   bounded=True)
    assert engine._fused_gru_ok(g, a) == 'fused'
    engine.TIMER.reset()
    out = layer.run(a)
    torch.cuda.synchronize()
    got = out.data.cpu().numpy()
    for b in list(range(0, B, 7)) + [B - 1]:
        n = T if lengths is None else lengths[b]
        ref = _oracle(layer, x[:n, b:b + 1])
        assert np.abs(got[:n, b] - ref[:, 0]).max() < 5e-5, b
        assert np.all(got[n:, b] == 0)


@pytest.mark.parametrize('scale,reverse', [(1.0, False), (3.0e4, True)])
def test_gru_gated_between_fused_and_projection(scale, reverse, monkeypatch):
    """An input whose range only the device knows (the elu convolution's output, max |x| in `Act.absmax`): both forms of
    the layer are enqueued (sloika_gru_fwd_gated) and the device word picks the fused launch (inside the fp16 range) or
    the tf32 projection + recurrence (outside); either way the result is the oracle's."""
    monkeypatch.setenv('SLOIKA_B200_FUSED_GRU', '1')
    np.random.seed(77)
    I, H, T, B = 96, 96, 40, 70
    g = layers.Gru(I, H, init=_init(), has_bias=True)
    if scale > 1:
        g.iW.set_value(g.iW.get_value() / scale * 4)              # keep the gates out of saturation for huge inputs
    layer = layers.Reverse(g) if reverse else g
    x = (np.random.standard_normal((T, B, I)) * scale).astype(np.float32)
    xd = torch.from_numpy(x).to(DEV)
    a = engine.Act(xd, None, bounded=False, absmax=xd.abs().max().reshape(1))
    assert engine._fused_gru_ok(g, a) == 'gated'
    engine.TIMER.reset()
    got = layer.run(a).data.cpu().numpy()
    assert engine.TIMER.launches == 3
    ref = _oracle(layer, x)
    tol = 5e-5 if scale == 1.0 else 2e-4                          # tf32-split products of inputs of magnitude 1e5
    assert np.abs(got - ref).max() < tol
    if scale == 1.0:
        # the other verdict on the same input: projection + recurrence
        other = engine.Act(xd, None, bounded=False, absmax=torch.tensor([1.0e9], dtype=torch.float32, device=DEV))
        got2 = layer.run(other).data.cpu().numpy()
        assert np.abs(got2 - ref).max() < 5e-5


def _to_blocked(x):
    """NumPy restatement of the blocked layout of include/sloika_b200.h (sloika_gru_seq_fwd)."""
    T, B, F = x.shape
    nblk, fg = (B + 127) // 128, (F + 3) // 4
    out = np.zeros((T, nblk, fg, 128, 4), dtype=np.float32)
    pad = np.zeros((T, nblk * 128, fg * 4), dtype=np.float32)
    pad[:, :B, :F] = x
    out[:] = pad.reshape(T, nblk, 128, fg, 4).transpose(0, 1, 3, 2, 4)
    return out.reshape(-1)


def _from_blocked(flat, T, B, F):
    nblk, fg = (B + 127) // 128, (F + 3) // 4
    v = flat.reshape(T, nblk, fg, 128, 4).transpose(0, 1, 3, 2, 4).reshape(T, nblk * 128, fg * 4)
    return v[:, :B, :F]


@pytest.mark.parametrize('T,B,F', [(5, 300, 96), (3, 128, 50), (4, 7, 1), (2, 129, 110)])
def test_block_layout_round_trip(T, B, F):
    lib = cabi.load()
    rng = np.random.default_rng(T * B + F)
    x = rng.standard_normal((T, B, F)).astype(np.float32)
    pitch = (F + 7) // 8 * 8
    xd = torch.zeros((T, B, pitch), dtype=torch.float32, device=DEV)
    xd[:, :, :F] = torch.from_numpy(x).to(DEV)
    n = lib.sloika_blocked_bytes(T, B, F) // 4
    bd = torch.full((n,), 9.0, dtype=torch.float32, device=DEV)
    assert lib.sloika_block_layout_fwd(cabi.ptr(xd), cabi.ptr(bd), pitch, T, B, F, 1, cabi.stream_ptr(DEV)) == 0
    assert np.array_equal(bd.cpu().numpy(), _to_blocked(x))
    yd = torch.full((T, B, F), -3.0, dtype=torch.float32, device=DEV)          # dense rows: the unaligned store path for odd F
    assert lib.sloika_block_layout_fwd(cabi.ptr(bd), cabi.ptr(yd), F, T, B, F, 0, cabi.stream_ptr(DEV)) == 0
    assert np.array_equal(yd.cpu().numpy(), x)


@pytest.mark.parametrize('layout', [0, 3])
@pytest.mark.parametrize('I,H,T,B,reverse,ragged', [(96, 96, 40, 130, False, False), (96, 96, 33, 128, True, True),
                                                    (32, 96, 20, 5, False, True), (64, 64, 25, 200, True, False),
                                                    (96, 80, 20, 33, False, True), (40, 50, 20, 17, True, True),
                                                    (20, 32, 15, 300, False, False), (96, 96, 1, 3, False, False)])
def test_gru_sequences_on_lanes(I, H, T, B, reverse, ragged, layout):
    """The GRU layer with 128 sequences on the tensor-memory lanes (csrc/gru_seq.cu: activations as TMEM operands,
    weights in shared memory, projection accumulated in place) through the C ABI against the oracle, with row-major and
    with blocked tensors; batches that do not fill a CTA, ragged and reversed included."""
    np.random.seed(I + H + B + 1)
    g = layers.Gru(I, H, init=_init(), has_bias=True)
    g.sW.set_value(g.sW.get_value() * 5)
    g.sW2.set_value(g.sW2.get_value() * 5)
    layer = layers.Reverse(g) if reverse else g
    lengths = [int(v) for v in np.random.randint(1, T + 1, size=B)] if ragged else None
    if lengths:
        lengths[0] = T
    x = np.tanh(np.random.standard_normal((T, B, I))).astype(np.float32)
    lib = cabi.load()
    pitch = (I + 3) // 4 * 4
    if layout & 1:
        xd = torch.from_numpy(_to_blocked(x)).to(DEV)
    else:
        xd = torch.zeros((T, B, pitch), dtype=torch.float32, device=DEV)
        xd[:, :, :I] = torch.from_numpy(x).to(DEV)
    if layout & 2:
        yd = torch.full((lib.sloika_blocked_bytes(T, B, H) // 4,), 7.0, dtype=torch.float32, device=DEV)
    else:
        yd = torch.full((T, B, H), 7.0, dtype=torch.float32, device=DEV)
    ld = None if lengths is None else torch.as_tensor(lengths, dtype=torch.int32, device=DEV)
    rc = lib.sloika_gru_seq_fwd(cabi.ptr(xd), pitch, cabi.ptr(g.iW.device(DEV)), cabi.ptr(g.sW.device(DEV)),
                                cabi.ptr(g.sW2.device(DEV)), cabi.ptr(g.b.device(DEV)), cabi.ptr(yd), H, cabi.ptr(ld),
                                T, B, I, H, 1 if reverse else 0, 1, 2, layout, cabi.stream_ptr(DEV))
    assert rc == 0
    torch.cuda.synchronize()
    got = _from_blocked(yd.cpu().numpy(), T, B, H) if layout & 2 else yd.cpu().numpy()
    for b in list(range(0, B, 7)) + [B - 1]:
        n = T if lengths is None else lengths[b]
        ref = _oracle(layer, x[:n, b:b + 1])
        assert np.abs(got[:n, b] - ref[:, 0]).max() < 5e-5, b
        assert np.all(got[n:, b] == 0)


@pytest.mark.parametrize('scale', [1.0, 3.0e4])
def test_network_through_sequences_on_lanes(scale, monkeypatch):
    """A conv -> Reverse(Gru) -> Gru -> Gru stack with the GRU layers forced onto the sequences-on-lanes launch: the first
    one takes the convolution's row-major output through the gated entry (either verdict), the others pass blocked
    activations to each other, the caller reads a row-major result; ragged batch."""
    monkeypatch.setenv('SLOIKA_B200_FUSED_GRU', '1')
    monkeypatch.setenv('SLOIKA_B200_GRU_SEQ', '1')
    np.random.seed(5)
    net = layers.Serial([layers.Convolution(1, 96, 11, 5, init=_init(), has_bias=True, fun=act.elu),
                         layers.Reverse(layers.Gru(96, 96, init=_init(), has_bias=True)),
                         layers.Gru(96, 64, init=_init(), has_bias=True),
                         layers.Reverse(layers.Gru(64, 80, init=_init(), has_bias=True))])
    if scale > 1:
        W = net.layers[0].W.get_value()
        net.layers[0].W.set_value(W * scale)
        net.layers[1].layer.iW.set_value(net.layers[1].layer.iW.get_value() / scale * 4)
    T, B = 400, 150
    x = np.random.standard_normal((T, B, 1)).astype(np.float32)
    lengths = np.random.randint(50, T + 1, size=B).astype(np.int32)
    lengths[0] = T
    engine.TIMER.reset()
    out = net.run(engine.Act(torch.from_numpy(x).to(DEV), torch.from_numpy(lengths).to(DEV)))
    assert out.blocked is not None
    got = out.data.cpu().numpy()
    olen = out.lengths.cpu().numpy()
    for b in (0, 1, 77, B - 1):
        ref = _oracle(net, x[:lengths[b], b:b + 1])
        assert ref.shape[0] == olen[b]
        assert np.abs(got[:olen[b], b] - ref[:, 0]).max() < (5e-5 if scale == 1.0 else 3e-4), b
        assert np.all(got[olen[b]:, b] == 0)


def test_softmax_logits_from_blocked_input(monkeypatch):
    """The logits GEMM fed by a blocked activation (what the last sequences-on-lanes GRU layer leaves behind) gives the
    logits and row statistics of the row-major path, and the fused decode the same paths."""
    monkeypatch.setenv('SLOIKA_B200_FUSED_GRU', '1')
    monkeypatch.setenv('SLOIKA_B200_GRU_SEQ', '1')
    np.random.seed(23)
    T, B, K = 30, 256, 96
    layer = layers.Softmax(K, 1025, init=_init(), has_bias=True)
    layer.W.set_value(layer.W.get_value() * 4)
    x = np.tanh(np.random.standard_normal((T, B, K))).astype(np.float32)
    xb = torch.from_numpy(_to_blocked(x)).to(DEV)
    a_blk = engine.Act.from_blocked(xb, (T, B, K), None, False, bounded=True)
    a_row = engine.Act(torch.from_numpy(x).to(DEV), None, bounded=True)
    l_blk = engine.run_softmax_logits(layer, a_blk)
    assert a_blk._data is None                                   # no row-major copy was made
    l_row = engine.run_softmax_logits(layer, a_row)
    torch.cuda.synchronize()
    assert torch.equal(l_blk.data, l_row.data)
    assert torch.equal(l_blk.stats, l_row.stats)
    # a whole network: conv -> 5 GRU (sequences on lanes) -> logits from the blocked activation -> fused decode
    net = zoo.raw_rgrgr()
    net.layers[-1].W.set_value(net.layers[-1].W.get_value() * 4)
    calc = net.compile()
    xs = torch.randn((1000, 128, 1), device=DEV)
    engine.TIMER.reset()
    fused = calc.forward_device(xs, None, fused_decode=True)
    monkeypatch.setenv('SLOIKA_B200_GRU_SEQ', '0')
    monkeypatch.setenv('SLOIKA_B200_FUSED_GRU', '0')
    plain = calc.forward_device(xs, None, fused_decode=True)
    s1, p1 = decode.viterbi_batch(fused, None, min_prob=1e-5)
    s2, p2 = decode.viterbi_batch(plain, None, min_prob=1e-5)
    assert (fused.data - plain.data).abs().max().item() < 2e-4
    assert sum(a != b for a, b in zip(p1, p2)) <= 1            # two float32 roundings of the same network: near-ties may flip
    np.testing.assert_allclose(s1, s2, rtol=1e-4, atol=5e-2)


@pytest.mark.parametrize('C,stride,B,ragged', [(96, 5, 150, False), (32, 2, 33, True), (64, 5, 700, True)])
def test_convolution_blocked_output(C, stride, B, ragged, monkeypatch):
    """The raw-signal convolution writing the blocked layout (lanes = sequences, `sloika_conv1d_fwd_ex` with ldy = -1) gives
    bit for bit the values of the row-major kernel, and the same range report."""
    np.random.seed(C + B)
    layer = layers.Convolution(1, C, 11, stride, init=_init(), has_bias=True, fun=act.elu)
    T = 333
    x = torch.from_numpy(np.random.standard_normal((T, B, 1)).astype(np.float32)).to(DEV)
    lens = torch.from_numpy(np.random.randint(1, T + 1, size=B).astype(np.int32)).to(DEV) if ragged else None
    monkeypatch.setenv('SLOIKA_B200_GRU_SEQ', '0')
    row = layer.run(engine.Act(x, lens))
    assert row.blocked is None
    monkeypatch.setenv('SLOIKA_B200_GRU_SEQ', '1')
    blk = layer.run(engine.Act(x, lens))
    assert blk.blocked is not None and blk._data is None
    torch.cuda.synchronize()
    assert torch.equal(blk.absmax, row.absmax)
    if lens is not None:
        assert torch.equal(blk.lengths, row.lengths)
    got = _from_blocked(blk.blocked.cpu().numpy(), blk.T, B, C)
    assert np.array_equal(got, row.data.cpu().numpy())
    assert torch.equal(blk.data, row.data)                       # and through the conversion kernel
