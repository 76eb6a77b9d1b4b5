// Strided 1-D convolution over the time axis with fused bias + activation.
//
// Reference semantics: Convolution.run (sloika/layers.py:417-419) -> conv.conv_1d
// (sloika/conv.py:90-111): zero-pad time by (pad_l, pad_r), valid cross-correlation
// (filter_flip=False) with step `stride`, add bias, activation.  Theano materialises the padded copy,
// two transposes and an im2col GEMM; here padding is a predicate on the load and the layout
// [time, batch, feature] is consumed and produced in place.
//
// HBM-bound: 4*Cin bytes in and 4*Cout/stride bytes out per raw sample (Cin = 1 for raw signal).
#include "common.cuh"

namespace sloika {

// Raw-signal case Cin == 1, Cout % 4 == 0.  One CTA handles TT output steps x BT sequences.
//   smem xs[(TT-1)*stride + WIN][BT]  input window, zero filled outside [0, len_b)
//   thread = (c4, bq): fixed group of 4 output channels (weights + bias live in registers for the
//   whole CTA), loops over (t, b) pairs; a warp writes consecutive float4 of the dense [B, Cout]
//   slab of one time step -> fully coalesced 128-bit stores.
// Epilogue activations with the MUFU-based forms (absolute error ~1e-7, as in the recurrence epilogues): the
// accurate expm1f / tanhf of libdevice cost 25-40 instructions per value and made this kernel issue-bound at 24 % of
// the HBM rate (271 warp instructions per float4 of output, profiles/r2_step_kernels_ncu.txt).
template <int ACT>
__device__ __forceinline__ float conv_act(float v) {
    if constexpr (ACT == SLOIKA_ACT_ELU) return v > 0.0f ? v : ex2_ftz(v * SLOIKA_LOG2E) - 1.0f;
    else if constexpr (ACT == SLOIKA_ACT_TANH) return tanh_fast(v);
    else if constexpr (ACT == SLOIKA_ACT_SIGMOID) return sigmoid_fast(v);
    else return v;
}

template <int WIN, int ACT>
__global__ void __launch_bounds__(256, 2)
conv1d_raw_kernel(const float *__restrict__ x, const float *__restrict__ W, const float *__restrict__ bias,
                  float *__restrict__ y, long ldy, const int32_t *__restrict__ lengths, int T, int B, int Cout,
                  int stride, int pad_l, int Tout, int TT, int BT, int act, unsigned *__restrict__ absmax_bits)
{
    extern __shared__ float xs[];                    // [rows][BT]
    const int b0 = blockIdx.y * BT;           // time tiles on grid.x (2^31 - 1 of them), batch tiles on grid.y
    const int t0 = blockIdx.x * TT;
    const int rows = (TT - 1) * stride + WIN;
    const int nc4 = Cout >> 2;
    const int bq = blockDim.x / nc4;                 // sequences handled per pass
    const int c4 = threadIdx.x % nc4;
    const int bl = threadIdx.x / nc4;

    for (int e = threadIdx.x; e < rows * BT; e += blockDim.x) {
        const int r = e / BT, b = e - r * BT;
        const int tin = t0 * stride - pad_l + r;
        const int bg = b0 + b;
        float v = 0.0f;
        if (bg < B && tin >= 0 && tin < T) {
            const int len = lengths ? lengths[bg] : T;
            if (tin < len) v = __ldg(x + (long)tin * B + bg);
        }
        xs[e] = v;
    }

    float2 wa[WIN], wb[WIN];                         // channels (4c4, 4c4+1) and (4c4+2, 4c4+3): packed fp32x2 FMAs
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bl < bq) {
#pragma unroll
        for (int k = 0; k < WIN; k++) {
            wa[k].x = __ldg(W + (4 * c4 + 0) * WIN + k);
            wa[k].y = __ldg(W + (4 * c4 + 1) * WIN + k);
            wb[k].x = __ldg(W + (4 * c4 + 2) * WIN + k);
            wb[k].y = __ldg(W + (4 * c4 + 3) * WIN + k);
        }
        bv = __ldg(reinterpret_cast<const float4 *>(bias) + c4);
    }
    __syncthreads();
    if (bl >= bq) return;

    float amax = 0.0f;                               // max |y| of this thread (for the optional range report)
    for (int tl = 0; tl < TT; tl++) {
        const int t = t0 + tl;
        if (t >= Tout) break;
        for (int b = bl; b < BT; b += bq) {
            const int bg = b0 + b;
            if (bg >= B) break;
            float2 a0 = make_float2(bv.x, bv.y), a1 = make_float2(bv.z, bv.w);
            const float *xp = xs + (tl * stride) * BT + b;
#pragma unroll
            for (int k = 0; k < WIN; k++) {
                const float xv = xp[k * BT];
                const float2 x2 = make_float2(xv, xv);
                a0 = fma2(wa[k], x2, a0);
                a1 = fma2(wb[k], x2, a1);
            }
            float4 acc;
            acc.x = conv_act<ACT>(a0.x);
            acc.y = conv_act<ACT>(a0.y);
            acc.z = conv_act<ACT>(a1.x);
            acc.w = conv_act<ACT>(a1.y);
            float *yp = y + ((long)t * B + bg) * ldy + 4 * c4;
            __stcs(reinterpret_cast<float4 *>(yp), acc);     // streamed once, read by the next layer
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(acc.x), fabsf(acc.y)), fmaxf(fabsf(acc.z), fabsf(acc.w))));
        }
    }
    if (absmax_bits) {
        // non-negative floats order like their bit patterns; a NaN has the largest pattern, so it also reads as "huge"
#pragma unroll
        // the block size (a multiple of Cout / 4) need not be a multiple of 32 and threads leave the loops above at
        // different times: reduce over whoever is here with the hardware reduction under the real mask
        const unsigned mask = __activemask();
        const unsigned wmax = __reduce_max_sync(mask, __float_as_uint(amax));
        if ((threadIdx.x & 31) == (unsigned)(__ffs(mask) - 1)) atomicMax(absmax_bits, wmax);
    }
}

// The same convolution writing the BLOCKED layout of the sequences-on-lanes GRU (include/sloika_b200.h, sloika_gru_seq_fwd):
// element (t, b, c) at (((t * nblk + b / 128) * (Cout / 4) + c / 4) * 128 + b % 128) * 4 + c % 4.  For the stores to be
// contiguous the lanes of a warp must be consecutive SEQUENCES, so the mapping is turned: lane = sequence (BT = 32), warp =
// group of four channels; a warp walks the channel groups, reloading its 44 weights per group, and the time tile inside.
template <int WIN, int ACT>
__global__ void __launch_bounds__(256, 2)
conv1d_raw_blocked_kernel(const float *__restrict__ x, const float *__restrict__ W, const float *__restrict__ bias,
                          float *__restrict__ y, const int32_t *__restrict__ lengths, int T, int B, int Cout, int stride,
                          int pad_l, int Tout, int TT, unsigned *__restrict__ absmax_bits)
{
    constexpr int BT = 32;
    extern __shared__ float xs[];                    // [rows][BT]
    const int b0 = blockIdx.y * BT;
    const int t0 = blockIdx.x * TT;
    const int rows = (TT - 1) * stride + WIN;
    const int nc4 = Cout >> 2;
    for (int e = threadIdx.x; e < rows * BT; e += blockDim.x) {
        const int r = e / BT, b = e - r * BT;
        const int tin = t0 * stride - pad_l + r;
        const int bg = b0 + b;
        float v = 0.0f;
        if (bg < B && tin >= 0 && tin < T) {
            const int len = lengths ? lengths[bg] : T;
            if (tin < len) v = __ldg(x + (long)tin * B + bg);
        }
        xs[e] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int bg = b0 + lane;
    const long nblk = (B + 127) / 128;
    float4 *y4 = reinterpret_cast<float4 *>(y);
    float amax = 0.0f;
    for (int c4 = warp; c4 < nc4; c4 += nwarps) {
        float2 wa[WIN], wb[WIN];
#pragma unroll
        for (int k = 0; k < WIN; k++) {
            wa[k].x = __ldg(W + (4 * c4 + 0) * WIN + k);
            wa[k].y = __ldg(W + (4 * c4 + 1) * WIN + k);
            wb[k].x = __ldg(W + (4 * c4 + 2) * WIN + k);
            wb[k].y = __ldg(W + (4 * c4 + 3) * WIN + k);
        }
        const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias) + c4);
        if (bg >= B) continue;
        for (int tl = 0; tl < TT; tl++) {
            const int t = t0 + tl;
            if (t >= Tout) break;
            float2 a0 = make_float2(bv.x, bv.y), a1 = make_float2(bv.z, bv.w);
            const float *xp = xs + (tl * stride) * BT + lane;
#pragma unroll
            for (int k = 0; k < WIN; k++) {
                const float xv = xp[k * BT];
                const float2 x2 = make_float2(xv, xv);
                a0 = fma2(wa[k], x2, a0);
                a1 = fma2(wb[k], x2, a1);
            }
            float4 acc;
            acc.x = conv_act<ACT>(a0.x);
            acc.y = conv_act<ACT>(a0.y);
            acc.z = conv_act<ACT>(a1.x);
            acc.w = conv_act<ACT>(a1.y);
            __stcs(y4 + (((long)t * nblk + (bg >> 7)) * nc4 + c4) * 128 + (bg & 127), acc);
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(acc.x), fabsf(acc.y)), fmaxf(fabsf(acc.z), fabsf(acc.w))));
        }
    }
    if (absmax_bits) {
        const unsigned mask = __activemask();
        const unsigned wmax = __reduce_max_sync(mask, __float_as_uint(amax));
        if ((threadIdx.x & 31) == (unsigned)(__ffs(mask) - 1)) atomicMax(absmax_bits, wmax);
    }
}

// General case (any Cin / winlen / Cout): one thread per output element, operands through L1/L2.
// Not on the raw-signal hot path; kept so the operator covers the reference's full signature.
__global__ void conv1d_generic_kernel(const float *__restrict__ x, const float *__restrict__ W,
                                      const float *__restrict__ bias, float *__restrict__ y, long ldy,
                                      const int32_t *__restrict__ lengths, int T, int B, int Cin, int Cout,
                                      int winlen, int stride, int pad_l, int Tout, int act,
                                      unsigned *__restrict__ absmax_bits)
{
    const long total = (long)Tout * B * Cout;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int o = (int)(e % Cout);
        const long tb = e / Cout;
        const int b = (int)(tb % B);
        const int t = (int)(tb / B);
        const int len = lengths ? lengths[b] : T;
        float acc = bias[o];
        for (int i = 0; i < Cin; i++) {
            for (int k = 0; k < winlen; k++) {
                const int tin = t * stride - pad_l + k;
                if (tin >= 0 && tin < T && tin < len)
                    acc = fmaf(W[((long)o * Cin + i) * winlen + k], x[((long)tin * B + b) * Cin + i], acc);
            }
        }
        const float v = apply_act(acc, act);
        y[tb * ldy + o] = v;
        if (absmax_bits) atomicMax(absmax_bits, __float_as_uint(fabsf(v)));
    }
}

}  // namespace sloika

using namespace sloika;

extern "C" int sloika_conv1d_fwd_ex(const float *x, const float *W, const float *bias, float *y, long ldy,
                                    const int32_t *lengths, int T, int B, int Cin, int Cout, int winlen, int stride,
                                    int pad_l, int pad_r, int act, float *absmax, void *stream);

extern "C" int sloika_conv1d_fwd(const float *x, const float *W, const float *bias, float *y, long ldy,
                                 const int32_t *lengths, int T, int B, int Cin, int Cout, int winlen, int stride,
                                 int pad_l, int pad_r, int act, void *stream)
{
    return sloika_conv1d_fwd_ex(x, W, bias, y, ldy, lengths, T, B, Cin, Cout, winlen, stride, pad_l, pad_r, act, nullptr,
                                stream);
}

extern "C" int sloika_conv1d_fwd_ex(const float *x, const float *W, const float *bias, float *y, long ldy,
                                    const int32_t *lengths, int T, int B, int Cin, int Cout, int winlen, int stride,
                                    int pad_l, int pad_r, int act, float *absmax, void *stream)
{
    unsigned *absmax_bits = reinterpret_cast<unsigned *>(absmax);
    if (!x || !W || !bias || !y) return SLOIKA_ERR_ARG;
    const bool blocked_out = ldy == -1;               // y in the blocked layout of sloika_gru_seq_fwd (raw-signal shape only)
    if (T < 0 || B <= 0 || Cin <= 0 || Cout <= 0 || winlen <= 0 || stride <= 0 || pad_l < 0 || pad_r < 0 ||
        (!blocked_out && ldy < Cout))
        return SLOIKA_ERR_ARG;
    if (!act_known(act)) return SLOIKA_ERR_UNSUPPORTED;
    const long span = (long)T + pad_l + pad_r - winlen;
    const int Tout = span < 0 ? 0 : (int)(span / stride + 1);
    if (Tout == 0) return SLOIKA_OK;
    cudaStream_t st = (cudaStream_t)stream;

    if (blocked_out) {
        const int TT = 16;
        const size_t smem = sizeof(float) * ((size_t)(TT - 1) * stride + winlen) * 32;
        if (Cin != 1 || winlen != 11 || Cout % 4 != 0 || ((uintptr_t)y & 15) != 0 || ((uintptr_t)bias & 15) != 0 ||
            smem > 48 * 1024 || ceil_div(B, 32) > 65535)
            return SLOIKA_ERR_UNSUPPORTED;
        dim3 grid((unsigned)ceil_div(Tout, TT), (unsigned)ceil_div(B, 32));
#define CONV_BLK(A) conv1d_raw_blocked_kernel<11, A><<<grid, 256, smem, st>>>(x, W, bias, y, lengths, T, B, Cout, stride, pad_l, \
                                                                               Tout, TT, absmax_bits)
        switch (act) {
            case SLOIKA_ACT_ELU: CONV_BLK(SLOIKA_ACT_ELU); break;
            case SLOIKA_ACT_TANH: CONV_BLK(SLOIKA_ACT_TANH); break;
            case SLOIKA_ACT_SIGMOID: CONV_BLK(SLOIKA_ACT_SIGMOID); break;
            default: CONV_BLK(SLOIKA_ACT_LINEAR); break;
        }
#undef CONV_BLK
        SLOIKA_RETURN_LAUNCH_STATUS();
    }

    const bool aligned = (Cout % 4 == 0) && (ldy % 4 == 0) && (((uintptr_t)y & 15) == 0) &&
                         (((uintptr_t)bias & 15) == 0) && (Cout / 4 <= 256);
    if (Cin == 1 && winlen == 11 && aligned) {
        const int BT = 32, TT = 16;
        const int nc4 = Cout / 4;
        const int threads = (256 / nc4) * nc4;
        dim3 grid((unsigned)ceil_div(Tout, TT), (unsigned)ceil_div(B, BT));
        const size_t smem = sizeof(float) * ((size_t)(TT - 1) * stride + winlen) * BT;
        if (smem <= 48 * 1024 && ceil_div(B, BT) <= 65535) {
#define CONV_RAW(A) conv1d_raw_kernel<11, A><<<grid, threads, smem, st>>>(x, W, bias, y, ldy, lengths, T, B, Cout, stride, \
                                                                       pad_l, Tout, TT, BT, act, absmax_bits)
            switch (act) {
                case SLOIKA_ACT_ELU: CONV_RAW(SLOIKA_ACT_ELU); break;
                case SLOIKA_ACT_TANH: CONV_RAW(SLOIKA_ACT_TANH); break;
                case SLOIKA_ACT_SIGMOID: CONV_RAW(SLOIKA_ACT_SIGMOID); break;
                default: CONV_RAW(SLOIKA_ACT_LINEAR); break;
            }
#undef CONV_RAW
            SLOIKA_RETURN_LAUNCH_STATUS();
        }
    }
    const long total = (long)Tout * B * Cout;
    const int threads = 256;
    long blocks = ceil_div(total, threads);
    if (blocks > 148L * 32) blocks = 148L * 32;
    conv1d_generic_kernel<<<(unsigned)blocks, threads, 0, st>>>(x, W, bias, y, ldy, lengths, T, B, Cin, Cout,
                                                               winlen, stride, pad_l, Tout, act, absmax_bits);
    SLOIKA_RETURN_LAUNCH_STATUS();
}
