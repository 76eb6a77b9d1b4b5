// The events-route operators: Lstm recurrence and Window.
//
// Lstm.step scanned by Lstm.run (reference sloika/layers.py:677-697), input projection vW = x iW' + b precomputed for
// all steps by the tensor-core GEMM.  Stored parameter layout is the reference's: sW [4H, H], row 4*j + g belongs to
// unit j and gate g (the step reshapes the 4H pre-activations as (H, 4): g = 0 update input, 1 update gate,
// 2 forget gate, 3 output gate); peepholes p [3, H] (p[0] update gate, p[1] forget gate, p[2] output gate).
//     state' = state * gate(s2 + state * p1) + fun(s0) * gate(s1 + state * p0)
//     out'   = fun(state') * gate(s3 + state' * p2)            s = vW_t + out sW'
// Persistent CTA per 4 sequences, thread = pre-activation row, sW transposed into shared memory when it fits
// (H <= 96; the shipped events models use H = 64), fp32 FMA, two CTA barriers per step.  Not the throughput path
// of this repository (that is the raw GRU stack): correctness first, one launch per layer.
//
// Window.run (layers.py:346-351): out[t, b, k*F + f] = xpad[t + k, b, f] with w//2 zero steps either side.
#include "common.cuh"

namespace sloika {

constexpr int LSTM_BT = 4;

__global__ void __launch_bounds__(1024, 1)
lstm_recurrence_kernel(const float *__restrict__ vW, long ldv, const float *__restrict__ sW, const float *__restrict__ peep,
                       float *__restrict__ y, long ldy, const int32_t *__restrict__ lengths, int T, int B, int H,
                       int reverse, int act, int gate_act, int w_in_smem)
{
    extern __shared__ __align__(16) float smem[];
    const int R = 4 * H;
    float *outp = smem;                          // [H][4]  out_{t-1}, sequence fastest (one 128-bit load per k)
    float *pre = outp + (size_t)H * LSTM_BT;     // [4][R]  pre-activations of this step
    float *Wt = pre + (size_t)LSTM_BT * R;       // [H][R]  sW transposed (k-major), when it fits
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int b0 = blockIdx.x * LSTM_BT;

    if (w_in_smem)
        for (int e = tid; e < R * H; e += nthr) {
            const int row = e / H, k = e - row * H;
            Wt[(size_t)k * R + row] = __ldg(sW + e);
        }
    for (int e = tid; e < H * LSTM_BT; e += nthr) outp[e] = 0.0f;
    // elementwise owner: thread e < H * 4 -> (unit j, sequence b)
    const int ej = tid / LSTM_BT, eb = tid - ej * LSTM_BT;
    const bool owner = tid < H * LSTM_BT;
    float state = 0.0f;
    float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
    int len = 0;
    if (owner) {
        p0 = peep[ej]; p1 = peep[H + ej]; p2 = peep[2 * H + ej];
        const int bg = b0 + eb;
        len = bg < B ? (lengths ? min(lengths[bg], T) : T) : 0;
    }
    __syncthreads();

    for (int s = 0; s < T; s++) {
        const int t = reverse ? T - 1 - s : s;
        // ---- pre-activations: row r of out sW' for the 4 sequences, plus the projection ----
        for (int r = tid; r < R; r += nthr) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            if (w_in_smem) {
#pragma unroll 4
                for (int k = 0; k < H; k++) {
                    const float w = Wt[(size_t)k * R + r];
                    const float4 o = *reinterpret_cast<const float4 *>(outp + k * LSTM_BT);
                    a0 = fmaf(w, o.x, a0); a1 = fmaf(w, o.y, a1); a2 = fmaf(w, o.z, a2); a3 = fmaf(w, o.w, a3);
                }
            } else {
                const float *wr = sW + (size_t)r * H;
#pragma unroll 4
                for (int k = 0; k < H; k++) {
                    const float w = __ldg(wr + k);
                    const float4 o = *reinterpret_cast<const float4 *>(outp + k * LSTM_BT);
                    a0 = fmaf(w, o.x, a0); a1 = fmaf(w, o.y, a1); a2 = fmaf(w, o.z, a2); a3 = fmaf(w, o.w, a3);
                }
            }
            const float acc[4] = {a0, a1, a2, a3};
#pragma unroll
            for (int b = 0; b < LSTM_BT; b++) {
                const float v = b0 + b < B ? __ldg(vW + ((long)t * B + b0 + b) * ldv + r) : 0.0f;
                pre[b * R + r] = acc[b] + v;
            }
        }
        __syncthreads();
        // ---- gates, state and output of (unit ej, sequence eb) ----
        if (owner) {
            const float *sp = pre + eb * R + 4 * ej;
            const float s0 = sp[0], s1 = sp[1], s2 = sp[2], s3 = sp[3];
            float st = state * apply_act(s2 + state * p1, gate_act);
            st += apply_act(s0, act) * apply_act(s1 + state * p0, gate_act);
            float out = apply_act(st, act) * apply_act(s3 + st * p2, gate_act);
            const bool live = t < len;                  // ragged batch: state and output stay 0 outside the read
            state = live ? st : 0.0f;
            out = live ? out : 0.0f;
            outp[ej * LSTM_BT + eb] = out;
            if (b0 + eb < B) y[((long)t * B + b0 + eb) * ldy + ej] = out;
        }
        __syncthreads();
    }
}

__global__ void window_kernel(const float *__restrict__ x, long ldx, float *__restrict__ y, long ldy,
                              const int32_t *__restrict__ lengths, int T, int B, int F, int w)
{
    const long total = (long)T * B * w * F;
    const int half = w / 2;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int c = (int)(e % (w * F));
        const long tb = e / (w * F);
        const int b = (int)(tb % B), t = (int)(tb / B);
        const int k = c / F, f = c - k * F;
        const int tin = t + k - half;
        const int len = lengths ? min(lengths[b], T) : T;
        // each read is padded on its own (the reference feeds one read per call): zeros outside [0, len)
        y[tb * ldy + c] = (tin >= 0 && tin < len && t < len) ? x[((long)tin * B + b) * ldx + f] : 0.0f;
    }
}

}  // namespace sloika

using namespace sloika;

extern "C" int sloika_lstm_recurrence_fwd(const float *vW, long ld_vw, const float *sW, const float *peep, float *y, long ldy,
                                          const int32_t *lengths, int T, int B, int H, int reverse, int act, int gate_act,
                                          void *stream)
{
    if (!vW || !sW || !peep || !y || T < 0 || B <= 0 || H <= 0 || ldy < H || ld_vw < 4L * H) return SLOIKA_ERR_ARG;
    if (!act_known(act) || !act_known(gate_act)) return SLOIKA_ERR_UNSUPPORTED;
    if (H > 256) return SLOIKA_ERR_UNSUPPORTED;
    if (T == 0) return SLOIKA_OK;
    const size_t base = sizeof(float) * ((size_t)H * LSTM_BT + (size_t)LSTM_BT * 4 * H);
    const size_t wbytes = sizeof(float) * (size_t)4 * H * H;
    const int w_in_smem = base + wbytes <= 220 * 1024 ? 1 : 0;
    const size_t smem = base + (w_in_smem ? wbytes : 0);
    cudaError_t err = cudaFuncSetAttribute(lstm_recurrence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    int threads = (int)ceil_div(4 * H, 32) * 32;
    if (threads > 1024) threads = 1024;
    if (threads < H * LSTM_BT) threads = (int)ceil_div(H * LSTM_BT, 32) * 32;      // = 4H: one owner per (unit, sequence)
    lstm_recurrence_kernel<<<(unsigned)ceil_div(B, LSTM_BT), threads, smem, (cudaStream_t)stream>>>(
        vW, ld_vw, sW, peep, y, ldy, lengths, T, B, H, reverse, act, gate_act, w_in_smem);
    SLOIKA_RETURN_LAUNCH_STATUS();
}

extern "C" int sloika_window_fwd(const float *x, long ldx, float *y, long ldy, const int32_t *lengths, int T, int B, int F,
                                 int w, void *stream)
{
    if (!x || !y || T < 0 || B <= 0 || F <= 0 || w <= 0 || (w & 1) == 0 || ldx < F || ldy < (long)w * F) return SLOIKA_ERR_ARG;
    if (T == 0) return SLOIKA_OK;
    const long total = (long)T * B * w * F;
    long blocks = ceil_div(total, 256);
    if (blocks > 148L * 16) blocks = 148L * 16;
    window_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, lengths, T, B, F, w);
    SLOIKA_RETURN_LAUNCH_STATUS();
}
