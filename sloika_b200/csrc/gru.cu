// GRU recurrence as one persistent kernel per layer.
//
// Reference semantics: Gru.step (sloika/layers.py:1010-1021) scanned by RNN.run (:85-88) from h0 = 0;
// Reverse(Gru) (:1449-1450) is the same scan walked from each sequence's last valid step down to 0.
//     vI = x_t iW' + b            (hoisted: one GEMM over all steps, see linear.cu / gemm path)
//     vS = h sW'                  z = gate(vI_z + vS_z)     r = gate(vI_r + vS_r)
//     y  = (r*h) sW2'             hbar = act(vI_c + y)      h' = z*h + (1-z)*hbar
// Theano runs this as T sequential scan dispatches of two tiny dependent GEMMs plus elementwise ops.
// Here a CTA owns BT = 8 sequences and ALL of sW | sW2 for the whole scan (batch-partitioned, so the
// two dependent reductions of a step need only __syncthreads, never a grid barrier):
//
//   * sW (2H x H) lives in shared memory, sW2 (H x H) in shared memory or -- when H is too large for
//     both to fit in 227 KB -- (partly) in registers, for all T steps;
//   * thread (jg, ks) accumulates rows {z_j0, z_j1, r_j0, r_j1} (j0 = 2jg) x 8 sequences over the
//     k-slice ks with packed FFMA2 (pairs along k), then the S k-slices are combined by a shuffle
//     reduce-scatter that leaves every lane with the finished pre-activations of its own (j, b) pairs,
//     so the gate math, r*h, the blend and the h' store are spread over all threads;
//   * h and r*h are exchanged through two small shared arrays; two __syncthreads per step;
//   * the vI loads of a step are issued before its first reduction (consumed ~1k cycles later) and
//     step t+6 is prefetched into L2.
//
// Latency-bound by design (T dependent steps); per step a CTA issues 6*H^2*BT FMAs.
#include <cstdlib>
#include "common.cuh"

namespace sloika {

constexpr int GRU_BT = 8;   // sequences per CTA

template <int HP, int S>
struct GruCfg {
    static constexpr int NJG = HP / 2;             // row-pair groups
    static constexpr int THREADS = NJG * S;
    static constexpr int NG = HP / (4 * S);        // float4 granules per k-slice
    static constexpr int NJ = (S == 8) ? 1 : 2;    // rows owned per lane after the reduce-scatter
    static_assert(S == 4 || S == 8, "S must be 4 or 8");
    static_assert(HP % (4 * S) == 0, "HP must be a multiple of 4*S");
};

__device__ __forceinline__ float2 lo2(const float4 &v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4 &v) { return make_float2(v.z, v.w); }

// Halve the sequence range held per lane: lanes whose `bit` is set keep the upper half.
template <int R, int NB>
__device__ __forceinline__ void scatter_b(float (&v)[R][NB], float (&w)[R][NB / 2], bool upper, int mask)
{
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int b = 0; b < NB / 2; b++) {
            const float mine = upper ? v[r][NB / 2 + b] : v[r][b];
            const float send = upper ? v[r][b] : v[r][NB / 2 + b];
            w[r][b] = mine + __shfl_xor_sync(0xffffffffu, send, mask);
        }
}

// Halve the row set held per lane (S == 8 only): lanes with jsel keep the odd rows of each pair.
template <int R, int NB>
__device__ __forceinline__ void scatter_j(float (&v)[R][NB], float (&w)[R / 2][NB], bool odd, int mask)
{
#pragma unroll
    for (int r = 0; r < R / 2; r++)
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const float mine = odd ? v[2 * r + 1][b] : v[2 * r][b];
            const float send = odd ? v[2 * r][b] : v[2 * r + 1][b];
            w[r][b] = mine + __shfl_xor_sync(0xffffffffu, send, mask);
        }
}

// Reduce R row-sums x 8 sequences across the S lanes of a k-slice group.
// Rows are ordered (pair p, j parity): index 2*p + parity.  Result: out[p*NJ + jj][e] for the lane's
// own rows (jj) and its two sequences b = 2*bp + e.
template <int S, int R>
__device__ __forceinline__ void reduce_scatter(float (&v)[R][GRU_BT], float (&out)[(S == 8) ? R / 2 : R][2], int ks)
{
    if constexpr (S == 8) {
        float a[R / 2][GRU_BT];
        scatter_j<R, GRU_BT>(v, a, (ks & 4) != 0, 4);
        float b4[R / 2][4];
        scatter_b<R / 2, GRU_BT>(a, b4, (ks & 2) != 0, 2);
        scatter_b<R / 2, 4>(b4, out, (ks & 1) != 0, 1);
    } else {
        float b4[R][4];
        scatter_b<R, GRU_BT>(v, b4, (ks & 2) != 0, 2);
        scatter_b<R, 4>(b4, out, (ks & 1) != 0, 1);
    }
}

template <int HP, int S, int W2MODE>
__global__ void __launch_bounds__(GruCfg<HP, S>::THREADS, 1)
gru_recurrence_kernel(const float *__restrict__ vI, long ldv, const float *__restrict__ sW, const float *__restrict__ sW2,
                      float *__restrict__ y, long ldy, const int32_t *__restrict__ lengths, int T, int B, int H,
                      int reverse, int act, int gate_act)
{
    using Cfg = GruCfg<HP, S>;
    constexpr int NG = Cfg::NG, NJ = Cfg::NJ, HP4 = HP / 4;
    extern __shared__ __align__(16) float smem[];
    float *W1 = smem;                                   // [2*HP][HP]  rows: z_0..z_HP-1, r_0..r_HP-1
    float *W2 = W1 + 2 * HP * HP;                       // W2MODE 0: [HP][HP]; 2: odd rows [HP/2][HP]; 1: absent
    float *hs = W2 + (W2MODE == 0 ? HP * HP : (W2MODE == 2 ? HP * HP / 2 : 0));   // [BT][HP]  h_{t-1}
    float *rh = hs + GRU_BT * HP;                       // [BT][HP]    r * h_{t-1}

    const int tid = threadIdx.x;
    const int ks = tid % S, jg = tid / S;
    const int j0 = 2 * jg;
    const int b_base = blockIdx.x * GRU_BT;

    // ---- one-time: stage the recurrent weights (zero padded to HP) and clear the state ----
    for (int e = tid; e < 2 * HP * HP; e += Cfg::THREADS) {
        const int row = e / HP, k = e - row * HP;
        const int gate = row / HP, j = row - gate * HP;
        W1[e] = (j < H && k < H) ? __ldg(sW + ((long)gate * H + j) * H + k) : 0.0f;
    }
    if constexpr (W2MODE == 0) {
        for (int e = tid; e < HP * HP; e += Cfg::THREADS) {
            const int j = e / HP, k = e - j * HP;
            W2[e] = (j < H && k < H) ? __ldg(sW2 + (long)j * H + k) : 0.0f;
        }
    } else if constexpr (W2MODE == 2) {
        for (int e = tid; e < HP * HP / 2; e += Cfg::THREADS) {
            const int j = 2 * (e / HP) + 1, k = e % HP;
            W2[e] = (j < H && k < H) ? __ldg(sW2 + (long)j * H + k) : 0.0f;
        }
    }
    for (int e = tid; e < 2 * GRU_BT * HP; e += Cfg::THREADS) hs[e] = 0.0f;   // hs and rh are adjacent

    constexpr int W2ROWS = (W2MODE == 1) ? 2 : 1;       // rows of the pair kept in registers
    float4 w2r[W2ROWS][(W2MODE != 0) ? NG : 1];
    if constexpr (W2MODE != 0) {
#pragma unroll
        for (int p = 0; p < W2ROWS; p++)
#pragma unroll
            for (int g = 0; g < NG; g++) {
                const int j = j0 + p, k = 4 * (g * S + ks);
                float t4[4];
#pragma unroll
                for (int c = 0; c < 4; c++)
                    t4[c] = (j < H && k + c < H) ? __ldg(sW2 + (long)j * H + k + c) : 0.0f;
                w2r[p][g] = make_float4(t4[0], t4[1], t4[2], t4[3]);
            }
    }

    // ---- ownership after the reduce-scatter: rows jown[0..NJ), sequences bl0, bl0+1 ----
    const int bp = ks & 3;
    const int bl0 = 2 * bp;
    int jown[NJ];
    if constexpr (S == 8) jown[0] = j0 + (ks >> 2);
    else { jown[0] = j0; jown[1] = j0 + 1; }
    int len[2];
    bool bok[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const int bgl = b_base + bl0 + e;
        bok[e] = bgl < B;
        len[e] = bok[e] ? (lengths ? min(lengths[bgl], T) : T) : 0;
    }

    auto vi_ptr = [&](int t, int e, int jj, int gate) -> const float * {
        return vI + ((long)t * B + (b_base + bl0 + e)) * ldv + (long)gate * H + jown[jj];
    };
    // vI registers for the current step: [jj][e][gate]
    float vcur[NJ][2][3];
    auto load_vi = [&](int t, float (&dst)[NJ][2][3]) {
#pragma unroll
        for (int jj = 0; jj < NJ; jj++)
#pragma unroll
            for (int e = 0; e < 2; e++)
#pragma unroll
                for (int gte = 0; gte < 3; gte++)
                    dst[jj][e][gte] = (bok[e] && jown[jj] < H && t >= 0 && t < T) ? __ldg(vi_ptr(t, e, jj, gte)) : 0.0f;
    };
    const int tstep = reverse ? -1 : 1;
    int t = reverse ? T - 1 : 0;
    __syncthreads();

    const float4 *W1v = reinterpret_cast<const float4 *>(W1);
    const float4 *W2v = reinterpret_cast<const float4 *>(W2);
    const float4 *hsv = reinterpret_cast<const float4 *>(hs);
    const float4 *rhv = reinterpret_cast<const float4 *>(rh);

    for (int s = 0; s < T; s++, t += tstep) {
        // vI of this step is consumed only after phase 1 (~1k cycles away); the loads are issued now
        // and step t+6 is pulled towards L2 so that they hit there.
        load_vi(t, vcur);
        {
            const int tp = t + 6 * tstep;
            if (tp >= 0 && tp < T && jown[0] < H) {
#pragma unroll
                for (int e = 0; e < 2; e++)
                    if (bok[e]) {
#pragma unroll
                        for (int gte = 0; gte < 3; gte++)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(vi_ptr(tp, e, 0, gte)));
                        if constexpr (NJ == 2) {
#pragma unroll
                            for (int gte = 0; gte < 3; gte++)
                                if (jown[1] < H) asm volatile("prefetch.global.L2 [%0];" ::"l"(vi_ptr(tp, e, 1, gte)));
                        }
                    }
            }
        }

        // ---------------- phase 1: vS = h sW'  (rows z_j0 z_j1 r_j0 r_j1) ----------------
        float2 acc[4][GRU_BT];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int b = 0; b < GRU_BT; b++) acc[r][b] = make_float2(0.f, 0.f);
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const int kq = g * S + ks;
            const float4 w0 = W1v[(j0)*HP4 + kq];
            const float4 w1 = W1v[(j0 + 1) * HP4 + kq];
            const float4 w2 = W1v[(HP + j0) * HP4 + kq];
            const float4 w3 = W1v[(HP + j0 + 1) * HP4 + kq];
#pragma unroll
            for (int b = 0; b < GRU_BT; b++) {
                const float4 hv = hsv[b * HP4 + kq];
                const float2 hl = lo2(hv), hh = hi2(hv);
                acc[0][b] = fma2(lo2(w0), hl, acc[0][b]); acc[0][b] = fma2(hi2(w0), hh, acc[0][b]);
                acc[1][b] = fma2(lo2(w1), hl, acc[1][b]); acc[1][b] = fma2(hi2(w1), hh, acc[1][b]);
                acc[2][b] = fma2(lo2(w2), hl, acc[2][b]); acc[2][b] = fma2(hi2(w2), hh, acc[2][b]);
                acc[3][b] = fma2(lo2(w3), hl, acc[3][b]); acc[3][b] = fma2(hi2(w3), hh, acc[3][b]);
            }
        }
        float part[4][GRU_BT];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int b = 0; b < GRU_BT; b++) part[r][b] = acc[r][b].x + acc[r][b].y;
        float red1[2 * NJ][2];                  // [gate*NJ + jj][e]
        reduce_scatter<S, 4>(part, red1, ks);

        float zg[NJ][2], hold[NJ][2];
#pragma unroll
        for (int jj = 0; jj < NJ; jj++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int j = jown[jj];
                const float z = apply_act(red1[0 * NJ + jj][e] + vcur[jj][e][0], gate_act);
                const float r = apply_act(red1[1 * NJ + jj][e] + vcur[jj][e][1], gate_act);
                const float h = (j < HP) ? hs[(bl0 + e) * HP + j] : 0.0f;
                zg[jj][e] = z;
                hold[jj][e] = h;
                if (j < H) rh[(bl0 + e) * HP + j] = r * h;
            }
        __syncthreads();

        // ---------------- phase 2: y = (r*h) sW2'  (rows c_j0 c_j1) ----------------
        float2 acc2[2][GRU_BT];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int b = 0; b < GRU_BT; b++) acc2[r][b] = make_float2(0.f, 0.f);
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const int kq = g * S + ks;
            float4 w0, w1;
            if constexpr (W2MODE == 1) { w0 = w2r[0][g]; w1 = w2r[1][g]; }
            else if constexpr (W2MODE == 2) { w0 = w2r[0][g]; w1 = W2v[jg * HP4 + kq]; }
            else { w0 = W2v[(j0)*HP4 + kq]; w1 = W2v[(j0 + 1) * HP4 + kq]; }
#pragma unroll
            for (int b = 0; b < GRU_BT; b++) {
                const float4 hv = rhv[b * HP4 + kq];
                const float2 hl = lo2(hv), hh = hi2(hv);
                acc2[0][b] = fma2(lo2(w0), hl, acc2[0][b]); acc2[0][b] = fma2(hi2(w0), hh, acc2[0][b]);
                acc2[1][b] = fma2(lo2(w1), hl, acc2[1][b]); acc2[1][b] = fma2(hi2(w1), hh, acc2[1][b]);
            }
        }
        float part2[2][GRU_BT];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int b = 0; b < GRU_BT; b++) part2[r][b] = acc2[r][b].x + acc2[r][b].y;
        float red2[NJ][2];
        reduce_scatter<S, 2>(part2, red2, ks);

#pragma unroll
        for (int jj = 0; jj < NJ; jj++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int j = jown[jj];
                const float hbar = apply_act(red2[jj][e] + vcur[jj][e][2], act);
                const float z = zg[jj][e];
                float hn = z * hold[jj][e] + (1.0f - z) * hbar;
                hn = (t < len[e]) ? hn : 0.0f;          // ragged batch: state stays 0 outside the read
                if (j < H) {
                    hs[(bl0 + e) * HP + j] = hn;
                    if (bok[e]) y[((long)t * B + (b_base + bl0 + e)) * ldy + j] = hn;
                }
            }
        __syncthreads();
    }
}

template <int HP, int S, int W2MODE>
static int launch_gru(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths,
                      int T, int B, int H, int reverse, int act, int gate_act, cudaStream_t st)
{
    using Cfg = GruCfg<HP, S>;
    const size_t w2 = W2MODE == 0 ? (size_t)HP * HP : (W2MODE == 2 ? (size_t)HP * HP / 2 : 0);
    const size_t smem = sizeof(float) * ((size_t)2 * HP * HP + w2 + 2 * GRU_BT * HP);
    auto kern = gru_recurrence_kernel<HP, S, W2MODE>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    const unsigned grid = (unsigned)ceil_div(B, GRU_BT);
    kern<<<grid, Cfg::THREADS, smem, st>>>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act);
    SLOIKA_RETURN_LAUNCH_STATUS();
}

namespace gru4 {
int dispatch(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths, int T,
             int B, int H, int reverse, int act, int gate_act, cudaStream_t st);
int dispatch_cluster(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths,
                     int T, int B, int H, int reverse, int act, int gate_act, cudaStream_t st);
}
namespace gru5 {
int dispatch(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths, int T,
             int B, int H, int reverse, int act, int gate_act, long seqs_in_flight, cudaStream_t st);
}
namespace gru3 {
int dispatch(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths, int T, int B,
             int H, int reverse, int act, int gate_act, cudaStream_t st);
}

// ---------------------------------------------------------------------------------------------------
// Hidden sizes beyond what one SM can keep on chip (H > 144): the scan is run step by step -- two GEMMs
// (h sW', (r*h) sW2': the tensor-core kernel when the batch is large enough, else the SIMT one) and two small
// gate kernels per step, all enqueued on the caller's stream.  Same arithmetic as the persistent kernels
// (Gru.step, sloika/layers.py:1010-1021); far slower (four launches per step), there so that the operator has
// no size limit.  No shipped raw model is this wide.
__global__ void gru_step_gates_kernel(const float *__restrict__ vI_t, long ldv, float *__restrict__ vS /* [B][2H] */,
                                      const float *__restrict__ hprev, long ldh, float *__restrict__ rh /* [B][H] */,
                                      int B, int H, int act_gate, bool first)
{
    const long n = (long)B * H;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const int b = (int)(e / H), j = (int)(e - (long)b * H);
        const float sz = first ? 0.0f : vS[(long)b * 2 * H + j], sr = first ? 0.0f : vS[(long)b * 2 * H + H + j];
        const float z = apply_act(vI_t[(long)b * ldv + j] + sz, act_gate);
        const float r = apply_act(vI_t[(long)b * ldv + H + j] + sr, act_gate);
        const float h = first ? 0.0f : hprev[(long)b * ldh + j];
        vS[(long)b * 2 * H + j] = z;                       // z parked where its pre-activation was
        rh[(long)b * H + j] = r * h;
    }
}

__global__ void gru_step_blend_kernel(const float *__restrict__ vI_t, long ldv, const float *__restrict__ vS,
                                      const float *__restrict__ y2 /* [B][H] */, const float *__restrict__ hprev,
                                      long ldh, float *__restrict__ y_t, long ldy, const int32_t *__restrict__ lengths,
                                      int t, int T, int B, int H, int act, bool first)
{
    const long n = (long)B * H;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const int b = (int)(e / H), j = (int)(e - (long)b * H);
        const float z = vS[(long)b * 2 * H + j];
        const float hbar = apply_act(vI_t[(long)b * ldv + 2 * H + j] + (first ? 0.0f : y2[(long)b * H + j]), act);
        const float h = first ? 0.0f : hprev[(long)b * ldh + j];
        float hn = z * h + (1.0f - z) * hbar;
        const int len = lengths ? min(lengths[b], T) : T;
        y_t[(long)b * ldy + j] = t < len ? hn : 0.0f;       // state stays 0 outside the read
    }
}

}  // namespace sloika

using namespace sloika;

extern "C" int sloika_linear_fwd_ex(const float *x, long ldx, const float *W, const float *bias, float *y, long ldy,
                                    long M, int K, int N, int act, int algo, void *stream);

static int gru_stepwise(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy,
                        const int32_t *lengths, int T, int B, int H, int reverse, int act, int gate_act, cudaStream_t st)
{
    float *ws = nullptr;                                         // vS [B][2H] | rh [B][H] | y2 [B][H], stream ordered
    const size_t nfl = (size_t)B * 4 * H;
    cudaError_t err = cudaMallocAsync(reinterpret_cast<void **>(&ws), nfl * sizeof(float), st);
    if (err != cudaSuccess) return (int)err;
    float *vS = ws, *rh = ws + (size_t)B * 2 * H, *y2 = rh + (size_t)B * H;
    const long n = (long)B * H;
    const unsigned blocks = (unsigned)(ceil_div(n, 256) > 148L * 16 ? 148L * 16 : ceil_div(n, 256));
    int rc = SLOIKA_OK;
    for (int s = 0; s < T && rc == SLOIKA_OK; s++) {
        const int t = reverse ? T - 1 - s : s;
        const int tp = reverse ? t + 1 : t - 1;                  // the step whose output is h_{t-1} of the scan
        const bool first = s == 0;
        const float *vI_t = vI + (long)t * B * ldv;
        const float *hprev = first ? nullptr : y + (long)tp * B * ldy;
        float *y_t = y + (long)t * B * ldy;
        if (!first) rc = sloika_linear_fwd_ex(hprev, ldy, sW, nullptr, vS, 2L * H, B, H, 2 * H, SLOIKA_ACT_LINEAR, SLOIKA_GEMM_AUTO, st);
        if (rc != SLOIKA_OK) break;
        gru_step_gates_kernel<<<blocks, 256, 0, st>>>(vI_t, ldv, vS, hprev, ldy, rh, B, H, gate_act, first);
        if (!first) rc = sloika_linear_fwd_ex(rh, H, sW2, nullptr, y2, H, B, H, H, SLOIKA_ACT_LINEAR, SLOIKA_GEMM_AUTO, st);
        if (rc != SLOIKA_OK) break;
        gru_step_blend_kernel<<<blocks, 256, 0, st>>>(vI_t, ldv, vS, y2, hprev, ldy, y_t, ldy, lengths, t, T, B, H, act, first);
    }
    cudaFreeAsync(ws, st);
    if (rc != SLOIKA_OK) return rc;
    SLOIKA_RETURN_LAUNCH_STATUS();
}

extern "C" int sloika_gru_recurrence_fwd(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy,
                                         const int32_t *lengths, int T, int B, int H, int reverse, int act,
                                         int gate_act, void *stream)
{
    return sloika_gru_recurrence_fwd_ex(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, 0, stream);
}

extern "C" int sloika_gru_recurrence_fwd_ex(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy,
                                            const int32_t *lengths, int T, int B, int H, int reverse, int act,
                                            int gate_act, long seqs_in_flight, void *stream)
{
    if (!vI || !sW || !sW2 || !y || T < 0 || B <= 0 || H <= 0 || ldy < H || ldv < 3L * H) return SLOIKA_ERR_ARG;
    if (!act_known(act) || !act_known(gate_act)) return SLOIKA_ERR_UNSUPPORTED;
    if (T == 0) return SLOIKA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    {   // kernel generations, newest first (tanh / sigmoid): tcgen05 kernel with the weights in tensor memory
        // (gru_tc.cu, H <= 128), fp16x3 mma.sync kernel (gru_h16.cu, H <= 144), 3xTF32 mma.sync kernel (gru_mma.cu);
        // then the FFMA2 kernel of this file (any activation pair).
        // SLOIKA_B200_GRU=v1|v3|v4 caps the choice (A/B measurements, tests).
        const char *sel = getenv("SLOIKA_B200_GRU");
        const int cap = (sel && sel[0] == 'v' && sel[1] >= '1' && sel[1] <= '5') ? sel[1] - '0' : 5;
        if (cap >= 5) {
            const int rc = gru5::dispatch(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, seqs_in_flight, st);
            if (rc != SLOIKA_ERR_UNSUPPORTED) return rc;
        }
        if (cap >= 4) {
            int rc = gru4::dispatch(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, st);
            if (rc != SLOIKA_ERR_UNSUPPORTED) return rc;
            if (!getenv("SLOIKA_B200_GRU_STEPWISE")) {          // 144 < H <= 256: weights spread over a 4-CTA cluster
                rc = gru4::dispatch_cluster(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, st);
                if (rc != SLOIKA_ERR_UNSUPPORTED) return rc;
            }
        }
        if (cap >= 3) {
            const int rc = gru3::dispatch(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, st);
            if (rc != SLOIKA_ERR_UNSUPPORTED) return rc;
        }
    }
#define GRU_CASE(HP_, S_, W2R_) \
    return launch_gru<HP_, S_, W2R_>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, st)
    if (H <= 16) GRU_CASE(16, 4, 0);
    if (H <= 32) GRU_CASE(32, 8, 0);
    if (H <= 48) GRU_CASE(48, 4, 0);
    if (H <= 64) GRU_CASE(64, 8, 0);
    if (H <= 80) GRU_CASE(80, 4, 0);
    if (H <= 96) GRU_CASE(96, 8, 0);
    if (H <= 112) GRU_CASE(112, 4, 0);
    if (H <= 128) GRU_CASE(128, 4, 1);
    if (H <= 144) GRU_CASE(144, 4, 2);
#undef GRU_CASE
    return gru_stepwise(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, st);
}

extern "C" size_t sloika_gru_workspace_bytes(int T, int B, int H)
{
    if (T < 0 || B < 0 || H < 0) return 0;
    return sizeof(float) * (size_t)T * (size_t)B * 3 * (size_t)H;
}

extern "C" int sloika_gru_fwd(const float *x, long ldx, const float *iW, const float *sW, const float *sW2,
                              const float *b, float *y, long ldy, void *ws, size_t ws_bytes, const int32_t *lengths,
                              int T, int B, int I, int H, int reverse, int act, int gate_act, void *stream)
{
    if (!x || !iW || !sW || !sW2 || !b || !y || T < 0 || B <= 0 || I <= 0 || H <= 0) return SLOIKA_ERR_ARG;
    if (T == 0) return SLOIKA_OK;
    if (!ws || ws_bytes < sloika_gru_workspace_bytes(T, B, H)) return SLOIKA_ERR_WORKSPACE;
    float *vI = static_cast<float *>(ws);
    int rc = sloika_linear_fwd(x, ldx, iW, b, vI, 3L * H, (long)T * B, I, 3 * H, SLOIKA_ACT_LINEAR, stream);
    if (rc != SLOIKA_OK) return rc;
    return sloika_gru_recurrence_fwd(vI, 3L * H, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, stream);
}
