// GRU recurrence, second generation: recurrent weights resident in REGISTERS for the whole scan.
//
// Same semantics and CTA decomposition as gru.cu (one persistent CTA per 8 sequences, all of sW | sW2
// on chip, two __syncthreads per step; reference sloika/layers.py:1010-1021, :85-88, :1449-1450).
// What changed, from the ncu profile of the first kernel (smem wavefronts 58 %, mio/short_sb stalls,
// fma pipe 36 %): the shared-memory pipe, not the FMA pipe, was the limiter, because every thread
// re-read its slice of sW | sW2 from shared memory every step and re-read h once per 4 rows.
//
//   * 256 threads = 32 row groups x 8 k-slices.  Thread (jg, ks) owns rows j = RPT*jg .. +RPT-1 of
//     each gate (RPT = HP/32) and the k-slice {4*(8g + ks) .. +3 : g < NG}; its RPT*3 x KS weights
//     are loaded ONCE into registers (H <= 96), or sW stays in shared memory and only sW2 is
//     register resident (H = 128).
//   * per step a thread issues only the h loads (8 sequences x NG LDS.128 per phase) and
//     2*RPT*8*KS/2 + RPT*8*KS/2 FFMA2; sequences are processed in two halves so the accumulators
//     stay at 2*RPT*4 float2.
//   * the 8 k-slice partials are combined with a shuffle reduce-scatter over the SEQUENCE index, so
//     lane ks ends up owning sequence b = ks for its 3*RPT rows: gates, r*h, blend and state stay in
//     that lane's registers.
//   * vI of step t+1 is brought into shared memory with cp.async during step t (coalesced 16-byte
//     copies, no register staging); h_t is written to HBM with coalesced 128-bit stores from the
//     shared state buffer.
#include <cstdlib>
#include "common.cuh"

namespace sloika {
namespace gru2 {

constexpr int BT = 8;        // sequences per CTA
constexpr int S = 8;         // k-slices (lanes that share a row group)

__device__ __forceinline__ float2 lo2(const float4 &v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4 &v) { return make_float2(v.z, v.w); }

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Reduce R rows x 8 sequences over the 8 k-slice lanes; lane ks keeps sequence b = ks.
template <int R>
__device__ __forceinline__ void reduce_to_own_sequence(float (&v)[R][BT], float (&out)[R], int ks)
{
    float a[R][4];
    const bool up4 = (ks & 4) != 0;
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const float mine = up4 ? v[r][4 + b] : v[r][b];
            const float send = up4 ? v[r][b] : v[r][4 + b];
            a[r][b] = mine + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    float c[R][2];
    const bool up2 = (ks & 2) != 0;
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const float mine = up2 ? a[r][2 + b] : a[r][b];
            const float send = up2 ? a[r][b] : a[r][2 + b];
            c[r][b] = mine + __shfl_xor_sync(0xffffffffu, send, 2);
        }
    const bool up1 = (ks & 1) != 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const float mine = up1 ? c[r][1] : c[r][0];
        const float send = up1 ? c[r][0] : c[r][1];
        out[r] = mine + __shfl_xor_sync(0xffffffffu, send, 1);
    }
}

template <int HP, int THREADS>
struct Cfg {
    static_assert(HP % 32 == 0, "HP must be a multiple of 32");
    static_assert(HP % (THREADS / S) == 0, "row groups must divide HP");
    static constexpr int RPT = HP / (THREADS / S);   // rows per gate per thread
    static constexpr int NG = HP / 32;           // float4 granules per k-slice (KS = 4*NG = HP/8)
    static constexpr int VLD = 3 * HP + 4;       // smem row pitch of the staged vI (floats, 16-byte multiple)
};

// W1REG: sW (2H x H) register resident (else read from shared memory every step); sW2 always in registers.
template <int HP, int THREADS, bool W1REG, int ACT, int GATE>
__global__ void __launch_bounds__(THREADS, 1)
gru_recurrence_v2_kernel(const float *__restrict__ vI, const float *__restrict__ sW, const float *__restrict__ sW2,
                         float *__restrict__ y, long ldy, const int32_t *__restrict__ lengths, int T, int B, int H,
                         int reverse, int dbg)
{
    using C = Cfg<HP, THREADS>;
    constexpr int RPT = C::RPT, NG = C::NG, HP4 = HP / 4, VLD = C::VLD;
    extern __shared__ __align__(16) float smem[];
    float *hs = smem;                            // [BT][HP]   h_{t-1}
    float *rh = hs + BT * HP;                    // [BT][HP]   r * h_{t-1}
    float *vbuf = rh + BT * HP;                  // [2][BT][VLD] staged vI (double buffered)
    float *W1 = vbuf + 2 * BT * VLD;             // [2*HP][HP] only when !W1REG

    const int tid = threadIdx.x;
    const int ks = tid & 7, jg = tid >> 3;
    const int j0 = RPT * jg;
    const int b_base = blockIdx.x * BT;
    const int b_own = ks;                        // sequence owned after the reduce-scatter
    const int bg_own = b_base + b_own;
    const bool b_ok = bg_own < B;
    const int len_own = b_ok ? (lengths ? min(lengths[bg_own], T) : T) : 0;

    // ---- weights -> registers (zero padded), once ----
    float4 w1[W1REG ? 2 * RPT : 1][W1REG ? NG : 1];
    float4 w2[RPT][NG];
    auto load_w4 = [&](const float *Wm, int row, int k) -> float4 {
        float t4[4];
#pragma unroll
        for (int c = 0; c < 4; c++) t4[c] = (row >= 0 && k + c < H) ? __ldg(Wm + (long)row * H + k + c) : 0.0f;
        return make_float4(t4[0], t4[1], t4[2], t4[3]);
    };
#pragma unroll
    for (int g = 0; g < NG; g++) {
        const int k = 4 * (g * S + ks);
#pragma unroll
        for (int i = 0; i < RPT; i++) {
            const int j = j0 + i;
            if constexpr (W1REG) {
                w1[i][g] = load_w4(sW, j < H ? j : -1, k);                  // z rows
                w1[RPT + i][g] = load_w4(sW, j < H ? H + j : -1, k);        // r rows
            }
            w2[i][g] = load_w4(sW2, j < H ? j : -1, k);
        }
    }
    if constexpr (!W1REG) {
        for (int e = tid; e < 2 * HP * HP; e += THREADS) {
            const int row = e / HP, k = e - row * HP;
            const int gate = row / HP, j = row - gate * HP;
            W1[e] = (j < H && k < H) ? __ldg(sW + ((long)gate * H + j) * H + k) : 0.0f;
        }
    }
    for (int e = tid; e < 2 * BT * HP; e += THREADS) hs[e] = 0.0f;        // hs and rh
    for (int e = tid; e < 2 * BT * VLD; e += THREADS) vbuf[e] = 0.0f;

    // ---- cooperative staging of vI[t] (BT rows of 3H floats) into vbuf[slot] ----
    const long H3 = 3L * H;
    const bool vec_vi = ((H3 & 3) == 0) && (((uintptr_t)vI & 15) == 0);
    auto stage_vi = [&](int t, int slot) {
        if (t < 0 || t >= T) return;
        float *dst = vbuf + slot * BT * VLD;
        if (vec_vi) {
            const int n4 = (int)(H3 >> 2);
            for (int e = tid; e < BT * n4; e += THREADS) {
                const int b = e / n4, c = e - b * n4;
                if (b_base + b < B) cp_async16(dst + b * VLD + 4 * c, vI + ((long)t * B + b_base + b) * H3 + 4 * c);
            }
        } else {
            for (int e = tid; e < BT * (int)H3; e += THREADS) {
                const int b = e / (int)H3, c = e - b * (int)H3;
                if (b_base + b < B) dst[b * VLD + c] = __ldg(vI + ((long)t * B + b_base + b) * H3 + c);
            }
        }
    };
    const int tstep = reverse ? -1 : 1;
    int t = reverse ? T - 1 : 0;
    __syncthreads();
    stage_vi(t, 0);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();

    const float4 *hsv = reinterpret_cast<const float4 *>(hs);
    const float4 *rhv = reinterpret_cast<const float4 *>(rh);
    const float4 *W1v = reinterpret_cast<const float4 *>(W1);
    const bool vec_y = ((ldy & 3) == 0) && (((uintptr_t)y & 15) == 0) && (HP == H);

    float h_own[RPT];                            // state of (rows j0.., sequence b_own), kept in registers
#pragma unroll
    for (int i = 0; i < RPT; i++) h_own[i] = 0.0f;

    for (int s = 0; s < T; s++, t += tstep) {
        const int slot = s & 1;
        if (!(dbg & 16)) stage_vi(t + tstep, slot ^ 1);           // lands during this step; consumed next step
        cp_async_commit();

        // ---------------- phase 1: vS = h sW'  (rows z_j, r_j : 2*RPT rows) ----------------
        float part[2 * RPT][BT];
#pragma unroll
        for (int half = 0; half < 2; half++) {
            float2 acc[2 * RPT][4];
            if (dbg & 1) { for (int r = 0; r < 2 * RPT; r++) for (int b = 0; b < 4; b++) part[r][4 * half + b] = hs[r + b + ks]; continue; }
#pragma unroll
            for (int r = 0; r < 2 * RPT; r++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[r][b] = make_float2(0.f, 0.f);
#pragma unroll
            for (int g = 0; g < NG; g++) {
                const int kq = g * S + ks;
                float4 wz[RPT], wr[RPT];
#pragma unroll
                for (int i = 0; i < RPT; i++) {
                    if constexpr (W1REG) { wz[i] = w1[i][g]; wr[i] = w1[RPT + i][g]; }
                    else { wz[i] = W1v[(j0 + i) * HP4 + kq]; wr[i] = W1v[(HP + j0 + i) * HP4 + kq]; }
                }
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const float4 hv = hsv[(4 * half + b) * HP4 + kq];
                    const float2 hl = lo2(hv), hh = hi2(hv);
#pragma unroll
                    for (int i = 0; i < RPT; i++) {
                        acc[i][b] = fma2(lo2(wz[i]), hl, acc[i][b]);
                        acc[i][b] = fma2(hi2(wz[i]), hh, acc[i][b]);
                        acc[RPT + i][b] = fma2(lo2(wr[i]), hl, acc[RPT + i][b]);
                        acc[RPT + i][b] = fma2(hi2(wr[i]), hh, acc[RPT + i][b]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 2 * RPT; r++)
#pragma unroll
                for (int b = 0; b < 4; b++) part[r][4 * half + b] = acc[r][b].x + acc[r][b].y;
        }
        float pre1[2 * RPT];
        if (dbg & 4) { for (int r = 0; r < 2 * RPT; r++) pre1[r] = part[r][0] + part[r][7]; }
        else reduce_to_own_sequence<2 * RPT>(part, pre1, ks);

        const float *vrow = vbuf + slot * BT * VLD + b_own * VLD;
        float zg[RPT];
#pragma unroll
        for (int i = 0; i < RPT; i++) {
            const int j = j0 + i;
            const float z = (dbg & 8) ? 0.5f + 0.001f * (pre1[i] + vrow[j]) : apply_act_fast<GATE>(pre1[i] + vrow[j]);
            const float r = (dbg & 8) ? 0.5f + 0.001f * (pre1[RPT + i] + vrow[H + j]) : apply_act_fast<GATE>(pre1[RPT + i] + vrow[H + j]);
            zg[i] = z;
            if (j < H) rh[b_own * HP + j] = r * h_own[i];
        }
        __syncthreads();

        // ---------------- phase 2: y = (r*h) sW2'  (rows c_j : RPT rows) ----------------
        float part2[RPT][BT];
#pragma unroll
        for (int half = 0; half < 2; half++) {
            float2 acc[RPT][4];
            if (dbg & 2) { for (int r = 0; r < RPT; r++) for (int b = 0; b < 4; b++) part2[r][4 * half + b] = rh[r + b + ks]; continue; }
#pragma unroll
            for (int r = 0; r < RPT; r++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[r][b] = make_float2(0.f, 0.f);
#pragma unroll
            for (int g = 0; g < NG; g++) {
                const int kq = g * S + ks;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const float4 hv = rhv[(4 * half + b) * HP4 + kq];
                    const float2 hl = lo2(hv), hh = hi2(hv);
#pragma unroll
                    for (int i = 0; i < RPT; i++) {
                        acc[i][b] = fma2(lo2(w2[i][g]), hl, acc[i][b]);
                        acc[i][b] = fma2(hi2(w2[i][g]), hh, acc[i][b]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RPT; r++)
#pragma unroll
                for (int b = 0; b < 4; b++) part2[r][4 * half + b] = acc[r][b].x + acc[r][b].y;
        }
        float pre2[RPT];
        if (dbg & 4) { for (int r = 0; r < RPT; r++) pre2[r] = part2[r][0] + part2[r][7]; }
        else reduce_to_own_sequence<RPT>(part2, pre2, ks);

        const bool live = t < len_own;           // ragged batch: state stays 0 outside the read
#pragma unroll
        for (int i = 0; i < RPT; i++) {
            const int j = j0 + i;
            const float hbar = (dbg & 8) ? 0.001f * (pre2[i] + vrow[2 * H + j]) : apply_act_fast<ACT>(pre2[i] + vrow[2 * H + j]);
            float hn = zg[i] * h_own[i] + (1.0f - zg[i]) * hbar;
            hn = (live && j < H) ? hn : 0.0f;
            h_own[i] = hn;
            if (j < HP) hs[b_own * HP + j] = hn;
        }
        cp_async_wait_all();                     // vI of the next step has landed (this thread's copies)
        __syncthreads();

        // ---------------- h_t -> HBM, coalesced from the shared state ----------------
        if (dbg & 32) continue;
        if (vec_y) {
            for (int e = tid; e < BT * HP4; e += THREADS) {
                const int b = e / HP4, c = e - b * HP4;
                if (b_base + b < B)
                    *reinterpret_cast<float4 *>(y + ((long)t * B + b_base + b) * ldy + 4 * c) = hsv[b * HP4 + c];
            }
        } else {
            for (int e = tid; e < BT * H; e += THREADS) {
                const int b = e / H, j = e - b * H;
                if (b_base + b < B) y[((long)t * B + b_base + b) * ldy + j] = hs[b * HP + j];
            }
        }
    }
}

template <int HP, int THREADS, bool W1REG>
static int launch(const float *vI, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths, int T,
                  int B, int H, int reverse, int act, int gate_act, cudaStream_t st)
{
    using C = Cfg<HP, THREADS>;
    const size_t smem = sizeof(float) * ((size_t)2 * BT * HP + (size_t)2 * BT * C::VLD + (W1REG ? 0 : (size_t)2 * HP * HP));
    // the models on the path use tanh / sigmoid (every models/*.py); other pairs go to the generic kernel
    if (act != SLOIKA_ACT_TANH || gate_act != SLOIKA_ACT_SIGMOID) return SLOIKA_ERR_UNSUPPORTED;
    auto kern = gru_recurrence_v2_kernel<HP, THREADS, W1REG, SLOIKA_ACT_TANH, SLOIKA_ACT_SIGMOID>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    const unsigned grid = (unsigned)ceil_div(B, BT);
    const char *dbg = getenv("SLOIKA_B200_GRU_DBG");
    kern<<<grid, THREADS, smem, st>>>(vI, sW, sW2, y, ldy, lengths, T, B, H, reverse, dbg ? atoi(dbg) : 0);
    SLOIKA_RETURN_LAUNCH_STATUS();
}

// Returns SLOIKA_ERR_UNSUPPORTED for sizes this kernel does not cover (caller uses gru.cu).
int dispatch(const float *vI, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths, int T, int B,
             int H, int reverse, int act, int gate_act, cudaStream_t st)
{
    const char *geo = getenv("SLOIKA_B200_GRU_THREADS");           // tuning experiments
    const int want = geo ? atoi(geo) : 0;
    if (H <= 32) return launch<32, 256, true>(vI, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, st);
    if (H <= 64) return launch<64, 256, true>(vI, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, st);
    if (H > 80 && H <= 96) {
        if (want == 256) return launch<96, 256, true>(vI, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, st);
        return launch<96, 384, true>(vI, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, st);
    }
    if (H > 112 && H <= 128) return launch<128, 256, false>(vI, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, st);
    return SLOIKA_ERR_UNSUPPORTED;
}

}  // namespace gru2
}  // namespace sloika
