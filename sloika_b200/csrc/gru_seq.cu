// GRU layer, throughput form with the SEQUENCES on the tensor-memory lanes: one CTA = 128 sequences, one launch = the
// whole layer including its input projection; nothing but x and h touches HBM.
//
// Gru.step scanned by RNN.run (reference sloika/layers.py:1010-1021, :85-88; Reverse :1449-1450; the projection line
// `vI = T.tensordot(in_vec, self.iW) + self.b` is :1011).  Same results as sloika_gru_fwd / sloika_gru_fused_fwd.
//
// gru_tc.cu / gru_fused.cu put GATE ROWS on the 128 lanes (96 used) and sequences on the MMA's N dimension with N = 16:
// an instruction then costs 9-12 cycles where its arithmetic needs 4, a step's dependent chain is ~2900 cycles whatever
// N is, and the projection weights need the tensor memory of a third SM.  Here the roles are swapped:
//   D[128 sequences][gate rows] (+)= A[128 sequences][K] . B[gate rows][K]^T
//   A  the activations (x_t, h_{t-1}, r*h) as fp16 hi / lo pairs IN TENSOR MEMORY: lane = sequence, two k per 32-bit
//      column -- exactly what the thread that owns (sequence, units) produces, written with tcgen05.st
//   B  the weights iW, sW, sW2 as fp16 hi / lo pairs in SHARED memory (K-major SWIZZLE_64B tiles, 216 KB at H = I = 96),
//      pre-scaled by the ex2 constants of the gate functions
//   D  z | r | candidate pre-activations, fp32, 3 HP columns; the projection MMAs of step t + 1 write them FIRST
//      (accumulate = 0), the recurrent MMAs of step t + 1 accumulate on top: vI exists only inside the accumulator
// Tensor memory: 3 HP accumulator columns + HP (h hi / lo; r*h overwrites it once phase 1 has read h) + IP (x hi / lo)
// = 480 of 512 at H = I = 96.  Per step the tensor pipe does 18 MMAs of N = 192 and 18 of N = 96, twice (projection,
// recurrence): ~2600 cycles for 128 sequences (the lanes-are-gate-rows form: ~1300 for 32).
//
// Warps: 16 compute warps (warp w: lanes 32 (w & 3).., units HP/4 (w >> 2)..; a thread owns one sequence x HP/4 units,
// its h values live in registers) + one issuing warp.  Per step t:
//   compute  wait d1 (r complete) -> r = sigmoid -> r*h -> operand (tcgen05.st) -> arrive RH;  load x_{t+1} (in L2 since t - 2)
//            wait d1z -> z raw;  wait dx (projection of step t has read the x operand) -> x_{t+1} -> operand;
//            arrive ZFREE (z | r accumulators and the x operand are handed over);  prefetch x_{t+3};  z -> denominators
//            wait d2 (candidate complete) -> blend -> h_t -> operand, arrive H;  h_t -> HBM
//   issuer   sync H: r += sW_r . h_{t-1}, commit d1;  z += sW_z . h_{t-1}, commit d1z;  c = iW_c . x_t (accumulate = 0), commit dx
//            sync RH: c += sW2 . (r*h), commit d2
//            sync ZFREE: z | r = iW_zr . x_{t+1} (accumulate = 0)      -- off the dependent chain
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "gru_tc_common.cuh"

namespace sloika {
namespace gru7 {

using namespace tc;
using namespace gru5;

#ifdef GRU_TC_TRACE
__device__ long long g_strace[64 * 16];
#define STRACE(slot) do { if (blockIdx.x == 0 && s >= 100 && s < 164) g_strace[(s - 100) * 16 + (slot)] = clock64(); } while (0)
#else
#define STRACE(slot) do { } while (0)
#endif

constexpr int MSEQ = 128;             // sequences per CTA = TMEM lanes
constexpr int CW = 16;                // compute warps
constexpr int NTHREADS = (CW + 1) * 32;

struct SBars {
    uint64_t d1, dx, d2, d1z;
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ void tmem_st_32x32b_x4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
                 : "memory");
}

// 8 values -> 4 + 4 columns of the hi / lo operand (two fp16 per column, even k in the low half)
__device__ __forceinline__ void store_operand8(uint32_t hi_col, uint32_t lo_col, const float (&v)[8]) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; i++) split_pair(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
    tmem_st_32x32b_x4(hi_col, hi);
    tmem_st_32x32b_x4(lo_col, lo);
}

// shared-memory bytes of the weight tiles (hi + lo): rows x 64 bytes per K block of 32
__host__ __device__ constexpr int tile_bytes(int rows, int kblocks) { return rows * 64 * kblocks; }

// HP / IP: hidden / input size padded to a multiple of 32 (32, 64, 96)
// BLK: both tensors are in the blocked layout (known at compile time: the row-major access code and its registers are gone)
template <int HP, int IP, bool BLK>
__global__ void __launch_bounds__(NTHREADS, 1)   // 96 registers: the fifth warp of a sub-partition has to fit its 16 K file
gru_seq_kernel(const float *__restrict__ x, long ldx, const float *__restrict__ iW, const float *__restrict__ bias,
               const float *__restrict__ sW, const float *__restrict__ sW2, float *__restrict__ y, long ldy,
               const int32_t *__restrict__ lengths, int T, int B, int I, int H, int reverse, int xblocked_rt, int yblocked_rt,
               const Gate gate)
{
    if (gate_closed(gate)) return;
    const bool xblocked = BLK || xblocked_rt != 0, yblocked = BLK || yblocked_rt != 0;
    constexpr int KBH = HP / 32, KBI = IP / 32;          // K blocks of 32
    constexpr int UPT = HP / 4, XPT = IP / 4;            // units / input features per thread
    constexpr int NCH = UPT / 8, NXC = XPT / 8;          // chunks of 8
    // tensor memory columns
    constexpr int D_Z = 0, D_R = HP, D_C = 2 * HP;
    constexpr int A_H_HI = 3 * HP, A_H_LO = A_H_HI + HP / 2;
    constexpr int A_X_HI = A_H_LO + HP / 2, A_X_LO = A_X_HI + IP / 2;
    static_assert(A_X_LO + IP / 2 <= 512, "tensor memory: 512 columns");
    // shared memory: weight tiles
    constexpr int SZ_SWZR = tile_bytes(2 * HP, KBH), SZ_SW2 = tile_bytes(HP, KBH), SZ_IW = tile_bytes(3 * HP, KBI);

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *swzr_hi = smem, *swzr_lo = swzr_hi + SZ_SWZR;
    uint8_t *sw2_hi = swzr_lo + SZ_SWZR, *sw2_lo = sw2_hi + SZ_SW2;
    uint8_t *iw_hi = sw2_lo + SZ_SW2, *iw_lo = iw_hi + SZ_IW;
    float *bias_s = reinterpret_cast<float *>(iw_lo + SZ_IW);                 // [3 HP], scaled like the weights
    SBars *bars = reinterpret_cast<SBars *>(bias_s + 3 * HP);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b_cta = blockIdx.x * MSEQ;

    // ---------------- prologue ----------------
    if (tid == 0) {
        mbar_init(&bars->d1, 1); mbar_init(&bars->dx, 1); mbar_init(&bars->d2, 1); mbar_init(&bars->d1z, 1);
        mbar_fence_init();
    }
    if (warp == CW) tmem_alloc(&bars->tmem_base, 512);
    // weights -> fp16 hi / lo, K-major SWIZZLE_64B tiles, scaled: z, r rows by -log2 e, candidate rows by 2 log2 e
    auto fill = [&](uint8_t *hi, uint8_t *lo, int rows, int kblocks, auto src) {
        for (int e = tid; e < rows * kblocks * 32; e += NTHREADS) {
            const int k = e % (kblocks * 32), n = e / (kblocks * 32);
            const float w = src(n, k);
            const __half h = __float2half_rn(w);
            const uint32_t off = (uint32_t)(k / 32) * (uint32_t)(rows * 64) + sw64_offset(n, k % 32);
            *reinterpret_cast<__half *>(hi + off) = h;
            *reinterpret_cast<__half *>(lo + off) = __float2half_rn(w - __half2float(h));
        }
    };
    constexpr float SG = -SLOIKA_LOG2E, SC = 2.0f * SLOIKA_LOG2E;
    fill(swzr_hi, swzr_lo, 2 * HP, KBH, [&](int n, int k) {
        const int g = n / HP, u = n - g * HP;
        return (u < H && k < H) ? SG * __ldg(sW + (long)(g * H + u) * H + k) : 0.0f;
    });
    fill(sw2_hi, sw2_lo, HP, KBH, [&](int n, int k) { return (n < H && k < H) ? SC * __ldg(sW2 + (long)n * H + k) : 0.0f; });
    fill(iw_hi, iw_lo, 3 * HP, KBI, [&](int n, int k) {
        const int g = n / HP, u = n - g * HP;
        return (u < H && k < I) ? (g < 2 ? SG : SC) * __ldg(iW + (long)(g * H + u) * I + k) : 0.0f;
    });
    for (int n = tid; n < 3 * HP; n += NTHREADS) {
        const int g = n / HP, u = n - g * HP;
        bias_s[n] = u < H ? (g < 2 ? SG : SC) * __ldg(bias + g * H + u) : 0.0f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    constexpr int NB_COUNT = NTHREADS;
    constexpr int NB_H = 1, NB_RH = 2, NB_ZFREE = 3;

    if (warp < CW) {
        // =====================================================================================================
        // compute warps: thread = sequence m (TMEM lane) x units [u0, u0 + UPT)
        // =====================================================================================================
        const int q = warp & 3, ug = warp >> 2;
        const int m = 32 * q + lane;
        const int u0 = ug * UPT, f0 = ug * XPT;
        const int bg = b_cta + m;
        const bool live = bg < B;
        const int len = live ? (lengths ? min(lengths[bg], T) : T) : 0;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16);
        const float *xrow = x + ((long)(reverse ? T - 1 : 0) * B + (live ? bg : 0)) * ldx + f0;
        const long xstep = (long)(reverse ? -1 : 1) * B * ldx;
        float *yrow = y + ((long)(reverse ? T - 1 : 0) * B + (live ? bg : 0)) * ldy + u0;
        const long ystep = (long)(reverse ? -1 : 1) * B * ldy;
        const bool yvec = ((ldy & 3) == 0) && (((uintptr_t)y & 15) == 0);

        const bool all_units = u0 + UPT <= H;
        float h[UPT];
#pragma unroll
        for (int i = 0; i < UPT; i++) h[i] = 0.0f;
        float xn[XPT];                                    // raw x of the step whose operand is written next
        // BLOCKED layout (between two layers of this kernel; include/sloika_b200.h): groups of 4 features of one sequence are
        // 16 contiguous bytes, the 128 sequences of a block follow each other, then the next group of 4 features --
        // element (t, b, f) at (((t * nblk + b / 128) * ceil(F / 4) + f / 4) * 128 + b % 128) * 4 + f % 4 -- so that the 32
        // lanes of a warp (consecutive sequences) read or write 512 contiguous bytes per 128-bit access.  With row-major rows
        // such an access touches 32 different lines; the 96 + 96 of them per step kept the L1 / shared-memory pipeline busy
        // for ~4600 cycles and starved the tensor core's operand reads (tools/gru_seq_trace.py: 16 800 -> 9 900 cycles per step).
        const long nblk = gridDim.x;
        const int xg = (I + 3) / 4, yg = (H + 3) / 4;            // groups of 4 features per sequence
        const float4 *xblk = reinterpret_cast<const float4 *>(x) + (((long)(reverse ? T - 1 : 0) * nblk + blockIdx.x) * xg + f0 / 4) * MSEQ + m;
        const long xbstep = (long)(reverse ? -1 : 1) * nblk * xg * MSEQ;
        float4 *yblk = reinterpret_cast<float4 *>(y) + (((long)(reverse ? T - 1 : 0) * nblk + blockIdx.x) * yg + u0 / 4) * MSEQ + m;
        const long ybstep = (long)(reverse ? -1 : 1) * nblk * yg * MSEQ;
        auto load_x = [&](int s) {                        // x of scan step s -> xn (zeros outside the batch / the input width)
            if (xblocked) {
                const float4 *p = xblk + (long)s * xbstep;
                const bool go = live && s < T;
#pragma unroll
                for (int g = 0; g < XPT / 4; g++) {        // padding features of the last group are zeros in memory
                    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (go && f0 / 4 + g < xg) v = __ldg(p + g * MSEQ);
                    xn[4 * g] = v.x; xn[4 * g + 1] = v.y; xn[4 * g + 2] = v.z; xn[4 * g + 3] = v.w;
                }
                return;
            }
#pragma unroll
            for (int c = 0; c < NXC; c++) {
#pragma unroll
                for (int v4 = 0; v4 < 2; v4++) {
                    const int f = f0 + 8 * c + 4 * v4;
                    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (live && s < T) {
                        const float *p = xrow + (long)s * xstep + 8 * c + 4 * v4;
                        if (f + 3 < I) v = __ldg(reinterpret_cast<const float4 *>(p));
                        else {
                            if (f < I) v.x = __ldg(p);
                            if (f + 1 < I) v.y = __ldg(p + 1);
                            if (f + 2 < I) v.z = __ldg(p + 2);
                        }
                    }
                    xn[8 * c + 4 * v4 + 0] = v.x; xn[8 * c + 4 * v4 + 1] = v.y;
                    xn[8 * c + 4 * v4 + 2] = v.z; xn[8 * c + 4 * v4 + 3] = v.w;
                }
            }
        };
        // x of scan step s -> L2, a step before it is loaded (blocked layout: one line per 8 lanes, so every 8th lane asks)
        auto prefetch_x = [&](int s) {
            if (!xblocked || !live || s >= T || (lane & 7) != 0) return;
            const float4 *p = xblk + (long)s * xbstep;
#pragma unroll
            for (int g = 0; g < XPT / 4; g++)
                if (f0 / 4 + g < xg) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + g * MSEQ));
        };
        auto store_x = [&]() {                            // xn -> x operand columns of this thread
#pragma unroll
            for (int c = 0; c < NXC; c++) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = xn[8 * c + i];
                store_operand8(lane_addr + (uint32_t)(A_X_HI + (f0 + 8 * c) / 2), lane_addr + (uint32_t)(A_X_LO + (f0 + 8 * c) / 2), v);
            }
        };
        auto store_h_operand = [&](const float (&v)[UPT]) {
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                float w[8];
#pragma unroll
                for (int i = 0; i < 8; i++) w[i] = v[8 * c + i];
                store_operand8(lane_addr + (uint32_t)(A_H_HI + (u0 + 8 * c) / 2), lane_addr + (uint32_t)(A_H_LO + (u0 + 8 * c) / 2), w);
            }
        };

        // h_{-1} = 0 and x_0 as operands; x_1 into registers
        load_x(0);
        store_h_operand(h);
        store_x();
        prefetch_x(1);
        prefetch_x(2);
        tmem_st_wait();
        tc_fence_before();
        nbar_arrive(NB_H, NB_COUNT);

        for (int s = 0; s < T; s++) {
            const int t = reverse ? T - 1 - s : s;
            const uint32_t par = (uint32_t)(s & 1);
            // ---- r -> r*h ----
            wait_bar(&bars->d1, par);
            if (tid == 0) STRACE(0);
            tc_fence_after();
            uint32_t rhi[UPT / 2], rlo[UPT / 2];          // r*h as packed fp16 hi / lo pairs
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                uint32_t d[8];
                tmem_ld_32x32b_x8(lane_addr + (uint32_t)(D_R + u0 + 8 * c), d);
                tmem_ld_wait();
                const float4 b0 = *reinterpret_cast<const float4 *>(&bias_s[HP + u0 + 8 * c]);
                const float4 b1 = *reinterpret_cast<const float4 *>(&bias_s[HP + u0 + 8 * c + 4]);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float rh[8];
#pragma unroll
                for (int i = 0; i < 8; i++) rh[i] = sigmoid_pre(__uint_as_float(d[i]) + bb[i]) * h[8 * c + i];
#pragma unroll
                for (int i = 0; i < 4; i++) split_pair(rh[2 * i], rh[2 * i + 1], rhi[4 * c + i], rlo[4 * c + i]);
            }
            // r*h takes the place of h in tensor memory, and the z MMAs (issued behind the r MMAs) still read h: they have
            // completed by now (they run while the r gate is being computed), the wait makes it certain
            wait_bar(&bars->d1z, par);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const uint32_t hi4[4] = {rhi[4 * c], rhi[4 * c + 1], rhi[4 * c + 2], rhi[4 * c + 3]};
                const uint32_t lo4[4] = {rlo[4 * c], rlo[4 * c + 1], rlo[4 * c + 2], rlo[4 * c + 3]};
                tmem_st_32x32b_x4(lane_addr + (uint32_t)(A_H_HI + (u0 + 8 * c) / 2), hi4);
                tmem_st_32x32b_x4(lane_addr + (uint32_t)(A_H_LO + (u0 + 8 * c) / 2), lo4);
            }
            tmem_st_wait();
            tc_fence_before();
            nbar_arrive(NB_RH, NB_COUNT);
            if (tid == 0) STRACE(1);
            // ---- x_{s+1} on its way from L2 (prefetched two steps ago; row-major inputs come from HBM here) ----
            if (s + 1 < T) load_x(s + 1);
            // ---- z raw: with it the z | r accumulators are read ----
            float za[UPT];
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                uint32_t d[8];
                tmem_ld_32x32b_x8(lane_addr + (uint32_t)(D_Z + u0 + 8 * c), d);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; i++) za[8 * c + i] = __uint_as_float(d[i]);
            }
            // ---- x_{s+1} -> operand (the projection of step s has read x_s); hand both over EARLY: the z | r projection of
            // step s + 1 then runs behind phase 2 of this step instead of in front of phase 1 of the next one ----
            wait_bar(&bars->dx, par);
            if (tid == 0) STRACE(2);
            tc_fence_after();
            if (s + 1 < T) store_x();
            if (tid == 0) STRACE(6);
            tmem_st_wait();
            tc_fence_before();
            nbar_arrive(NB_ZFREE, NB_COUNT);               // z | r accumulators read, x_{s+1} operand written
            if (tid == 0) STRACE(3);
            prefetch_x(s + 3);
            if (tid == 0) STRACE(7);
            // ---- z: denominators 1 + 2^(-z log2 e), while phase 2 runs ----
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const float4 b0 = *reinterpret_cast<const float4 *>(&bias_s[u0 + 8 * c]);
                const float4 b1 = *reinterpret_cast<const float4 *>(&bias_s[u0 + 8 * c + 4]);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; i++) za[8 * c + i] = gate_denominator(za[8 * c + i] + bb[i]);
            }
            if (tid == 0) STRACE(14);
            // ---- candidate -> blend -> h ----
            wait_bar(&bars->d2, par);
            if (tid == 0) STRACE(4);
            tc_fence_after();
            const bool on = t < len;
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                uint32_t d[8];
                tmem_ld_32x32b_x8(lane_addr + (uint32_t)(D_C + u0 + 8 * c), d);
                tmem_ld_wait();
                const float4 b0 = *reinterpret_cast<const float4 *>(&bias_s[2 * HP + u0 + 8 * c]);
                const float4 b1 = *reinterpret_cast<const float4 *>(&bias_s[2 * HP + u0 + 8 * c + 4]);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                float hc[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float hn = gru_blend(za[8 * c + i], __uint_as_float(d[i]) + bb[i], h[8 * c + i]);
                    hc[i] = h[8 * c + i] = (on && (all_units || u0 + 8 * c + i < H)) ? hn : 0.0f;
                }
                store_operand8(lane_addr + (uint32_t)(A_H_HI + (u0 + 8 * c) / 2), lane_addr + (uint32_t)(A_H_LO + (u0 + 8 * c) / 2), hc);
            }
            tmem_st_wait();
            tc_fence_before();
            nbar_arrive(NB_H, NB_COUNT);
            if (tid == 0) STRACE(5);
            if (live && yblocked) {
                float4 *yp = yblk + (long)s * ybstep;
#pragma unroll
                for (int g = 0; g < UPT / 4; g++)           // h is 0 for padding units
                    if (u0 / 4 + g < yg) yp[g * MSEQ] = make_float4(h[4 * g], h[4 * g + 1], h[4 * g + 2], h[4 * g + 3]);
            } else if (live) {
                float *yp = yrow + (long)s * ystep;
#pragma unroll
                for (int c = 0; c < UPT / 4; c++) {
                    const int u = u0 + 4 * c;
                    if (yvec && u + 3 < H) {
                        *reinterpret_cast<float4 *>(yp + 4 * c) = make_float4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            if (u + i < H) yp[4 * c + i] = h[4 * c + i];
                    }
                }
            }
        }
    } else {
        // =====================================================================================================
        // issuing warp
        // =====================================================================================================
        const uint32_t id_zr = umma_idesc_f16_m128(2 * HP), id_c = umma_idesc_f16_m128(HP);
        const uint32_t d_zr = tmem_base + D_Z, d_c = tmem_base + D_C;
        const uint32_t a_hh = tmem_base + A_H_HI, a_hl = tmem_base + A_H_LO, a_xh = tmem_base + A_X_HI, a_xl = tmem_base + A_X_LO;
        const uint32_t s_swzr_hi = smem_u32(swzr_hi), s_swzr_lo = smem_u32(swzr_lo), s_sw2_hi = smem_u32(sw2_hi), s_sw2_lo = smem_u32(sw2_lo);
        const uint32_t s_iw_hi = smem_u32(iw_hi), s_iw_lo = smem_u32(iw_lo);
        // D (+)= A . B^T over K = 32 KB: per K = 16 step the three products hi.hi + lo.hi + hi.lo
        auto mma_set = [&](uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int rows, int row0, int kblocks,
                           uint32_t idesc, bool accumulate) {
#pragma unroll
            for (int kc = 0; kc < 2 * kblocks; kc++) {
                const uint32_t boff = (uint32_t)(kc >> 1) * (uint32_t)(rows * 64) + (uint32_t)(row0 >> 3) * 512u + (uint32_t)(kc & 1) * 32u;
                const uint64_t bh = umma_desc_sw64_kmajor(b_hi + boff), bl = umma_desc_sw64_kmajor(b_lo + boff);
                umma_f16_ts(d, a_hi + (uint32_t)(kc * 8), bh, idesc, accumulate || kc != 0);
                umma_f16_ts(d, a_lo + (uint32_t)(kc * 8), bh, idesc, true);
                umma_f16_ts(d, a_hi + (uint32_t)(kc * 8), bl, idesc, true);
            }
        };
        for (int s = 0; s < T; s++) {
            nbar_sync(NB_H, NB_COUNT);                     // h_{s-1} operand written, candidate accumulator read
            if (lane == 0) STRACE(8);
            tc_fence_after();
            if (elect_one()) {
                if (s == 0) mma_set(d_zr, a_xh, a_xl, s_iw_hi, s_iw_lo, 3 * HP, 0, KBI, id_zr, false);   // z | r = iW_zr . x_0
                // r first and on its own: the dependent chain (r -> r*h -> candidate) waits for it, z is needed later.
                // (Tried: r, candidate projection, z -- the x operand is free and phase 2 can start 860 cycles earlier, but r*h
                // may replace h only after the z MMAs: 4.20 -> 4.38 ms per layer.)
                mma_set(d_zr + HP, a_hh, a_hl, s_swzr_hi, s_swzr_lo, 2 * HP, HP, KBH, id_c, true);        // r += sW_r . h
                umma_commit(&bars->d1);
                mma_set(d_zr, a_hh, a_hl, s_swzr_hi, s_swzr_lo, 2 * HP, 0, KBH, id_c, true);              // z += sW_z . h
                umma_commit(&bars->d1z);
                mma_set(d_c, a_xh, a_xl, s_iw_hi, s_iw_lo, 3 * HP, 2 * HP, KBI, id_c, false);              // c = iW_c . x_s
                umma_commit(&bars->dx);
                STRACE(9);
            }
            __syncwarp();
            nbar_sync(NB_RH, NB_COUNT);                    // r*h operand written
            if (lane == 0) STRACE(10);
            tc_fence_after();
            if (elect_one()) {
                mma_set(d_c, a_hh, a_hl, s_sw2_hi, s_sw2_lo, HP, 0, KBH, id_c, true);                      // c += sW2 . (r*h)
                umma_commit(&bars->d2);
                STRACE(11);
            }
            __syncwarp();
            nbar_sync(NB_ZFREE, NB_COUNT);                 // z | r accumulators read, x_{s+1} operand written
            if (lane == 0) STRACE(12);
            tc_fence_after();
            if (s + 1 < T && elect_one()) {
                mma_set(d_zr, a_xh, a_xl, s_iw_hi, s_iw_lo, 3 * HP, 0, KBI, id_zr, false);
                STRACE(13);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == CW) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int HP, int IP>
static int launch(const float *x, long ldx, const float *iW, const float *bias, const float *sW, const float *sW2, float *y,
                  long ldy, const int32_t *lengths, int T, int B, int I, int H, int reverse, int xblocked, int yblocked,
                  cudaStream_t st, Gate gate)
{
    const size_t smem = 1024 + 2 * (size_t)(tile_bytes(2 * HP, HP / 32) + tile_bytes(HP, HP / 32) + tile_bytes(3 * HP, IP / 32)) +
                        (size_t)3 * HP * 4 + sizeof(SBars) + 64;
    size_t ask = smem < 116 * 1024 ? 116 * 1024 : smem;     // one CTA per SM: it owns the SM's tensor memory
    auto kern = (xblocked && yblocked) ? gru_seq_kernel<HP, IP, true> : gru_seq_kernel<HP, IP, false>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ask);
    if (err != cudaSuccess) return (int)err;
    kern<<<(unsigned)ceil_div(B, MSEQ), NTHREADS, ask, st>>>(x, ldx, iW, bias, sW, sW2, y, ldy, lengths, T, B, I, H, reverse, xblocked, yblocked, gate);
    SLOIKA_RETURN_LAUNCH_STATUS();
}

int dispatch(const float *x, long ldx, const float *iW, const float *sW, const float *sW2, const float *b, float *y, long ldy,
             const int32_t *lengths, int T, int B, int I, int H, int reverse, int act, int gate_act, int layout, cudaStream_t st,
             Gate gate)
{
    const int xblocked = layout & 1, yblocked = (layout >> 1) & 1;
    if (!x || !iW || !sW || !sW2 || !b || !y || T < 0 || B <= 0 || I <= 0 || H <= 0) return SLOIKA_ERR_ARG;
    if ((!xblocked && ldx < I) || (!yblocked && ldy < H)) return SLOIKA_ERR_ARG;
    if (act != SLOIKA_ACT_TANH || gate_act != SLOIKA_ACT_SIGMOID) return SLOIKA_ERR_UNSUPPORTED;
    if (H > 96 || I > 96) return SLOIKA_ERR_UNSUPPORTED;
    if (!xblocked && ((ldx & 3) != 0 || ((uintptr_t)x & 15) != 0)) return SLOIKA_ERR_UNSUPPORTED;
    if ((xblocked && ((uintptr_t)x & 15) != 0) || (yblocked && ((uintptr_t)y & 15) != 0)) return SLOIKA_ERR_UNSUPPORTED;
    if (T == 0) return SLOIKA_OK;
    const int HP = (H + 31) / 32 * 32, IP = (I + 31) / 32 * 32;
#define SEQ_CASE(HP_, IP_) \
    if (HP == HP_ && IP == IP_) return launch<HP_, IP_>(x, ldx, iW, b, sW, sW2, y, ldy, lengths, T, B, I, H, reverse, xblocked, yblocked, st, gate)
    SEQ_CASE(32, 32); SEQ_CASE(32, 64); SEQ_CASE(32, 96);
    SEQ_CASE(64, 32); SEQ_CASE(64, 64); SEQ_CASE(64, 96);
    SEQ_CASE(96, 32); SEQ_CASE(96, 64); SEQ_CASE(96, 96);
#undef SEQ_CASE
    return SLOIKA_ERR_UNSUPPORTED;
}


// row-major [T][B][F] (row pitch ld) <-> blocked (see above): one CTA per (t, block of 128 sequences); the tile goes through
// shared memory so that both sides are accessed in 16-byte pieces, contiguous across the threads of a warp.
__global__ void __launch_bounds__(256)
block_layout_kernel(const float *__restrict__ src, float *__restrict__ dst, long ld, int T, int B, int F, int to_blocked,
                    const Gate gate)
{
    if (gate_closed(gate)) return;
    extern __shared__ float4 tile4[];                       // [fg][128 + 1] float4 (row-major side indexed [b][g])
    const int fg = (F + 3) / 4;
    const long nblk = (B + MSEQ - 1) / MSEQ;
    const long blk = blockIdx.x % nblk, t = blockIdx.x / nblk;
    const int b0 = (int)blk * MSEQ;
    const int nb = min(MSEQ, B - b0);
    float4 *blocked = reinterpret_cast<float4 *>(to_blocked ? dst : const_cast<float *>(src)) + (t * nblk + blk) * fg * MSEQ;
    const float *rows = (to_blocked ? src : dst) + (t * B + b0) * ld;
    const int tid = threadIdx.x;
    const bool vec = (ld & 3) == 0 && (((uintptr_t)(to_blocked ? src : dst)) & 15) == 0;
    constexpr int TP = MSEQ + 1;
    if (to_blocked) {
        for (int e = tid; e < MSEQ * fg; e += 256) {        // consecutive threads: consecutive groups of one row
            const int b = e / fg, g = e - b * fg;
            float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (b < nb) {
                const float *p = rows + (long)b * ld + 4 * g;
                if (vec && 4 * g + 3 < F) v = __ldg(reinterpret_cast<const float4 *>(p));
                else {
                    if (4 * g < F) v.x = __ldg(p);
                    if (4 * g + 1 < F) v.y = __ldg(p + 1);
                    if (4 * g + 2 < F) v.z = __ldg(p + 2);
                    if (4 * g + 3 < F) v.w = __ldg(p + 3);
                }
            }
            tile4[g * TP + b] = v;
        }
        __syncthreads();
        for (int e = tid; e < MSEQ * fg; e += 256) {        // consecutive threads: consecutive sequences of one group
            const int g = e / MSEQ, b = e - g * MSEQ;
            blocked[(long)g * MSEQ + b] = tile4[g * TP + b];
        }
    } else {
        for (int e = tid; e < MSEQ * fg; e += 256) {
            const int g = e / MSEQ, b = e - g * MSEQ;
            tile4[g * TP + b] = blocked[(long)g * MSEQ + b];
        }
        __syncthreads();
        for (int e = tid; e < MSEQ * fg; e += 256) {
            const int b = e / fg, g = e - b * fg;
            if (b >= nb) continue;
            const float4 v = tile4[g * TP + b];
            float *p = const_cast<float *>(rows) + (long)b * ld + 4 * g;
            if (vec && 4 * g + 3 < F) *reinterpret_cast<float4 *>(p) = v;
            else {
                if (4 * g < F) p[0] = v.x;
                if (4 * g + 1 < F) p[1] = v.y;
                if (4 * g + 2 < F) p[2] = v.z;
                if (4 * g + 3 < F) p[3] = v.w;
            }
        }
    }
}

int block_layout(const float *src, float *dst, long ld, int T, int B, int F, int to_blocked, cudaStream_t st, Gate gate)
{
    if (!src || !dst || T < 0 || B <= 0 || F <= 0 || ld < F) return SLOIKA_ERR_ARG;
    if ((((uintptr_t)(to_blocked ? dst : src)) & 15) != 0) return SLOIKA_ERR_UNSUPPORTED;
    if (T == 0) return SLOIKA_OK;
    const long nblk = (B + MSEQ - 1) / MSEQ;
    const long grid = nblk * T;
    if (grid > 0x7fffffffL) return SLOIKA_ERR_ARG;
    const int fg = (F + 3) / 4;
    const size_t smem = (size_t)fg * (MSEQ + 1) * sizeof(float4);
    cudaError_t err = cudaFuncSetAttribute(block_layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    block_layout_kernel<<<(unsigned)grid, 256, smem, st>>>(src, dst, ld, T, B, F, to_blocked, gate);
    SLOIKA_RETURN_LAUNCH_STATUS();
}

}  // namespace gru7
}  // namespace sloika

using namespace sloika;

#ifdef GRU_TC_TRACE
extern "C" int sloika_debug_gru_seq_trace(long long *out)
{
    return (int)cudaMemcpyFromSymbol(out, sloika::gru7::g_strace, sizeof(long long) * 64 * 16);
}
#endif

extern "C" int sloika_gru_seq_fwd(const float *x, long ldx, const float *iW, const float *sW, const float *sW2,
                                  const float *b, float *y, long ldy, const int32_t *lengths, int T, int B, int I, int H,
                                  int reverse, int act, int gate_act, int layout, void *stream)
{
    return gru7::dispatch(x, ldx, iW, sW, sW2, b, y, ldy, lengths, T, B, I, H, reverse, act, gate_act, layout,
                          (cudaStream_t)stream, gru5::Gate{nullptr, 0u, 0});
}

extern "C" size_t sloika_blocked_bytes(int T, int B, int F)
{
    if (T < 0 || B <= 0 || F <= 0) return 0;
    return (size_t)T * (size_t)((B + 127) / 128) * (size_t)((F + 3) / 4) * 128 * 4 * sizeof(float);
}

extern "C" int sloika_block_layout_fwd(const float *src, float *dst, long ld, int T, int B, int F, int to_blocked, void *stream)
{
    return gru7::block_layout(src, dst, ld, T, B, F, to_blocked, (cudaStream_t)stream, gru5::Gate{nullptr, 0u, 0});
}

namespace sloika {
namespace gemm_tc {
int launch(const float *x, long ldx, const float *W, const float *bias, float *y, long ldy, long M, int K, int N, int act,
           float2 *stats, int rot, bool f16, cudaStream_t st, const unsigned *gate, unsigned gate_limit, int gate_mode);
}
namespace gru5 {
int dispatch_gated(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths,
                   int T, int B, int H, int reverse, int act, int gate_act, long seqs_in_flight, cudaStream_t st, Gate gate);
}
}  // namespace sloika

// The layer for an input whose range only the device knows (the elu convolution's output; `absmax` as in
// sloika_gru_fwd_gated), with the output in the BLOCKED layout `yb` either way.  x is row-major with pitch ldx
// (x_blocked = 0; `scratch` then holds its blocked copy) or already blocked (x_blocked = 1; `scratch` is a row-major
// [T][B] buffer of pitch ldx that is written only if the range check fails):
//   max |x| <  limit : (x -> blocked scratch,) then the sequences-on-lanes launch -> yb
//   max |x| >= limit : (x -> row-major scratch,) tf32-split projection GEMM into vI, the recurrence kernel into the row-major
//                      scratch y, y -> yb
// Every launch of the form that is ruled out returns at once.
extern "C" int sloika_gru_seq_fwd_gated(const float *x, long ldx, int x_blocked, const float *iW, const float *sW,
                                        const float *sW2, const float *b, float *yb, float *scratch, float *y, long ldy,
                                        float *vI, long ldv, const int32_t *lengths, int T, int B, int I, int H, int reverse,
                                        int act, int gate_act, long seqs_in_flight, const float *absmax, float limit,
                                        void *stream)
{
    if (!absmax || !vI || !scratch || !yb || !y || !(limit > 0.0f) || ldv < 3L * H || ldy < H || ldx < I) return SLOIKA_ERR_ARG;
    if ((ldv & 3) != 0 || ((uintptr_t)vI & 15) != 0 || (ldx & 3) != 0 || (long)T * B < 128) return SLOIKA_ERR_UNSUPPORTED;
    unsigned limit_bits;
    memcpy(&limit_bits, &limit, sizeof(limit_bits));
    const unsigned *word = reinterpret_cast<const unsigned *>(absmax);
    const gru5::Gate below{word, limit_bits, 1}, above{word, limit_bits, 2};
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    const float *xb = x_blocked ? x : scratch;
    const float *xr = x_blocked ? scratch : x;
    if (!x_blocked) {
        rc = gru7::block_layout(x, scratch, ldx, T, B, I, 1, st, below);
        if (rc != SLOIKA_OK || T == 0) return rc;
    }
    rc = gru7::dispatch(xb, 0, iW, sW, sW2, b, yb, 0, lengths, T, B, I, H, reverse, act, gate_act, 3, st, below);
    if (rc != SLOIKA_OK || T == 0) return rc;
    if (x_blocked) {
        rc = gru7::block_layout(x, scratch, ldx, T, B, I, 0, st, above);
        if (rc != SLOIKA_OK) return rc;
    }
    rc = gemm_tc::launch(xr, ldx, iW, b, vI, ldv, (long)T * B, I, 3 * H, SLOIKA_ACT_LINEAR, nullptr, 0, false, st, word, limit_bits, 2);
    if (rc != SLOIKA_OK) return rc;
    rc = gru5::dispatch_gated(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, seqs_in_flight, st, above);
    if (rc != SLOIKA_OK) return rc;
    return gru7::block_layout(y, yb, ldy, T, B, H, 1, st, above);
}
