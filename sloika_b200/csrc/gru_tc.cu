// GRU recurrence on the 5th-generation tensor cores: tcgen05.mma with the recurrent weights RESIDENT IN TENSOR MEMORY.
//
// Semantics: Gru.step scanned by RNN.run (reference sloika/layers.py:1010-1021, :85-88; Reverse :1449-1450), input
// projection vI = x iW' + b precomputed for all steps (gemm_tc.cu).  Same results as gru_h16.cu / gru.cu.
//
// Why this shape (profiles/r2_umma_ts_probe.txt): with the weight operand A in TMEM a `tcgen05.mma kind::f16`
// (M = 128, K = 16) costs 8 cycles at N <= 16 and 16 at N = 32 -- its 128 x N / 256 floor -- where the same
// instruction with A in shared memory costs 39 (a 4 KB operand read each) and legacy mma.sync needs 2.16 cycles per
// m16n8k16, i.e. 700 cycles per step for 8 sequences.  A step here is 54 MMAs (3 gate tiles x 6 K chunks x 3 split
// products) = 440 cycles at N = 8, 500 at N = 16.  What remains is the latency of the two dependent phases of a GRU
// step (~360 cycles each: smem write -> fence -> mbarrier -> MMA -> commit -> tcgen05.ld), which G independent
// groups of sequences per CTA overlap.
//
// Decomposition: one persistent CTA per G x N sequences.  Gate rows are TMEM lanes: M tile m (z, r, candidate) holds
// row j of that gate in lane j, as fp16 hi and lo parts (6 tiles of HP/2 columns; H <= 128).  A group is 4 compute
// warps (warp q owns lanes 32q..32q+31, thread = hidden unit j, all N sequences of the group in registers) plus one
// issuing warp.  Per step:
//   issuer   wait h_{t-1} operand -> [z | r] = sW . h   (2 tiles x KC x 3 MMAs, one commit) -> refill a vI ring slot
//            with 1-D bulk copies (TMA) -> wait r*h operand -> c = sW2 . (r*h) (KC x 3 MMAs, commit)
//   compute  tcgen05.ld r, z pre-activations; r = sigmoid(. + vI_r); r*h -> fp16 hi/lo operand in shared memory
//            (K-major SWIZZLE_64B rows = sequences) -> fence.proxy.async -> arrive; z = sigmoid(. + vI_z) while the
//            phase-2 MMAs run; tcgen05.ld c; h = z h + (1 - z) tanh(. + vI_c); h -> y (HBM) and -> operand -> arrive
// Accuracy: operands are split x = hi + lo (fp16 pairs, |h| <= 1, weights O(1)) and three products
// hi.hi + lo.hi + hi.lo accumulate in fp32 in TMEM: 2^-22 relative, as in gru_h16.cu.
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "gru_tc_common.cuh"

namespace sloika {
namespace gru5 {

using namespace tc;

#ifdef GRU_TC_TRACE
__device__ long long g_trace[64 * 16];
#define TRACE(slot) do { if (blockIdx.x == 0 && s >= 100 && s < 164) g_trace[(s - 100) * 16 + (slot)] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif

struct Bars {                 // per group
    uint64_t h, d1, rh, d2, v[3];
    uint64_t d1z;             // z pre-activations (committed after r's, which the dependent chain needs first)
};

// G groups of N sequences per CTA (N = 8 or 16 = the N of every MMA: at N = 16 an instruction costs 9 cycles instead
// of 8, so the tensor pipe -- the limiter once several groups share an SM -- serves twice the sequences); CW compute
// warps per group (CW / 4 of them share a lane quarter and split the group's sequences), one issuing warp per group.
template <int HP, int G, int CW, int N>
__global__ void __launch_bounds__(G *(CW + 1) * 32, 1)
gru_tc_kernel(const float *__restrict__ vI, long ldv, const float *__restrict__ sW, const float *__restrict__ sW2,
              float *__restrict__ y, long ldy, const int32_t *__restrict__ lengths, int T, int B, int H, int reverse, const Gate gate)
{
    if (gate_closed(gate)) return;
    constexpr int KC = HP / 16;                   // K = 16 chunks
    constexpr int ACOLS = HP / 2;                 // TMEM columns of one A tile
    constexpr int D_BASE = 6 * ACOLS;             // accumulators behind the 6 weight tiles
    constexpr int OPB = HP * 2 * N;               // bytes of one operand array: [k][N sequences] fp16, MN-major core matrices
    constexpr int NCW = G * CW;                   // compute warps
    constexpr int NS = N / (CW / 4);              // sequences per compute thread
    constexpr int VLD = 3 * HP;                   // floats per staged vI row
    constexpr int NTHREADS = G * (CW + 1) * 32;
    static_assert(D_BASE + G * 3 * N <= 512, "tensor memory: 512 columns");
    static_assert(CW == 4 || CW == 8 || CW == 16, "compute warps per group");
    static_assert(N == 8 || N == 16, "sequences per group");
    static_assert(N / (CW / 4) >= 2 && N / (CW / 4) <= 8, "sequences per compute thread");

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint8_t *ops = smem;                                               // [G][4][OPB]: h hi, h lo, r*h hi, r*h lo
    float *vring = reinterpret_cast<float *>(ops + (size_t)G * 4 * OPB);   // [G][3][N][VLD]
    Bars *bars = reinterpret_cast<Bars *>(vring + (size_t)G * 3 * N * VLD);
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(bars + G);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b_cta = blockIdx.x * (G * N);

    // ---------------- prologue ----------------
    if (tid == 0) {
        for (int g = 0; g < G; g++) {
            mbar_init(&bars[g].d1, 1); mbar_init(&bars[g].d2, 1); mbar_init(&bars[g].d1z, 1);
            for (int i = 0; i < 3; i++) mbar_init(&bars[g].v[i], 1);
        }
        mbar_fence_init();
    }
    if (warp == NCW) tmem_alloc(tmem_base_s, 512);
    {
        uint32_t *z = reinterpret_cast<uint32_t *>(smem);
        const int nz = (int)(((uint8_t *)bars - smem) / 4);
        for (int e = tid; e < nz; e += NTHREADS) z[e] = 0u;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp < NCW) {
        // weights -> TMEM (split once): lane = gate row j, tile 2m + p (m: z, r, candidate; p: hi, lo), chunk kc at
        // columns kc*8 .. kc*8+7, column i = k pair (16kc + 2i, 16kc + 2i + 1).  Chunks are dealt over the warps
        // that share a lane quarter.
        const int q = warp & 3, j = 32 * q + lane;
        const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
        for (int c = warp >> 2; c < 3 * KC; c += NCW / 4) {
            const int m = c / KC, kc = c - m * KC;
            const float *row = m < 2 ? sW + (long)(m * H + j) * H : sW2 + (long)j * H;
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int k = 16 * kc + 2 * i;
                // the gate pre-activations feed ex2 directly: z, r rows carry -log2 e, candidate rows 2 log2 e
                // (sigmoid(x) = 1 / (1 + 2^(-x log2 e)), tanh(x) = 1 - 2 / (1 + 2^(2 x log2 e))); the projection term
                // is scaled by the same constant when it is added (one FFMA instead of FADD + FMUL per value)
                const float gs = m < 2 ? -SLOIKA_LOG2E : 2.0f * SLOIKA_LOG2E;
                const float w0 = (j < H && k < H) ? gs * __ldg(row + k) : 0.0f;
                const float w1 = (j < H && k + 1 < H) ? gs * __ldg(row + k + 1) : 0.0f;
                const __half2 h2 = __floats2half2_rn(w0, w1);
                const float2 hb = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(w0 - hb.x, w1 - hb.y);
                hi[i] = *reinterpret_cast<const uint32_t *>(&h2);
                lo[i] = *reinterpret_cast<const uint32_t *>(&l2);
            }
            tmem_st_32x32b_x8(tmem_base + lane_addr + (uint32_t)((2 * m) * ACOLS + kc * 8), hi);
            tmem_st_32x32b_x8(tmem_base + lane_addr + (uint32_t)((2 * m + 1) * ACOLS + kc * 8), lo);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp < NCW) {
        // ================= compute warps: thread = hidden unit j, NS sequences =================
        const int g = warp / CW, wg = warp - g * CW, q = wg & 3, j = 32 * q + lane;
        const int n0 = (wg >> 2) * NS;                                     // first sequence of this thread
        const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
        const uint32_t dcol = tmem_base + lane_addr + (uint32_t)(D_BASE + g * 3 * N + n0);
        Bars &bar = bars[g];
        const int b0 = b_cta + g * N;
        const bool jop = j < HP;                       // this unit has a k row in the operands
        const bool jv = j < H;
        const int jc = jv ? j : H - 1;                 // clamped column for vI reads of padding rows
        // operand element (k, n): core matrices of 8 k rows x 8 sequences (128 bytes, a k row = 16 bytes); the N / 8 core
        // matrices of a k block follow each other, k blocks after that
        const int jk = jop ? j : 0;
        uint8_t *op = ops + (size_t)g * 4 * OPB + (size_t)(jk >> 3) * (16 * N) + (size_t)(n0 >> 3) * 128 + (jk & 7) * 16 + (n0 & 7) * 2;
        const float *vbase = vring + (size_t)g * 3 * N * VLD + n0 * VLD + jc;
        int len[NS];
#pragma unroll
        for (int n = 0; n < NS; n++) {
            const int bg = b0 + n0 + n;
            len[n] = (jv && bg < B) ? (lengths ? min(lengths[bg], T) : T) : 0;
        }
        float h[NS];
#pragma unroll
        for (int n = 0; n < NS; n++) h[n] = 0.0f;
        float *yp = y + ((long)(reverse ? T - 1 : 0) * B + b0 + n0) * ldy + j;
        const long ystep = (long)(reverse ? -1 : 1) * B * ldy;
        const bool live = b0 < B;                      // a group wholly past the batch end has no issuer either

        constexpr int NB_COUNT = (CW + 1) * 32;        // compute warps + the issuing warp
        const int nb_h = 1 + 2 * g, nb_rh = 2 + 2 * g;  // named barriers of this group (0 is __syncthreads)
        if (live) nbar_arrive(nb_h, NB_COUNT);         // h_{-1} = 0 is in place
        for (int s = 0; s < (live ? T : 0); s++) {
            const int t = reverse ? T - 1 - s : s;
            const float *vrow = vbase + (size_t)(s % 3) * N * VLD;
            const uint32_t par = (uint32_t)(s & 1);
            // ---- phase 1: r (needed at once) and z pre-activations ----
            wait_bar(&bar.v[s % 3], (uint32_t)((s / 3) & 1));
            float vr[NS];
#pragma unroll
            for (int n = 0; n < NS; n++) vr[n] = vrow[n * VLD + H];
            wait_bar(&bar.d1, par);
            if (warp == 0 && lane == 0) TRACE(4);
            tc_fence_after();
            uint32_t dr[NS], dz[NS];
            tmem_ld_cols<NS>(dcol + N, dr);
            tmem_ld_wait();
            if (warp == 0 && lane == 0) TRACE(5);
            if (jop) {
                float rh[NS];
#pragma unroll
                for (int n = 0; n < NS; n++) rh[n] = sigmoid_pre(fmaf(vr[n], -SLOIKA_LOG2E, __uint_as_float(dr[n]))) * h[n];
                uint32_t whi[NS / 2], wlo[NS / 2];
#pragma unroll
                for (int n = 0; n < NS; n += 2) split_pair(rh[n], rh[n + 1], whi[n / 2], wlo[n / 2]);
                store_halves<NS>(op + 2 * OPB, whi);
                store_halves<NS>(op + 3 * OPB, wlo);
            }
            fence_proxy_async();
            tc_fence_before();
            nbar_arrive(nb_rh, NB_COUNT);
            if (warp == 0 && lane == 0) TRACE(6);
            wait_bar(&bar.d1z, par);                   // long since complete: issued right behind r's MMAs
            tc_fence_after();
            tmem_ld_cols<NS>(dcol, dz);
            tmem_ld_wait();
            float z[NS], vc[NS];
#pragma unroll
            for (int n = 0; n < NS; n++) {
                z[n] = gate_denominator(fmaf(vrow[n * VLD], -SLOIKA_LOG2E, __uint_as_float(dz[n])));   // 1 + 2^(-z log2 e)
                vc[n] = vrow[n * VLD + 2 * H];
            }
            // ---- phase 2: candidate, blend, publish h_t ----
            wait_bar(&bar.d2, par);
            if (warp == 0 && lane == 0) TRACE(7);
            tc_fence_after();
            uint32_t dc[NS];
            tmem_ld_cols<NS>(dcol + 2 * N, dc);
            tmem_ld_wait();
            if (warp == 0 && lane == 0) TRACE(8);
#pragma unroll
            for (int n = 0; n < NS; n++) {
                const float hn = gru_blend(z[n], fmaf(vc[n], 2.0f * SLOIKA_LOG2E, __uint_as_float(dc[n])), h[n]);
                h[n] = t < len[n] ? hn : 0.0f;                   // ragged batch: state stays 0 outside the read
            }
            if (jop) {
                uint32_t whi[NS / 2], wlo[NS / 2];
#pragma unroll
                for (int n = 0; n < NS; n += 2) split_pair(h[n], h[n + 1], whi[n / 2], wlo[n / 2]);
                store_halves<NS>(op, whi);
                store_halves<NS>(op + OPB, wlo);
            }
            fence_proxy_async();
            tc_fence_before();
            nbar_arrive(nb_h, NB_COUNT);
            if (warp == 0 && lane == 0) TRACE(9);
            if (jv) {
#pragma unroll
                for (int n = 0; n < NS; n++)
                    if (b0 + n0 + n < B) yp[(long)n * ldy] = h[n];
            }
            yp += ystep;
        }
    } else {
        // ================= issuing warps: one per group =================
        const int g = warp - NCW;
        Bars &bar = bars[g];
        const int b0 = b_cta + g * N;
        const int nrows = min(N, B - b0);                                  // sequences of this group that exist
        const uint32_t op = smem_u32(ops + (size_t)g * 4 * OPB);
        float *vr0 = vring + (size_t)g * 3 * N * VLD;
        const uint32_t idesc = umma_idesc_f16_m128_bmn(N);
        const uint32_t dz = tmem_base + (uint32_t)(D_BASE + g * 3 * N), dr = dz + N, dc = dz + 2 * N;
        const uint64_t b_hh = umma_desc_mn_noswizzle(op, N), b_hl = umma_desc_mn_noswizzle(op + OPB, N);
        const uint64_t b_rh = umma_desc_mn_noswizzle(op + 2 * OPB, N), b_rl = umma_desc_mn_noswizzle(op + 3 * OPB, N);
        const uint32_t rowbytes = (uint32_t)((3 * H + 3) / 4 * 4) * 4u;     // <= ldv * 4: vI rows are 16-byte multiples
        auto load_vi = [&](int st) {                                       // elected lane: vI rows of scan step st
            if (st >= T) return;
            const int t = reverse ? T - 1 - st : st;
            uint64_t *vb = &bar.v[st % 3];
            mbar_arrive_expect_tx(vb, rowbytes * (uint32_t)nrows);
            const float *src = vI + ((long)t * B + b0) * ldv;
            float *dst = vr0 + (size_t)(st % 3) * N * VLD;
            for (int n = 0; n < nrows; n++) bulk_load_1d(dst + (size_t)n * VLD, src + (long)n * ldv, rowbytes, vb);
        };
        if (nrows > 0) {
            if (elect_one()) { load_vi(0); load_vi(1); load_vi(2); }
            __syncwarp();
            for (int s = 0; s < T; s++) {
                const uint32_t par = (uint32_t)(s & 1);
                nbar_sync(1 + 2 * g, (CW + 1) * 32);                       // h_{s-1} operand complete, step s-1 done with its vI slot
                if (g == 0 && lane == 0) TRACE(0);
                tc_fence_after();
                if (elect_one()) {
                    // r first: the dependent chain (r -> r*h -> candidate) waits for it; z is needed only at the blend
#pragma unroll
                    for (int kc = 0; kc < KC; kc++) {
                        const uint64_t koff = (uint64_t)(kc * 2 * N);      // 32 N bytes per K = 16 step
                        const uint32_t acol = tmem_base + (uint32_t)(kc * 8);
                        umma_f16_ts(dr, acol + 2 * ACOLS, b_hh + koff, idesc, kc != 0);
                        umma_f16_ts(dr, acol + 3 * ACOLS, b_hh + koff, idesc, true);
                        umma_f16_ts(dr, acol + 2 * ACOLS, b_hl + koff, idesc, true);
                    }
                    umma_commit(&bar.d1);
#pragma unroll
                    for (int kc = 0; kc < KC; kc++) {
                        const uint64_t koff = (uint64_t)(kc * 2 * N);
                        const uint32_t acol = tmem_base + (uint32_t)(kc * 8);
                        umma_f16_ts(dz, acol + 0 * ACOLS, b_hh + koff, idesc, kc != 0);
                        umma_f16_ts(dz, acol + 1 * ACOLS, b_hh + koff, idesc, true);
                        umma_f16_ts(dz, acol + 0 * ACOLS, b_hl + koff, idesc, true);
                    }
                    umma_commit(&bar.d1z);
                    if (g == 0) TRACE(1);
                    if (s >= 1) load_vi(s + 2);                            // slot (s-1) % 3 was released by step s-1
                }
                __syncwarp();
                nbar_sync(2 + 2 * g, (CW + 1) * 32);
                if (g == 0 && lane == 0) TRACE(2);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kc = 0; kc < KC; kc++) {
                        const uint64_t koff = (uint64_t)(kc * 2 * N);
                        const uint32_t acol = tmem_base + (uint32_t)(kc * 8);
                        umma_f16_ts(dc, acol + 4 * ACOLS, b_rh + koff, idesc, kc != 0);
                        umma_f16_ts(dc, acol + 5 * ACOLS, b_rh + koff, idesc, true);
                        umma_f16_ts(dc, acol + 4 * ACOLS, b_rl + koff, idesc, true);
                    }
                    umma_commit(&bar.d2);
                    if (g == 0) TRACE(3);
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NCW) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int HP, int G, int CW, int N>
static int launch(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths,
                  int T, int B, int H, int reverse, cudaStream_t st, Gate gate)
{
    size_t smem = 128 + (size_t)G * 4 * HP * 2 * N + (size_t)G * 3 * N * 3 * HP * 4 + (size_t)G * sizeof(Bars) + 64;
    // every CTA of this kernel owns the whole tensor memory of its SM: ask for more than half of the shared memory
    // so that a second one (another stream's batch) is placed on a free SM instead of stalling in tcgen05.alloc
    if (smem < 116 * 1024) smem = 116 * 1024;
    auto kern = gru_tc_kernel<HP, G, CW, N>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    const unsigned grid = (unsigned)ceil_div(B, G * N);
    kern<<<grid, G *(CW + 1) * 32, smem, st>>>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, gate);
    SLOIKA_RETURN_LAUNCH_STATUS();
}

template <int HP>
static int launch_hp(int g, int cw, int n, const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy,
                     const int32_t *lengths, int T, int B, int H, int reverse, cudaStream_t st, Gate gate)
{
#define TC_SHAPE(G_, CW_, N_) \
    if (g == G_ && cw == CW_ && n == N_) return launch<HP, G_, CW_, N_>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st, gate)
    TC_SHAPE(1, 4, 8); TC_SHAPE(1, 8, 8); TC_SHAPE(1, 16, 8);
    TC_SHAPE(2, 4, 8); TC_SHAPE(2, 8, 8);
    TC_SHAPE(4, 4, 8);
    TC_SHAPE(1, 8, 16); TC_SHAPE(2, 8, 16);
#undef TC_SHAPE
    return SLOIKA_ERR_UNSUPPORTED;
}

// tanh / sigmoid GRUs with H <= 128 whose vI rows are 16-byte aligned; SLOIKA_ERR_UNSUPPORTED otherwise (the caller
// falls back to gru_h16.cu).  `seqs_in_flight` is the number of sequences the caller keeps on the device at once
// (this batch times the batches it pipelines on other streams): it picks how many groups of 8 sequences share a
// CTA, i.e. whether the SMs are spread over one batch (latency) or packed (throughput).
// SLOIKA_B200_GRU_TC="G,CW[,N]" overrides (groups per CTA, compute warps per group, sequences per group).
int dispatch_gated(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths,
                   int T, int B, int H, int reverse, int act, int gate_act, long seqs_in_flight, cudaStream_t st, Gate gate)
{
    if (act != SLOIKA_ACT_TANH || gate_act != SLOIKA_ACT_SIGMOID) return SLOIKA_ERR_UNSUPPORTED;
    if (H > 128 || (ldv & 3) != 0 || ((uintptr_t)vI & 15) != 0) return SLOIKA_ERR_UNSUPPORTED;
    const long load = seqs_in_flight > B ? seqs_in_flight : B;
    int g = 1, cw = 8, n = 8;
    if (load > 8L * 148) { g = 2; cw = 8; }
    if (load > 16L * 148) { g = 2; cw = 8; n = 16; }       // 32 sequences per CTA, half the MMAs of 4 groups of 8
    if (const char *ov = getenv("SLOIKA_B200_GRU_TC")) {
        int og = 0, ocw = 0, on = 8;
        const int got = sscanf(ov, "%d,%d,%d", &og, &ocw, &on);
        if (got >= 2) { g = og; cw = ocw; n = got == 3 ? on : 8; }
    }
#define TC_CASE(HP_) return launch_hp<HP_>(g, cw, n, vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st, gate)
    if (H <= 32) TC_CASE(32);
    if (H <= 64) TC_CASE(64);
    if (H <= 96) TC_CASE(96);
    if (H <= 112) TC_CASE(112);
    TC_CASE(128);
#undef TC_CASE
}

int dispatch(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths, int T,
             int B, int H, int reverse, int act, int gate_act, long seqs_in_flight, cudaStream_t st)
{
    return dispatch_gated(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, seqs_in_flight, st, Gate{nullptr, 0u, 0});
}

}  // namespace gru5
}  // namespace sloika

#ifdef GRU_TC_TRACE
extern "C" int sloika_debug_gru_trace(long long *out)
{
    return (int)cudaMemcpyFromSymbol(out, sloika::gru5::g_trace, sizeof(long long) * 64 * 16);
}
#endif
