// GRU layer with the input projection INSIDE the recurrence launch: vI = x iW' + b never reaches HBM.
//
// Gru.step scanned by RNN.run (reference sloika/layers.py:1010-1021, :85-88; Reverse :1449-1450), including its first
// line `vI = T.tensordot(in_vec, self.iW) + self.b` (:1011).  Same results as sloika_gru_fwd.
//
// Why a cluster: the recurrence (gru_tc.cu) keeps sW / sW2 as fp16 hi / lo pairs in 288 of an SM's 512 tensor-memory
// columns; iW needs 288 more, and from shared memory an MMA costs 39 cycles instead of 9 (profiles/r2_umma_ts_probe.txt).
// So a thread-block cluster of three CTAs shares the work:
//   ranks 0, 1   RECURRENCE CTAs: G groups of N = 16 sequences each, exactly the scheme of gru_tc.cu (weights in their
//                SM's tensor memory, two dependent MMA phases per time step)
//   rank 2       PROJECTION CTA: iW in ITS SM's tensor memory; for each of the 2 G groups and each time step it stages
//                x_t (1-D TMA bulk copies), splits it into the fp16 hi / lo operand, runs the 54 TS-MMAs of
//                [z | r | c] = iW . x_t, adds the bias and writes the N x 3H values into a small ring in GLOBAL memory
// The ring (RING steps per group, ~300 KB per cluster, ~5 MB per batch) lives in L2: the projection CTA runs at most
// RING steps ahead, the recurrence CTAs read a slot with ld.global.cg as soon as it is published, and the slot is
// rewritten before it is ever evicted.  Hand-offs are cluster-scope mbarriers in the consumer's shared memory:
// "slot ready" (projection -> recurrence, after the writers' CTA barrier and a device fence) and "slot free"
// (recurrence -> projection, when the step that read it has published its h).  One projection CTA serves two
// recurrence CTAs because its step (216 MMAs, no dependent phases) is shorter than theirs.
// (Tried: a CTA-wide lock so that only one issuing warp at a time feeds the tensor pipe -- the two panels' / groups' MMAs
// then no longer alternate between accumulators.  The issue of a phase did not get faster -- an N = 32 instruction costs
// ~24 cycles here with or without interleaving, the projection CTA's tensor pipe is ~80 % busy -- and the lock's
// serialisation added 130 cycles per step: 3061 -> 3191.  Not kept.)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "gru_tc_common.cuh"

namespace sloika {
namespace gru6 {

using namespace tc;
using namespace gru5;

#ifdef GRU_TC_TRACE
__device__ long long g_ftrace[64 * 24];
#define FTRACE(block, slot) do { if (blockIdx.x == (block) && s >= 100 && s < 164) g_ftrace[(s - 100) * 24 + (slot)] = clock64(); } while (0)
#else
#define FTRACE(block, slot) do { } while (0)
#endif

constexpr int N = 16;                 // sequences per group
constexpr int RING = 8;               // projection steps in flight per group
constexpr int CWR = 8;                // compute warps per recurrence group (thread: unit j, 8 sequences)
constexpr int CWP = 4;                // compute warps per projection group (thread: row j, 16 sequences)
constexpr int VSLOTS = 3;             // recurrence: projected steps staged in shared memory
constexpr int NCLUSTER = 3;

struct FBars {                        // one per group index (recurrence CTAs use the first G, the projection CTA all 2 G)
    uint64_t d1, d1z, d2;             // recurrence: MMA phases complete
    uint64_t ready[RING];             // recurrence: ring slot published by the projection CTA (remote arrive)
    uint64_t v[VSLOTS];               // recurrence: ring slot copied into this CTA's shared memory (TMA byte count)
    uint64_t pd;                      // projection: MMAs complete
    uint64_t xs[2];                   // projection: x_t staged (TMA byte count)
    uint64_t freeb[RING];             // projection: ring slot consumed by the recurrence CTA (remote arrive)
    uint64_t pad;
};

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster.  RELAXED on purpose: a
// release at cluster scope is MEMBAR.ALL.GPU + ERRBAR + CCTL.IVALL in SASS and cost 2000-3000 cycles per time step here
// (profiles/r2_gru_fused_trace.txt).  Nothing needs it: the data a "ready" arrival announces was written by a bulk
// async store whose completion the same thread has just waited for (cp.async.bulk.wait_group), and a "free" arrival
// follows reads that have completed (the consumers waited for the copy's mbarrier and joined the issuing warp at a CTA
// barrier).  The waiting side still acquires at cluster scope.
__device__ __forceinline__ void remote_arrive(uint64_t *bar, uint32_t rank) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// wait with acquire at cluster scope (the arrival came from another CTA of the cluster)
__device__ __forceinline__ void wait_bar_cluster(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1, 200000;\n\t"
        "@p bra DONEC_%=;\n\t"
        "bra WAITC_%=;\n\t"
        "DONEC_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// HP: hidden size padded (32 / 64 / 96); IP: input size padded (32 / 64 / 96); G groups per recurrence CTA.
template <int HP, int IP, int G>
__global__ void __launch_bounds__(2 * G *(CWP + 1) * 32, 1)
gru_fused_kernel(const float *__restrict__ x, long ldx, const float *__restrict__ iW, const float *__restrict__ bias,
                 const float *__restrict__ sW, const float *__restrict__ sW2, float *__restrict__ y, long ldy,
                 float *__restrict__ ring, const int32_t *__restrict__ lengths, int T, int B, int I, int H, int reverse,
                 const Gate gate)
{
    if (gate_closed(gate)) return;                // every CTA of every cluster alike: nobody is left waiting
    constexpr int KC = HP / 16, KCP = IP / 16;
    constexpr int ACOLS = HP / 2, ACOLS_P = IP / 2;
    constexpr int OPB = HP * 2 * N;               // bytes of one recurrence operand array [k][N]
    constexpr int OPB_P = IP * 2 * N;             // bytes of one projection operand array [k][N]
    constexpr int VLD = 3 * HP;                   // floats per ring row
    constexpr int NTHREADS = 2 * G * (CWP + 1) * 32;
    constexpr int NWARPS = NTHREADS / 32;
    static_assert(6 * ACOLS + G * 3 * N <= 512 && 6 * ACOLS_P + 2 * G * 3 * N <= 512, "tensor memory: 512 columns");
    static_assert(G * (CWR + 1) <= NWARPS, "block size is set by the projection CTA");

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    // one layout for all three CTAs (remote mbarrier addresses are local offsets): barriers first
    FBars *bars = reinterpret_cast<FBars *>(smem);                      // [2 G]
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(bars + 2 * G);
    uint8_t *work = smem + ((sizeof(FBars) * 2 * G + 64 + 127) / 128) * 128;
    // recurrence CTA: operands [G][4][OPB], then projected steps [G][VSLOTS][N][VLD] floats;
    // projection CTA: operands [2G][2 buffers][hi, lo][OPB_P], then x staging [2G][2][N][IP] floats, then result staging
    // [2G][N][VLD] floats
    uint8_t *ops = work;
    float *vring = reinterpret_cast<float *>(work + (size_t)G * 4 * OPB);
    float *xstage = reinterpret_cast<float *>(work + (size_t)2 * G * 4 * OPB_P);
    float *ostage = reinterpret_cast<float *>(work + (size_t)2 * G * 4 * OPB_P + (size_t)2 * G * 2 * N * IP * 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    const bool is_proj = rank == 2;
    const int cluster_id = blockIdx.x / NCLUSTER;
    const int b_cl = cluster_id * (2 * G * N);                          // first sequence of this cluster
    float *ring_cl = ring + (size_t)cluster_id * (2 * G) * RING * N * VLD;

    // ---------------- prologue ----------------
    if (tid == 0) {
        for (int g = 0; g < 2 * G; g++) {
            mbar_init(&bars[g].d1, 1); mbar_init(&bars[g].d1z, 1); mbar_init(&bars[g].d2, 1); mbar_init(&bars[g].pd, 1);
            mbar_init(&bars[g].xs[0], 1); mbar_init(&bars[g].xs[1], 1);
            for (int i = 0; i < RING; i++) { mbar_init(&bars[g].ready[i], 1); mbar_init(&bars[g].freeb[i], 1); }
            for (int i = 0; i < VSLOTS; i++) mbar_init(&bars[g].v[i], 1);
        }
        mbar_fence_init();
    }
    if (warp == NWARPS - 1) tmem_alloc(tmem_base_s, 512);
    {
        const size_t wbytes = is_proj ? (size_t)2 * G * 4 * OPB_P + (size_t)2 * G * 2 * N * IP * 4 + (size_t)2 * G * N * VLD * 4
                                      : (size_t)G * 4 * OPB + (size_t)G * VSLOTS * N * VLD * 4;
        uint32_t *z = reinterpret_cast<uint32_t *>(work);
        for (int e = tid; e < (int)(wbytes / 4); e += NTHREADS) z[e] = 0u;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    // weights -> TMEM: lane = gate row j; recurrence CTAs: sW (z, r) and sW2 (c) pre-scaled for ex2 as in gru_tc.cu;
    // projection CTA: iW (z, r, c rows), unscaled (the consumer scales vI when it adds it)
    {
        const int q = warp & 3, j = 32 * q + lane;
        const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
        const int kcs = is_proj ? KCP : KC, acols = is_proj ? ACOLS_P : ACOLS, kdim = is_proj ? I : H;
        for (int c = warp >> 2; c < 3 * kcs; c += NWARPS / 4) {
            const int m = c / kcs, kc = c - m * kcs;
            const float *row = is_proj ? iW + (long)(m * H + j) * I : (m < 2 ? sW + (long)(m * H + j) * H : sW2 + (long)j * H);
            const float gs = is_proj ? 1.0f : (m < 2 ? -SLOIKA_LOG2E : 2.0f * SLOIKA_LOG2E);
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int k = 16 * kc + 2 * i;
                const float w0 = (j < H && k < kdim) ? gs * __ldg(row + k) : 0.0f;
                const float w1 = (j < H && k + 1 < kdim) ? gs * __ldg(row + k + 1) : 0.0f;
                const __half2 h2 = __floats2half2_rn(w0, w1);
                const float2 hb = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(w0 - hb.x, w1 - hb.y);
                hi[i] = *reinterpret_cast<const uint32_t *>(&h2);
                lo[i] = *reinterpret_cast<const uint32_t *>(&l2);
            }
            if (warp < (NWARPS / 4) * 4) {
                tmem_st_32x32b_x8(tmem_base + lane_addr + (uint32_t)((2 * m) * acols + kc * 8), hi);
                tmem_st_32x32b_x8(tmem_base + lane_addr + (uint32_t)((2 * m + 1) * acols + kc * 8), lo);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync_all();                                   // every CTA's barriers are initialised before anyone arrives remotely

    if (!is_proj) {
        // =====================================================================================================
        // RECURRENCE CTA (rank 0 / 1): groups g = 0 .. G-1 <-> projection groups pg = rank * G + g
        // =====================================================================================================
        constexpr int NCW = G * CWR;
        constexpr int D_BASE = 6 * ACOLS;
        constexpr int NS = N / (CWR / 4);                 // 8 sequences per compute thread
        if (warp < NCW) {
            const int g = warp / CWR, wg = warp - g * CWR, q = wg & 3, j = 32 * q + lane;
            const int n0 = (wg >> 2) * NS;
            const int pg = (int)rank * G + g;
            const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
            const uint32_t dcol = tmem_base + lane_addr + (uint32_t)(D_BASE + g * 3 * N + n0);
            FBars &bar = bars[g];
            const int b0 = b_cl + pg * N;
            const bool jop = j < HP, jv = j < H;
            const int jc = jv ? j : H - 1;
            const int jk = jop ? j : 0;
            uint8_t *op = ops + (size_t)g * 4 * OPB + (size_t)(jk >> 3) * (16 * N) + (size_t)(n0 >> 3) * 128 + (jk & 7) * 16 + (n0 & 7) * 2;
            const float *vbase = vring + (size_t)g * VSLOTS * N * VLD + (size_t)n0 * VLD + jc;
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
            const bool live = b0 < B;
            constexpr int NB_COUNT = (CWR + 1) * 32;
            const int nb_h = 1 + 2 * g, nb_rh = 2 + 2 * g;
            if (live) nbar_arrive(nb_h, NB_COUNT);
            for (int s = 0; s < (live ? T : 0); s++) {
                const int t = reverse ? T - 1 - s : s;
                const uint32_t par = (uint32_t)(s & 1);
                // the projection of this step: copied from the ring by the issuing warp two steps ago
                const float *vrow = vbase + (size_t)(s % VSLOTS) * N * VLD;
                wait_bar(&bar.v[s % VSLOTS], (uint32_t)((s / VSLOTS) & 1));
                if (warp == 0 && lane == 0) FTRACE(0, 5);
                float vr[NS];
#pragma unroll
                for (int n = 0; n < NS; n++) vr[n] = vrow[n * VLD + H];
                // ---- phase 1 ----
                wait_bar(&bar.d1, par);
                if (warp == 0 && lane == 0) FTRACE(0, 6);
                tc_fence_after();
                uint32_t dr[NS], dz[NS];
                tmem_ld_cols<NS>(dcol + N, dr);
                tmem_ld_wait();
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
                if (warp == 0 && lane == 0) FTRACE(0, 7);
                wait_bar(&bar.d1z, par);
                tc_fence_after();
                tmem_ld_cols<NS>(dcol, dz);
                tmem_ld_wait();
                float z[NS], vc[NS];
#pragma unroll
                for (int n = 0; n < NS; n++) {
                    z[n] = gate_denominator(fmaf(vrow[n * VLD], -SLOIKA_LOG2E, __uint_as_float(dz[n])));   // 1 + 2^(-z log2 e)
                    vc[n] = vrow[n * VLD + 2 * H];
                }
                // ---- phase 2 ----
                wait_bar(&bar.d2, par);
                if (warp == 0 && lane == 0) FTRACE(0, 8);
                tc_fence_after();
                uint32_t dc[NS];
                tmem_ld_cols<NS>(dcol + 2 * N, dc);
                tmem_ld_wait();
#pragma unroll
                for (int n = 0; n < NS; n++) {
                    const float hn = gru_blend(z[n], fmaf(vc[n], 2.0f * SLOIKA_LOG2E, __uint_as_float(dc[n])), h[n]);
                    h[n] = t < len[n] ? hn : 0.0f;
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
                if (warp == 0 && lane == 0) FTRACE(0, 9);
                if (jv) {
#pragma unroll
                    for (int n = 0; n < NS; n++)
                        if (b0 + n0 + n < B) yp[(long)n * ldy] = h[n];
                }
                yp += ystep;
            }
        } else if (warp < NCW + G) {
            const int g = warp - NCW;
            const int pg = (int)rank * G + g;
            FBars &bar = bars[g];
            const int b0 = b_cl + pg * N;
            const uint32_t op = smem_u32(ops + (size_t)g * 4 * OPB);
            const uint32_t idesc = umma_idesc_f16_m128_bmn(N);
            const uint32_t dz = tmem_base + (uint32_t)(D_BASE + g * 3 * N), dr = dz + N, dc = dz + 2 * N;
            const uint64_t b_hh = umma_desc_mn_noswizzle(op, N), b_hl = umma_desc_mn_noswizzle(op + OPB, N);
            const uint64_t b_rh = umma_desc_mn_noswizzle(op + 2 * OPB, N), b_rl = umma_desc_mn_noswizzle(op + 3 * OPB, N);
            const float *ring_g = ring_cl + (size_t)pg * RING * N * VLD;
            float *vr0 = vring + (size_t)g * VSLOTS * N * VLD;
            auto load_vi = [&](int st) {                  // elected lane: one bulk copy, ring slot (L2) -> shared memory
                if (st >= T) return;
                wait_bar_cluster(&bar.ready[st % RING], (uint32_t)((st / RING) & 1));
                fence_proxy_async();                      // the slot was written through the generic proxy (other SM)
                uint64_t *vb = &bar.v[st % VSLOTS];
                mbar_arrive_expect_tx(vb, (uint32_t)(N * VLD * 4));
                bulk_load_1d(vr0 + (size_t)(st % VSLOTS) * N * VLD, ring_g + (size_t)(st % RING) * N * VLD, (uint32_t)(N * VLD * 4), vb);
            };
            if (b0 < B) {
                if (elect_one()) { load_vi(0); load_vi(1); load_vi(2); }
                __syncwarp();
                for (int s = 0; s < T; s++) {
                    nbar_sync(1 + 2 * g, (CWR + 1) * 32);                  // h_{s-1} published: step s-1 is done with its vI slot
                    if (g == 0 && lane == 0) FTRACE(0, 0);
                    tc_fence_after();
                    if (elect_one()) {
                        if (s >= 1) remote_arrive(&bars[pg].freeb[(s - 1) % RING], 2);   // its ring slot was copied long ago
#pragma unroll
                        for (int kc = 0; kc < KC; kc++) {
                            const uint64_t koff = (uint64_t)(kc * 2 * N);
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
                        if (g == 0) FTRACE(0, 1);
                        if (s >= 1) load_vi(s + 2);                        // into the shared-memory slot step s-1 released
                        if (g == 0) FTRACE(0, 2);
                    }
                    __syncwarp();
                    nbar_sync(2 + 2 * g, (CWR + 1) * 32);
                    if (g == 0 && lane == 0) FTRACE(0, 3);
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
                        if (g == 0) FTRACE(0, 4);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // =====================================================================================================
        // PROJECTION CTA (rank 2): groups pg = 0 .. 2G-1
        // =====================================================================================================
        // The G groups of one recurrence CTA form a PANEL of PW = G N sequences: one MMA covers the panel (the tensor pipe
        // runs an N = 32 instruction in 16 cycles but an N = 16 one in 9-14, profiles/r2_umma_ts_probe.txt), everything
        // around the MMAs (staging, accumulator read-out, ring slot) stays per group.
        constexpr int NPG = 2 * G;                        // groups
        constexpr int NCW = NPG * CWP;
        constexpr int PW = G * N;                         // sequences per panel
        constexpr int OPB_PP = IP * 2 * PW;               // bytes of one panel operand array [k][PW]
        constexpr int D_BASE = 6 * ACOLS_P;
        static_assert(4 * OPB_PP == G * 4 * OPB_P, "operand footprint");
        const uint32_t rowbytes = (uint32_t)((I + 3) / 4 * 4) * 4u;        // <= ldx * 4: x rows are 16-byte multiples
        if (warp < NCW) {
            const int pg = warp / CWP, q = warp & 3, j = 32 * q + lane;
            const int pp = pg / G, gh = pg - pp * G;      // panel (= destination CTA rank), group inside it
            FBars &bar = bars[pg];                        // group level: freeb
            FBars &pbar = bars[pp * G];                   // panel level: pd, xs
            const int b0 = b_cl + pg * N;
            const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
            const uint32_t dcol = tmem_base + lane_addr + (uint32_t)(D_BASE + pp * 3 * PW + gh * N);
            const bool kop = j < IP;                      // this thread owns operand row k = j
            const bool jv = j < H;
            const int jk = kop ? j : 0;
            uint8_t *xop = ops + (size_t)pp * 4 * OPB_PP + (size_t)(jk >> 3) * (16 * PW) + (size_t)((gh * N) >> 3) * 128 + (jk & 7) * 16;
            const float *xs0 = xstage + (size_t)pp * 2 * PW * IP + (size_t)gh * N * IP + jk;
            float *ring_g = ring_cl + (size_t)pg * RING * N * VLD;
            float *ostage_g = ostage + (size_t)pg * N * VLD;
            float *orow = ostage_g + (jv ? j : 0);
            const bool leader = q == 0 && lane == 0;
            const float bz = jv ? __ldg(bias + j) : 0.0f, br = jv ? __ldg(bias + H + j) : 0.0f, bc = jv ? __ldg(bias + 2 * H + j) : 0.0f;
            const uint32_t dst_rank = (uint32_t)pp;
            constexpr int NB_X = (G * CWP + 1) * 32;
            const int nb_x = 1 + pp, nb_w = 3 + pg, nb_o = 3 + NPG + pg;
            const bool plive = b_cl + pp * PW < B;        // the panel exists
            const bool glive = b0 < B;                    // this group exists (its recurrence group consumes the ring)
            auto stage_operand = [&](int st) {            // x of scan step st -> fp16 hi / lo operand [k][n], buffer st & 1
                wait_bar(&pbar.xs[st & 1], (uint32_t)((st >> 1) & 1));
                if (kop) {
                    const float *xs = xs0 + (size_t)(st & 1) * PW * IP;
                    uint8_t *dst = xop + (size_t)(st & 1) * 2 * OPB_PP;
                    uint32_t whi[N / 2], wlo[N / 2];
#pragma unroll
                    for (int n = 0; n < N; n += 2) split_pair(xs[n * IP], xs[(n + 1) * IP], whi[n / 2], wlo[n / 2]);
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(whi[0], whi[1], whi[2], whi[3]);
                    *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(whi[4], whi[5], whi[6], whi[7]);
                    *reinterpret_cast<uint4 *>(dst + OPB_PP) = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]);
                    *reinterpret_cast<uint4 *>(dst + OPB_PP + 128) = make_uint4(wlo[4], wlo[5], wlo[6], wlo[7]);
                }
                fence_proxy_async();
            };
            if (plive && T > 0) {
                if (glive) stage_operand(0);
                tc_fence_before();
                nbar_arrive(nb_x, NB_X);
                if (glive && T > 1) stage_operand(1);     // the other operand buffer: ready before the accumulators are
            }
            for (int s = 0; s < (plive ? T : 0); s++) {
                // ---- [z | r | c] = iW . x_s is complete ----
                wait_bar(&pbar.pd, (uint32_t)(s & 1));
                if (warp == 0 && lane == 0) FTRACE(2, 13);
                if (!glive) {                             // no sequences here: only keep the panel's barrier counts whole
                    if (s + 1 < T) nbar_arrive(nb_x, NB_X);
                    continue;
                }
                tc_fence_after();
                uint32_t dz[N], dr[N], dc[N];
                tmem_ld_32x32b_x16(dcol, dz);
                tmem_ld_32x32b_x16(dcol + PW, dr);
                tmem_ld_32x32b_x16(dcol + 2 * PW, dc);
                tmem_ld_wait();
                tc_fence_before();
                if (s + 1 < T) nbar_arrive(nb_x, NB_X);   // accumulators read: the MMAs of step s + 1 may go
                if (warp == 0 && lane == 0) FTRACE(2, 15);
                // ---- while they run: + bias -> staging buffer -> ring slot (one bulk store per group) ... ----
                nbar_sync(nb_o, CWP * 32);                // the store of step s - 1 has read the staging buffer (warp 0 got here)
                if (jv) {
#pragma unroll
                    for (int n = 0; n < N; n++) {
                        orow[n * VLD] = __uint_as_float(dz[n]) + bz;
                        orow[n * VLD + H] = __uint_as_float(dr[n]) + br;
                        orow[n * VLD + 2 * H] = __uint_as_float(dc[n]) + bc;
                    }
                }
                fence_proxy_async();
                if (warp == 0 && lane == 0) FTRACE(2, 17);
                nbar_sync(nb_w, CWP * 32);                // the whole slot is staged
                if (warp == 0 && lane == 0) FTRACE(2, 18);
                if (leader) {
                    const int slot = s % RING;
                    if (s >= RING) wait_bar_cluster(&bar.freeb[slot], (uint32_t)(((s / RING) - 1) & 1));
                    if (warp == 0) FTRACE(2, 16);
                    bulk_store_1d(ring_g + (size_t)slot * N * VLD, ostage_g, (uint32_t)(N * VLD * 4));
                    bulk_commit();
                    bulk_wait_all1();                     // the store of step s - 1 is complete: publish that slot
                    if (s >= 1) remote_arrive(&bars[gh].ready[(s - 1) % RING], dst_rank);
                }
                // ---- ... and the operand of step s + 2 into the buffer the MMAs of step s have released ----
                if (s + 2 < T) stage_operand(s + 2);
                if (warp == 0 && lane == 0) FTRACE(2, 14);
                if (leader) bulk_wait_read();             // before warp 0 joins nb_o: this step's store has read the staging buffer
                if (warp == 0 && lane == 0) FTRACE(2, 19);
            }
            if (leader && glive && T > 0) {
                bulk_wait_all();
                remote_arrive(&bars[gh].ready[(T - 1) % RING], dst_rank);
            }
        } else if (warp < NCW + 2) {
            const int pp = warp - NCW;                    // one issuing warp per panel
            FBars &pbar = bars[pp * G];
            const int b0 = b_cl + pp * PW;
            const int nrows = min(PW, B - b0);
            const uint32_t xop = smem_u32(ops + (size_t)pp * 4 * OPB_PP);
            const uint32_t idesc = umma_idesc_f16_m128_bmn(PW);
            const uint32_t dz = tmem_base + (uint32_t)(D_BASE + pp * 3 * PW), dr = dz + PW, dc = dz + 2 * PW;
            float *xs0 = xstage + (size_t)pp * 2 * PW * IP;
            const bool one_copy = ldx == IP && I == IP;   // the panel's rows of a time step are one contiguous block
            auto load_x = [&](int st) {                   // elected lane: x rows of scan step st
                if (st >= T) return;
                const int t = reverse ? T - 1 - st : st;
                uint64_t *xb = &pbar.xs[st & 1];
                mbar_arrive_expect_tx(xb, rowbytes * (uint32_t)nrows);
                const float *src = x + ((long)t * B + b0) * ldx;
                float *dst = xs0 + (size_t)(st & 1) * PW * IP;
                if (one_copy) bulk_load_1d(dst, src, rowbytes * (uint32_t)nrows, xb);
                else
                    for (int n = 0; n < nrows; n++) bulk_load_1d(dst + (size_t)n * IP, src + (long)n * ldx, rowbytes, xb);
            };
            if (nrows > 0) {
                if (elect_one()) { load_x(0); load_x(1); }
                __syncwarp();
                for (int s = 0; s < T; s++) {
                    nbar_sync(1 + pp, (G * CWP + 1) * 32);    // accumulators free, operand of step s staged, staging of x_{s+1} read
                    if (pp == 0 && lane == 0) FTRACE(2, 10);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t xb = xop + (uint32_t)(s & 1) * 2 * OPB_PP;
                        const uint64_t b_xh = umma_desc_mn_noswizzle(xb, PW), b_xl = umma_desc_mn_noswizzle(xb + OPB_PP, PW);
                        // gate by gate: consecutive MMAs into ONE accumulator run at the pipe's rate, a change of
                        // accumulator after every instruction does not (profiles/r2_gru_fused_trace.txt)
#pragma unroll
                        for (int m = 0; m < 3; m++) {
                            const uint32_t dm = dz + (uint32_t)(m * PW);
#pragma unroll
                            for (int kc = 0; kc < KCP; kc++) {
                                const uint64_t koff = (uint64_t)(kc * 2 * PW);
                                const uint32_t acol = tmem_base + (uint32_t)(kc * 8 + 2 * m * ACOLS_P);
                                umma_f16_ts(dm, acol, b_xh + koff, idesc, kc != 0);
                                umma_f16_ts(dm, acol + ACOLS_P, b_xh + koff, idesc, true);
                                umma_f16_ts(dm, acol, b_xl + koff, idesc, true);
                            }
                        }
                        umma_commit(&pbar.pd);
                        if (pp == 0) FTRACE(2, 11);
                        load_x(s + 2);                    // its staging buffer held x_s, which was split a step ago
                        if (pp == 0) FTRACE(2, 12);
                    }
                    __syncwarp();
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                   // nobody leaves while a neighbour may still arrive on its barriers
    if (warp == NWARPS - 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int HP, int IP, int G>
static size_t smem_bytes()
{
    const size_t rec = (size_t)G * 4 * HP * 2 * N + (size_t)G * VSLOTS * N * 3 * HP * 4;
    const size_t proj = (size_t)2 * G * 4 * IP * 2 * N + (size_t)2 * G * 2 * N * IP * 4 + (size_t)2 * G * N * 3 * HP * 4;
    size_t s = 128 + ((sizeof(FBars) * 2 * G + 64 + 127) / 128) * 128 + (rec > proj ? rec : proj);
    if (s < 116 * 1024) s = 116 * 1024;                   // one CTA per SM: each owns its SM's tensor memory
    return s;
}

template <int HP, int IP, int G>
static int launch(const float *x, long ldx, const float *iW, const float *bias, const float *sW, const float *sW2, float *y,
                  long ldy, float *ring, const int32_t *lengths, int T, int B, int I, int H, int reverse, cudaStream_t st, Gate gate)
{
    const size_t smem = smem_bytes<HP, IP, G>();
    auto kern = gru_fused_kernel<HP, IP, G>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(NCLUSTER * ceil_div(B, 2 * G * N)));
    cfg.blockDim = dim3(2 * G * (CWP + 1) * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    err = cudaLaunchKernelEx(&cfg, kern, x, ldx, iW, bias, sW, sW2, y, ldy, ring, lengths, T, B, I, H, reverse, gate);
    if (err != cudaSuccess) return (int)err;
    SLOIKA_RETURN_LAUNCH_STATUS();
}

constexpr int GF = 2;                 // groups per recurrence CTA

static size_t ring_bytes(int B, int HP)
{
    const size_t clusters = (size_t)ceil_div(B, 2 * GF * N);
    return clusters * (2 * GF) * RING * N * (size_t)(3 * HP) * sizeof(float);
}

}  // namespace gru6
}  // namespace sloika

using namespace sloika;

static int padded16(int v) { return (v + 15) / 16 * 16; }
static int hp_of(int H) { return H <= 32 ? 32 : H <= 64 ? 64 : 96; }

#ifdef GRU_TC_TRACE
extern "C" int sloika_debug_gru_fused_trace(long long *out)
{
    return (int)cudaMemcpyFromSymbol(out, sloika::gru6::g_ftrace, sizeof(long long) * 64 * 24);
}
#endif

extern "C" size_t sloika_gru_fused_workspace_bytes(int B, int H)
{
    if (B <= 0 || H <= 0 || H > 96) return 0;
    return gru6::ring_bytes(B, hp_of(H));
}

static int fused_fwd(const float *x, long ldx, const float *iW, const float *sW, const float *sW2, const float *b, float *y,
                     long ldy, void *ws, size_t ws_bytes, const int32_t *lengths, int T, int B, int I, int H, int reverse,
                     int act, int gate_act, void *stream, gru5::Gate gate)
{
    if (!x || !iW || !sW || !sW2 || !b || !y || T < 0 || B <= 0 || I <= 0 || H <= 0 || ldx < I || ldy < H) return SLOIKA_ERR_ARG;
    if (act != SLOIKA_ACT_TANH || gate_act != SLOIKA_ACT_SIGMOID) return SLOIKA_ERR_UNSUPPORTED;
    // iW must fit the projection CTA's tensor memory beside 4 groups of accumulators (I <= 96) and its double-buffered
    // result staging the shared memory (H <= 96)
    if (H > 96 || I > 96 || (ldx & 3) != 0 || ((uintptr_t)x & 15) != 0) return SLOIKA_ERR_UNSUPPORTED;
    if (T == 0) return SLOIKA_OK;
    if (!ws || ws_bytes < sloika_gru_fused_workspace_bytes(B, H)) return SLOIKA_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float *ring = static_cast<float *>(ws);
    const int HP = hp_of(H), IP = padded16(I) <= 32 ? 32 : padded16(I) <= 64 ? 64 : 96;
#define FUSED_CASE(HP_, IP_) \
    if (HP == HP_ && IP == IP_) return gru6::launch<HP_, IP_, gru6::GF>(x, ldx, iW, b, sW, sW2, y, ldy, ring, lengths, T, B, I, H, reverse, st, gate)
    FUSED_CASE(32, 32); FUSED_CASE(32, 64); FUSED_CASE(32, 96);
    FUSED_CASE(64, 32); FUSED_CASE(64, 64); FUSED_CASE(64, 96);
    FUSED_CASE(96, 32); FUSED_CASE(96, 64); FUSED_CASE(96, 96);
#undef FUSED_CASE
    return SLOIKA_ERR_UNSUPPORTED;
}

extern "C" int sloika_gru_fused_fwd(const float *x, long ldx, const float *iW, const float *sW, const float *sW2,
                                    const float *b, float *y, long ldy, void *ws, size_t ws_bytes, const int32_t *lengths,
                                    int T, int B, int I, int H, int reverse, int act, int gate_act, void *stream)
{
    return fused_fwd(x, ldx, iW, sW, sW2, b, y, ldy, ws, ws_bytes, lengths, T, B, I, H, reverse, act, gate_act, stream,
                     gru5::Gate{nullptr, 0u, 0});
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

// The layer for an input whose range is known only on the device (`absmax`: max |x| written by the producing kernel, e.g.
// sloika_conv1d_fwd_ex for an elu convolution): BOTH forms are enqueued and the device word picks one --
//   max |x| <  limit : the fused launch above (fp16 hi / lo operands are exact enough only inside the fp16 range)
//   max |x| >= limit : tf32-split projection GEMM into `vI`, then the recurrence kernel
// The CTAs of the form that is ruled out return at once (a few microseconds per launch).
extern "C" int sloika_gru_fwd_gated(const float *x, long ldx, const float *iW, const float *sW, const float *sW2,
                                    const float *b, float *y, long ldy, float *vI, long ldv, void *ws, size_t ws_bytes,
                                    const int32_t *lengths, int T, int B, int I, int H, int reverse, int act, int gate_act,
                                    long seqs_in_flight, const float *absmax, float limit, void *stream)
{
    if (!absmax || !vI || !(limit > 0.0f) || ldv < 3L * H) return SLOIKA_ERR_ARG;
    if ((ldv & 3) != 0 || ((uintptr_t)vI & 15) != 0 || (long)T * B < 128) return SLOIKA_ERR_UNSUPPORTED;
    unsigned limit_bits;
    memcpy(&limit_bits, &limit, sizeof(limit_bits));
    const unsigned *word = reinterpret_cast<const unsigned *>(absmax);
    int rc = fused_fwd(x, ldx, iW, sW, sW2, b, y, ldy, ws, ws_bytes, lengths, T, B, I, H, reverse, act, gate_act, stream,
                       gru5::Gate{word, limit_bits, 1});
    if (rc != SLOIKA_OK || T == 0) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = gemm_tc::launch(x, ldx, iW, b, vI, ldv, (long)T * B, I, 3 * H, SLOIKA_ACT_LINEAR, nullptr, 0, false, st, word, limit_bits, 2);
    if (rc != SLOIKA_OK) return rc;
    return gru5::dispatch_gated(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, act, gate_act, seqs_in_flight, st,
                                gru5::Gate{word, limit_bits, 2});
}
