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

namespace sloika {
namespace gru5 {

using namespace tc;

// mbarrier wait that turns a protocol bug into a trap after ~2 s instead of a hung device
__device__ __forceinline__ void wait_bar(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 20000;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && spins > 100000u) __trap();
    }
}

template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[N]) {
    if constexpr (N == 8) tmem_ld_32x32b_x8(taddr, r);
    else tmem_ld_32x32b_x16(taddr, r);
}

// x -> fp16 hi (low half) | fp16 lo (high half): hi = x with the low 13 mantissa bits cleared (exactly representable
// in fp16 for |x| in [2^-14, 65504]), lo = x - hi (exact), both converted by one packed cvt.  Below 2^-14 the fp16
// subnormal rounding of hi is not compensated: an absolute error <= 2^-25, far under the fp32 noise of the sums.
__device__ __forceinline__ uint32_t split_pack(float x) {
    const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    const __half2 p = __floats2half2_rn(hi, x - hi);
    return *reinterpret_cast<const uint32_t *>(&p);
}

struct Bars {                 // per group
    uint64_t h, d1, rh, d2, v[3];
    uint64_t pad;
};

template <int HP, int N, int G>
__global__ void __launch_bounds__(G * 160, 1)
gru_tc_kernel(const float *__restrict__ vI, long ldv, const float *__restrict__ sW, const float *__restrict__ sW2,
              float *__restrict__ y, long ldy, const int32_t *__restrict__ lengths, int T, int B, int H, int reverse)
{
    constexpr int KC = HP / 16;                   // K = 16 chunks
    constexpr int NKB = (HP + 31) / 32;           // K blocks of 32 halves (one 64-byte swizzle row each)
    constexpr int ACOLS = HP / 2;                 // TMEM columns of one A tile
    constexpr int D_BASE = 6 * ACOLS;             // accumulators behind the 6 weight tiles
    constexpr int OPB = NKB * N * 64;             // bytes of one operand array (N sequences x K halves, swizzled)
    constexpr int NCW = 4 * G;                    // compute warps
    static_assert(D_BASE + G * 3 * N <= 512, "tensor memory: 512 columns");
    static_assert(N == 8 || N == 16, "sequences per group");

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int vld = (3 * H + 3) / 4 * 4;                               // floats per staged vI row (16-byte multiple)
    uint8_t *ops = smem;                                               // [G][4][OPB]: h hi, h lo, r*h hi, r*h lo
    float *vring = reinterpret_cast<float *>(ops + (size_t)G * 4 * OPB);   // [G][3][N][vld]
    Bars *bars = reinterpret_cast<Bars *>(vring + (size_t)G * 3 * N * vld);
    int *lens_s = reinterpret_cast<int *>(bars + G);                   // [G][N]
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(lens_s + G * N);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthreads = G * 160;
    const int b_cta = blockIdx.x * (G * N);

    // ---------------- prologue ----------------
    if (tid == 0) {
        for (int g = 0; g < G; g++) {
            mbar_init(&bars[g].h, 128); mbar_init(&bars[g].rh, 128);
            mbar_init(&bars[g].d1, 1); mbar_init(&bars[g].d2, 1);
            for (int i = 0; i < 3; i++) mbar_init(&bars[g].v[i], 1);
        }
        mbar_fence_init();
    }
    if (warp == NCW) tmem_alloc(tmem_base_s, 512);
    {
        uint32_t *z = reinterpret_cast<uint32_t *>(smem);
        const int nz = (int)(((uint8_t *)bars - smem) / 4);
        for (int e = tid; e < nz; e += nthreads) z[e] = 0u;
    }
    for (int e = tid; e < G * N; e += nthreads) {
        const int bg = b_cta + e;
        lens_s[e] = bg < B ? (lengths ? min(lengths[bg], T) : T) : 0;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp < NCW) {
        // weights -> TMEM (split once): lane = gate row j, tile 2m + p (m: z, r, candidate; p: hi, lo), chunk kc at
        // columns kc*8 .. kc*8+7, column i = k pair (16kc + 2i, 16kc + 2i + 1).  Chunks are dealt over the groups.
        const int q = warp & 3, j = 32 * q + lane, g = warp >> 2;
        const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
        for (int c = g; c < 3 * KC; c += G) {
            const int m = c / KC, kc = c - m * KC;
            const float *row = m < 2 ? sW + (long)(m * H + j) * H : sW2 + (long)j * H;
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int k = 16 * kc + 2 * i;
                const float w0 = (j < H && k < H) ? __ldg(row + k) : 0.0f;
                const float w1 = (j < H && k + 1 < H) ? __ldg(row + k + 1) : 0.0f;
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
        // ================= compute warps: thread = hidden unit j, N sequences =================
        const int g = warp >> 2, q = warp & 3, j = 32 * q + lane;
        const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
        const uint32_t dcol = tmem_base + lane_addr + (uint32_t)(D_BASE + g * 3 * N);
        Bars &bar = bars[g];
        const int b0 = b_cta + g * N;
        uint8_t *op = ops + (size_t)g * 4 * OPB;
        const float *vr0 = vring + (size_t)g * 3 * N * vld;
        const int *lens = lens_s + g * N;
        const bool jop = j < HP;                       // this unit has a k column in the operands
        const bool jv = j < H;
        const int jc = jv ? j : H - 1;                 // clamped column for vI reads of padding rows
        // operand byte offsets of element (sequence n, k = j): sw64 layout, K block j >> 5
        const int kk = j & 31;
        const uint32_t obase = (uint32_t)((j >> 5) * (N * 64) + ((kk & 7) << 1));
        uint32_t oswz[4];
#pragma unroll
        for (int mth = 0; mth < 4; mth++) oswz[mth] = (uint32_t)((((kk >> 3) ^ mth) & 3) << 4);
        float h[N];
#pragma unroll
        for (int n = 0; n < N; n++) h[n] = 0.0f;
        float *yp = y + ((long)(reverse ? T - 1 : 0) * B + b0) * ldy + j;
        const long ystep = (long)(reverse ? -1 : 1) * B * ldy;

        if (b0 < B) mbar_arrive(&bar.h);               // completion 0: h_{-1} = 0 is in place
        for (int s = 0; s < (b0 < B ? T : 0); s++) {      // a group wholly past the batch end has no issuer either
            const int t = reverse ? T - 1 - s : s;
            const float *vrow = vr0 + (size_t)(s % 3) * N * vld;
            const uint32_t par = (uint32_t)(s & 1);
            // ---- phase 1: r (needed at once) and z pre-activations ----
            wait_bar(&bar.d1, par);
            tc_fence_after();
            uint32_t dr[N], dz[N];
            tmem_ld_cols<N>(dcol + N, dr);
            tmem_ld_cols<N>(dcol, dz);
            tmem_ld_wait();
            wait_bar(&bar.v[s % 3], (uint32_t)((s / 3) & 1));
            if (jop) {
                float vr[N];
#pragma unroll
                for (int n = 0; n < N; n++) vr[n] = vrow[n * vld + H + jc];
                uint32_t pk[N];
#pragma unroll
                for (int n = 0; n < N; n++) pk[n] = split_pack(sigmoid_fast(__uint_as_float(dr[n]) + vr[n]) * h[n]);
#pragma unroll
                for (int n = 0; n < N; n++) {
                    const uint32_t o = obase + (uint32_t)((n >> 3) * 512 + (n & 7) * 64) + oswz[(n >> 1) & 3];
                    *reinterpret_cast<uint16_t *>(op + 2 * OPB + o) = (uint16_t)(pk[n] & 0xffffu);
                    *reinterpret_cast<uint16_t *>(op + 3 * OPB + o) = (uint16_t)(pk[n] >> 16);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bar.rh);
            float z[N];
            if (jop) {
#pragma unroll
                for (int n = 0; n < N; n++) z[n] = sigmoid_fast(__uint_as_float(dz[n]) + vrow[n * vld + jc]);
            }
            // ---- phase 2: candidate, blend, publish h_t ----
            wait_bar(&bar.d2, par);
            tc_fence_after();
            uint32_t dc[N];
            tmem_ld_cols<N>(dcol + 2 * N, dc);
            tmem_ld_wait();
            if (jop) {
                float vc[N];
#pragma unroll
                for (int n = 0; n < N; n++) vc[n] = vrow[n * vld + 2 * H + jc];
#pragma unroll
                for (int n = 0; n < N; n++) {
                    const float hbar = tanh_fast(__uint_as_float(dc[n]) + vc[n]);
                    const float hn = z[n] * h[n] + (1.0f - z[n]) * hbar;
                    h[n] = (jv && t < lens[n]) ? hn : 0.0f;          // ragged batch: state stays 0 outside the read
                }
#pragma unroll
                for (int n = 0; n < N; n++) {
                    const uint32_t pk = split_pack(h[n]);
                    const uint32_t o = obase + (uint32_t)((n >> 3) * 512 + (n & 7) * 64) + oswz[(n >> 1) & 3];
                    *reinterpret_cast<uint16_t *>(op + o) = (uint16_t)(pk & 0xffffu);
                    *reinterpret_cast<uint16_t *>(op + OPB + o) = (uint16_t)(pk >> 16);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bar.h);
            if (jv) {
#pragma unroll
                for (int n = 0; n < N; n++)
                    if (b0 + n < B) yp[(long)n * ldy] = h[n];
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
        float *vr0 = vring + (size_t)g * 3 * N * vld;
        const uint32_t idesc = umma_idesc_f16_m128(N);
        const uint32_t dz = tmem_base + (uint32_t)(D_BASE + g * 3 * N), dr = dz + N, dc = dz + 2 * N;
        const uint64_t b_hh = umma_desc_sw64_kmajor(op), b_hl = umma_desc_sw64_kmajor(op + OPB);
        const uint64_t b_rh = umma_desc_sw64_kmajor(op + 2 * OPB), b_rl = umma_desc_sw64_kmajor(op + 3 * OPB);
        const uint32_t rowbytes = (uint32_t)vld * 4u;
        auto load_vi = [&](int st) {                                       // elected lane: vI rows of scan step st
            if (st >= T || nrows <= 0) return;
            const int t = reverse ? T - 1 - st : st;
            uint64_t *vb = &bar.v[st % 3];
            mbar_arrive_expect_tx(vb, rowbytes * (uint32_t)nrows);
            const float *src = vI + ((long)t * B + b0) * ldv;
            float *dst = vr0 + (size_t)(st % 3) * N * vld;
            for (int n = 0; n < nrows; n++) bulk_load_1d(dst + (size_t)n * vld, src + (long)n * ldv, rowbytes, vb);
        };
        if (nrows > 0) {
            if (elect_one()) { load_vi(0); load_vi(1); load_vi(2); }
            __syncwarp();
            for (int s = 0; s < T; s++) {
                const uint32_t par = (uint32_t)(s & 1);
                wait_bar(&bar.h, par);                                     // h_{s-1} operand complete, step s-1 done with its vI slot
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kc = 0; kc < KC; kc++) {
                        const uint64_t koff = (uint64_t)(((kc >> 1) * (N * 64) + (kc & 1) * 32) >> 4);
                        const uint32_t acol = tmem_base + (uint32_t)(kc * 8);
                        umma_f16_ts(dz, acol + 0 * ACOLS, b_hh + koff, idesc, kc != 0);
                        umma_f16_ts(dr, acol + 2 * ACOLS, b_hh + koff, idesc, kc != 0);
                        umma_f16_ts(dz, acol + 1 * ACOLS, b_hh + koff, idesc, true);
                        umma_f16_ts(dr, acol + 3 * ACOLS, b_hh + koff, idesc, true);
                        umma_f16_ts(dz, acol + 0 * ACOLS, b_hl + koff, idesc, true);
                        umma_f16_ts(dr, acol + 2 * ACOLS, b_hl + koff, idesc, true);
                    }
                    umma_commit(&bar.d1);
                    if (s >= 1) load_vi(s + 2);                            // slot (s-1) % 3 was released by step s-1
                }
                __syncwarp();
                wait_bar(&bar.rh, par);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kc = 0; kc < KC; kc++) {
                        const uint64_t koff = (uint64_t)(((kc >> 1) * (N * 64) + (kc & 1) * 32) >> 4);
                        const uint32_t acol = tmem_base + (uint32_t)(kc * 8);
                        umma_f16_ts(dc, acol + 4 * ACOLS, b_rh + koff, idesc, kc != 0);
                        umma_f16_ts(dc, acol + 5 * ACOLS, b_rh + koff, idesc, true);
                        umma_f16_ts(dc, acol + 4 * ACOLS, b_rl + koff, idesc, true);
                    }
                    umma_commit(&bar.d2);
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NCW) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int HP, int N, int G>
static int launch(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths,
                  int T, int B, int H, int reverse, cudaStream_t st)
{
    constexpr int NKB = (HP + 31) / 32, OPB = NKB * N * 64;
    const int vld = (3 * H + 3) / 4 * 4;
    size_t smem = 1024 + (size_t)G * 4 * OPB + (size_t)G * 3 * N * vld * 4 + (size_t)G * sizeof(Bars) + (size_t)G * N * 4 + 64;
    // every CTA of this kernel owns the whole tensor memory of its SM: ask for more than half of the shared memory
    // so that a second one (another stream's batch) is placed on a free SM instead of stalling in tcgen05.alloc
    if (smem < 116 * 1024) smem = 116 * 1024;
    auto kern = gru_tc_kernel<HP, N, G>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    const unsigned grid = (unsigned)ceil_div(B, G * N);
    kern<<<grid, G * 160, smem, st>>>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse);
    SLOIKA_RETURN_LAUNCH_STATUS();
}

template <int HP>
static int launch_hp(int n, int g, const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy,
                     const int32_t *lengths, int T, int B, int H, int reverse, cudaStream_t st)
{
    if (n == 8 && g == 1) return launch<HP, 8, 1>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st);
    if (n == 8 && g == 2) return launch<HP, 8, 2>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st);
    if (n == 16 && g == 1) return launch<HP, 16, 1>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st);
    if (n == 16 && g == 2) return launch<HP, 16, 2>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st);
    return SLOIKA_ERR_UNSUPPORTED;
}

// tanh / sigmoid GRUs with H <= 128 whose vI rows are 16-byte aligned; SLOIKA_ERR_UNSUPPORTED otherwise (the caller
// falls back to gru_h16.cu).  `seqs_in_flight` is the number of sequences the caller keeps on the device at once
// (this batch times the batches it pipelines on other streams): it picks how many sequences share a CTA, i.e.
// whether the SMs are spread over one batch (latency) or packed (throughput).  SLOIKA_B200_GRU_TC="N,G" overrides.
int dispatch(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths, int T,
             int B, int H, int reverse, int act, int gate_act, long seqs_in_flight, cudaStream_t st)
{
    if (act != SLOIKA_ACT_TANH || gate_act != SLOIKA_ACT_SIGMOID) return SLOIKA_ERR_UNSUPPORTED;
    if (H > 128 || (ldv & 3) != 0 || ((uintptr_t)vI & 15) != 0) return SLOIKA_ERR_UNSUPPORTED;
    int n = 8, g = 1;
    const long load = seqs_in_flight > B ? seqs_in_flight : B;
    if (load > 8L * 148) g = 2;
    if (load > 16L * 148) n = 16;
    if (const char *ov = getenv("SLOIKA_B200_GRU_TC")) {
        int on = 0, og = 0;
        if (sscanf(ov, "%d,%d", &on, &og) == 2 && (on == 8 || on == 16) && (og == 1 || og == 2)) { n = on; g = og; }
    }
#define TC_CASE(HP_) return launch_hp<HP_>(n, g, vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st)
    if (H <= 32) TC_CASE(32);
    if (H <= 64) TC_CASE(64);
    if (H <= 96) TC_CASE(96);
    if (H <= 112) TC_CASE(112);
    TC_CASE(128);
#undef TC_CASE
}

}  // namespace gru5
}  // namespace sloika
