// y = act(x . W' + bias) on the 5th-generation tensor cores with fp32-equivalent accuracy (three-product operand split).
//
// Used for the GRU input projection over all time steps (sloika/layers.py:1011: vI = x iW' + b),
// FeedForward.run (layers.py:157-158) and the logits of Softmax.run (layers.py:310).
// M = T*B rows is huge (819 200), K <= 512 and N <= ~1100 are small: the weights are tiny and stay
// resident on chip, x is streamed once per N slice (re-reads hit L2), y is written once.  HBM-bound by design.
//
// Accuracy: the reference multiplies in float32.  Plain TF32 (10-bit mantissa) would put ~1e-3 of error into the gate
// pre-activations, so every operand is split on chip into hi + lo and three MMAs  hi.hi + lo.hi + hi.lo  are
// accumulated in fp32 in TMEM (the dropped lo.lo term is 2^-22 relative).  Two operand formats:
//   tf32  hi = top 19 bits, lo = x - hi (exact); kind::tf32, K = 8 per MMA, SWIZZLE_128B tiles; any input
//   f16   hi = fp16(x), lo = fp16(x - hi); kind::f16, K = 16 per MMA (twice the rate), SWIZZLE_64B tiles, half the
//         shared memory per weight (wider slices); only for inputs known to be far inside the fp16 range
//         (sloika_linear_fwd_ex / _gated decide; see include/sloika_b200.h)
//
// Structure (one persistent CTA per SM, 448 threads, warp specialised):
//   grid = n_slices x ctas_per_slice; a CTA owns output columns [n0, n0 + BN) and walks m-tiles.
//   prologue   all warps: W slice -> smem as W_hi / W_lo in the UMMA K-major swizzled layout
//   warp 4     TMA producer: x tile [128 rows x 32 k] per stage (cp.async.bulk.tensor, SWIZZLE_128B)
//   warps 6-9  transform: raw fp32 tile -> hi + lo (tf32: hi in place; f16: both in place of the raw tile, so a stage
//              is 16 KB), fence.proxy.async, signal
//   warp 5     MMA issuer: 3 tcgen05.mma (M=128, N=BN) per K step, accumulators in TMEM (2 stages x BN columns);
//              tcgen05.commit frees the smem stage / publishes the tile
//   warps 0-3, 10-13  epilogue (two groups on alternate 32-column chunks): tcgen05.ld (thread = row), + bias,
//              activation, optional softmax row statistics, then the 32 x 32 box goes to a swizzled staging tile and
//              out through a TMA store (cp.async.bulk.tensor shared -> global); rows whose pitch is not 16-byte
//              aligned take the transposing LDS + STG path instead
#include <cstdlib>
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace sloika {

namespace gemm_tc {

int sm_budget = 0;                 // 0 = all SMs; set through sloika_b200_set_gemm_sm_budget()

constexpr int BM = 128;            // rows per tile (UMMA M)
constexpr int KB = 32;             // k elements per K block (one 128-byte swizzle row)
constexpr int MAX_STAGES = 12;     // smem stages of x-tiles (as many as fit)
constexpr int NACC = 2;            // TMEM accumulator stages
constexpr int THREADS = 448;        // 14 warps: 0-3 + 10-13 epilogue, 4 TMA, 5 MMA, 6-9 transform
constexpr int A_TILE_BYTES = BM * KB * 4;        // 16 KB
constexpr int TMEM_COLS = 512;
constexpr int STG_LD = 36;           // staging row pitch (floats): 16-byte aligned rows, conflict-free float4 phases

struct Params {
    const float *W;
    const float *bias;
    float *y;
    long ldy;
    long M;
    int K, N, act;
    int BN, n_slices, ctas_per_slice, nkb;
    float2 *stats;    // optional [M][n_slices][2] (row max, sum of exp(t - max)) per slice and epilogue group
    int rot;          // output column n is computed from weight/bias row (n + rot) % N  (stay-last logits layout)
    int stages;       // smem pipeline depth
    int vec_out;      // rows of y are 16-byte aligned: 128-bit stores
    int direct;       // vec_out && store rows straight from registers (no smem transpose)
    int tma_out;      // vec_out: 32 x 32 output boxes leave through cp.async.bulk.tensor stores (tmap_y)
    const unsigned *gate;   // optional device word (bits of a non-negative float, e.g. max |x| from the producer):
    unsigned gate_limit;    // gate_mode 1: run only if *gate < gate_limit; 2: only if *gate >= gate_limit; 0: always
    int gate_mode;
    int x_blocked;    // x is in the blocked layout of gru_seq.cu (F16 form, M a multiple of 128): a raw tile is [k / 4][128 rows][4 k]
    int dbg;          // timing experiments only (SLOIKA_B200_GEMM_DBG): 1 no split math, 2 no stores, 4 no MMA
};

__host__ __device__ inline size_t smem_bytes(int BN, int nkb, int stages, bool f16)
{
    size_t w = (size_t)2 * nkb * BN * (f16 ? 64 : 128);    // W_hi + W_lo
    size_t a = (size_t)stages * (f16 ? 1 : 2) * A_TILE_BYTES;   // tf32: (hi, lo) per stage; f16: hi16 | lo16 in place of the raw tile
    // epilogue staging per epilogue warp: 32 x STG_LD floats (transposing fallback) or one 32 x 32 swizzled box for the TMA
    // stores.  (Tried: two boxes per warp so that a store drains while the next box is filled.  The plain 1025-column GEMM
    // gained 7 %, the softmax form with its row statistics -- the one the path uses -- lost 12 %: one box.)
    size_t stg = (size_t)8 * 32 * STG_LD * 4;
    size_t misc = (size_t)BN * 4 + 512;                    // bias slice + barriers
    return w + a + stg + misc + 1024;                      // + alignment slack
}

// F16: operands split into fp16 hi / lo pairs (kind::f16, K = 16 per MMA, SWIZZLE_64B tiles) instead of tf32
// hi / lo: half the shared-memory footprint of the weights (wider column slices) and twice the MMA rate.  Only for
// callers that know |x| and |W| stay far below the fp16 range (bounded activations): see sloika_linear_fwd_ex.
template <int ACT, bool STATS, bool F16>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y, const Params p)
{
    if (p.gate_mode != 0) {                       // device-side choice between two enqueued forms of the same GEMM
        const bool below = *p.gate < p.gate_limit;
        if (below != (p.gate_mode == 1)) return;
    }
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128-byte swizzle; plain pointer arithmetic keeps the shared address space
    uint8_t *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int BN = p.BN, nkb = p.nkb, STAGES = p.stages;
    constexpr int WROW = F16 ? 64 : 128;                                  // bytes of one weight row per K block
    uint8_t *Whi = smem;
    uint8_t *Wlo = Whi + (size_t)nkb * BN * WROW;
    constexpr int STAGE_BYTES = (F16 ? 1 : 2) * A_TILE_BYTES;
    uint8_t *Abase = Wlo + (size_t)nkb * BN * WROW;                       // stage s: raw tile, converted in place to hi; lo behind it
                                                                          // (F16: raw tile replaced by hi16 | lo16, 8 KB each)
    float *stg_all = reinterpret_cast<float *>(Abase + (size_t)STAGES * STAGE_BYTES);
    float *bias_s = stg_all + 8 * 32 * STG_LD;
    uint64_t *bars = reinterpret_cast<uint64_t *>(bias_s + BN);
    uint64_t *full_raw = bars;                     // [STAGES] TMA -> transform
    uint64_t *full_split = bars + MAX_STAGES;      // [STAGES] transform -> MMA
    uint64_t *empty = bars + 2 * MAX_STAGES;       // [STAGES] MMA -> TMA
    uint64_t *tmem_full = bars + 3 * MAX_STAGES;   // [NACC]   MMA -> epilogue
    uint64_t *tmem_empty = tmem_full + NACC;       // [NACC]   epilogue -> MMA
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(tmem_empty + NACC);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slice = blockIdx.x % p.n_slices;             // slices interleaved: the CTAs that share an x tile run together
    const int cta_in_slice = blockIdx.x / p.n_slices;
    const int mt_stride = p.ctas_per_slice;
    const int n0 = slice * BN;
    const long m_tiles = (p.M + BM - 1) / BM;

    // ---------------- prologue ----------------
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { tc::mbar_init(&full_raw[s], 1); tc::mbar_init(&full_split[s], 128); tc::mbar_init(&empty[s], 1); }
        for (int a = 0; a < NACC; a++) { tc::mbar_init(&tmem_full[a], 1); tc::mbar_init(&tmem_empty[a], 256); }
        tc::mbar_fence_init();
    }
    if (warp == 4 && lane == 0) { tc::tma_prefetch_desc(&tmap_x); tc::tma_prefetch_desc(&tmap_y); }
    if (warp == 5) tc::tmem_alloc(tmem_base_s, TMEM_COLS);
    // weights of this slice: split into hi / lo and stored K-major, 128-byte swizzled, zero padded
    for (int e = tid; e < BN * nkb * KB; e += THREADS) {
        const int k = e % (nkb * KB), n = e / (nkb * KB);
        float w = 0.0f;
        if (n0 + n < p.N && k < p.K) w = __ldg(p.W + (long)((n0 + n + p.rot) % p.N) * p.K + k);
        if constexpr (F16) {
            const __half hi = __float2half_rn(w);
            const uint32_t off = (uint32_t)(k / KB) * (uint32_t)(BN * 64) + tc::sw64_offset(n, k % KB);
            *reinterpret_cast<__half *>(Whi + off) = hi;
            *reinterpret_cast<__half *>(Wlo + off) = __float2half_rn(w - __half2float(hi));
        } else {
            const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
            const uint32_t off = (uint32_t)(k / KB) * (uint32_t)(BN * 128) + tc::sw128_offset(n, k % KB);
            *reinterpret_cast<float *>(Whi + off) = hi;
            *reinterpret_cast<float *>(Wlo + off) = w - hi;
        }
    }
    for (int n = tid; n < BN; n += THREADS) bias_s[n] = (p.bias && n0 + n < p.N) ? __ldg(p.bias + (n0 + n + p.rot) % p.N) : 0.0f;
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp == 4) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t it = 0;
            for (long mt = cta_in_slice; mt < m_tiles; mt += mt_stride) {
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    tc::mbar_wait(&empty[s], ph ^ 1);
                    tc::mbar_arrive_expect_tx(&full_raw[s], A_TILE_BYTES);
                    if (p.x_blocked) tc::tma_load_4d(Abase + (size_t)s * STAGE_BYTES, &tmap_x, &full_raw[s], 0, 0, kb * (KB / 4), (int)mt);
                    else tc::tma_load_2d(Abase + (size_t)s * STAGE_BYTES, &tmap_x, &full_raw[s], kb * KB, (int)(mt * BM));
                }
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = F16 ? tc::umma_idesc_f16_m128(BN) : tc::umma_idesc_tf32_m128(BN);
            uint32_t it = 0, tile = 0;
            for (long mt = cta_in_slice; mt < m_tiles; mt += mt_stride, tile++) {
                const int a = tile % NACC;
                const uint32_t aph = (tile / NACC) & 1;
                tc::mbar_wait(&tmem_empty[a], aph ^ 1);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    tc::mbar_wait(&full_split[s], ph);
                    tc::tc_fence_after();
                    const int krem = p.K - kb * KB;
                    if constexpr (F16) {
                        const uint32_t a_hi = tc::smem_u32(Abase + (size_t)s * STAGE_BYTES);
                        const uint32_t a_lo = a_hi + A_TILE_BYTES / 2;
                        const uint32_t b_hi = tc::smem_u32(Whi) + (uint32_t)kb * (uint32_t)(BN * 64);
                        const uint32_t b_lo = tc::smem_u32(Wlo) + (uint32_t)kb * (uint32_t)(BN * 64);
                        const int ksteps = krem >= KB ? 2 : (krem + 15) / 16;
#pragma unroll 1
                        for (int ks = 0; ks < ksteps; ks++) {
                            const uint64_t dah = tc::umma_desc_sw64_kmajor(a_hi + ks * 32);
                            const uint64_t dal = tc::umma_desc_sw64_kmajor(a_lo + ks * 32);
                            const uint64_t dbh = tc::umma_desc_sw64_kmajor(b_hi + ks * 32);
                            const uint64_t dbl = tc::umma_desc_sw64_kmajor(b_lo + ks * 32);
                            if (p.dbg & 4) continue;
                            tc::umma_f16_ss(d_tmem, dah, dbh, idesc, (kb | ks) != 0);
                            tc::umma_f16_ss(d_tmem, dal, dbh, idesc, true);
                            tc::umma_f16_ss(d_tmem, dah, dbl, idesc, true);
                        }
                        tc::umma_commit(&empty[s]);
                        continue;
                    }
                    const uint32_t a_hi = tc::smem_u32(Abase + (size_t)s * STAGE_BYTES);
                    const uint32_t a_lo = a_hi + A_TILE_BYTES;
                    const uint32_t b_hi = tc::smem_u32(Whi) + (uint32_t)kb * (uint32_t)(BN * 128);
                    const uint32_t b_lo = tc::smem_u32(Wlo) + (uint32_t)kb * (uint32_t)(BN * 128);
                    const int ksteps = krem >= KB ? 4 : (krem + 7) / 8;
#pragma unroll 1
                    for (int ks = 0; ks < ksteps; ks++) {
                        const uint64_t dah = tc::umma_desc_sw128_kmajor(a_hi + ks * 32);
                        const uint64_t dal = tc::umma_desc_sw128_kmajor(a_lo + ks * 32);
                        const uint64_t dbh = tc::umma_desc_sw128_kmajor(b_hi + ks * 32);
                        const uint64_t dbl = tc::umma_desc_sw128_kmajor(b_lo + ks * 32);
                        if (p.dbg & 4) continue;
                        tc::umma_tf32_ss(d_tmem, dah, dbh, idesc, (kb | ks) != 0);
                        tc::umma_tf32_ss(d_tmem, dal, dbh, idesc, true);
                        tc::umma_tf32_ss(d_tmem, dah, dbl, idesc, true);
                    }
                    tc::umma_commit(&empty[s]);                 // smem stage reusable once these MMAs retire
                }
                tc::umma_commit(&tmem_full[a]);                 // accumulator tile complete
            }
        }
    } else if (warp >= 6 && warp < 10) {
        // ================= transform: fp32 -> (hi, lo) =================
        const int tt = tid - 6 * 32;                            // 0..127
        uint32_t it = 0;
        for (long mt = cta_in_slice; mt < m_tiles; mt += mt_stride) {
            for (int kb = 0; kb < nkb; kb++, it++) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                tc::mbar_wait(&full_raw[s], ph);
                if constexpr (F16) {
                    // thread -> (row r, 32-byte pair p of the raw 128-byte row): 8 fp32 in, 8 fp16 hi + 8 fp16 lo out
                    // in place: every thread first pulls its 8 x 4 floats into registers, the four warps meet on a
                    // named barrier, then the fp16 tiles overwrite the raw one (a 16 KB stage instead of 32 KB
                    // doubles the number of TMA loads in flight)
                    const uint8_t *raw = Abase + (size_t)s * STAGE_BYTES;
                    uint8_t *h16 = Abase + (size_t)s * STAGE_BYTES;
                    if (p.dbg & 1) {                                // timing experiments: no split at all
                        tc::fence_proxy_async();
                        tc::mbar_arrive(&full_split[s]);
                        continue;
                    }
                    float4 va[4], vb[4];
                    // blocked input: the raw tile is [k / 4][128 rows][4 k] (no swizzle); thread = row tt, pair i
                    const bool xb = p.x_blocked != 0;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int q = tt + 128 * i, r = xb ? tt : q >> 2, pp = xb ? i : q & 3;
                        if (xb) {
                            va[i] = *reinterpret_cast<const float4 *>(raw + (((2 * pp) * 128 + r) << 4));
                            vb[i] = *reinterpret_cast<const float4 *>(raw + (((2 * pp + 1) * 128 + r) << 4));
                        } else {
                            const uint32_t rb = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
                            va[i] = *reinterpret_cast<const float4 *>(raw + rb + (((2 * pp) ^ (r & 7)) << 4));
                            vb[i] = *reinterpret_cast<const float4 *>(raw + rb + (((2 * pp + 1) ^ (r & 7)) << 4));
                        }
                    }
                    asm volatile("bar.sync 3, 128;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int q = tt + 128 * i, r = xb ? tt : q >> 2, pp = xb ? i : q & 3;
                        const float v[8] = {va[i].x, va[i].y, va[i].z, va[i].w, vb[i].x, vb[i].y, vb[i].z, vb[i].w};
                        uint32_t hw[4], lw[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                            const float2 hb = __half22float2(h);
                            const __half2 l = __floats2half2_rn(v[2 * e] - hb.x, v[2 * e + 1] - hb.y);
                            hw[e] = *reinterpret_cast<const uint32_t *>(&h);
                            lw[e] = *reinterpret_cast<const uint32_t *>(&l);
                        }
                        const uint32_t off = (uint32_t)((r >> 3) * 512 + (r & 7) * 64 + (((pp ^ (r >> 1)) & 3) << 4));
                        *reinterpret_cast<uint4 *>(h16 + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                        *reinterpret_cast<uint4 *>(h16 + A_TILE_BYTES / 2 + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    }
                    tc::fence_proxy_async();
                    tc::mbar_arrive(&full_split[s]);
                    continue;
                }
                float4 *hi = reinterpret_cast<float4 *>(Abase + (size_t)s * STAGE_BYTES);
                float4 *lo = reinterpret_cast<float4 *>(Abase + (size_t)s * STAGE_BYTES + A_TILE_BYTES);
#pragma unroll
                for (int i = 0; i < ((p.dbg & 1) ? 0 : A_TILE_BYTES / 16 / 128); i++) {
                    const int c = tt + 128 * i;
                    const float4 v = hi[c];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
                    hi[c] = h;
                    lo[c] = l;
                }
                tc::fence_proxy_async();                        // generic-proxy writes -> visible to the tensor core
                tc::mbar_arrive(&full_split[s]);
            }
        }
    } else if (warp < 4 || warp >= 10) {
        // ================= epilogue: two groups of 4 warps (0-3 and 10-13) =================
        // A warp may only read the TMEM lane quarter 32*(warp % 4); the two groups take alternate
        // 32-column chunks of the tile so that every scheduler has two epilogue warps to interleave.
        const int eg = warp < 4 ? 0 : 1;
        const int lq = warp & 3;
        float *stg = stg_all + (eg * 4 + lq) * 32 * STG_LD;
        // TMA-store path: a dense 32 x 32 fp32 box per warp in the SWIZZLE_128B layout (1024-byte aligned)
        uint8_t *stg_t = reinterpret_cast<uint8_t *>(stg_all) + (eg * 4 + lq) * 4096;
        const int nchunks = (BN + 31) / 32;
        uint32_t tile = 0;
        for (long mt = cta_in_slice; mt < m_tiles; mt += mt_stride, tile++) {
            const int a = tile % NACC;
            const uint32_t aph = (tile / NACC) & 1;
            tc::mbar_wait(&tmem_full[a], aph);
            tc::tc_fence_after();
            const long row0 = mt * BM + lq * 32;
            float m_run = -INFINITY, s_run = 0.0f;              // STATS: online (max, sum exp) of this thread's row
            bool released = false;
            for (int ci = eg; ci < nchunks; ci += 2) {
                const int c0 = ci * 32;
                const int width = (BN - c0) >= 32 ? 32 : 16;
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(a * BN + c0);
                if (width == 32) tc::tmem_ld_32x32b_x32(taddr, v);
                else tc::tmem_ld_32x32b_x16(taddr, v);
                tc::tmem_ld_wait();
                if (ci + 2 >= nchunks) {                         // last read of this accumulator: hand it back
                    tc::tc_fence_before();
                    tc::mbar_arrive(&tmem_empty[a]);
                    released = true;
                }
                if (p.tma_out) {                                 // the previous box of this warp has left shared memory
                    if (lane == 0) tc::bulk_wait_read();
                    __syncwarp();
                }
                // thread = row: bias + activation
                float o[32];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    if (i < width) {
                        const float4 b4 = *reinterpret_cast<const float4 *>(&bias_s[c0 + i]);
                        o[i + 0] = apply_act_t<ACT>(__uint_as_float(v[i + 0]) + b4.x);
                        o[i + 1] = apply_act_t<ACT>(__uint_as_float(v[i + 1]) + b4.y);
                        o[i + 2] = apply_act_t<ACT>(__uint_as_float(v[i + 2]) + b4.z);
                        o[i + 3] = apply_act_t<ACT>(__uint_as_float(v[i + 3]) + b4.w);
                        if (p.tma_out)
                            *reinterpret_cast<float4 *>(stg_t + lane * 128 + ((((i >> 2) ^ lane) & 7) << 4)) =
                                make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
                        else if (!p.direct)
                            *reinterpret_cast<float4 *>(&stg[lane * STG_LD + i]) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
                    }
                }
                if (p.tma_out) {
                    // one elected lane hands the box to the TMA engine: no shared-memory read-back, no store
                    // instructions, columns >= N and rows >= M are clipped by the tensor map
                    tc::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && !(p.dbg & 2)) {
                        tc::tma_store_2d(stg_t, &tmap_y, n0 + c0, (int)row0);
                        tc::bulk_commit();
                    }
                }
                if (p.direct && !(p.dbg & 2)) {
                    // thread = row: 128-bit stores straight from registers (each row is written 16 bytes at a
                    // time by consecutive instructions of the same thread; L2 merges the sectors)
                    const long m = row0 + lane;
                    if (m < p.M) {
                        float *dst = p.y + m * p.ldy + n0 + c0;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            if (i < width) {
                                const int n = n0 + c0 + i;
                                if (n + 3 < p.N) {
                                    __stcs(reinterpret_cast<float4 *>(dst + i), make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]));
                                } else {
                                    if (n < p.N) dst[i] = o[i];
                                    if (n + 1 < p.N) dst[i + 1] = o[i + 1];
                                    if (n + 2 < p.N) dst[i + 2] = o[i + 2];
                                }
                            }
                        }
                    }
                }
                if constexpr (STATS) {
                    const int nvalid = min(width, p.N - (n0 + c0));      // padded columns must not enter the max
                    if (nvalid == 32) {
                        float cm = o[0];
#pragma unroll
                        for (int i = 1; i < 32; i++) cm = fmaxf(cm, o[i]);
                        const float nm = fmaxf(m_run, cm);
                        const float nms = nm * SLOIKA_LOG2E;                 // exp(o - nm) = ex2(o*log2e - nm*log2e)
                        float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            acc0 += ex2_ftz(fmaf(o[i], SLOIKA_LOG2E, -nms));
                            acc1 += ex2_ftz(fmaf(o[i + 1], SLOIKA_LOG2E, -nms));
                        }
                        s_run = s_run * ex2_ftz(fmaf(m_run, SLOIKA_LOG2E, -nms)) + (acc0 + acc1);
                        m_run = nm;
                    } else if (nvalid > 0) {
                        float cm = -INFINITY;
#pragma unroll
                        for (int i = 0; i < 32; i++)
                            if (i < nvalid) cm = fmaxf(cm, o[i]);
                        const float nm = fmaxf(m_run, cm);
                        float acc = 0.0f;
#pragma unroll
                        for (int i = 0; i < 32; i++)
                            if (i < nvalid) acc += __expf(o[i] - nm);
                        s_run = s_run * __expf(m_run - nm) + acc;
                        m_run = nm;
                    }
                }
                __syncwarp();
                if (!p.direct && !p.tma_out && !(p.dbg & 2)) {
                    if (p.vec_out) {
                        // 8 lanes x float4 cover the 32 columns of one row: 4 rows (4 x 128 B) per instruction
                        const int cq = (lane & 7) * 4, rsub = lane >> 3;
                        const int n = n0 + c0 + cq;
                        if (cq < width && n < p.N) {
#pragma unroll
                            for (int r4 = 0; r4 < 8; r4++) {
                                const int r = r4 * 4 + rsub;
                                const long m = row0 + r;
                                if (m < p.M) {
                                    const float4 q = *reinterpret_cast<const float4 *>(&stg[r * STG_LD + cq]);
                                    float *dst = p.y + m * p.ldy + n;
                                    if (n + 3 < p.N) {
                                        __stcs(reinterpret_cast<float4 *>(dst), q);
                                    } else {
                                        dst[0] = q.x;
                                        if (n + 1 < p.N) dst[1] = q.y;
                                        if (n + 2 < p.N) dst[2] = q.z;
                                    }
                                }
                            }
                        }
                    } else {
                        const int n = n0 + c0 + lane;
                        if (lane < width && n < p.N) {
#pragma unroll 8
                            for (int r = 0; r < 32; r++) {
                                const long m = row0 + r;
                                if (m < p.M) p.y[m * p.ldy + n] = stg[r * STG_LD + lane];
                            }
                        }
                    }
                }
                __syncwarp();
            }
            if (!released) {                                     // a group without chunks still frees the stage
                tc::tc_fence_before();
                tc::mbar_arrive(&tmem_empty[a]);
            }
            if constexpr (STATS) {
                const long m = row0 + lane;
                if (m < p.M) p.stats[(m * p.n_slices + slice) * 2 + eg] = make_float2(m_run, s_run);
            }
        }
        if (p.tma_out && lane == 0) tc::bulk_wait_all();         // every box has been written before the CTA retires
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---- host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// Returns SLOIKA_ERR_UNSUPPORTED when the shape/alignment cannot use the tensor path (caller falls back).
// Column slices the kernel would use for (K, N) (0 = shape not supported): lets callers size `stats`.
int plan_slices(int K, int N, int *bn_out, bool f16)
{
    if (K <= 0 || K > 512 || N <= 0) return 0;
    const int nkb = (K + KB - 1) / KB;
    // slice widths are multiples of 32 columns: the epilogue works in 32-column boxes (TMA stores clip at N).
    // (Tried: multiples of 16 with the last half chunk stored from registers -- 1025 logits as 5 x 208 computed columns instead
    // of 5 x 224.  The scattered 64-byte row stores cost more than the 7 % of columns saved: 0.90 -> 1.04 ms with statistics.)
    int bn_max = 256;
    const char *bn_env = getenv("SLOIKA_B200_GEMM_BN");
    if (bn_env && atoi(bn_env) >= 32) bn_max = atoi(bn_env) / 32 * 32;
    while (bn_max >= 32 && smem_bytes(bn_max, nkb, 3, f16) > 227 * 1024) bn_max -= 32;
    if (bn_max < 32) return 0;
    const int n_slices = (N + bn_max - 1) / bn_max;
    if (bn_out) *bn_out = ((N + n_slices - 1) / n_slices + 31) / 32 * 32;
    return n_slices;
}

int launch(const float *x, long ldx, const float *W, const float *bias, float *y, long ldy, long M, int K, int N,
           int act, float2 *stats, int rot, bool f16, cudaStream_t st, const unsigned *gate = nullptr,
           unsigned gate_limit = 0, int gate_mode = 0)
{
    // ldx == -1: x is in the blocked layout (sloika_b200.h, sloika_gru_seq_fwd): fp16-split form only, whole blocks of 128 rows
    const bool x_blocked = ldx == -1;
    if (x_blocked) {
        if (!f16 || (M % BM) != 0 || ((uintptr_t)x & 15) != 0 || K > 512 || M > 0x7fffffffL) return SLOIKA_ERR_UNSUPPORTED;
    } else if ((ldx & 3) != 0 || ((uintptr_t)x & 15) != 0 || K > 512 || M < BM || M > 0x7fffffffL) return SLOIKA_ERR_UNSUPPORTED;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return SLOIKA_ERR_UNSUPPORTED;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return SLOIKA_ERR_UNSUPPORTED;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) return SLOIKA_ERR_UNSUPPORTED;

    const int nkb = (K + KB - 1) / KB;
    // widest slice whose (hi, lo) weights fit beside (at least) 3 x stages in 227 KB of shared memory
    int BN = 0;
    const int n_slices = plan_slices(K, N, &BN, f16);
    if (n_slices <= 0 || n_slices > sms) return SLOIKA_ERR_UNSUPPORTED;
    if (stats && act != SLOIKA_ACT_LINEAR) return SLOIKA_ERR_UNSUPPORTED;

    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t gstride[1] = {(cuuint64_t)ldx * 4};
    const cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    if (x_blocked) {
        // [M / 128 blocks][K / 4 groups][128 rows][4 floats]: one box = 8 groups x 128 rows x 16 bytes = a 16 KB raw tile
        const cuuint64_t groups = (cuuint64_t)((K + 3) / 4);
        const cuuint64_t bdim[4] = {4, 128, groups, (cuuint64_t)(M / BM)};
        const cuuint64_t bstride[3] = {16, 128 * 16, groups * 128 * 16};
        const cuuint32_t bbox[4] = {4, 128, (cuuint32_t)(KB / 4), 1};
        const cuuint32_t bestr[4] = {1, 1, 1, 1};
        if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(x), bdim, bstride, bbox, bestr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return SLOIKA_ERR_UNSUPPORTED;
    } else if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(x), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return SLOIKA_ERR_UNSUPPORTED;

    Params p;
    p.W = W; p.bias = bias; p.y = y; p.ldy = ldy; p.M = M; p.K = K; p.N = N; p.act = act;
    p.BN = BN; p.n_slices = n_slices; p.nkb = nkb;
    p.stats = stats; p.rot = rot;
    p.gate = gate; p.gate_limit = gate_limit; p.gate_mode = gate ? gate_mode : 0;
    p.x_blocked = x_blocked ? 1 : 0;
    p.vec_out = ((ldy & 3) == 0) && (((uintptr_t)y & 15) == 0);
    const char *direct = getenv("SLOIKA_B200_GEMM_DIRECT");
    p.direct = (p.vec_out && direct && atoi(direct) != 0) ? 1 : 0;
    CUtensorMap tmap_y = tmap;                                 // placeholder when the TMA-store path is off
    p.tma_out = 0;
    if (p.vec_out && !p.direct && !getenv("SLOIKA_B200_GEMM_NO_TMA_OUT")) {
        const cuuint64_t ydim[2] = {(cuuint64_t)N, (cuuint64_t)M};
        const cuuint64_t ystride[1] = {(cuuint64_t)ldy * 4};
        const cuuint32_t ybox[2] = {32, 32};
        if (enc(&tmap_y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, y, ydim, ystride, ybox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
            p.tma_out = 1;
    }
    const char *dbg = getenv("SLOIKA_B200_GEMM_DBG");
    p.dbg = dbg ? atoi(dbg) : 0;
    const long m_tiles = (M + BM - 1) / BM;
    // SM budget: a caller that pipelines several batches on different streams keeps part of the SMs busy with the
    // long recurrence kernels (their CTAs own the tensor memory of their SM, so a GEMM CTA cannot share it);
    // a grid no larger than the SMs that are free starts at once instead of queueing behind them.
    if (sm_budget > 0 && sm_budget < sms) sms = sm_budget < n_slices ? n_slices : sm_budget;
    long per = sms / n_slices;
    if (per > m_tiles) per = m_tiles;
    p.ctas_per_slice = (int)per;
    // (Tried: giving a narrower last slice -- 1025 logits = 4 x 224 + 129 -- CTAs in proportion to its width so that
    // it does not finish early.  The slices then walk the m-tiles at different paces, their x tiles stop meeting in L2,
    // and both GEMM shapes get 20-30 % slower: every slice keeps the same number of CTAs.)
    // Pipeline depth (measured sweep, tools/gemm_bench.py SHAPES=sweep): with few slices a deeper ring of x tiles
    // hides more HBM latency (best at ~5); with many slices (softmax logits: write dominated, every x tile is
    // re-read by each slice's CTA out of L2) a deep ring lets the slices drift apart and costs up to 25 %.
    const int want_stages = n_slices >= 4 ? 3 : 5;
    int stages = 3;
    while (stages < want_stages && stages < MAX_STAGES && smem_bytes(BN, nkb, stages + 1, f16) <= 227 * 1024) stages++;
    {
        const char *cap = getenv("SLOIKA_B200_GEMM_STAGES");       // tuning experiments
        if (cap && atoi(cap) >= 2 && atoi(cap) <= MAX_STAGES && smem_bytes(BN, nkb, atoi(cap), f16) <= 227 * 1024) stages = atoi(cap);
    }
    p.stages = stages;
    const size_t smem = smem_bytes(BN, nkb, stages, f16);
    const unsigned grid = (unsigned)(n_slices * p.ctas_per_slice);
#define LAUNCH_KERNEL(KERN)                                                                                    \
    {                                                                                                         \
        cudaError_t err = cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (err != cudaSuccess) return (int)err;                                                              \
        KERN<<<grid, THREADS, smem, st>>>(tmap, tmap_y, p);                                                           \
    }
#define LAUNCH_ACT(A)                                                                                         \
    case A:                                                                                                   \
        if (f16) LAUNCH_KERNEL((gemm_tf32x3_kernel<A, false, true>))                                          \
        else LAUNCH_KERNEL((gemm_tf32x3_kernel<A, false, false>))                                             \
        break;
    if (stats) {
        if (f16) LAUNCH_KERNEL((gemm_tf32x3_kernel<SLOIKA_ACT_LINEAR, true, true>))
        else LAUNCH_KERNEL((gemm_tf32x3_kernel<SLOIKA_ACT_LINEAR, true, false>))
        SLOIKA_RETURN_LAUNCH_STATUS();
    }
    switch (act) {
        LAUNCH_ACT(SLOIKA_ACT_LINEAR)
        LAUNCH_ACT(SLOIKA_ACT_TANH)
        LAUNCH_ACT(SLOIKA_ACT_SIGMOID)
        LAUNCH_ACT(SLOIKA_ACT_ELU)
        default: return SLOIKA_ERR_UNSUPPORTED;
    }
#undef LAUNCH_KERNEL
#undef LAUNCH_ACT
    SLOIKA_RETURN_LAUNCH_STATUS();
}

}  // namespace gemm_tc
}  // namespace sloika
