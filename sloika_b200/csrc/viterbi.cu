// Transducer Viterbi best-path decode, batched: one CTA per read.
//
// Reference semantics: decode.prepare_post (sloika/decode.py:21-36) + decode.viterbi (:39-93), called
// per read from basecall.decode_post (sloika/basecall.py:44-46).  The reference is a Python loop over
// events with ~15 NumPy calls over the 1024 k-mer states and an int32 traceback matrix.
//
// State j in [0,K) is k-mer j (posterior column j+1); column 0 is "stay".  With p = v_{i-1}:
//     step[j] = max_{a<nb}   p[a*K/nb   + j/nb  ]              (first maximum)
//     skip[j] = max_{a<nb^2} p[a*K/nb^2 + j/nb^2] - skip_pen   (first maximum)
//     move[j] = lpost[i][1+j] + max(step, skip)     from: step if step > skip else skip
//     stay[j] = p[j] + lpost[i][0]                  v[j] = max(move, stay), traceback stay unless move > stay
// All nb states 'nb*r .. nb*r+nb-1' share the same step predecessor set (index r) so thread r owns
// them: the step max stays in registers, the skip max is nb^2 conflict-free shared loads, v_{i-1} and
// v_i ping-pong in shared memory (one __syncthreads per event).  Only 1 + nb + nb^2 traceback outcomes
// exist per state, so the generic kernel's traceback is ONE BYTE per state per event (code 0 = stay, 1+a = step
// from a, 1+nb+a = skip from a) in global memory -- 1 KB/event instead of the reference's 4 KB.  The K = 1024
// kernel goes further: the move predecessor is common to the four states a thread owns, only "moved or
// stayed" is per state, so a quad is one uint16 (bits 0-3 = moved flags, bits 4-8 = the shared code):
// 512 bytes per event.  Arithmetic is float32 add/max in the reference's order, so given identical
// log-posteriors scores and paths are bit-identical.
//
// HBM-bound: 4*(K+1) bytes of posterior read + K/2 (K = 1024 kernel; K otherwise) bytes of traceback written per event.
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include "common.cuh"

namespace sloika {

constexpr float VIT_ETA = 1e-10f;

// lpost = log((min_prob + (1 - min_prob) * post) + 1e-10), each operation rounded to float32 exactly as
// NumPy evaluates decode.py:36 and :56 (no FMA contraction).
__device__ __forceinline__ float to_lpost(float v, int mode, float c0, float c1)
{
    if (mode == SLOIKA_VIT_LOG) return v;
    const float p = __fadd_rn(c0, __fmul_rn(c1, v));
    return logf(__fadd_rn(p, VIT_ETA));
}

template <int NB>
__global__ void __launch_bounds__(256)
viterbi_kernel(const float *__restrict__ post, long ld_t, long ld_b, const int32_t *__restrict__ lengths, int T, int B,
               int K, float skip_pen, float c0, float c1, int mode, uint8_t *__restrict__ tb, int32_t *__restrict__ path_out,
               int32_t *__restrict__ path_len, float *__restrict__ score_out)
{
    constexpr int NS = NB * NB;
    extern __shared__ __align__(16) float vbuf[];       // [2][K]
    __shared__ float red_v[32];
    __shared__ int red_i[32];
    __shared__ int s_best;

    const int b = blockIdx.x;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int nev = lengths ? min(lengths[b], T) : T;
    if (nev < 1) {
        if (tid == 0) { path_len[b] = 0; score_out[b] = 0.0f; }
        return;
    }
    const int rstep = K / NB, rskip = K / NS;
    const float *pb = post + (long)b * ld_b;
    uint8_t *tbb = tb + (size_t)b * (size_t)T * (size_t)K;

    // v_0 = lpost[0][1:]   (decode.py:57)
    for (int j = tid; j < K; j += nthr) vbuf[j] = to_lpost(__ldg(pb + 1 + j), mode, c0, c1);
    __syncthreads();

    int cur = 0;
    for (int i = 1; i < nev; i++) {
        const float *row = pb + (long)i * ld_t;
        const float *p = vbuf + cur * K;
        float *vn = vbuf + (cur ^ 1) * K;
        const float lp0 = to_lpost(__ldg(row), mode, c0, c1);
        for (int r = tid; r < rstep; r += nthr) {
            float lp[NB];
#pragma unroll
            for (int c = 0; c < NB; c++) lp[c] = __ldg(row + 1 + NB * r + c);
            // step: first maximum over a of p[a*rstep + r]
            float ss = p[r];
            int as = 0;
#pragma unroll
            for (int a = 1; a < NB; a++) {
                const float c = p[a * rstep + r];
                if (c > ss) { ss = c; as = a; }
            }
            // skip: first maximum over a of p[a*rskip + q], q = r / NB
            const int q = r / NB;
            float sk = p[q];
            int ak = 0;
#pragma unroll
            for (int a = 1; a < NS; a++) {
                const float c = p[a * rskip + q];
                if (c > sk) { sk = c; ak = a; }
            }
            sk = __fsub_rn(sk, skip_pen);
            const bool use_step = ss > sk;                       // tie -> skip (decode.py:76)
            const float best = use_step ? ss : sk;
            const unsigned code = use_step ? (1u + as) : (1u + NB + ak);
            unsigned packed = 0;
#pragma unroll
            for (int c = 0; c < NB; c++) {
                const int j = NB * r + c;
                const float move = __fadd_rn(to_lpost(lp[c], mode, c0, c1), best);
                const float stay = __fadd_rn(p[j], lp0);
                const bool mv = move > stay;                     // tie -> stay (decode.py:81)
                vn[j] = mv ? move : stay;
                const unsigned cd = mv ? code : 0u;
                if (NB == 4) packed |= cd << (8 * c);
                else tbb[(size_t)i * K + j] = (uint8_t)cd;
            }
            if (NB == 4) reinterpret_cast<unsigned *>(tbb + (size_t)i * K)[r] = packed;
        }
        cur ^= 1;
        __syncthreads();
    }

    // ---- argmax of v_T (first maximum) ----
    const float *v = vbuf + cur * K;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = tid; j < K; j += nthr) {
        const float c = v[j];
        if (c > bv || (c == bv && j < bi)) { bv = c; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { red_v[tid >> 5] = bv; red_i[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
        const int nw = (nthr + 31) >> 5;
        for (int w = 1; w < nw; w++)
            if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi)) { bv = red_v[w]; bi = red_i[w]; }
        score_out[b] = bv;
        // ---- backtrace (decode.py:84-91): states with stays removed, written right-aligned ----
        int32_t *out = path_out + (size_t)b * T;
        int pos = nev - 1;
        int state = bi;
        out[pos] = state;
        for (int i = nev - 1; i > 0; i--) {
            const unsigned cd = tbb[(size_t)i * K + state];
            if (cd != 0) {
                if (cd <= (unsigned)NB) state = (int)(cd - 1) * rstep + state / NB;
                else state = (int)(cd - 1 - NB) * rskip + state / NS;
                out[--pos] = state;
            }
        }
        path_len[b] = nev - pos;
        s_best = pos;
    }
    __syncthreads();
    // ---- left-align the path ----
    const int off = s_best;
    const int n = nev - off;
    if (off > 0) {
        int32_t *out = path_out + (size_t)b * T;
        for (int base = 0; base < n; base += nthr) {
            const int idx = base + tid;
            int32_t val = 0;
            if (idx < n) val = out[off + idx];
            __syncthreads();
            if (idx < n) out[idx] = val;
            __syncthreads();
        }
    }
}


// ---------------------------------------------------------------------------------------------------
// Specialisation for the basecaller's shape: 4 bases, k = 5 -> K = 1024 states, 128 threads, thread r owns the
// eight states 8r..8r+7, i.e. the two quads 2r and 2r+1 (a quad = the four states that share a step predecessor
// set) which also share their skip predecessor set (index r >> 1).  Per event and thread: eight log-posteriors,
// two step maxima, ONE skip search -- the per-thread fixed work (row staging, barriers, addresses, the stay term)
// is paid once per eight states; round 1's 256-thread form (four states per thread, 32 registers, spilling) issued
// 1520 warp instructions per event, this one about 700.  Compile-time strides, 128-bit accesses to the own-state
// octet of v, the row of the NEXT event staged with cp.async while the current one is processed.
//   IN_POST / IN_LOG : rows laid out [stay, kmer 0..1023] (the network's posterior layout), any stride
//   IN_LOGITS        : rows laid out [kmer 0..1023, stay] (16-byte aligned), un-normalised logits plus
//                      the per-slice (max, sum exp) pairs of the softmax GEMM epilogue; the softmax
//                      division, the min_prob floor and the log are all applied here, so the
//                      posterior matrix is never written to HBM on the fused basecall path.
constexpr int IN_POST = 0, IN_LOG = 1, IN_LOGITS = 2;
constexpr int VIT_THREADS = 128;

__device__ __forceinline__ void cp_async16_v(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
// 16-byte copy of which only the first `nbytes` are read from global memory (the rest is zero filled)
__device__ __forceinline__ void cp_async16_zfill_v(void *smem_dst, const void *gsrc, int nbytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async4_v(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
// the same with the destination given as a 32-bit shared-memory address (computed once outside the event loop)
__device__ __forceinline__ void cp_async16_sa(uint32_t d, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill_sa(uint32_t d, const void *gsrc, int nbytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async4_sa(uint32_t d, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
// 1-D bulk copy (TMA without a tensor map) of one whole row, issued by ONE thread; completion is counted in bytes on an
// mbarrier that every thread then waits on (try_wait suspends the thread in hardware)
__device__ __forceinline__ void vit_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void vit_bulk_row(uint32_t dst, const void *gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool vit_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void vit_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "VWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 200000;\n\t"
        "@p bra VDONE_%=;\n\t"
        "bra VWAIT_%=;\n\t"
        "VDONE_%=:\n\t}"
        ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void cp_async_wait0_v() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit_v() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1_v() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(VIT_THREADS, 7)      // 7 CTAs/SM: all 1024 reads of a batch resident in one wave on 148 SMs
viterbi_k1024_kernel(const float *__restrict__ post, long ld_t, long ld_b, const float *__restrict__ rowms,
                     const int32_t *__restrict__ lengths, int T, int B, float skip_pen, float c0, float c1,
                     uint8_t *__restrict__ tb, int32_t *__restrict__ path_out,
                     int32_t *__restrict__ path_len, float *__restrict__ score_out)
{
    constexpr int K = 1024, RS = 256, RK = 64, NT = VIT_THREADS;
    __shared__ __align__(16) float vbuf[2][K];
    __shared__ __align__(16) float2 m4_s[256];       // per quad q4: (max_a p[a*256 + q4], 4 * argmax) of the current event
    __shared__ float red_v[NT / 32];
    __shared__ int red_i[NT / 32];
    __shared__ int s_best, s_state;
    // the next event's row, staged with cp.async.  [0, K) k-mer columns, [K] stay column.  The storage doubles as
    // the second traceback chunk buffer of the backtrace.
    constexpr int XROW = K + 12;                     // K k-mer columns (+ up to 3 floats of alignment phase + 1 chunk), stay at [XROW - 1]
    __shared__ __align__(16) float xrow_s[2][XROW];
    __shared__ __align__(8) uint64_t xbar[2];         // LOGITS mode: a row staged by one bulk copy (TMA) has landed
    uint8_t *tb_s2 = reinterpret_cast<uint8_t *>(&xrow_s[0][0]);
    static_assert(sizeof(float) * 2 * XROW >= 8 * 1024, "xrow_s doubles as a traceback chunk buffer");
    constexpr int TBROW = K / 4;                     // uint16 traceback entries (one per quad of states) per event

    const int b = blockIdx.x;
    const int r = threadIdx.x;
    const int nev = lengths ? min(lengths[b], T) : T;
    if (nev < 1) {
        if (r == 0) { path_len[b] = 0; score_out[b] = 0.0f; }
        return;
    }
    const float *pb = post + (long)b * ld_b;
    uint16_t *tbb = reinterpret_cast<uint16_t *>(tb) + (size_t)b * (size_t)T * (size_t)TBROW;
    // The thread's two quads 2r and 2r+1 are 32 bytes apart, so a warp's 128-bit accesses "quad 2r of every lane"
    // would touch only every other 16-byte chunk: a 2-way bank conflict on each.  Lanes with bit 2 set therefore take
    // their odd quad first (slot A) and the even one second (slot B): any 8 consecutive lanes then cover 8 different
    // chunks modulo 128 bytes in both accesses.
    const int swp = (r >> 2) & 1;
    const int qa = 2 * r + swp, qb = 2 * r + 1 - swp;                 // quad index (16-byte chunk) of slot A / B

    // LOGITS mode: -(row max * log2 e + log2 sum exp(logit - max)) of the softmax row, combined from the GEMM's per-slice
    // statistics by softmax_rowms_kernel, sits in the first padding column of the row (column K + 1) and is staged with it.
    // (It used to be reduced here, by warp 0 with ten shuffles per event, which every other warp then waited for at the
    // barrier.)
    // stage the row of the next event.  The k-mer columns of a row start at `a` = row (logits layout) or row + 1
    // (posterior layout), which is 16-byte aligned only in the first case; in general the row is copied as the
    // 16-byte ALIGNED chunks that cover it (257 128-bit cp.async per row instead of 1024 32-bit ones), keeping its
    // alignment phase ph = (a / 4 bytes) mod 4 in shared memory: column j lands at xrow[ph + j].  The last chunk is
    // copied with its valid byte count only (zero filled), so nothing is read past the row's end; the first chunk
    // starts at most 3 floats before `a`, inside the tensor because its base is 16-byte aligned (checked: otherwise
    // 32-bit copies).  Thread 0 also stages the stay column at xrow[XROW - 1].
    //
    // The event loop below is unrolled by two so that every buffer index (row staging, v ping-pong) is a compile-time
    // constant, and the staging addresses are 32-bit shared-memory addresses computed once: with run-time buffer
    // pointers the compiler rebuilt generic addresses from the shared window base every event (~50 of the ~285
    // instructions a thread issued per event; the kernel is issue-bound).
    const float *rowp = pb;                                           // advanced by ld_t per event
    constexpr int KOFF = MODE == IN_LOGITS ? 0 : 1;                    // first k-mer column of a row
    const bool span_rows = (((uintptr_t)post & 15) == 0);
    // LOGITS rows are 16-byte aligned by contract (sloika_viterbi_logits_fwd checks base and pitches): phase 0
    auto phase_of = [&](const float *row) -> int {
        return MODE == IN_LOGITS ? 0 : (span_rows ? (int)(((uintptr_t)(row + KOFF) >> 2) & 3) : 0);
    };
    const uint32_t xs_sa = (uint32_t)__cvta_generic_to_shared(&xrow_s[0][0]);
    const uint32_t xs_own = xs_sa + 32u * (uint32_t)r;                // this thread's two chunks of a staged row
    constexpr uint32_t XROWB = XROW * sizeof(float);
    const uint32_t xbar_sa = (uint32_t)__cvta_generic_to_shared(&xbar[0]);
    if (MODE == IN_LOGITS) {
        if (r == 0) {
            vit_mbar_init(xbar_sa, 1);
            vit_mbar_init(xbar_sa + 8u, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    // row j is the (j >> 1)-th fill of buffer j & 1: its barrier phase has parity (j >> 1) & 1
    auto wait_row = [&](auto buf_c, int j) {
        if (MODE == IN_LOGITS) vit_mbar_wait(xbar_sa + 8u * (uint32_t)decltype(buf_c)::value, (uint32_t)(j >> 1) & 1u);
    };
    constexpr int MSCOL = MODE == IN_LOGITS ? K + 1 : XROW - 2;       // where the row statistic of a staged row lands
    auto stage_row = [&](auto buf_c) {
        constexpr uint32_t boff = (uint32_t)decltype(buf_c)::value * XROWB;
        const float *row = rowp;
        rowp += ld_t;
        if (MODE == IN_LOGITS) {
            // aligned rows [kmer 0..1023, stay, row statistic, pad, pad] (the statistic was put into the first padding
            // column by softmax_rowms_kernel): ONE bulk copy of 4112 bytes by thread 0 instead of two cp.async per
            // thread (each of which the compiler pads with three dummy shared-memory loads) plus two special cases
            // (warp-uniform condition + elect.sync: with `r == 0` the compiler wraps the uniform-datapath copy in a loop over lanes)
            if (r < 32 && vit_elect_one())
                vit_bulk_row(xs_sa + boff, row, (uint32_t)(K + 4) * 4u, xbar_sa + 8u * (uint32_t)decltype(buf_c)::value);
            return;
        }
        const float *a = row + KOFF;
        if (span_rows) {
            const int ph = (int)(((uintptr_t)a >> 2) & 3);
            const float *a0 = a - ph;                                  // 16-byte aligned
            cp_async16_sa(xs_own + boff, a0 + 8 * r);
            cp_async16_sa(xs_own + boff + 16u, a0 + 8 * r + 4);
            if (r == 0 && ph != 0) cp_async16_zfill_sa(xs_sa + boff + 4u * K, a0 + K, 4 * ph);
        } else {
#pragma unroll
            for (int c = 0; c < 8; c++) cp_async4_sa(xs_own + boff + 4u * c, a + 8 * r + c);
        }
        if (r == 0) cp_async4_sa(xs_sa + boff + 4u * (XROW - 1), row);
        cp_async_commit_v();
    };
    constexpr int STAY = MODE == IN_LOGITS ? K : XROW - 1;            // where the stay column of a staged row lands
    // this thread's 8 k-mer columns of a staged row: three aligned 128-bit loads (conflict free; 32-bit loads at a
    // stride of 8 floats would hit 4 banks), then a shift by the row's phase (uniform over the CTA)
    auto load_cols = [&](const float *xr, int ph, float (&x)[8]) {
        if (MODE == IN_LOGITS || ph == 0) {          // slot order (A then B), see qa / qb
            const float4 xa = reinterpret_cast<const float4 *>(xr)[qa], xb = reinterpret_cast<const float4 *>(xr)[qb];
            x[0] = xa.x; x[1] = xa.y; x[2] = xa.z; x[3] = xa.w; x[4] = xb.x; x[5] = xb.y; x[6] = xb.z; x[7] = xb.w;
            return;
        }
        const float4 xa = reinterpret_cast<const float4 *>(xr)[2 * r], xb = reinterpret_cast<const float4 *>(xr)[2 * r + 1];
        const float4 xc = reinterpret_cast<const float4 *>(xr)[2 * r + 2];
        const float w[12] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w, xc.x, xc.y, xc.z, xc.w};
        float y[8];
        if (ph == 1) {
#pragma unroll
            for (int c = 0; c < 8; c++) y[c] = w[c + 1];
        } else if (ph == 2) {
#pragma unroll
            for (int c = 0; c < 8; c++) y[c] = w[c + 2];
        } else {
#pragma unroll
            for (int c = 0; c < 8; c++) y[c] = w[c + 3];
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {           // state order -> slot order
            x[c] = swp ? y[4 + c] : y[c];
            x[4 + c] = swp ? y[c] : y[4 + c];
        }
    };
    auto lpost_of = [&](float v, float2 ms) -> float {
        if (MODE == IN_LOG) return v;
        if (MODE == IN_LOGITS) {
            // fused path: softmax (exp(t - m) * 1/rowsum), min_prob floor and log with the MUFU ex2 / lg2
            // approximations (|error| ~1e-6 on a log-posterior, inside what libm-vs-device logf already allows);
            // ms = (-(m * log2e + log2 rowsum), unused): posterior = 2^(v*log2e + ms.x), 5 instructions per value
            const float pr = ex2_ftz(fmaf(v, SLOIKA_LOG2E, ms.x));
            return lg2_ftz(fmaf(c1, pr, c0)) * SLOIKA_LN2;             // c0 carries min_prob + 1e-10 on this path
        }
        return logf(__fadd_rn(__fadd_rn(c0, __fmul_rn(c1, v)), VIT_ETA));
    };
    // inside the event loop the LOGITS form keeps log2 of the floored posterior and folds the factor ln 2 into the
    // addition of the predecessor's score (one FFMA instead of FMUL + FADD per state; the log-posterior is a MUFU
    // approximation on this path anyway, and the fused product is the more accurate one)
    auto lpost2_of = [&](float v, float2 ms) -> float {
        if (MODE == IN_LOGITS) return lg2_ftz(fmaf(c1, ex2_ftz(fmaf(v, SLOIKA_LOG2E, ms.x)), c0));
        return lpost_of(v, ms);
    };
    auto add_lpost = [&](float lq, float best) -> float {
        return MODE == IN_LOGITS ? fmaf(lq, SLOIKA_LN2, best) : __fadd_rn(lq, best);
    };
    using buf0 = std::integral_constant<int, 0>;
    using buf1 = std::integral_constant<int, 1>;

    const float *rowq = pb;                                           // row of the event being consumed (phase only)
    stage_row(buf0());
    cp_async_wait0_v();
    wait_row(buf0(), 0);
    __syncthreads();
    {
        const float2 ms = make_float2(xrow_s[0][MSCOL], 0.0f);
        float q[8];
        load_cols(xrow_s[0], phase_of(rowq), q);
        rowq += ld_t;
        float4 va, vb;
        va.x = lpost_of(q[0], ms); va.y = lpost_of(q[1], ms); va.z = lpost_of(q[2], ms); va.w = lpost_of(q[3], ms);
        vb.x = lpost_of(q[4], ms); vb.y = lpost_of(q[5], ms); vb.z = lpost_of(q[6], ms); vb.w = lpost_of(q[7], ms);
        reinterpret_cast<float4 *>(vbuf[0])[qa] = va;                // v_0 = lpost[0][1:]   (decode.py:57); q[] is in slot order
        reinterpret_cast<float4 *>(vbuf[0])[qb] = vb;
    }
    if (nev > 1) stage_row(buf1());
    cp_async_wait0_v();
    __syncthreads();

    uint32_t *tbp = reinterpret_cast<uint32_t *>(tbb) + r;            // this thread's two traceback entries, advanced per event
    // one event: its row is staged in xrow_s[XB] (XB = i & 1), v_{i-1} is vbuf[XB ^ 1], v_i goes to vbuf[XB]
    auto event = [&](auto xb_c, const int i, const bool more) {
        constexpr int XB = decltype(xb_c)::value;
        constexpr int CUR = XB ^ 1;
        const float *xr = xrow_s[XB];
        wait_row(xb_c, i);                                           // LOGITS mode; the others waited an event ago
        float x[8];
        load_cols(xr, phase_of(rowq), x);
        rowq += ld_t;
        const float x0 = xr[STAY];
        const float2 ms = make_float2(xr[MSCOL], 0.0f);
        if (more) stage_row(std::integral_constant<int, CUR>());     // next event, asynchronous
        const float *p = vbuf[CUR];
        // step: first maximum over a of p[a*256 + q4] for the two quads q4 = 2r, 2r+1; published as (value, 4*a)
        float2 ss = reinterpret_cast<const float2 *>(p)[r];
        int as0 = 0, as1 = 0;
#pragma unroll
        for (int a = 1; a < 4; a++) {
            const float2 c = reinterpret_cast<const float2 *>(p + a * RS)[r];
            if (c.x > ss.x) { ss.x = c.x; as0 = a; }
            if (c.y > ss.y) { ss.y = c.y; as1 = a; }
        }
        reinterpret_cast<float4 *>(m4_s)[r] = make_float4(ss.x, __int_as_float(4 * as0), ss.y, __int_as_float(4 * as1));
        const float4 pa = reinterpret_cast<const float4 *>(p)[qa], pb4 = reinterpret_cast<const float4 *>(p)[qb];
        // step maximum / argument of slot A and slot B
        const float ssA = swp ? ss.y : ss.x, ssB = swp ? ss.x : ss.y;
        const int asA = swp ? as1 : as0, asB = swp ? as0 : as1;
        const float lp0 = lpost_of(x0, ms);
        float lp[8];
#pragma unroll
        for (int c = 0; c < 8; c++) lp[c] = lpost2_of(x[c], ms);
        __syncthreads();
        // skip: predecessor a*64 + q with a = 4*a_hi + a_lo is p[a_hi*256 + (a_lo*64 + q)], so its maximum is
        // the maximum over a_lo of the published step maxima of quads a_lo*64 + q; "first maximum" = smallest a
        const int q = r >> 1;
        float sk;
        int ak;
        {
            const float2 e0 = m4_s[q];
            sk = e0.x;
            ak = __float_as_int(e0.y);
#pragma unroll
            for (int al = 1; al < 4; al++) {
                const float2 e = m4_s[al * RK + q];
                const int key = __float_as_int(e.y) + al;
                if (e.x > sk || (e.x == sk && key < ak)) { sk = e.x; ak = key; }
            }
        }
        sk = __fsub_rn(sk, skip_pen);
        const float pj[8] = {pa.x, pa.y, pa.z, pa.w, pb4.x, pb4.y, pb4.z, pb4.w};
        float vo[8];
        unsigned ent[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {                                    // h = slot (A, B)
            const float ssh = h == 0 ? ssA : ssB;
            const bool use_step = ssh > sk;                              // tie -> skip (decode.py:76)
            const float best = use_step ? ssh : sk;
            unsigned e = (use_step ? (1u + (unsigned)(h == 0 ? asA : asB)) : (5u + (unsigned)ak)) << 4;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float move = add_lpost(lp[4 * h + c], best);
                const float stay = __fadd_rn(pj[4 * h + c], lp0);
                const bool mv = move > stay;                             // tie -> stay (decode.py:81)
                vo[4 * h + c] = mv ? move : stay;
                e |= (mv ? 1u : 0u) << c;
            }
            ent[h] = e;
        }
        const unsigned packed = swp ? (ent[1] | (ent[0] << 16)) : (ent[0] | (ent[1] << 16));      // even quad in the low half
        reinterpret_cast<float4 *>(vbuf[XB])[qa] = make_float4(vo[0], vo[1], vo[2], vo[3]);
        reinterpret_cast<float4 *>(vbuf[XB])[qb] = make_float4(vo[4], vo[5], vo[6], vo[7]);
        tbp += TBROW / 2;
        *tbp = packed;
        if (MODE != IN_LOGITS) cp_async_wait0_v();                   // the next row has landed (issued an event ago)
        __syncthreads();
    };
    {
        int i = 1;
        for (; i + 2 < nev; i += 2) {                                  // events i (odd) and i + 1, both with a successor
            event(buf1(), i, true);
            event(buf0(), i + 1, true);
        }
        if (i < nev) event(buf1(), i, i + 1 < nev);
        if (i + 1 < nev) event(buf0(), i + 1, false);
    }
    const int cur = (nev - 1) & 1;                                       // v of the last event

    // ---- argmax of v_T (first maximum), backtrace, left-align: as in the generic kernel ----
    const float *v = vbuf[cur];
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const float val = v[8 * r + c];
        if (val > bv) { bv = val; bi = 8 * r + c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((r & 31) == 0) { red_v[r >> 5] = bv; red_i[r >> 5] = bi; }
    __syncthreads();
    if (r == 0) {
        for (int w = 1; w < NT / 32; w++)
            if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi)) { bv = red_v[w]; bi = red_i[w]; }
        score_out[b] = bv;
        s_state = bi;
        s_best = nev - 1;
        path_out[(size_t)b * T + nev - 1] = bi;
    }
    // ---- backtrace (decode.py:84-91).  The walk itself is sequential (one thread), but each step would
    // be a dependent ~1 us HBM access; instead the whole CTA streams the traceback rows backwards in chunks
    // of 16 events through shared memory (cp.async, next chunk in flight while this one is walked). ----
    {
        constexpr int CH = 16;
        constexpr int ROWB = TBROW * 2;                                    // bytes per traceback row
        // two 8 KB chunk buffers: vbuf (free now that v_T has been reduced) and tb_s2
        uint8_t *bufs[2] = {reinterpret_cast<uint8_t *>(&vbuf[0][0]), tb_s2};
        auto issue = [&](int hi, int which) {                              // rows (hi-CH, hi], clipped at 1
            uint8_t *dst = bufs[which];
            for (int e = r; e < CH * (ROWB / 16); e += NT) {
                const int row = hi - e / (ROWB / 16);
                if (row >= 1)
                    cp_async16_v(dst + (size_t)e * 16,
                                 reinterpret_cast<const uint8_t *>(tbb + (size_t)row * TBROW) + (size_t)(e % (ROWB / 16)) * 16);
            }
            cp_async_commit_v();
        };
        __syncthreads();
        int hi = nev - 1, which = 0;
        if (hi >= 1) issue(hi, 0);
        while (hi >= 1) {
            const int nxt = hi - CH;
            if (nxt >= 1) issue(nxt, which ^ 1); else cp_async_commit_v();
            cp_async_wait1_v();
            __syncthreads();
            if (r == 0) {
                const uint16_t *src = reinterpret_cast<const uint16_t *>(bufs[which]);
                int state = s_state, pos = s_best;
                int32_t *out = path_out + (size_t)b * T;
                for (int k = 0; k < CH && hi - k >= 1; k++) {
                    const unsigned e = src[k * TBROW + (state >> 2)];
                    if ((e >> (state & 3)) & 1u) {
                        const unsigned cd = e >> 4;
                        state = cd <= 4u ? (int)(cd - 1) * RS + (state >> 2) : (int)(cd - 5) * RK + (state >> 4);
                        out[--pos] = state;
                    }
                }
                s_state = state;
                s_best = pos;
            }
            __syncthreads();
            hi = nxt;
            which ^= 1;
        }
        if (r == 0) path_len[b] = nev - s_best;
    }
    __syncthreads();
    const int off = s_best;
    const int n = nev - off;
    if (off > 0) {
        int32_t *out = path_out + (size_t)b * T;
        for (int base = 0; base < n; base += NT) {
            const int idx = base + r;
            int32_t val = 0;
            if (idx < n) val = out[off + idx];
            __syncthreads();
            if (idx < n) out[idx] = val;
            __syncthreads();
        }
    }
}

// One float per softmax row from the GEMM's per-slice (max, sum exp) pairs: -(M log2 e + log2 S) with M the row maximum
// and S = sum_j exp(logit_j - M), so that posterior_j = 2^(logit_j * log2 e + rowms).
// It is stored in the first PADDING column of the row it belongs to (column K + 1 = 1025 of a row pitch that is a multiple
// of 4 floats), so that the decoder stages k-mer logits, stay logit and statistic with one contiguous copy.
__global__ void softmax_rowms_kernel(const float2 *__restrict__ stats, int n_slices, long M, float *__restrict__ logits,
                                     long ld_t, long ld_b, int B, int col)
{
    for (long m = blockIdx.x * (long)blockDim.x + threadIdx.x; m < M; m += (long)gridDim.x * blockDim.x) {
        const float2 *st = stats + m * n_slices;
        float mx = -INFINITY;
        for (int s = 0; s < n_slices; s++) mx = fmaxf(mx, __ldg(&st[s]).x);
        float tot = 0.0f;
        for (int s = 0; s < n_slices; s++) {
            const float2 v = __ldg(&st[s]);
            tot += v.y * ex2_ftz((v.x - mx) * SLOIKA_LOG2E);
        }
        logits[(m / B) * ld_t + (m % B) * ld_b + col] = -(mx * SLOIKA_LOG2E + lg2_ftz(tot));
    }
}

static bool ipow_ok(int nbase, int klen, long *K)
{
    long k = 1;
    for (int i = 0; i < klen; i++) {
        k *= nbase;
        if (k > (1L << 20)) return false;
    }
    *K = k;
    return true;
}

static bool use_k1024(int nbase, long K) { return nbase == 4 && K == 1024 && !getenv("SLOIKA_B200_VITERBI_GENERIC"); }

}  // namespace sloika

using namespace sloika;

extern "C" size_t sloika_viterbi_workspace_bytes(int T, int B, int nbase, int klen)
{
    long K;
    if (T < 0 || B < 0 || nbase < 2 || klen < 1 || !ipow_ok(nbase, klen, &K)) return 0;
    // K = 1024 kernel: one uint16 per quad of states; generic kernel: one byte per state
    // (+ one float per event and read for the combined softmax row statistics of the fused-logits path)
    if (use_k1024(nbase, K)) return (size_t)T * (size_t)B * (size_t)(K / 2) + (size_t)T * (size_t)B * sizeof(float);
    return (size_t)T * (size_t)B * (size_t)K;
}

extern "C" int sloika_viterbi_fwd(const float *post, long ld_t, long ld_b, const int32_t *lengths, int T, int B,
                                  int nbase, int klen, double skip_pen, double min_prob, int mode, void *tb_ws,
                                  size_t ws_bytes, int32_t *path_out, int32_t *path_len, float *score_out,
                                  void *stream)
{
    if (!post || !path_out || !path_len || !score_out || T < 0 || B <= 0) return SLOIKA_ERR_ARG;
    if (klen < 3) return SLOIKA_ERR_ARG;                 // decode.py:50
    long K;
    if (nbase < 2 || !ipow_ok(nbase, klen, &K)) return SLOIKA_ERR_ARG;
    if (mode != SLOIKA_VIT_POST && mode != SLOIKA_VIT_LOG) return SLOIKA_ERR_ARG;
    if (nbase != 4 && nbase != 5) return SLOIKA_ERR_UNSUPPORTED;
    if (2 * K * sizeof(float) > 200 * 1024) return SLOIKA_ERR_UNSUPPORTED;
    if (!tb_ws || ws_bytes < sloika_viterbi_workspace_bytes(T, B, nbase, klen)) return SLOIKA_ERR_WORKSPACE;
    if (T == 0) return SLOIKA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = 2 * (size_t)K * sizeof(float);
    // python-float constants enter the float32 arithmetic exactly as NumPy casts them (decode.py:36, :72)
    const float c0 = (float)min_prob, c1 = (float)(1.0 - min_prob), sp = (float)skip_pen;
    long threads = K / nbase;
    threads = threads > 256 ? 256 : (threads < 32 ? 32 : ceil_div(threads, 32) * 32);
    cudaError_t err;
    if (use_k1024(nbase, K)) {
        if (mode == SLOIKA_VIT_POST)
            viterbi_k1024_kernel<IN_POST><<<B, VIT_THREADS, 0, st>>>(post, ld_t, ld_b, nullptr, lengths, T, B, sp, c0, c1,
                                                              (uint8_t *)tb_ws, path_out, path_len, score_out);
        else
            viterbi_k1024_kernel<IN_LOG><<<B, VIT_THREADS, 0, st>>>(post, ld_t, ld_b, nullptr, lengths, T, B, sp, c0, c1,
                                                             (uint8_t *)tb_ws, path_out, path_len, score_out);
        SLOIKA_RETURN_LAUNCH_STATUS();
    }
    if (nbase == 4) {
        err = cudaFuncSetAttribute(viterbi_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        viterbi_kernel<4><<<B, (unsigned)threads, smem, st>>>(post, ld_t, ld_b, lengths, T, B, (int)K, sp,
                                                             c0, c1, mode, (uint8_t *)tb_ws, path_out, path_len,
                                                             score_out);
    } else {
        err = cudaFuncSetAttribute(viterbi_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        viterbi_kernel<5><<<B, (unsigned)threads, smem, st>>>(post, ld_t, ld_b, lengths, T, B, (int)K, sp,
                                                             c0, c1, mode, (uint8_t *)tb_ws, path_out, path_len,
                                                             score_out);
    }
    SLOIKA_RETURN_LAUNCH_STATUS();
}

extern "C" int sloika_viterbi_logits_fwd(float *logits, long ld_t, long ld_b, const float *stats, int n_slices,
                                         const int32_t *lengths, int T, int B, int nbase, int klen, double skip_pen,
                                         double min_prob, void *tb_ws, size_t ws_bytes, int32_t *path_out,
                                         int32_t *path_len, float *score_out, void *stream)
{
    if (!logits || !stats || !path_out || !path_len || !score_out || T <= 0 || B <= 0 || n_slices <= 0 || n_slices > 32)
        return SLOIKA_ERR_ARG;
    if (nbase != 4 || klen != 5) return SLOIKA_ERR_UNSUPPORTED;          // K = 1024 specialisation only
    if ((ld_t & 3) != 0 || (ld_b & 3) != 0 || ((uintptr_t)logits & 15) != 0) return SLOIKA_ERR_UNSUPPORTED;
    if (!tb_ws || ws_bytes < sloika_viterbi_workspace_bytes(T, B, nbase, klen)) return SLOIKA_ERR_WORKSPACE;
    // on this path the kernel evaluates log(c0 + c1 * p) with one fused multiply-add: c0 carries the + 1e-10 of
    // decode.prepare_post (decode.py:36) as well
    const float c0 = (float)min_prob + 1e-10f, c1 = (float)(1.0 - min_prob), sp = (float)skip_pen;
    // the logits rows are (t, b) ordered: ld_t == B * ld_b is what the engine produces and what the statistics index
    if (ld_t != (long)B * ld_b && B > 1) return SLOIKA_ERR_UNSUPPORTED;
    if (ld_b < 1028 && B > 1) return SLOIKA_ERR_UNSUPPORTED;             // the row statistic lives in padding column 1025
    if (ld_t < 1028) return SLOIKA_ERR_UNSUPPORTED;
    const long M = (long)T * B;
    long blocks = ceil_div(M, 256);
    if (blocks > 148L * 16) blocks = 148L * 16;
    softmax_rowms_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(stats), n_slices, M,
                                                                             logits, ld_t, ld_b, B, 1025);
    viterbi_k1024_kernel<IN_LOGITS><<<B, VIT_THREADS, 0, (cudaStream_t)stream>>>(
        logits, ld_t, ld_b, nullptr, lengths, T, B, sp, c0, c1, (uint8_t *)tb_ws, path_out, path_len, score_out);
    SLOIKA_RETURN_LAUNCH_STATUS();
}
