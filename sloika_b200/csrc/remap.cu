// Remap decode: Viterbi path of a transducer posterior through a known k-mer sequence, batched, one CTA per read.
//
// Reference: sloika/transducer.py:14-73 (map_to_sequence) and its only native helper,
// sloika/viterbi_helpers.pyx:12-35 (slip_update); callers sloika/tools/chunkify_raw.py:262-274 and
// sloika/batch.py:141-155 (one read at a time, NumPy + Cython on the CPU).
//
// Per event i the reference updates, over the positions j of the sequence,
//     stay   cur[j] = prev[j] + lt[i][0]                              from j
//     step   prev[j-1] + lt[i][seq[j]]  if strictly greater          from j-1
//     slip   fs[j] + lt[i][seq[j]]      unless fs[j] + .. <= cur[j]  from fp[j]   (any position <= j-2)
// where (fs, fp) = slip_update(prev, slip) is a running maximum with a geometric penalty:
//     fs[j] = (fs[j-1] >= prev[j-2] ? fs[j-1] : prev[j-2]) - slip.
// The running maximum is a chain of float32 subtractions whose roundings the result depends on, so it is kept
// sequential (one thread walks it: ~10 cycles per position) and overlapped with the stay / step work of the other
// threads; the parallelism of the kernel is across reads (a 64-thread CTA per read, many CTAs per SM).
// All arithmetic is float32 in the reference's order, comparisons are written so that NaN behaves as in NumPy
// (the reference turns slip=None into a NaN penalty, see oracle/remap_ref.py), priors are added in double and
// rounded, the final arg-max takes the first maximum with NaN counting as largest: paths and scores are
// bit-identical to the reference's (tests/golden/remap_cases.npz) when the log-transducer is given (log=True);
// with log=False the only difference is device logf vs NumPy's.
//
// Traceback: uint16 source position per (event, position) in HBM (P <= 65535), walked backwards through shared
// memory in chunks of rows like the basecall Viterbi kernel.
#include <cstdlib>
#include "common.cuh"

namespace sloika {
namespace remap {

constexpr int MAX_THREADS = 256;       // 64 threads per read for chunk-sized sequences, 256 for long ones
constexpr int MAX_SEG = 32;            // segments of the slip scan (threads 0..nseg-1): 16 or 32; the walk over them is
                                       // sequential, so few and long beats many and short (measured)
constexpr int TB_CHUNK = 8;            // traceback rows staged per backtrace step: 8 * P * 2 bytes = the 16 P bytes of nxt .. src

__device__ __forceinline__ void cp_async4_r(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_r() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0_r() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// How the running maximum of slip_update is parallelised without changing a bit of it.  Sequentially
//     F[j] = (F[j-1] >= x[j-2] ? F[j-1] : x[j-2]) - slip,       source kept on ties.
// Rounded subtraction of a constant is monotone, so F[j] is the maximum over the candidates k <= j-2 of x[k] with slip
// subtracted (j-1-k) times, each with its own roundings.  Cut the positions in segments: every thread scans ONE
// segment as if nothing came before it (L, "local"); the value carried into a segment decays by one subtraction
// per position for as long as it is still >= every new x it meets (exactly the comparisons the sequential code makes
// while that candidate is its running best); where it survives it overrides L, and the first time it loses, the
// sequential code and the local scan both switch to that same x and agree from there on.  One thread walks the
// segments in order and applies the carries: usually a handful of positions each, at worst one subtract + compare
// per position.
template <int THREADS, int NSEG>
__global__ void __launch_bounds__(THREADS)
remap_kernel(const float *__restrict__ trans, long ld_t, long ld_b, const int32_t *__restrict__ nev_p, int T, int nstate,
             const int32_t *__restrict__ seq, long ld_seq, const int32_t *__restrict__ npos_p, int P, float slip,
             const double *__restrict__ prior0, const double *__restrict__ prior1, long ld_prior, int is_log,
             uint16_t *__restrict__ tb, int32_t *__restrict__ path_out, float *__restrict__ score_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *prev = reinterpret_cast<float *>(smem_raw);     // [P] scores of the previous event
    float *nxt = prev + P;                                 // [P] scores being built
    float *e = nxt + P;                                    // [P] emission lt[i][seq[j]]
    float *fs = e + P;                                     // [P] slip_update scores
    uint16_t *fp = reinterpret_cast<uint16_t *>(fs + P);   // [P] slip_update sources
    uint16_t *src = fp + P;                                // [P] stay / step source
    int32_t *sq = reinterpret_cast<int32_t *>(src + P);    // [P] the sequence (fp + src = 4 P bytes: stays 4-byte aligned)
    float *rowbuf = reinterpret_cast<float *>(sq + P);     // [2][nstate] transducer rows, staged one event ahead
    __shared__ float seg_f[MAX_SEG];                       // end state of every local scan
    __shared__ int seg_p[MAX_SEG];
    __shared__ int s_pos;

    const int b = blockIdx.x, tid = threadIdx.x;
    const int nev = nev_p ? min(nev_p[b], T) : T;
    const int npos = npos_p ? min(npos_p[b], P) : P;
    if (nev < 1 || npos < 3) {                             // the reference indexes from_score[2] unconditionally
        if (tid == 0) score_out[b] = nanf("");
        return;
    }
    const float *tr = trans + (long)b * ld_b;
    uint16_t *tbb = tb + (size_t)b * (size_t)T * (size_t)P;
    auto stage_row = [&](int i) {                          // row i -> rowbuf[i & 1] (asynchronous)
        const float *row = tr + (long)i * ld_t;
        float *dst = rowbuf + (i & 1) * nstate;
        for (int c = tid; c < nstate; c += THREADS) cp_async4_r(dst + c, row + c);
        cp_async_commit_r();
    };
    auto lval = [&](const float *row, int col) -> float { return is_log ? row[col] : logf(row[col]); };

    stage_row(0);
    for (int j = tid; j < npos; j += THREADS) sq[j] = __ldg(seq + (long)b * ld_seq + j);
    cp_async_wait0_r();
    __syncthreads();
    if (nev > 1) stage_row(1);
    {   // transducer.py:41-44
        const float stay0 = lval(rowbuf, 0);
        for (int j = tid; j < npos; j += THREADS) {
            const float p = prior0 ? (float)(0.0 + prior0[(long)b * ld_prior + j]) : 0.0f;
            prev[j] = __fadd_rn(p, fmaxf(lval(rowbuf, sq[j]), stay0));
        }
    }
    cp_async_wait0_r();
    __syncthreads();

    // segments of the positions 3 .. npos-1 for the local scans
    const int seg_len = (max(npos - 3, 0) + NSEG - 1) / NSEG;
    const int seg_a = 3 + tid * seg_len, seg_b = tid < NSEG ? min(seg_a + seg_len, npos) : 0;

    for (int i = 1; i < nev; i++) {
        const float *row = rowbuf + (i & 1) * nstate;
        // ---- stay / step (transducer.py:48-55): threads NSEG.. ----
        if (tid >= NSEG) {
            const float stay = lval(row, 0);
            for (int j = tid - NSEG; j < npos; j += THREADS - NSEG) {
                const float em = lval(row, sq[j]);
                e[j] = em;
                float c = __fadd_rn(prev[j], stay);
                int s = j;
                if (j >= 1) {
                    const float st = __fadd_rn(prev[j - 1], em);
                    if (st > c) { c = st; s = j - 1; }      // tie -> stay
                }
                nxt[j] = c;
                src[j] = (uint16_t)s;
            }
        }
        // ---- slip_update (viterbi_helpers.pyx:22-33): local scan of this thread's segment, meanwhile ----
        if (seg_a < seg_b) {
            float f = __fsub_rn(prev[seg_a - 2], slip);
            int pos = seg_a - 2;
            fs[seg_a] = f; fp[seg_a] = (uint16_t)pos;
            for (int j = seg_a + 1; j < seg_b; j++) {
                const float x = prev[j - 2];
                const bool keep = f >= x;                   // tie keeps the older source; NaN takes the new one
                pos = keep ? pos : j - 2;
                f = __fsub_rn(keep ? f : x, slip);
                fs[j] = f;
                fp[j] = (uint16_t)pos;
            }
            seg_f[tid] = f;
            seg_p[tid] = pos;
        }
        __syncthreads();
        if (i + 1 < nev && tid != 0) stage_row(i + 1);      // the row buffer of event i-1 is free: next row, asynchronously
        if (tid == 0) {
            // ---- carries across the segments, in order ----
            float cf = __fsub_rn(prev[0], slip);            // F[2]
            int cp = 0;
            fs[0] = -1e38f; fs[1] = -1e38f; fp[0] = 0; fp[1] = 0;
            fs[2] = cf; fp[2] = 0;
            for (int a = 3, t = 0; a < npos; a += seg_len, t++) {
                const int bnd = min(a + seg_len, npos);
                float d = cf;
                bool alive = true;
                for (int j = a; j < bnd; j++) {
                    alive = d >= prev[j - 2];               // the comparison the sequential code makes at j
                    if (!alive) break;
                    d = __fsub_rn(d, slip);
                    fs[j] = d;
                    fp[j] = (uint16_t)cp;
                }
                if (alive) cf = d;                          // still the running best at the end of the segment
                else { cf = seg_f[t]; cp = seg_p[t]; }      // lost inside it: the local scan's end state is the truth
            }
            if (i + 1 < nev) {                              // thread 0's share of the row staging
                const float *rown = tr + (long)(i + 1) * ld_t;
                float *dst = rowbuf + ((i + 1) & 1) * nstate;
                for (int c = 0; c < nstate; c += THREADS) cp_async4_r(dst + c, rown + c);
                cp_async_commit_r();
            }
        }
        __syncthreads();
        // ---- slip (transducer.py:56-60) ----
        uint16_t *tbi = tbb + (size_t)i * P;
        for (int j = tid; j < npos; j += THREADS) {
            const float from = __fadd_rn(fs[j], e[j]);
            float c = nxt[j];
            uint16_t s = src[j];
            if (!(from <= c)) { c = from; s = fp[j]; }      // tie -> no slip; NaN -> slip (np.where(from <= c, ...))
            prev[j] = c;
            tbi[j] = s;
        }
        cp_async_wait0_r();
        __syncthreads();
    }

    if (prior1)                                             // transducer.py:66-67: float64 add, float32 store
        for (int j = tid; j < npos; j += THREADS) prev[j] = (float)((double)prev[j] + prior1[(long)b * ld_prior + j]);
    __syncthreads();
    if (tid == 0) {                                         // np.argmax: first maximum, NaN counts as the maximum
        int best = 0;
        for (int j = 1; j < npos; j++) {
            if (isnan(prev[best])) break;
            if (isnan(prev[j]) || prev[j] > prev[best]) best = j;
        }
        score_out[b] = prev[best];
        s_pos = best;
        path_out[(size_t)b * T + nev - 1] = best;
    }
    __syncthreads();
    // backtrace (transducer.py:70-73): rows (hi - TB_CHUNK, hi] staged through shared memory, then walked by thread 0
    uint16_t *stage = reinterpret_cast<uint16_t *>(nxt);    // nxt, e, fs, fp, src are free now: 16 P bytes = TB_CHUNK rows
    for (int hi = nev - 1; hi >= 1; hi -= TB_CHUNK) {
        const int lo = max(hi - TB_CHUNK + 1, 1);
        for (int k = tid; k < (hi - lo + 1) * npos; k += THREADS) {
            const int rr = k / npos, j = k - rr * npos;
            stage[rr * npos + j] = tbb[(size_t)(lo + rr) * P + j];
        }
        __syncthreads();
        if (tid == 0) {
            int p = s_pos;
            for (int i = hi; i >= lo; i--) {
                p = stage[(i - lo) * npos + p];
                path_out[(size_t)b * T + i - 1] = p;
            }
            s_pos = p;
        }
        __syncthreads();
    }
}

// slip_update alone (API parity with sloika.viterbi_helpers.slip_update; one thread)
__global__ void slip_update_kernel(const float *__restrict__ x, int n, float slip, float *__restrict__ from_score,
                                   long long *__restrict__ from_pos)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int j = 0; j < n && j < 2; j++) { from_score[j] = -1e38f; from_pos[j] = 0; }
    if (n < 3) return;
    float f = __fsub_rn(x[0], slip);
    long long pos = 0;
    from_score[2] = f; from_pos[2] = 0;
    for (int j = 3; j < n; j++) {
        const float v = x[j - 2];
        const bool keep = f >= v;
        pos = keep ? pos : j - 2;
        f = __fsub_rn(keep ? f : v, slip);
        from_score[j] = f;
        from_pos[j] = pos;
    }
}

}  // namespace remap
}  // namespace sloika

using namespace sloika;

extern "C" size_t sloika_remap_workspace_bytes(int T, int B, int P)
{
    if (T < 0 || B < 0 || P < 0) return 0;
    return sizeof(uint16_t) * (size_t)T * (size_t)B * (size_t)P;
}

static size_t remap_smem_bytes(int P, int nstate)
{
    // prev, nxt, e, fs (float) + fp, src (uint16) + sq (int32) + two staged transducer rows
    return (size_t)P * (4 * 4 + 2 * 2 + 4) + (size_t)2 * nstate * 4 + 16;
}

extern "C" int sloika_remap_fwd(const float *trans, long ld_t, long ld_b, const int32_t *nev, int T, int B, int nstate,
                                const int32_t *seq, long ld_seq, const int32_t *npos, int P, double slip, int has_slip,
                                const double *prior_initial, const double *prior_final, long ld_prior, int is_log,
                                void *ws, size_t ws_bytes, int32_t *path_out, float *score_out, void *stream)
{
    if (!trans || !seq || !path_out || !score_out || T < 1 || B <= 0 || nstate < 2 || P < 3) return SLOIKA_ERR_ARG;
    if (has_slip && !(slip >= 0.0)) return SLOIKA_ERR_ARG;              // transducer.py:26
    if ((prior_initial || prior_final) && ld_prior < P) return SLOIKA_ERR_ARG;
    if (P > 65535) return SLOIKA_ERR_UNSUPPORTED;                       // uint16 traceback
    const size_t smem = remap_smem_bytes(P, nstate);
    if (smem > 200 * 1024) return SLOIKA_ERR_UNSUPPORTED;               // sequence too long for the on-chip score vectors
    if (!ws || ws_bytes < sloika_remap_workspace_bytes(T, B, P)) return SLOIKA_ERR_WORKSPACE;
    const float pen = has_slip ? (float)slip : nanf("");                // np.float32(None) is NaN in the reference
#define REMAP_LAUNCH(TH, NS)                                                                                          \
    {                                                                                                                \
        cudaError_t err = cudaFuncSetAttribute(remap::remap_kernel<TH, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               (int)smem);                                                           \
        if (err != cudaSuccess) return (int)err;                                                                     \
        remap::remap_kernel<TH, NS><<<B, TH, smem, (cudaStream_t)stream>>>(                                          \
            trans, ld_t, ld_b, nev, T, nstate, seq, ld_seq, npos, P, pen, prior_initial, prior_final, ld_prior, is_log, \
            static_cast<uint16_t *>(ws), path_out, score_out);                                                       \
    }
    if (P <= 1024) REMAP_LAUNCH(64, 16)                                 // chunk-sized sequences: many small CTAs per SM
    else REMAP_LAUNCH(remap::MAX_THREADS, remap::MAX_SEG)               // long sequences: more threads per read
#undef REMAP_LAUNCH
    SLOIKA_RETURN_LAUNCH_STATUS();
}

extern "C" int sloika_slip_update_fwd(const float *x, int n, float slip, float *from_score, long long *from_pos, void *stream)
{
    if (!x || !from_score || !from_pos || n < 3) return SLOIKA_ERR_ARG;
    remap::slip_update_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(x, n, slip, from_score, from_pos);
    SLOIKA_RETURN_LAUNCH_STATUS();
}
