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

constexpr int THREADS = 64;
constexpr int TB_CHUNK = 8;            // traceback rows staged per backtrace step: 8 * P * 2 bytes <= the 16 P bytes of e + nxt + fs

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
    int32_t *sq = reinterpret_cast<int32_t *>(src + P);     // [P] the sequence (fp + src = 4 P bytes: stays 4-byte aligned)
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
    auto lval = [&](const float *row, int col) -> float {
        const float v = __ldg(row + col);
        return is_log ? v : logf(v);
    };

    for (int j = tid; j < npos; j += THREADS) sq[j] = __ldg(seq + (long)b * ld_seq + j);
    __syncthreads();
    {   // transducer.py:41-44
        const float stay0 = lval(tr, 0);
        for (int j = tid; j < npos; j += THREADS) {
            const float p = prior0 ? (float)(0.0 + prior0[(long)b * ld_prior + j]) : 0.0f;
            prev[j] = __fadd_rn(p, fmaxf(lval(tr, sq[j]), stay0));
        }
    }
    __syncthreads();

    for (int i = 1; i < nev; i++) {
        const float *row = tr + (long)i * ld_t;
        if (tid == 0) {
            // slip_update (viterbi_helpers.pyx:22-33), sequential: value chain = compare/select + one subtraction
            fs[0] = -1e38f; fs[1] = -1e38f;
            fp[0] = 0; fp[1] = 0;
            float f = __fsub_rn(prev[0], slip);
            int pos = 0;
            fs[2] = f; fp[2] = 0;
            for (int j = 3; j < npos; j++) {
                const float x = prev[j - 2];
                const bool keep = f >= x;                   // tie keeps the older source; NaN takes the new one
                pos = keep ? pos : j - 2;
                f = __fsub_rn(keep ? f : x, slip);
                fs[j] = f;
                fp[j] = (uint16_t)pos;
            }
        } else {
            // stay / step (transducer.py:48-55) by the other 63 threads, meanwhile
            const float stay = lval(row, 0);
            for (int j = tid - 1; j < npos; j += THREADS - 1) {
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
        __syncthreads();
        uint16_t *tbi = tbb + (size_t)i * P;
        for (int j = tid; j < npos; j += THREADS) {         // slip (transducer.py:56-60)
            const float from = __fadd_rn(fs[j], e[j]);
            float c = nxt[j];
            uint16_t s = src[j];
            if (!(from <= c)) { c = from; s = fp[j]; }      // tie -> no slip; NaN -> slip (np.where(from <= c, ...))
            prev[j] = c;                                    // every read of prev for this event happened before the barrier
            tbi[j] = s;
        }
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

static size_t remap_smem_bytes(int P)
{
    // prev, nxt, e, fs (float) + fp, src (uint16, padded) + sq (int32)
    return (size_t)P * (4 * 4 + 2 * 2 + 4) + 8;
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
    const size_t smem = remap_smem_bytes(P);
    if (smem > 200 * 1024) return SLOIKA_ERR_UNSUPPORTED;               // sequence too long for the on-chip score vectors
    if (!ws || ws_bytes < sloika_remap_workspace_bytes(T, B, P)) return SLOIKA_ERR_WORKSPACE;
    cudaError_t err = cudaFuncSetAttribute(remap::remap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    const float pen = has_slip ? (float)slip : nanf("");                // np.float32(None) is NaN in the reference
    remap::remap_kernel<<<B, remap::THREADS, smem, (cudaStream_t)stream>>>(
        trans, ld_t, ld_b, nev, T, nstate, seq, ld_seq, npos, P, pen, prior_initial, prior_final, ld_prior, is_log,
        static_cast<uint16_t *>(ws), path_out, score_out);
    SLOIKA_RETURN_LAUNCH_STATUS();
}

extern "C" int sloika_slip_update_fwd(const float *x, int n, float slip, float *from_score, long long *from_pos, void *stream)
{
    if (!x || !from_score || !from_pos || n < 3) return SLOIKA_ERR_ARG;
    remap::slip_update_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(x, n, slip, from_score, from_pos);
    SLOIKA_RETURN_LAUNCH_STATUS();
}
