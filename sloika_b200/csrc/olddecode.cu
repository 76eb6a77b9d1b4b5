// The reference's non-transducer decoder: olddecode.decode_profile (sloika/olddecode.py:13-73) and
// olddecode.estimate_transitions (:94-118), reached from basecall.decode_post (sloika/basecall.py:47-50) for models
// that are not transducers.
//
// decode_profile: Viterbi over K = 4^k k-mer states with per-event log weights (stay, step, skip) and a "slip" from
// the best state.  For event ev with p = score of event ev-1:
//     stay   p[j] + w0                                  from j
//     slip   max_i p[i] + log(eta + slip)               from argmax (first maximum)        wins ties against stay
//     step   max_a p[a*K/4  + j/4 ] + w1                from that a (first maximum)        wins ties against the above
//     skip   max_a p[a*K/16 + j/16] + w2                from that a (first maximum)        wins ties against the above
//     score[j] = fmax of the four;  p'[j] = score[j] + lpost[ev][j];  the predecessor is stored per state and event
// and the whole state sequence (one state per event, stays included) is traced back from the first maximum of the
// last scores.  One CTA per read; the recursion runs in FLOAT64 on float32 log-posteriors, which is what the
// reference computes under NumPy >= 2 (it adds float64 weights to the float32 scores; see oracle/olddecode_ref.py).
// Not a throughput path: legacy models only.
#include <cmath>
#include "common.cuh"

namespace sloika {

__device__ __forceinline__ void argmax_merge(double &v, int &i, double ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// first maximum of p[0..K) over the CTA; result in *out_v / *out_i (shared), valid after the trailing barrier
__device__ void block_argmax(const double *p, int K, double *red_v, int *red_i, double *out_v, int *out_i)
{
    double bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = threadIdx.x; j < K; j += blockDim.x) argmax_merge(bv, bi, p[j], j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        argmax_merge(bv, bi, ov, oi);
    }
    if ((threadIdx.x & 31) == 0) { red_v[threadIdx.x >> 5] = bv; red_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) argmax_merge(bv, bi, red_v[w], red_i[w]);
        *out_v = bv;
        *out_i = bi;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
olddecode_kernel(const float *__restrict__ post, long ld_t, long ld_b, const double *__restrict__ ltrans, long ldw_b,
                 const int32_t *__restrict__ lengths, double log_slip, int T, int K, int log_mode,
                 int32_t *__restrict__ tb, int32_t *__restrict__ seq_out, double *__restrict__ score_out)
{
    extern __shared__ double ps[];                   // [2][K]
    __shared__ double red_v[8], s_max;
    __shared__ int red_i[8], s_arg;
    const int b = blockIdx.x;
    const int nev = lengths ? min(lengths[b], T) : T;
    if (nev < 1) {
        if (threadIdx.x == 0) score_out[b] = 0.0;
        return;
    }
    const float *pb = post + (long)b * ld_b;
    const double *wb = ltrans ? ltrans + (long)b * ldw_b : nullptr;
    int32_t *tbb = tb + (size_t)b * (size_t)T * (size_t)K;
    int32_t *seq = seq_out + (size_t)b * T;
    auto lp = [&](float v) -> double {               // olddecode.py:22-25, in the posteriors' own float32
        return log_mode ? (double)v : (double)logf(__fadd_rn(v, 1e-10f));
    };
    const int n4 = K / 4, n16 = K / 16;
    for (int j = threadIdx.x; j < K; j += blockDim.x) ps[j] = lp(pb[j]);
    __syncthreads();
    int cur = 0;
    for (int ev = 1; ev < nev; ev++) {
        const double *p = ps + cur * K;
        double *pn = ps + (cur ^ 1) * K;
        block_argmax(p, K, red_v, red_i, &s_max, &s_arg);
        const double w0 = wb ? wb[(long)(ev - 1) * 3 + 0] : 0.0;
        const double w1 = wb ? wb[(long)(ev - 1) * 3 + 1] : 0.0;     // the caller has subtracted log 4 / log 16 (:31-32)
        const double w2 = wb ? wb[(long)(ev - 1) * 3 + 2] : 0.0;
        const double slipv = s_max + log_slip;
        const int slipi = s_arg;
        const float *row = pb + (long)ev * ld_t;
        for (int j = threadIdx.x; j < K; j += blockDim.x) {
            double score = p[j] + w0;
            int idx = j;
            if (!(score > slipv)) idx = slipi;
            score = fmax(score, slipv);
            {
                const int c = j >> 2;
                double m = p[c];
                int am = 0;
#pragma unroll
                for (int a = 1; a < 4; a++) {
                    const double v = p[a * n4 + c];
                    if (v > m) { m = v; am = a; }
                }
                const double nw = m + w1;
                if (!(score > nw)) idx = n4 * am + c;
                score = fmax(score, nw);
            }
            {
                const int c = j >> 4;
                double m = p[c];
                int am = 0;
#pragma unroll
                for (int a = 1; a < 16; a++) {
                    const double v = p[a * n16 + c];
                    if (v > m) { m = v; am = a; }
                }
                const double nw = m + w2;
                if (!(score > nw)) idx = n16 * am + c;
                score = fmax(score, nw);
            }
            tbb[(size_t)(ev - 1) * K + j] = idx;
            pn[j] = score + lp(row[j]);
        }
        cur ^= 1;
        __syncthreads();
    }
    block_argmax(ps + cur * K, K, red_v, red_i, &s_max, &s_arg);
    if (threadIdx.x == 0) {
        score_out[b] = s_max;
        int state = s_arg;
        seq[nev - 1] = state;
        for (int ev = nev; ev > 1; ev--) {           // :69-71
            state = tbb[(size_t)(ev - 2) * K + state];
            seq[ev - 2] = state;
        }
    }
}

// estimate_transitions (:94-118), the sums of one event pair per CTA, in float64:
//   res[ev-1] = ( sum_j q[j] p[j],  sum_j q[j] S4[j mod K/4] / 4,  sum_j q[j] S16[j mod K/16] / 16 ),  q = post[ev-1],
//   p = post[ev], S4[c] = sum of p[4c .. 4c+3], S16[c] = sum of p[16c .. 16c+15]
__global__ void __launch_bounds__(256)
transitions_kernel(const float *__restrict__ post, long ld_t, int T, int K, double *__restrict__ res)
{
    extern __shared__ double sums[];                 // S4 [K/4] then S16 [K/16]
    __shared__ double red[3][8];
    const int ev = blockIdx.x + 1;
    const float *q = post + (long)(ev - 1) * ld_t, *p = post + (long)ev * ld_t;
    const int n4 = K / 4, n16 = K / 16;
    double *S4 = sums, *S16 = sums + n4;
    for (int c = threadIdx.x; c < n4; c += blockDim.x)
        S4[c] = ((double)p[4 * c] + (double)p[4 * c + 1]) + ((double)p[4 * c + 2] + (double)p[4 * c + 3]);
    __syncthreads();
    for (int c = threadIdx.x; c < n16; c += blockDim.x) S16[c] = (S4[4 * c] + S4[4 * c + 1]) + (S4[4 * c + 2] + S4[4 * c + 3]);
    __syncthreads();
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        const double qj = (double)q[j];
        a0 += qj * (double)p[j];
        a1 += qj * S4[j % n4];
        a2 += qj * S16[j % n16];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a0; red[1][threadIdx.x >> 5] = a1; red[2][threadIdx.x >> 5] = a2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { s0 += red[0][w]; s1 += red[1][w]; s2 += red[2][w]; }
        res[(long)(ev - 1) * 3 + 0] = s0;
        res[(long)(ev - 1) * 3 + 1] = s1 / 4.0;
        res[(long)(ev - 1) * 3 + 2] = s2 / 16.0;
    }
}

}  // namespace sloika

using namespace sloika;

extern "C" size_t sloika_olddecode_workspace_bytes(int T, int B, int K)
{
    if (T < 0 || B < 0 || K < 0) return 0;
    return sizeof(int32_t) * (size_t)T * (size_t)B * (size_t)K;
}

extern "C" int sloika_olddecode_fwd(const float *post, long ld_t, long ld_b, const double *ltrans, long ldw_b,
                                    const int32_t *lengths, int T, int B, int K, double slip, int log_mode, void *tb_ws,
                                    size_t ws_bytes, int32_t *seq_out, double *score_out, void *stream)
{
    if (!post || !seq_out || !score_out || T <= 0 || B <= 0 || K < 16 || (K & 15) != 0) return SLOIKA_ERR_ARG;
    if (K > 4096) return SLOIKA_ERR_UNSUPPORTED;
    if (!tb_ws || ws_bytes < sloika_olddecode_workspace_bytes(T, B, K)) return SLOIKA_ERR_WORKSPACE;
    const double log_slip = log(1e-10 + slip);                     // olddecode.py:34
    const size_t smem = 2 * (size_t)K * sizeof(double);
    cudaError_t err = cudaFuncSetAttribute(olddecode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    olddecode_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(post, ld_t, ld_b, ltrans, ldw_b, lengths, log_slip, T, K,
                                                            log_mode, (int32_t *)tb_ws, seq_out, score_out);
    SLOIKA_RETURN_LAUNCH_STATUS();
}

extern "C" int sloika_transitions_fwd(const float *post, long ld_t, int T, int K, double *res, void *stream)
{
    if (!post || !res || T <= 0 || K < 16 || (K & 15) != 0) return SLOIKA_ERR_ARG;
    if (T == 1) return SLOIKA_OK;
    const size_t smem = sizeof(double) * ((size_t)K / 4 + (size_t)K / 16);
    transitions_kernel<<<T - 1, 256, smem, (cudaStream_t)stream>>>(post, ld_t, T, K, res);
    SLOIKA_RETURN_LAUNCH_STATUS();
}
