// Best path -> base sequence on the device, batched, one CTA per read.
//
// Reference: sloika/bio.py:228-237 (kmers_to_sequence) = max_overlap (bio.py:160-178) + reduce_kmers
// (bio.py:208-225), called per read from SeqPrinter.write (sloika/basecall.py:141-149) on the k-mers of the
// Viterbi path.  On the integer states (k-mer = base-nbase digits, first letter most significant):
//   move(prev, next) = 0                      if prev == next and stays are allowed (always_move == 0)
//                    = smallest m in 1..k-1   with  prev mod nbase^(k-m) == next div nbase^m     (k1[m:] == k2[:-m])
//                    = k                      otherwise
//   sequence = letters(first k-mer) + for each move m > 0 the last min(m, k) letters of `next`.
// Every move is independent of the others, so a read is: moves in parallel -> block-wide exclusive scan of the
// move lengths -> letters written at their offsets.  Integer / byte work: results are identical to the reference's
// Python strings (tests/golden/bio_cases.json).
#include "common.cuh"

namespace sloika {
namespace bases {

constexpr int THREADS = 256;

__device__ __forceinline__ int move_of(int prev, int next, int klen, const int *pw, int always_move)
{
    if (!always_move && prev == next) return 0;
    for (int m = 1; m < klen; m++)
        if (prev % pw[klen - m] == next / pw[m]) return m;
    return klen;
}

__global__ void __launch_bounds__(THREADS)
path_to_bases_kernel(const int32_t *__restrict__ path, long ld_path, const int32_t *__restrict__ path_len, int klen,
                     int nbase, int always_move, const char *__restrict__ alphabet, char *__restrict__ out,
                     long ld_out, int32_t *__restrict__ out_len)
{
    __shared__ int pw[17];
    __shared__ char alpha[16];
    __shared__ int part[THREADS];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = path_len[b];
    if (tid <= klen) {
        int v = 1;
        for (int i = 0; i < tid; i++) v *= nbase;
        pw[tid] = v;
    }
    if (tid < nbase) alpha[tid] = alphabet[tid];
    __syncthreads();
    if (n <= 0) {
        if (tid == 0) out_len[b] = 0;
        return;
    }
    const int32_t *p = path + (long)b * ld_path;
    char *o = out + (long)b * ld_out;
    // positions 1..n-1 in contiguous blocks per thread
    const int per = (n - 1 + THREADS - 1) / THREADS;
    const int lo = 1 + tid * per, hi = min(lo + per, n);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += move_of(p[i - 1], p[i], klen, pw, always_move);
    part[tid] = sum;
    __syncthreads();
    // exclusive scan of the 256 partial sums (Hillis-Steele in shared memory)
    for (int d = 1; d < THREADS; d <<= 1) {
        const int v = tid >= d ? part[tid - d] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int off = klen + part[tid] - sum;                       // letters before this thread's first position
    if (tid == THREADS - 1) out_len[b] = klen + part[tid];
    if (tid == 0) {
        const int s = p[0];
        for (int i = 0; i < klen; i++) o[i] = alpha[(s / pw[klen - 1 - i]) % nbase];
    }
    for (int i = lo; i < hi; i++) {
        const int nxt = p[i];
        const int m = move_of(p[i - 1], nxt, klen, pw, always_move);
        for (int c = 0; c < m; c++) o[off + c] = alpha[(nxt / pw[m - 1 - c]) % nbase];   // last m letters of `next`
        off += m;
    }
}

}  // namespace bases
}  // namespace sloika

using namespace sloika;

extern "C" int sloika_path_to_bases_fwd(const int32_t *path, long ld_path, const int32_t *path_len, int B, int klen, int nbase,
                                        int always_move, const char *alphabet, char *out, long ld_out, int32_t *out_len,
                                        void *stream)
{
    if (!path || !path_len || !alphabet || !out || !out_len || B <= 0) return SLOIKA_ERR_ARG;
    if (klen < 1 || klen > 16 || nbase < 2 || nbase > 16) return SLOIKA_ERR_ARG;
    double states = 1.0;
    for (int i = 0; i < klen; i++) states *= nbase;
    if (states > 2147483647.0) return SLOIKA_ERR_ARG;
    bases::path_to_bases_kernel<<<B, bases::THREADS, 0, (cudaStream_t)stream>>>(path, ld_path, path_len, klen, nbase,
                                                                                always_move, alphabet, out, ld_out, out_len);
    SLOIKA_RETURN_LAUNCH_STATUS();
}
