// Forward-only scoring of a network on labelled chunks: the reductions of bin/validate_network.py:46-54
// over the posteriors the forward pass leaves on the device --
//     loss     = mean over (t, b) of  -log post[t, b, label[t, b]]      (T.nnet.categorical_crossentropy per row, :49)
//     ncorrect = number of (t, b) with argmax_s post[t, b, s] == label[t, b]   (first maximum, :50)
// One warp per row; the sums are accumulated in float64 with one atomic per warp.
#include "common.cuh"

namespace sloika {

__global__ void __launch_bounds__(256)
score_kernel(const float *__restrict__ post, long ld, const int32_t *__restrict__ labels, long M, int S,
             double *__restrict__ loss_sum, unsigned long long *__restrict__ ncorrect)
{
    const int lane = threadIdx.x & 31;
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    double lsum = 0.0;
    unsigned long long nc = 0;
    for (long m = warp0; m < M; m += nwarps) {
        const float *row = post + m * ld;
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int s = lane; s < S; s += 32) {
            const float v = row[s];
            if (v > bv) { bv = v; bi = s; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            const int lbl = labels[m];
            if (lbl >= 0 && lbl < S) lsum -= (double)logf(row[lbl]);
            nc += (bi == lbl);
        }
    }
    if (lane == 0) {
        atomicAdd(loss_sum, lsum);
        atomicAdd(ncorrect, nc);
    }
}

}  // namespace sloika

extern "C" int sloika_score_fwd(const float *post, long ld, const int32_t *labels, long M, int S, double *loss_sum,
                                unsigned long long *ncorrect, void *stream)
{
    if (!post || !labels || !loss_sum || !ncorrect || M < 0 || S <= 0 || ld < S) return SLOIKA_ERR_ARG;
    if (M == 0) return SLOIKA_OK;
    long blocks = sloika::ceil_div(M, 8);
    if (blocks > 148L * 8) blocks = 148L * 8;
    sloika::score_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(post, ld, labels, M, S, loss_sum, ncorrect);
    SLOIKA_RETURN_LAUNCH_STATUS();
}
