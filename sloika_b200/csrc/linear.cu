// y = act(x . W' + bias): FeedForward.run (sloika/layers.py:157-158), the GRU input projection
// vI = x iW' + b (layers.py:1011) for all time steps at once, and the logits of Softmax.run
// (layers.py:310) followed by the row softmax (layers.py:311-314).
//
// Round-1 implementation: fp32 SIMT GEMM (128x64 tile, 8x4 per thread) with fused bias/activation
// epilogue.  M = T*B rows is huge (819 200), K and N are small (<= 256 / <= 1025), so both operands'
// K extent is streamed in 16-wide slabs through shared memory.  This is the kernel the tcgen05
// (3xTF32) projection replaces; it stays as the exact-fp32 fallback for odd shapes.
#include <cstring>
#include "common.cuh"

namespace sloika {

constexpr int BM = 128, BN = 64, BK = 16, LIN_THREADS = 256;

__global__ void __launch_bounds__(LIN_THREADS, 2)
linear_kernel(const float *__restrict__ x, long ldx, const float *__restrict__ W, const float *__restrict__ bias,
              float *__restrict__ y, long ldy, long M, int K, int N, int act, int vec_in, int vec_out)
{
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const long m0 = (long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads: n = tx*4.., m = ty*8..
    const int lrow = tid >> 2, lkq = tid & 3;         // loader mapping: row, k-quad

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        const int k = k0 + lkq * 4;
        // A slab: 128 rows x 16 k (two rows per thread)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int r = lrow + 64 * h;
            const long m = m0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < M) {
                const float *p = x + m * ldx + k;
                if (vec_in && k + 3 < K) {
                    v = __ldg(reinterpret_cast<const float4 *>(p));
                } else {
                    if (k + 0 < K) v.x = __ldg(p + 0);
                    if (k + 1 < K) v.y = __ldg(p + 1);
                    if (k + 2 < K) v.z = __ldg(p + 2);
                    if (k + 3 < K) v.w = __ldg(p + 3);
                }
            }
            As[lkq * 4 + 0][r] = v.x; As[lkq * 4 + 1][r] = v.y;
            As[lkq * 4 + 2][r] = v.z; As[lkq * 4 + 3][r] = v.w;
        }
        {   // B slab: 64 rows of W x 16 k
            const int n = n0 + lrow;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < N) {
                const float *p = W + (long)n * K + k;
                if ((K & 3) == 0 && k + 3 < K) {
                    v = __ldg(reinterpret_cast<const float4 *>(p));
                } else {
                    if (k + 0 < K) v.x = __ldg(p + 0);
                    if (k + 1 < K) v.y = __ldg(p + 1);
                    if (k + 2 < K) v.z = __ldg(p + 2);
                    if (k + 3 < K) v.w = __ldg(p + 3);
                }
            }
            Bs[lkq * 4 + 0][lrow] = v.x; Bs[lkq * 4 + 1][lrow] = v.y;
            Bs[lkq * 4 + 2][lrow] = v.z; Bs[lkq * 4 + 3][lrow] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int i = 0; i < 8; i++) {
                acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
                acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
                acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
                acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
            }
        }
        __syncthreads();
    }

    const int n = n0 + tx * 4;
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (n + j < N) bv[j] = __ldg(bias + n + j);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const long m = m0 + ty * 8 + i;
        if (m >= M) break;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) o[j] = apply_act(acc[i][j] + bv[j], act);
        float *p = y + m * ldy + n;
        if (vec_out && n + 3 < N) {
            *reinterpret_cast<float4 *>(p) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (n + j < N) p[j] = o[j];
        }
    }
}

// Row softmax in place, one warp per row: m = max; e = exp(t - m); e / sum(e)  (layers.py:311-314).
__global__ void __launch_bounds__(256)
softmax_rows_kernel(float *__restrict__ post, long ldp, long M, int N)
{
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    float *p = post + row * ldp;
    float m = -INFINITY;
    for (int j = lane; j < N; j += 32) m = fmaxf(m, p[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.0f;
    for (int j = lane; j < N; j += 32) {
        const float e = expf(p[j] - m);
        p[j] = e;
        s += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    for (int j = lane; j < N; j += 32) p[j] = p[j] / s;
}

// post = exp(t - M) / S in place, from the per-slice (max, sum exp) pairs written by the GEMM epilogue:
// M = max_s m_s, S = sum_s s_s * exp(m_s - M)  (layers.py:311-314 with the row reductions pre-computed).
// One warp per row; rows are 16-byte aligned (ldp % 4 == 0) so the bulk moves as float4.
__global__ void __launch_bounds__(256)
softmax_normalise_kernel(float *__restrict__ post, long ldp, const float2 *__restrict__ stats, int n_slices, long M, int N)
{
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    float m = -INFINITY, s = 0.0f;
    if (lane < n_slices) {
        const float2 st = __ldg(stats + row * n_slices + lane);
        m = st.x;
        s = st.y;
    }
    float mx = m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float tot = lane < n_slices ? s * expf(m - mx) : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    float *p = post + row * ldp;
    const int n4 = N >> 2;
    float4 *p4 = reinterpret_cast<float4 *>(p);
    for (int c = lane; c < n4; c += 32) {
        float4 v = p4[c];
        v.x = expf(v.x - mx) / tot; v.y = expf(v.y - mx) / tot;
        v.z = expf(v.z - mx) / tot; v.w = expf(v.w - mx) / tot;
        p4[c] = v;
    }
    for (int j = 4 * n4 + lane; j < N; j += 32) p[j] = expf(p[j] - mx) / tot;
}

static int launch_linear(const float *x, long ldx, const float *W, const float *bias, float *y, long ldy, long M,
                         int K, int N, int act, cudaStream_t st)
{
    const int vec_in = ((ldx & 3) == 0) && ((K & 3) == 0) && (((uintptr_t)x & 15) == 0);
    const int vec_out = ((ldy & 3) == 0) && (((uintptr_t)y & 15) == 0);
    const long mt = ceil_div(M, BM);
    if (mt > 0x7fffffffL) return SLOIKA_ERR_ARG;
    dim3 grid((unsigned)mt, (unsigned)ceil_div(N, BN));
    linear_kernel<<<grid, LIN_THREADS, 0, st>>>(x, ldx, W, bias, y, ldy, M, K, N, act, vec_in, vec_out);
    SLOIKA_RETURN_LAUNCH_STATUS();
}

namespace gemm_tc {
int plan_slices(int K, int N, int *bn_out, bool f16);
int launch(const float *x, long ldx, const float *W, const float *bias, float *y, long ldy, long M, int K, int N,
           int act, float2 *stats, int rot, bool f16, cudaStream_t st, const unsigned *gate = nullptr,
           unsigned gate_limit = 0, int gate_mode = 0);
}

// algo: SLOIKA_GEMM_AUTO tries the tcgen05 3xTF32 kernel and falls back to the fp32 SIMT kernel when the
// shape or alignment rules it out; SLOIKA_GEMM_SIMT / SLOIKA_GEMM_TC force one of them.
static int dispatch_linear(const float *x, long ldx, const float *W, const float *bias, float *y, long ldy, long M,
                           int K, int N, int act, int algo, cudaStream_t st)
{
    if (algo != SLOIKA_GEMM_SIMT) {
        const int rc = gemm_tc::launch(x, ldx, W, bias, y, ldy, M, K, N, act, nullptr, 0, algo == SLOIKA_GEMM_TC_F16, st);
        if (rc != SLOIKA_ERR_UNSUPPORTED || algo == SLOIKA_GEMM_TC || algo == SLOIKA_GEMM_TC_F16) return rc;
    }
    return launch_linear(x, ldx, W, bias, y, ldy, M, K, N, act, st);
}

}  // namespace sloika

using namespace sloika;

extern "C" int sloika_linear_fwd_ex(const float *x, long ldx, const float *W, const float *bias, float *y, long ldy,
                                    long M, int K, int N, int act, int algo, void *stream)
{
    if (!x || !W || !y || M < 0 || K <= 0 || N <= 0 || ldx < K || ldy < N) return SLOIKA_ERR_ARG;
    if (algo < SLOIKA_GEMM_AUTO || algo > SLOIKA_GEMM_TC_F16) return SLOIKA_ERR_ARG;
    if (!act_known(act)) return SLOIKA_ERR_UNSUPPORTED;
    if (M == 0) return SLOIKA_OK;
    return dispatch_linear(x, ldx, W, bias, y, ldy, M, K, N, act, algo, (cudaStream_t)stream);
}

extern "C" int sloika_linear_fwd(const float *x, long ldx, const float *W, const float *bias, float *y, long ldy,
                                 long M, int K, int N, int act, void *stream)
{
    if (!x || !W || !y || M < 0 || K <= 0 || N <= 0 || ldx < K || ldy < N) return SLOIKA_ERR_ARG;
    if (!act_known(act)) return SLOIKA_ERR_UNSUPPORTED;
    if (M == 0) return SLOIKA_OK;
    return dispatch_linear(x, ldx, W, bias, y, ldy, M, K, N, act, SLOIKA_GEMM_AUTO, (cudaStream_t)stream);
}

extern "C" int sloika_linear_fwd_gated(const float *x, long ldx, const float *W, const float *bias, float *y, long ldy,
                                       long M, int K, int N, int act, const float *absmax, float limit, void *stream)
{
    if (!x || !W || !y || !absmax || M < 0 || K <= 0 || N <= 0 || ldx < K || ldy < N || !(limit > 0.0f)) return SLOIKA_ERR_ARG;
    if (!act_known(act)) return SLOIKA_ERR_UNSUPPORTED;
    if (M == 0) return SLOIKA_OK;
    unsigned limit_bits;
    memcpy(&limit_bits, &limit, sizeof(limit_bits));
    const unsigned *gate = reinterpret_cast<const unsigned *>(absmax);
    cudaStream_t st = (cudaStream_t)stream;
    // both forms are enqueued; the CTAs of the one the gate rules out return at once (a few microseconds)
    int rc = gemm_tc::launch(x, ldx, W, bias, y, ldy, M, K, N, act, nullptr, 0, true, st, gate, limit_bits, 1);
    if (rc != SLOIKA_OK) return rc;             // shape / alignment the tensor-core kernel cannot take: caller's problem
    return gemm_tc::launch(x, ldx, W, bias, y, ldy, M, K, N, act, nullptr, 0, false, st, gate, limit_bits, 2);
}

extern "C" int sloika_softmax_slices(int K, int N, int algo)
{
    if (N > 32 * 256) return 0;
    // the kernel writes one (max, sum exp) pair per column slice AND per epilogue warp group
    const int n = 2 * gemm_tc::plan_slices(K, N, nullptr, algo == SLOIKA_GEMM_TC_F16);
    return n > 32 ? 0 : n;
}

extern "C" int sloika_softmax_logits_fwd(const float *x, long ldx, const float *W, const float *bias, float *logits,
                                         long ldl, float *stats, long M, int K, int N, int stay_last, int algo, void *stream)
{
    if (!x || !W || !logits || !stats || M < 0 || K <= 0 || N <= 0 || ldx < K || ldl < N) return SLOIKA_ERR_ARG;
    if (M == 0) return SLOIKA_OK;
    if (sloika_softmax_slices(K, N, algo) <= 0) return SLOIKA_ERR_UNSUPPORTED;
    return gemm_tc::launch(x, ldx, W, bias, logits, ldl, M, K, N, SLOIKA_ACT_LINEAR, reinterpret_cast<float2 *>(stats),
                           stay_last ? 1 : 0, algo == SLOIKA_GEMM_TC_F16, (cudaStream_t)stream);
}

// the same with x in the blocked layout of sloika_gru_seq_fwd (M a multiple of 128, fp16-split form): the last GRU layer's
// output goes into the logits GEMM as it is
extern "C" int sloika_softmax_logits_blocked_fwd(const float *xb, const float *W, const float *bias, float *logits, long ldl,
                                                 float *stats, long M, int K, int N, int stay_last, void *stream)
{
    if (!xb || !W || !logits || !stats || M < 0 || K <= 0 || N <= 0 || ldl < N) return SLOIKA_ERR_ARG;
    if (M == 0) return SLOIKA_OK;
    if (sloika_softmax_slices(K, N, SLOIKA_GEMM_TC_F16) <= 0) return SLOIKA_ERR_UNSUPPORTED;
    return gemm_tc::launch(xb, -1, W, bias, logits, ldl, M, K, N, SLOIKA_ACT_LINEAR, reinterpret_cast<float2 *>(stats),
                           stay_last ? 1 : 0, true, (cudaStream_t)stream);
}

extern "C" int sloika_softmax_normalise_fwd(float *logits, long ldl, const float *stats, int n_slices, long M, int N,
                                            void *stream)
{
    if (!logits || !stats || M < 0 || N <= 0 || ldl < N || n_slices <= 0 || n_slices > 32) return SLOIKA_ERR_ARG;
    if ((ldl & 3) != 0 || ((uintptr_t)logits & 15) != 0) return SLOIKA_ERR_UNSUPPORTED;
    if (M == 0) return SLOIKA_OK;
    const int warps = 8;
    const long blocks = ceil_div(M, warps);
    if (blocks > 0x7fffffffL) return SLOIKA_ERR_ARG;
    softmax_normalise_kernel<<<(unsigned)blocks, warps * 32, 0, (cudaStream_t)stream>>>(
        logits, ldl, reinterpret_cast<const float2 *>(stats), n_slices, M, N);
    SLOIKA_RETURN_LAUNCH_STATUS();
}

extern "C" int sloika_softmax_fwd(const float *x, long ldx, const float *W, const float *bias, float *post,
                                  long ldp, long M, int K, int N, void *stream)
{
    if (!x || !W || !post || M < 0 || K <= 0 || N <= 0 || ldx < K || ldp < N) return SLOIKA_ERR_ARG;
    if (M == 0) return SLOIKA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = dispatch_linear(x, ldx, W, bias, post, ldp, M, K, N, SLOIKA_ACT_LINEAR, SLOIKA_GEMM_AUTO, st);
    if (rc != SLOIKA_OK) return rc;
    const int warps = 8;
    const long blocks = ceil_div(M, warps);
    if (blocks > 0x7fffffffL) return SLOIKA_ERR_ARG;
    softmax_rows_kernel<<<(unsigned)blocks, warps * 32, 0, st>>>(post, ldp, M, N);
    SLOIKA_RETURN_LAUNCH_STATUS();
}
