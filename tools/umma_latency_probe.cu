// Microbenchmark for the tensor-core GRU recurrence: latency of one dependent "phase"
//   128 threads write a small B operand (h) into swizzled smem -> fence.proxy.async -> mbarrier
//   -> one thread issues NMMA tcgen05.mma (M=128, N=NN, K=8, tf32) + commit -> 128 threads wait,
//   tcgen05.ld NN columns -> next phase depends on the loaded values.
// Reports cycles per phase as a function of NMMA and NN.
#include <cstdio>
#include "tc_common.cuh"
using namespace sloika::tc;

__global__ void __launch_bounds__(160, 1) probe(float *out, int iters, int nmma, int NN, long long *cycles)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *A = smem;                       // 128 rows x 32 k (one K block), reused for every MMA
    uint8_t *Bt = smem + 16384;              // NN rows x 32 k
    __shared__ uint64_t bar_b, bar_d;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bar_b, 128); mbar_init(&bar_d, 1); mbar_fence_init(); }
    if (warp == 4) tmem_alloc(&tmem_base_s, 64);
    for (int e = tid; e < 128 * 32; e += 160) *reinterpret_cast<float *>(A + sw128_offset(e / 32, e % 32)) = 1.0f / 64;
    for (int e = tid; e < NN * 32; e += 160) *reinterpret_cast<float *>(Bt + sw128_offset(e / 32, e % 32)) = 0.0f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = umma_idesc_tf32_m128(NN);
    long long t0 = clock64();
    if (warp == 4) {
        if (lane == 0) {
            for (int it = 0; it < iters; it++) {
                mbar_wait(&bar_b, it & 1);
                tc_fence_after();
                const uint32_t a0 = smem_u32(A), b0 = smem_u32(Bt);
#pragma unroll 4
                for (int m = 0; m < nmma; m++) {
                    const int ks = m & 3;
                    umma_tf32_ss(tmem_base + ((m >> 2) % 3) * 16, umma_desc_sw128_kmajor(a0 + ks * 32),
                                 umma_desc_sw128_kmajor(b0 + ks * 32), idesc, m >= 12);
                }
                umma_commit(&bar_d);
            }
        }
    } else {
        float h = 0.001f * tid;
        for (int it = 0; it < iters; it++) {
            // write this thread's value into the B operand (row = tid % NN, k = tid / NN ...)
            *reinterpret_cast<float *>(Bt + sw128_offset(tid % NN, (tid / NN) % 32)) = h;
            fence_proxy_async();
            mbar_arrive(&bar_b);
            mbar_wait(&bar_d, it & 1);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16), v);
            tmem_ld_wait();
            tc_fence_before();
            h = __uint_as_float(v[0]) * 0.5f + __uint_as_float(v[7]) * 0.25f + 0.001f;
        }
        out[blockIdx.x * 128 + tid] = h;
    }
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

int main() {
    float *out; long long *cyc; cudaMalloc(&out, 148 * 128 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    for (int NN : {16, 32}) {
        for (int nmma : {0, 12, 36, 72, 108, 216}) {
            probe<<<148, 160, 16384 + 4096 + 1024>>>(out, iters, nmma, NN, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double s = 0; for (int i = 0; i < 148; i++) s += h[i];
            printf("N=%2d nmma=%3d: %.0f cycles per phase (%s)\n", NN, nmma, s / 148 / iters, cudaGetErrorString(e));
        }
    }
    return 0;
}
