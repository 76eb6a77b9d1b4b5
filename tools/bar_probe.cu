// Microbenchmark: cost of the per-step skeleton of a persistent recurrent kernel on one SM:
//   LDS -> ALU -> STS -> __syncthreads, twice per iteration, for 128..512 threads per CTA.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(float *out, int iters, int mode, long long *cyc)
{
    __shared__ float a[1024], b[1024];
    const int tid = threadIdx.x;
    a[tid] = tid; b[tid] = 0;
    __syncthreads();
    long long t0 = clock64();
    float v = 0.f;
    for (int i = 0; i < iters; i++) {
        if (mode >= 1) { v = a[(tid + 1) % blockDim.x] * 0.5f + v; b[tid] = v; }
        __syncthreads();
        if (mode >= 1) { v = b[(tid + 7) % blockDim.x] * 0.25f + v; a[tid] = v; }
        if (mode >= 2) { v = __fdividef(1.0f, 1.0f + __expf(-v)); }
        __syncthreads();
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + tid] = v;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 20000;
    for (int mode = 0; mode < 3; mode++)
        for (int nt : {128, 256, 384, 512}) {
            probe<<<148, nt>>>(out, iters, mode, cyc);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double s = 0; for (int i = 0; i < 148; i++) s += h[i];
            printf("mode %d (0 = 2 barriers only, 1 = + LDS/ALU/STS round trips, 2 = + sigmoid) threads %3d: %.0f clk per iteration\n", mode, nt, s / 148 / iters);
        }
    return 0;
}
