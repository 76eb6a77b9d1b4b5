// Microbenchmark: legacy mma.sync throughput on sm_100a (TF32 m16n8k8, BF16 m16n8k16) and FFMA2,
// per SM, to size the tensor-core GRU recurrence.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_tf32(float *out, int iters) {
    float d[4][4] = {};
    unsigned a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f800000u};
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0; for (int j = 0; j < 4; j++) for (int c = 0; c < 4; c++) s += d[j][c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_bf16(float *out, int iters) {
    float d[4][4] = {};
    unsigned a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f803f80u};
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0; for (int j = 0; j < 4; j++) for (int c = 0; c < 4; c++) s += d[j][c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float *out, int iters) {
    float2 d[8]; for (int j = 0; j < 8; j++) d[j] = make_float2(0.f, 0.f);
    float2 a = make_float2(1.0001f, 0.9999f), b = make_float2(threadIdx.x * 1e-9f, 1e-9f);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) d[j] = __ffma2_rn(a, d[j], b);
    }
    float s = 0; for (int j = 0; j < 8; j++) s += d[j].x + d[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// dependent chain latency of one mma.sync
__global__ void k_tf32_lat(float *out, int iters) {
    float d[4] = {};
    unsigned a[4] = {0x3f800000u, 0, 0, 0}, b[2] = {0x3f800000u, 0};
    for (int i = 0; i < iters; i++)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    out[blockIdx.x * blockDim.x + threadIdx.x] = d[0] + d[1] + d[2] + d[3];
}

template <typename F> float run(F f, int warps_per_sm, int iters, float *out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f<<<148, warps_per_sm * 32>>>(out, 10); cudaDeviceSynchronize();
    cudaEventRecord(e0); f<<<148, warps_per_sm * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 1024 * 4);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    for (int w : {4, 8, 12, 16}) {
        float ms = run(k_tf32, w, iters, out);
        double mma = (double)iters * 4 * w;            // per SM
        printf("tf32 m16n8k8   warps/SM %2d: %.3f ms -> %.2f clk per MMA per SM (at %d kHz), %.0f FMA/clk/SM\n", w, ms, ms * 1e-3 * clk * 1e3 / mma, clk, mma * 1024 / (ms * 1e-3 * clk * 1e3));
        ms = run(k_bf16, w, iters, out);
        printf("bf16 m16n8k16  warps/SM %2d: %.3f ms -> %.2f clk per MMA per SM, %.0f FMA/clk/SM\n", w, ms, ms * 1e-3 * clk * 1e3 / mma, mma * 2048 / (ms * 1e-3 * clk * 1e3));
        ms = run(k_ffma2, w, iters, out);
        double ff = (double)iters * 8 * w;
        printf("ffma2          warps/SM %2d: %.3f ms -> %.2f clk per warp-FFMA2 per SM, %.0f FMA/clk/SM\n", w, ms, ms * 1e-3 * clk * 1e3 / ff, ff * 64 / (ms * 1e-3 * clk * 1e3));
    }
    float ms = run(k_tf32_lat, 1, iters, out);
    printf("tf32 mma.sync dependent-chain latency: %.1f clk\n", ms * 1e-3 * clk * 1e3 / iters);
    cudaError_t e = cudaDeviceSynchronize(); printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
