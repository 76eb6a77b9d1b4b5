// Second tcgen05 probe for the GRU recurrence ("What's weak" 5 / next-round 3b): the dependent phase
//   128 threads write the B operand (h: NN sequences x 16*KC k, fp16) into SWIZZLE_64B smem -> fence.proxy.async
//   -> mbarrier -> ONE thread issues nmma tcgen05.mma kind::f16 (M = 128, N = NN, K = 16) + one commit
//   -> 128 threads wait, tcgen05.ld NN columns -> the next phase's h depends on the loaded values
// with the weight operand A either in shared memory (SS form, as in r1's tf32 probe) or RESIDENT IN TMEM (TS form,
// `tcgen05.mma [d], [a_tmem], b_desc, ...`), which removes the 4 KB shared-memory read per instruction.
// Prints cycles per phase for nmma = 0 .. 108 and N = 8 .. 64; the slope is the issue cost per MMA.
#include <cstdio>
#include <cuda_fp16.h>
#include "tc_common.cuh"
using namespace sloika::tc;

constexpr int KB = 3;                        // K blocks of 32 halves (H = 96)
constexpr int A_COLS = KB * 16;              // TMEM columns of one 128 x 96 fp16 A tile
constexpr int NTILES = 6;                    // resident A tiles (hi / lo of 3 M tiles): 288 columns
constexpr int D_COL = 288;                   // accumulators above the weights (3 x 64 columns; N = 128 overlaps them, timing only)

// ts: A from TMEM; otherwise A from shared memory (SWIZZLE_64B tiles).  The MMA sequence is fully unrolled with
// compile-time operand offsets: a first version that computed tile / k indices per instruction in the single issuing
// thread measured that thread's ALU latency (125 cycles per MMA), not the tensor pipe.
template <int NMMA, int TS>
__global__ void __launch_bounds__(160, 1) probe(float *out, int iters, int NN, long long *cycles)
{
    constexpr int nmma = NMMA, ts = TS;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *Bt = smem;                              // [KB][128 rows x 64 B] (NN <= 128)
    uint8_t *A = smem + KB * 8192;                   // SS form: [NTILES][KB][128 rows x 64 B]
    __shared__ uint64_t bar_b, bar_d;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bar_b, 128); mbar_init(&bar_d, 1); mbar_fence_init(); }
    if (warp == 4) tmem_alloc(&tmem_base_s, 512);
    for (int e = tid; e < KB * 8192 / 4; e += 160) reinterpret_cast<uint32_t *>(Bt)[e] = 0u;
    if (!ts)
        for (int e = tid; e < NTILES * KB * 8192 / 4; e += 160) reinterpret_cast<uint32_t *>(A)[e] = 0x1c001c00u;   // 2^-8
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (ts && warp < 4) {                            // weights -> TMEM: lane = row, 8 columns per K = 16 step
        uint32_t v[8];
        for (int i = 0; i < 8; i++) v[i] = 0x1c001c00u;
        for (int c = 0; c < NTILES * A_COLS; c += 8) tmem_st_32x32b_x8(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16_m128(NN);
    long long t0 = clock64();
    if (warp == 4) {
        // the whole warp runs the loop and one ELECTED lane issues (warp-uniform operands stay in uniform registers;
        // with `if (lane == 0)` around the loop every MMA was wrapped in an R2UR / ELECT / BRA.U.ANY sequence)
        for (int it = 0; it < iters; it++) {
            mbar_wait(&bar_b, it & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = smem_u32(A), b0 = smem_u32(Bt);
#pragma unroll
                for (int m = 0; m < nmma; m++) {
                    const int tile = (m / (2 * KB)) % NTILES, kk = m % (2 * KB);      // 6 K = 16 steps per tile
                    const uint32_t d = tmem_base + D_COL + (tile % 3) * 64;
                    const uint64_t bd = umma_desc_sw64_kmajor(b0 + (kk >> 1) * 8192 + (kk & 1) * 32);
                    if (ts) umma_f16_ts(d, tmem_base + tile * A_COLS + kk * 8, bd, idesc, m >= 6 * KB);
                    else umma_f16_ss(d, umma_desc_sw64_kmajor(a0 + (tile * KB + (kk >> 1)) * 8192 + (kk & 1) * 32), bd, idesc,
                                     m >= 6 * KB);
                }
                umma_commit(&bar_d);
            }
            __syncwarp();
        }
    } else {
        float h = 0.001f * tid;
        for (int it = 0; it < iters; it++) {
            *reinterpret_cast<__half *>(Bt + sw64_offset(tid % NN, (tid / NN) % 32)) = __float2half_rn(h);
            fence_proxy_async();
            mbar_arrive(&bar_b);
            mbar_wait(&bar_d, it & 1);
            tc_fence_after();
            uint32_t v[8];
            tmem_ld_32x32b_x8(tmem_base + ((uint32_t)(warp * 32) << 16) + D_COL, v);
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
    if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
    float *out; long long *cyc; cudaMalloc(&out, 148 * 128 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    const int smem = KB * 8192 + NTILES * KB * 8192 + 1024;
    for (int ts : {1, 0}) {
        for (int NN : {8, 16, 32, 64}) {
            for (int nmma : {0, 18, 36, 54, 108}) {
                auto run = [&](auto kern) {
                    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    kern<<<148, 160, smem>>>(out, iters, NN, cyc);
                };
                if (ts) {
                    if (nmma == 0) run(probe<0, 1>); else if (nmma == 18) run(probe<18, 1>); else if (nmma == 36) run(probe<36, 1>);
                    else if (nmma == 54) run(probe<54, 1>); else run(probe<108, 1>);
                } else {
                    if (nmma == 0) run(probe<0, 0>); else if (nmma == 18) run(probe<18, 0>); else if (nmma == 36) run(probe<36, 0>);
                    else if (nmma == 54) run(probe<54, 0>); else run(probe<108, 0>);
                }
                cudaError_t e = cudaDeviceSynchronize();
                long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
                double s = 0; for (int i = 0; i < 148; i++) s += h[i];
                printf("kind::f16 A in %s N=%2d nmma=%3d: %.0f cycles per phase (%s)\n", ts ? "TMEM" : "smem", NN, nmma,
                       s / 148 / iters, cudaGetErrorString(e));
                if (e != cudaSuccess) return 1;
            }
        }
    }
    return 0;
}
