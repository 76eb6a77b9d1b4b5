// Microbenchmark: how fast does one SM pull tiles into shared memory through (a) 2-D tensor TMA with
// 128-byte swizzle (box 32 fp32 x 128 rows of a [M, K] row-major matrix) and (b) 1-D bulk copies
// (cp.async.bulk) of contiguous bytes, as a function of the number of tiles in flight.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../sloika_b200/csrc -I../include -o tma_probe tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace sloika::tc;

__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// mode 0: tensor TMA box {32, rows}; mode 1: 1-D bulk copy of `tile_bytes`
__global__ void probe(const __grid_constant__ CUtensorMap tmap, const float *x, int mode, int stages, int tile_bytes,
                      int rows, int nkb, long tiles_total, long ld_bytes)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[16];
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    uint32_t it = 0;
    // issue `stages` loads, then wait/reissue in a ring
    long per_cta = tiles_total / gridDim.x;
    long t0 = (long)blockIdx.x * per_cta;
    for (long i = 0; i < per_cta + stages; i++) {
        const int s = it % stages;
        const uint32_t ph = (it / stages) & 1;
        if (i >= stages) mbar_wait(&bars[s], ph ^ 1);           // previous load of this slot landed
        if (i < per_cta) {
            mbar_arrive_expect_tx(&bars[s], tile_bytes);
            const long tile = t0 + i;
            if (mode == 0) {
                const long mt = tile / nkb; const int kb = (int)(tile % nkb);
                tma_load_2d(smem + (size_t)s * tile_bytes, &tmap, &bars[s], kb * 32, (int)(mt * rows));
            } else {
                bulk_load_1d(smem + (size_t)s * tile_bytes, reinterpret_cast<const uint8_t *>(x) + tile * (long)tile_bytes, tile_bytes, &bars[s]);
            }
        }
        it++;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const long M = 819200;
    float *x; cudaMalloc(&x, M * 128 * 4); cudaMemset(x, 0, M * 128 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int K : {96, 32, 128}) {
        for (int rows : {128, 64}) {
            CUtensorMap tmap;
            cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M}; cuuint64_t gstr[1] = {(cuuint64_t)K * 4};
            cuuint32_t box[2] = {32, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
            CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            const int nkb = K / 32; const int tile_bytes = rows * 128;
            const long tiles = (M / rows) * nkb;
            for (int stages : {2, 4, 8}) {
                probe<<<148, 32, stages * tile_bytes>>>(tmap, x, 0, stages, tile_bytes, rows, nkb, tiles, K * 4);
                cudaEventRecord(e0);
                probe<<<148, 32, stages * tile_bytes>>>(tmap, x, 0, stages, tile_bytes, rows, nkb, tiles, K * 4);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                printf("tensor TMA K=%3d box 32x%3d stages %d: %.3f ms  %.0f GB/s  (%s)\n", K, rows, stages, ms,
                       (double)(tiles / 148 * 148) * tile_bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    CUtensorMap dummy; memset(&dummy, 0, sizeof(dummy));
    for (int tile_bytes : {16384, 49152}) {
        const long tiles = M * 96 * 4 / tile_bytes;
        for (int stages : {2, 4}) {
            if ((long)stages * tile_bytes > 200 * 1024) continue;
            probe<<<148, 32, stages * tile_bytes>>>(dummy, x, 1, stages, tile_bytes, 0, 1, tiles, 0);
            cudaEventRecord(e0);
            probe<<<148, 32, stages * tile_bytes>>>(dummy, x, 1, stages, tile_bytes, 0, 1, tiles, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("1-D bulk %5d B stages %d: %.3f ms  %.0f GB/s  (%s)\n", tile_bytes, stages, ms,
                   (double)(tiles / 148 * 148) * tile_bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
