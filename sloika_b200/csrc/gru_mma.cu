// GRU recurrence on the tensor cores (warp-level mma.sync, 3xTF32 split = fp32-equivalent accuracy).
//
// Same semantics and CTA decomposition as gru.cu (one persistent CTA per 8 sequences, all of
// sW | sW2 on chip, two CTA barriers per step; reference sloika/layers.py:1010-1021, :85-88, :1449-1450).
// Why a third kernel: measurements on B200 (tools/*_probe.cu, profiles/) showed
//   * the FFMA2 kernels are bounded by register-delivery of h from shared memory (every lane must
//     receive all 8 x H state values per phase: 128 B/clk/SM) plus ~1.4 us/step of shuffle / barrier /
//     epilogue latency -> 2.7 us/step;
//   * tcgen05.mma costs >= ~56 cycles per instruction however small N is, so 8 sequences per CTA cannot
//     feed it (108 MMAs/step -> 6k cycles);
//   * legacy mma.sync.m16n8k8.tf32 sustains one MMA per 2.16 cycles per SM with 23 cycles latency, and
//     its operand fragments cut the shared-memory traffic of a step to two 4-byte loads per k-chunk.
//
// Layout: rows of a gate are cut in 16-row tiles (jt); warp (jt, role) holds the A fragments (weights)
// of its tiles in REGISTERS for the whole scan:
//     role 0 ("owner")  z tile of sW in phase 1, c tile of sW2 in phase 2, keeps z and h in registers,
//                       does the blend and publishes h_t (fp32 + tf32 hi/lo split) to shared memory
//     role 1            r tile of sW in phase 1 (writes r*h split hi/lo); in phase 2 it computes the second
//                       half of the k range of the c tile (partial sums handed to the owner through shared
//                       memory and a 64-thread named barrier) and then does the HBM traffic: cp.async
//                       staging of vI[t+2] and the coalesced store of h_{t-1}
// D[16 x 8] tiles: N = 8 sequences.  Every product is hi*hi + lo*hi + hi*lo with hi = top 19 bits.
#include <cstdlib>
#include "common.cuh"

namespace sloika {
namespace gru3 {


constexpr int BT = 8;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
// named barrier 1: non-blocking arrive (producers) / blocking sync (consumers)
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void split_tf32(float w, uint32_t &hi, uint32_t &lo) {
    hi = __float_as_uint(w) & 0xffffe000u;
    lo = __float_as_uint(w - __uint_as_float(hi));
}

// D (16x8, fp32) += A (16x8, tf32, row) * B (8x8, tf32, col)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// B fragments (the 8 sequences' state, tf32 hi / lo parts in [8][P] K-major arrays) of all k chunks, loaded up
// front: with the loads batched ahead of the MMA chain the tensor pipe is never left waiting on a 30-cycle
// shared-memory round trip per chunk (which is what bounded the first version of this kernel).
template <int NKC, int P>
__device__ __forceinline__ void load_bfrags(const float *__restrict__ Bf, int g, int t4, float (&bf)[NKC][2])
{
#pragma unroll
    for (int kc = 0; kc < NKC; kc++) {
        const int o = g * P + 8 * kc + t4;
        bf[kc][0] = Bf[o];
        bf[kc][1] = Bf[o + 4];
    }
}

// One 16-row tile (A fragments in registers, split on the fly) times the 8 sequences.
template <int NKC, int P>
__device__ __forceinline__ void tile_matvec(const float (&w)[NKC][4], const float *__restrict__ Bf, int g, int t4,
                                            float (&out)[4])
{
    float bf[NKC][2];
    load_bfrags<NKC, P>(Bf, g, t4, bf);
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kc = 0; kc < NKC; kc++) {
        uint32_t ah[4], al[4], bh0, bl0, bh1, bl1;
#pragma unroll
        for (int i = 0; i < 4; i++) split_tf32(w[kc][i], ah[i], al[i]);
        split_tf32(bf[kc][0], bh0, bl0);
        split_tf32(bf[kc][1], bh1, bl1);
        mma_tf32(acc0, ah, bh0, bh1);
        mma_tf32(acc1, al, bh0, bh1);
        mma_tf32(acc2, ah, bl0, bl1);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = acc0[i] + (acc1[i] + acc2[i]);
}

// Phase-1 variant for layers too wide for register-resident weights (H > 96): the A fragments stay in shared
// memory as fp32 in fragment order (one 128-bit load per lane and chunk, fetched one chunk ahead) and are split
// into tf32 hi / lo on the fly like the B words.
template <int NKC, int P, bool ACC3>
__device__ __forceinline__ void tile_matvec_smemA32(const float4 *__restrict__ Af, int lane, const float *__restrict__ Bf,
                                                    int g, int t4, float (&out)[4])
{
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
    float4 a4 = Af[lane];
    float b0 = Bf[g * P + t4], b1 = Bf[g * P + t4 + 4];
#pragma unroll
    for (int kc = 0; kc < NKC; kc++) {
        const float w[4] = {a4.x, a4.y, a4.z, a4.w};
        const float c0 = b0, c1 = b1;
        if (kc + 1 < NKC) {
            a4 = Af[(kc + 1) * 32 + lane];
            b0 = Bf[g * P + 8 * (kc + 1) + t4];
            b1 = Bf[g * P + 8 * (kc + 1) + t4 + 4];
        }
        uint32_t ah[4], al[4], bh0, bl0, bh1, bl1;
#pragma unroll
        for (int i = 0; i < 4; i++) split_tf32(w[i], ah[i], al[i]);
        split_tf32(c0, bh0, bl0);
        split_tf32(c1, bh1, bl1);
        mma_tf32(acc0, ah, bh0, bh1);
        mma_tf32(acc1, al, bh0, bh1);
        if constexpr (ACC3) mma_tf32(acc2, ah, bl0, bl1); else mma_tf32(acc1, ah, bl0, bl1);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = acc0[i] + (acc1[i] + acc2[i]);
}

// Register-A variant over a sub-range of the k chunks (phase 2 of the wide kernel: half the range per warp).
template <int KC0, int N, int P, bool ACC3>
__device__ __forceinline__ void tile_matvec_range(const float (&w)[N][4], const float *__restrict__ Bf, int g, int t4,
                                                  float (&out)[4])
{
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
    float b0 = Bf[g * P + 8 * KC0 + t4], b1 = Bf[g * P + 8 * KC0 + t4 + 4];
#pragma unroll
    for (int c = 0; c < N; c++) {
        const float c0 = b0, c1 = b1;
        if (c + 1 < N) {
            b0 = Bf[g * P + 8 * (KC0 + c + 1) + t4];
            b1 = Bf[g * P + 8 * (KC0 + c + 1) + t4 + 4];
        }
        uint32_t ah[4], al[4], bh0, bl0, bh1, bl1;
#pragma unroll
        for (int i = 0; i < 4; i++) split_tf32(w[c][i], ah[i], al[i]);
        split_tf32(c0, bh0, bl0);
        split_tf32(c1, bh1, bl1);
        mma_tf32(acc0, ah, bh0, bh1);
        mma_tf32(acc1, al, bh0, bh1);
        if constexpr (ACC3) mma_tf32(acc2, ah, bl0, bl1); else mma_tf32(acc1, ah, bl0, bl1);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = acc0[i] + (acc1[i] + acc2[i]);
}

// Same, with the A fragments pre-split into tf32 hi / lo and stored in shared memory in fragment order
// (one 128-bit load per lane, chunk and part): used for the phase-2 tile so that a warp keeps only its
// phase-1 weights in registers.  A fragments are fetched one chunk ahead of the MMAs that use them.
template <int KC0, int KC1, int P>
__device__ __forceinline__ void tile_matvec_smemA(const uint4 *__restrict__ Ahi, const uint4 *__restrict__ Alo, int lane,
                                                  const float *__restrict__ Bf, int g, int t4, float (&out)[4])
{
    constexpr int N = KC1 - KC0;             // this warp's share of the k chunks
    float bf[N][2];
#pragma unroll
    for (int c = 0; c < N; c++) {
        const int o = g * P + 8 * (KC0 + c) + t4;
        bf[c][0] = Bf[o];
        bf[c][1] = Bf[o + 4];
    }
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
    uint4 h4 = Ahi[KC0 * 32 + lane], l4 = Alo[KC0 * 32 + lane];
#pragma unroll
    for (int c = 0; c < N; c++) {
        const uint32_t ah[4] = {h4.x, h4.y, h4.z, h4.w}, al[4] = {l4.x, l4.y, l4.z, l4.w};
        if (c + 1 < N) { h4 = Ahi[(KC0 + c + 1) * 32 + lane]; l4 = Alo[(KC0 + c + 1) * 32 + lane]; }
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32(bf[c][0], bh0, bl0);
        split_tf32(bf[c][1], bh1, bl1);
        mma_tf32(acc0, ah, bh0, bh1);
        mma_tf32(acc1, al, bh0, bh1);
        mma_tf32(acc2, ah, bl0, bl1);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = acc0[i] + (acc1[i] + acc2[i]);
}

// WIDE (H > 96): 2*NT warps can no longer keep a whole 16 x HP phase-1 tile in registers (the register file is
// 64K words per SM), so the phase-1 fragments of sW live in shared memory as fp32 (2*HP*HP words) and each warp
// keeps its HALF of the phase-2 (sW2) tile in registers instead.
template <int HP, bool WIDE>
__global__ void __launch_bounds__(HP * 4, 1)
gru_mma_kernel(const float *__restrict__ vI, long ldv, const float *__restrict__ sW, const float *__restrict__ sW2,
               float *__restrict__ y, long ldy, const int32_t *__restrict__ lengths, int T, int B, int H, int reverse)
{
    constexpr int NT = HP / 16;              // 16-row tiles per gate
    constexpr int NKC = HP / 8;              // 8-wide k chunks
    constexpr int P = HP + 4;                // pitch of the [8][P] state arrays: conflict-free fragment loads
    constexpr int VLD = 3 * HP + 4;          // pitch of the staged vI rows
    constexpr int NTHREADS = HP * 4;         // 2 * NT warps
    extern __shared__ __align__(16) float smem[];
    float *Hf = smem;                                            // [2][8][P]  h (fp32), double buffered by step parity
    float *RH = Hf + 2 * BT * P;                                 // [8][P]  r * h_{t-1} (consumers split to tf32 hi/lo on the fly)
    float *vbuf = RH + BT * P;                                   // [3][8][VLD] staged vI (ring: t, t+1, t+2)
    uint4 *Wchi = reinterpret_cast<uint4 *>(vbuf + 3 * BT * VLD);     // [NT][NKC][32] sW2 A fragments, tf32 hi
    uint4 *Wclo = Wchi + NT * NKC * 32;                          //                                      tf32 lo
    float4 *W1f = reinterpret_cast<float4 *>(Wchi);              // WIDE: [2*NT][NKC][32] sW A fragments, fp32 (same bytes)
    float4 *Cx = reinterpret_cast<float4 *>(Wclo + NT * NKC * 32);   // [NT][32] phase-2 partial sums of the role-1 warps
    float4 *Zx = Cx + NT * 32;                                   // WIDE: [NT][32] the owners' z, parked during phase 2

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int jt = warp % NT, role = warp / NT;
    const int b_base = blockIdx.x * BT;

    // ---- A fragments (weights) -> registers, zero padded ----
    auto wload = [&](const float *Wm, int row_in_gate, int gate_row0, int k) -> float {
        return (row_in_gate < H && k < H) ? __ldg(Wm + (long)(gate_row0 + row_in_gate) * H + k) : 0.0f;
    };
    constexpr int NWA = WIDE ? NKC / 2 : NKC;
    constexpr bool PIN_W = HP >= 144;        // 18 warps: 96 registers per thread
    float wa[NWA][4];                        // phase 1: z tile (role 0) or r tile (role 1) of sW; WIDE: half c tile of sW2
    for (int e = tid; e < 3 * BT * P + 3 * BT * VLD; e += NTHREADS) smem[e] = 0.0f;
    if constexpr (WIDE) {
        const int r0 = 16 * jt + g, r1 = r0 + 8;
        const int gate0 = role == 0 ? 0 : H;
        for (int kc = 0; kc < NKC; kc++) {
            const int k0 = 8 * kc + t4, k1 = k0 + 4;
            W1f[((role * NT + jt) * NKC + kc) * 32 + lane] =
                make_float4(wload(sW, r0, gate0, k0), wload(sW, r1, gate0, k0), wload(sW, r0, gate0, k1), wload(sW, r1, gate0, k1));
        }
#pragma unroll
        for (int c = 0; c < NWA; c++) {
            const int k0 = 8 * (role * NWA + c) + t4, k1 = k0 + 4;
            wa[c][0] = wload(sW2, r0, 0, k0);
            wa[c][1] = wload(sW2, r1, 0, k0);
            wa[c][2] = wload(sW2, r0, 0, k1);
            wa[c][3] = wload(sW2, r1, 0, k1);
        }
    } else {
        const int r0 = 16 * jt + g, r1 = r0 + 8;
        const int gate0 = role == 0 ? 0 : H;
#pragma unroll
        for (int kc = 0; kc < NKC; kc++) {
            const int k0 = 8 * kc + t4, k1 = k0 + 4;
            wa[kc][0] = wload(sW, r0, gate0, k0);
            wa[kc][1] = wload(sW, r1, gate0, k0);
            wa[kc][2] = wload(sW, r0, gate0, k1);
            wa[kc][3] = wload(sW, r1, gate0, k1);
            if (role == 0) {                 // phase-2 (sW2) fragments of this tile -> shared memory, pre-split
                uint4 h4, l4;
                split_tf32(wload(sW2, r0, 0, k0), h4.x, l4.x);
                split_tf32(wload(sW2, r1, 0, k0), h4.y, l4.y);
                split_tf32(wload(sW2, r0, 0, k1), h4.z, l4.z);
                split_tf32(wload(sW2, r1, 0, k1), h4.w, l4.w);
                Wchi[(jt * NKC + kc) * 32 + lane] = h4;
                Wclo[(jt * NKC + kc) * 32 + lane] = l4;
            }
        }
    }

    // fragment element i of this lane: row j = 16*jt + g + 8*(i >> 1), sequence b = 2*t4 + (i & 1)
    // len[i] = number of steps for which element i is live (0 for padding rows j >= H and sequences past B)
    int len[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int j = 16 * jt + g + 8 * (i >> 1);
        const int bg = b_base + 2 * t4 + (i & 1);
        len[i] = (bg < B && j < H) ? (lengths ? min(lengths[bg], T) : T) : 0;
    }

    // ---- I/O (role-1 warps): per-thread work lists are fixed for the whole scan, so the index math is done once ----
    const int io_tid = tid - NT * 32, io_n = NT * 32;
    const long H3 = 3L * H;
    const bool vec_vi = ((ldv & 3) == 0) && (((uintptr_t)vI & 15) == 0);    // then ldv >= roundup4(3H): whole float4s
    const bool vec_y = ((ldy & 3) == 0) && (((uintptr_t)y & 15) == 0);
    // aligned (128-bit) work split, fixed for the whole scan: vI row b = vq + 2k (k = 0..3), float4 column vc;
    // h row yb, float4 column yc.  io_n = 2*HP threads; 3H/4 < HP and 8 * (HP/4) = io_n.
    // (derived from the thread index inside the lambdas: the wide kernel recomputes them every step from a
    // laundered copy rather than holding them in registers across the matrix products)
    auto stage_vi = [&](int t, int slot, int who, int nwho) {          // generic (prologue / unaligned) path
        if (t < 0 || t >= T) return;
        float *dst = vbuf + slot * BT * VLD;
        for (int e = who; e < BT * (int)H3; e += nwho) {
            const int b = e / (int)H3, c = e - b * (int)H3;
            if (b_base + b < B) dst[b * VLD + c] = __ldg(vI + ((long)t * B + b_base + b) * ldv + c);
        }
    };
    auto stage_vi_fast = [&](int t, int slot, int io_tid) {                        // role-1 threads, cp.async
        if (t < 0 || t >= T) return;
        float *dst = vbuf + slot * BT * VLD;
        const float *src = vI + ((long)t * B + b_base) * ldv;
        if (!vec_vi) {                       // rows not 16-byte aligned: 4-byte async copies, <= 2 columns per thread and row
            const int c0 = io_tid, c1 = io_tid + io_n;                 // io_n = 2*HP >= 2*H, so 3H < 2*io_n
#pragma unroll
            for (int b = 0; b < BT; b++) {
                if (b_base + b < B) {
                    if (c0 < (int)H3) cp_async4(dst + b * VLD + c0, src + b * ldv + c0);
                    if (c1 < (int)H3) cp_async4(dst + b * VLD + c1, src + b * ldv + c1);
                }
            }
            return;
        }
        const int vq = io_tid / HP, vc = io_tid - vq * HP;
        if (4 * vc < (int)H3) {
#pragma unroll
            for (int k = 0; k < BT / 2; k++) {
                const int b = vq + 2 * k;
                if (b_base + b < B) cp_async16(dst + b * VLD + 4 * vc, src + b * ldv + 4 * vc);
            }
        }
    };
    auto store_h = [&](int t, int slot, int who, int nwho) {           // generic path: Hf[slot] -> y[t]
        const float *src = Hf + slot * BT * P;
        for (int e = who; e < BT * H; e += nwho) {
            const int b = e / H, j = e - b * H;
            if (b_base + b < B) y[((long)t * B + b_base + b) * ldy + j] = src[b * P + j];
        }
    };
    auto store_h_fast = [&](int t, int slot, int io_tid) {                         // role-1 threads, 128-bit
        const float *src = Hf + slot * BT * P;
        float *dst = y + ((long)t * B + b_base) * ldy;
        if (!vec_y) {                        // unaligned rows: one column per thread and row, coalesced 4-byte stores
            if (io_tid < H) {
#pragma unroll
                for (int b = 0; b < BT; b++)
                    if (b_base + b < B) dst[(long)b * ldy + io_tid] = src[b * P + io_tid];
            }
            return;
        }
        const int yb = io_tid / (HP / 4), yc = io_tid - yb * (HP / 4);
        if (4 * yc < H && b_base + yb < B) {
            float *d = dst + (long)yb * ldy + 4 * yc;
            const float *sp = src + yb * P + 4 * yc;
            if (4 * yc + 3 < H) {
                *reinterpret_cast<float4 *>(d) = *reinterpret_cast<const float4 *>(sp);
            } else {                         // last, partial quad of a row whose width is not a multiple of 4
                d[0] = sp[0];
                if (4 * yc + 1 < H) d[1] = sp[1];
                if (4 * yc + 2 < H) d[2] = sp[2];
            }
        }
    };

    const int tstep = reverse ? -1 : 1;
    int t = reverse ? T - 1 : 0;
    __syncthreads();
    stage_vi(t, 0, tid, NTHREADS);
    stage_vi(t + tstep, 1, tid, NTHREADS);
    __syncthreads();

    float hreg[4] = {0.f, 0.f, 0.f, 0.f};    // role 0: state of this lane's 4 (row, sequence) elements
    float zreg[4];

    // smem offsets of this lane's 4 fragment elements (row j, sequence b): fixed for the whole scan
    // (element i sits at compile-time offsets from element 0, so one base register serves all four)
    const int o_st0 = 2 * t4 * P + 16 * jt + g, o_vi0 = 2 * t4 * VLD + 16 * jt + g;
    int o_st[4], o_vi[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { o_st[i] = o_st0 + (i & 1) * P + 8 * (i >> 1); o_vi[i] = o_vi0 + (i & 1) * VLD + 8 * (i >> 1); }

    // The two roles run separate copies of the scan loop (same barrier sequence) so that neither carries the
    // other's loop state in registers: the owners keep z / h / lengths, the role-1 warps the I/O bookkeeping.
    // The epilogues are written load-all / compute-all / store-all with unconditional stores so that the four
    // elements' MUFU chains overlap.
    if (role == 0) {
        for (int s = 0; s < T; s++, t += tstep) {
            const int slot = s & 1;          // Hf[slot] receives h_t, Hf[slot^1] holds h_{t-1}
            const float *vrow = vbuf + (s % 3) * BT * VLD;
            if constexpr (PIN_W) {           // keep the compiler from hoisting the tf32 split of the weights (2x the registers)
#pragma unroll
                for (int c = 0; c < NWA; c++)
#pragma unroll
                    for (int i = 0; i < 4; i++) asm volatile("" : "+f"(wa[c][i]));
            }
            // ---- phase 1: z pre-activations ----
            float pre[4];
            if constexpr (WIDE) tile_matvec_smemA32<NKC, P, (HP < 160)>(W1f + jt * NKC * 32, lane, Hf + (slot ^ 1) * BT * P, g, t4, pre);
            else tile_matvec<NKC, P>(wa, Hf + (slot ^ 1) * BT * P, g, t4, pre);
            float vz[4];
#pragma unroll
            for (int i = 0; i < 4; i++) vz[i] = vrow[o_vi[i]];
#pragma unroll
            for (int i = 0; i < 4; i++) zreg[i] = sigmoid_fast(pre[i] + vz[i]);
            if constexpr (WIDE) Zx[jt * 32 + lane] = make_float4(zreg[0], zreg[1], zreg[2], zreg[3]);   // (registers are short)
            bar_sync(1, NTHREADS);           // wait for r*h of every row
            // ---- phase 2: first half of the k range of the c tile, blend ----
            float cpre[4];
            if constexpr (WIDE) tile_matvec_range<0, NWA, P, (HP < 160)>(wa, RH, g, t4, cpre);
            else tile_matvec_smemA<0, NKC / 2, P>(Wchi + jt * NKC * 32, Wclo + jt * NKC * 32, lane, RH, g, t4, cpre);
            bar_sync(2 + jt, 64);            // partner's half of the k range
            {
                const float4 px = Cx[jt * 32 + lane];
                cpre[0] += px.x; cpre[1] += px.y; cpre[2] += px.z; cpre[3] += px.w;
            }
            float *hout = Hf + slot * BT * P;
            float vc[4];
#pragma unroll
            for (int i = 0; i < 4; i++) vc[i] = vrow[o_vi[i] + 2 * H];
            if constexpr (WIDE) {            // z and h_{t-1} come back from shared memory (same lane wrote both)
                const float4 z4 = Zx[jt * 32 + lane];
                zreg[0] = z4.x; zreg[1] = z4.y; zreg[2] = z4.z; zreg[3] = z4.w;
                const float *hprev = Hf + (slot ^ 1) * BT * P;
#pragma unroll
                for (int i = 0; i < 4; i++) hreg[i] = hprev[o_st[i]];
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float hbar = tanh_fast(cpre[i] + vc[i]);
                float hn = zreg[i] * hreg[i] + (1.0f - zreg[i]) * hbar;
                hn = t < len[i] ? hn : 0.0f;     // ragged batch: state stays 0 outside the read
                hreg[i] = hn;
            }
#pragma unroll
            for (int i = 0; i < 4; i++) hout[o_st[i]] = hreg[i];
            bar_sync(0, NTHREADS);
        }
    } else {
        for (int s = 0; s < T; s++, t += tstep) {
            const int slot = s & 1;
            const float *vrow = vbuf + (s % 3) * BT * VLD;
            if constexpr (PIN_W) {
#pragma unroll
                for (int c = 0; c < NWA; c++)
#pragma unroll
                    for (int i = 0; i < 4; i++) asm volatile("" : "+f"(wa[c][i]));
            }
            // ---- phase 1: r pre-activations, r * h_{t-1} ----
            float pre[4];
            if constexpr (WIDE) tile_matvec_smemA32<NKC, P, (HP < 160)>(W1f + (NT + jt) * NKC * 32, lane, Hf + (slot ^ 1) * BT * P, g, t4, pre);
            else tile_matvec<NKC, P>(wa, Hf + (slot ^ 1) * BT * P, g, t4, pre);
            const float *hprev = Hf + (slot ^ 1) * BT * P;
            float vr[4], hp[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { vr[i] = vrow[o_vi[i] + H]; hp[i] = hprev[o_st[i]]; }
            float rh[4];
#pragma unroll
            for (int i = 0; i < 4; i++) rh[i] = sigmoid_fast(pre[i] + vr[i]) * hp[i];
#pragma unroll
            for (int i = 0; i < 4; i++) RH[o_st[i]] = rh[i];
            bar_sync(1, NTHREADS);           // r*h of every row: now every warp consumes it
            // ---- phase 2: second half of the k range of the c tile ----
            float cpart[4];
            if constexpr (WIDE) tile_matvec_range<NKC / 2, NWA, P, (HP < 160)>(wa, RH, g, t4, cpart);
            else tile_matvec_smemA<NKC / 2, NKC, P>(Wchi + jt * NKC * 32, Wclo + jt * NKC * 32, lane, RH, g, t4, cpart);
            Cx[jt * 32 + lane] = make_float4(cpart[0], cpart[1], cpart[2], cpart[3]);
            bar_arrive(2 + jt, 64);
            // HBM traffic: vI two steps ahead (slot read last in step s-1), h_{t-1} (complete since the last barrier) out
            int it = io_tid;
            if constexpr (WIDE) { it = (int)threadIdx.x - NT * 32; asm volatile("" : "+r"(it)); }
            stage_vi_fast(t + 2 * tstep, (s + 2) % 3, it);
            cp_async_commit();
            if (s > 0) store_h_fast(t - tstep, slot ^ 1, it);
            cp_async_wait_1();               // vI of step s+1 (issued one step ago) has landed
            bar_sync(0, NTHREADS);
        }
    }
    t = (reverse ? T - 1 : 0) + T * tstep;
    if (T > 0) store_h(t - tstep, (T - 1) & 1, tid, NTHREADS);           // last step's state
}

template <int HP>
static int launch(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths, int T,
                  int B, int H, int reverse, cudaStream_t st)
{
    constexpr int P = HP + 4, VLD = 3 * HP + 4;
    const size_t smem = sizeof(float) * ((size_t)3 * BT * P + (size_t)3 * BT * VLD) + (size_t)2 * (HP / 16) * (HP / 8) * 32 * 16 +
                        (size_t)(HP / 16) * 32 * 16 * (HP > 96 ? 2 : 1);
    auto kern = gru_mma_kernel<HP, (HP > 96)>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    const unsigned grid = (unsigned)ceil_div(B, BT);
    kern<<<grid, HP * 4, smem, st>>>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse);
    SLOIKA_RETURN_LAUNCH_STATUS();
}

// tanh / sigmoid GRUs with H <= 144; SLOIKA_ERR_UNSUPPORTED otherwise (caller falls back to gru.cu).
int dispatch(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths, int T, int B,
             int H, int reverse, int act, int gate_act, cudaStream_t st)
{
    if (act != SLOIKA_ACT_TANH || gate_act != SLOIKA_ACT_SIGMOID) return SLOIKA_ERR_UNSUPPORTED;
    if (H <= 32) return launch<32>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st);
    if (H <= 48) return launch<48>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st);
    if (H <= 64) return launch<64>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st);
    if (H <= 80) return launch<80>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st);
    if (H <= 96) return launch<96>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st);
    if (H <= 112) return launch<112>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st);
    if (H <= 128) return launch<128>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st);
    if (H <= 144) return launch<144>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st);
    return SLOIKA_ERR_UNSUPPORTED;
}

}  // namespace gru3
}  // namespace sloika
