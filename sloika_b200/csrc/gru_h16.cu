// GRU recurrence on the tensor cores: warp-level mma.sync.m16n8k16 on an fp16 hi/lo split (three products per
// term: hi*hi + lo*hi + hi*lo, fp32 accumulate), fp32-equivalent accuracy at twice the MMA rate of the 3xTF32
// kernel in gru_mma.cu and without any per-step splitting of the weights.
//
// Same semantics as gru.cu / gru_mma.cu (reference sloika/layers.py:1010-1021, :85-88, :1449-1450) and the same
// CTA decomposition as gru_mma.cu: one persistent CTA per 8 sequences, rows of a gate cut in 16-row tiles (jt),
// two warps per tile:
//     role 0 ("owner")  z tile of sW in phase 1, first part of the k range of the c tile of sW2 in phase 2, keeps
//                       z and h in registers, does the blend and publishes h_t
//     role 1            r tile of sW in phase 1 (publishes r*h); the rest of the k range of the c tile in phase 2
//                       (partial sums handed to the owner through shared memory and a 64-thread named barrier),
//                       then the HBM traffic: cp.async staging of vI[t+2] and the store of h_{t-1}
// Why fp16 pairs: every value entering a product is bounded (|h| <= 1, weights O(1)), x = hi + lo with
// hi = fp16(x), lo = fp16(x - hi) represents x to 2^-22 relative (absolute floor 2^-25 from fp16 subnormals,
// below the fp32 rounding noise of the sums), and mma.m16n8k16.f16 runs at twice the m16n8k8.tf32 rate
// (measured 947 vs 474 FMA/clk/SM, profiles/r1_mma_probe.txt).  Weights with |w| >= 65504 would overflow fp16:
// the dispatcher cannot see device values, so that (absurd) case is documented rather than detected.
//
// State in shared memory: h_{t-1} and r*h as packed half2 words [8 sequences][PW] (hi and lo arrays, K-major,
// pitch PW = HP/2 + 4 words: conflict-free B-fragment loads), plus h in fp32 (double buffered) for the blend,
// r*h and the output store.
#include <cstdlib>
#include <cuda_fp16.h>
#include "common.cuh"

namespace sloika {
namespace gru4 {

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
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// (x0, x1) -> packed half2 hi and lo words
__device__ __forceinline__ void split2_f16(float x0, float x1, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hb = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hb.x, x1 - hb.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
__device__ __forceinline__ void split1_f16(float x, __half &hi, __half &lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

// D (16x8, fp32) += A (16x16, f16, row) * B (16x8, f16, col)
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One 16-row tile times the 8 sequences over k chunks [KC0, KC0 + N), A fragments (hi / lo) in registers.  The B
// fragments (state, packed half2 [8][PW]) of all chunks are loaded up front so that the tensor pipe never waits on
// a shared-memory round trip per chunk.
template <int KC0, int N, int NW, int PW>
__device__ __forceinline__ void matvec_regA(const uint32_t (&wh)[NW][4], const uint32_t (&wl)[NW][4],
                                            const uint32_t *__restrict__ Bh, const uint32_t *__restrict__ Bl, int g, int t4,
                                            float (&out)[4])
{
    uint32_t bh[N][2], bl[N][2];
#pragma unroll
    for (int c = 0; c < N; c++) {
        const int o = g * PW + 8 * (KC0 + c) + t4;
        bh[c][0] = Bh[o]; bh[c][1] = Bh[o + 4];
        bl[c][0] = Bl[o]; bl[c][1] = Bl[o + 4];
    }
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < N; c++) {
        mma_f16(acc0, wh[c], bh[c][0], bh[c][1]);
        mma_f16(acc1, wl[c], bh[c][0], bh[c][1]);
        mma_f16(acc2, wh[c], bl[c][0], bl[c][1]);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = acc0[i] + (acc1[i] + acc2[i]);
}

// Same with the A fragments in shared memory in fragment order (one 128-bit load per lane, chunk and part);
// A and B of the next chunk are fetched while the MMAs of the current one issue.
template <int KC0, int N, int PW>
__device__ __forceinline__ void matvec_smemA(const uint4 *__restrict__ Ahi, const uint4 *__restrict__ Alo, int lane,
                                             const uint32_t *__restrict__ Bh, const uint32_t *__restrict__ Bl, int g, int t4,
                                             float (&out)[4])
{
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
    uint4 h4 = Ahi[KC0 * 32 + lane], l4 = Alo[KC0 * 32 + lane];
    const int o0 = g * PW + 8 * KC0 + t4;
    uint32_t nbh0 = Bh[o0], nbh1 = Bh[o0 + 4], nbl0 = Bl[o0], nbl1 = Bl[o0 + 4];
#pragma unroll
    for (int c = 0; c < N; c++) {
        const uint32_t ah[4] = {h4.x, h4.y, h4.z, h4.w}, al[4] = {l4.x, l4.y, l4.z, l4.w};
        const uint32_t bh0 = nbh0, bh1 = nbh1, bl0 = nbl0, bl1 = nbl1;
        if (c + 1 < N) {
            h4 = Ahi[(KC0 + c + 1) * 32 + lane];
            l4 = Alo[(KC0 + c + 1) * 32 + lane];
            const int o = o0 + 8 * (c + 1);
            nbh0 = Bh[o]; nbh1 = Bh[o + 4]; nbl0 = Bl[o]; nbl1 = Bl[o + 4];
        }
        mma_f16(acc0, ah, bh0, bh1);
        mma_f16(acc1, al, bh0, bh1);
        mma_f16(acc2, ah, bl0, bl1);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = acc0[i] + (acc1[i] + acc2[i]);
}

// WIDE (H > 96): 2*NT warps cannot keep a whole 16 x HP phase-1 tile (hi and lo) in registers, so the phase-1
// fragments of sW live in shared memory and each warp keeps its PART of the phase-2 (sW2) tile in registers.
template <int HP, bool WIDE>
__global__ void __launch_bounds__(HP * 4, 1)
gru_h16_kernel(const float *__restrict__ vI, long ldv, const float *__restrict__ sW, const float *__restrict__ sW2,
               float *__restrict__ y, long ldy, const int32_t *__restrict__ lengths, int T, int B, int H, int reverse)
{
    constexpr int NT = HP / 16;              // 16-row tiles per gate
    constexpr int NKC = HP / 16;             // 16-wide k chunks
    constexpr int KH0 = NKC / 2;             // phase 2: role 0 takes chunks [0, KH0), role 1 [KH0, NKC)
    constexpr int KH1 = NKC - KH0;
    constexpr int P = HP + 4;                // pitch (floats) of the fp32 [8][P] state arrays
    constexpr int PW = HP / 2 + 4;           // pitch (words) of the packed half2 [8][PW] state arrays
    constexpr int VLD = 3 * HP + 4;          // pitch of the staged vI rows
    constexpr int NTHREADS = HP * 4;         // 2 * NT warps
    constexpr int ZERO_WORDS = 2 * BT * P + 4 * BT * PW + 3 * BT * VLD;
    extern __shared__ __align__(16) float smem[];
    float *Hf = smem;                                                  // [2][8][P]  h (fp32), double buffered by step parity
    uint32_t *Hh = reinterpret_cast<uint32_t *>(Hf + 2 * BT * P);      // [8][PW]  h_{t-1}, fp16 hi pairs
    uint32_t *Hl = Hh + BT * PW;                                       //                   fp16 lo pairs
    uint32_t *RHh = Hl + BT * PW;                                      // [8][PW]  r * h_{t-1}, hi
    uint32_t *RHl = RHh + BT * PW;                                     //                       lo
    float *vbuf = reinterpret_cast<float *>(RHl + BT * PW);            // [3][8][VLD] staged vI (ring: t, t+1, t+2)
    uint4 *Wshi = reinterpret_cast<uint4 *>(vbuf + 3 * BT * VLD);      // A fragments in smem, hi: [NT][NKC][32] of sW2 (c tiles),
    uint4 *Wslo = Wshi + (WIDE ? 2 : 1) * NT * NKC * 32;               //   WIDE: [2*NT][NKC][32] of sW (z then r tiles); lo
    float4 *Cx = reinterpret_cast<float4 *>(Wslo + (WIDE ? 2 : 1) * NT * NKC * 32);   // [NT][32] phase-2 partial sums of role 1

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int jt = warp % NT, role = warp / NT;
    const int b_base = blockIdx.x * BT;

    // ---- A fragments (weights), zero padded, split once ----
    auto wload = [&](const float *Wm, int row_in_gate, int gate_row0, int k) -> float {
        return (row_in_gate < H && k < H) ? __ldg(Wm + (long)(gate_row0 + row_in_gate) * H + k) : 0.0f;
    };
    // fragment of chunk kc: a0 (row g, k 2t4..+1), a1 (row g+8, same k), a2 (row g, k 2t4+8..+9), a3 (row g+8, ...)
    auto wfrag = [&](const float *Wm, int gate_row0, int kc, uint32_t (&fh)[4], uint32_t (&fl)[4]) {
        const int r0 = 16 * jt + g, r1 = r0 + 8, k0 = 16 * kc + 2 * t4, k1 = k0 + 8;
        split2_f16(wload(Wm, r0, gate_row0, k0), wload(Wm, r0, gate_row0, k0 + 1), fh[0], fl[0]);
        split2_f16(wload(Wm, r1, gate_row0, k0), wload(Wm, r1, gate_row0, k0 + 1), fh[1], fl[1]);
        split2_f16(wload(Wm, r0, gate_row0, k1), wload(Wm, r0, gate_row0, k1 + 1), fh[2], fl[2]);
        split2_f16(wload(Wm, r1, gate_row0, k1), wload(Wm, r1, gate_row0, k1 + 1), fh[3], fl[3]);
    };
    constexpr int NWR = WIDE ? KH1 : NKC;    // chunks of register-resident A fragments per warp
    uint32_t wh[NWR][4], wl[NWR][4];         // phase-1 tile of sW (z / r by role); WIDE: this warp's part of the c tile of sW2
    for (int e = tid; e < ZERO_WORDS; e += NTHREADS) smem[e] = 0.0f;
    if constexpr (WIDE) {
        for (int kc = 0; kc < NKC; kc++) {
            uint32_t fh[4], fl[4];
            wfrag(sW, role == 0 ? 0 : H, kc, fh, fl);
            Wshi[((role * NT + jt) * NKC + kc) * 32 + lane] = make_uint4(fh[0], fh[1], fh[2], fh[3]);
            Wslo[((role * NT + jt) * NKC + kc) * 32 + lane] = make_uint4(fl[0], fl[1], fl[2], fl[3]);
        }
#pragma unroll
        for (int c = 0; c < NWR; c++) {
            const int kc = role == 0 ? c : KH0 + c;
            if (role == 0 && c >= KH0) {
#pragma unroll
                for (int i = 0; i < 4; i++) { wh[c][i] = 0u; wl[c][i] = 0u; }
            } else {
                wfrag(sW2, 0, kc, wh[c], wl[c]);
            }
        }
    } else {
#pragma unroll
        for (int kc = 0; kc < NKC; kc++) {
            wfrag(sW, role == 0 ? 0 : H, kc, wh[kc], wl[kc]);
            if (role == 0) {                 // phase-2 (sW2) fragments of this tile -> shared memory
                uint32_t fh[4], fl[4];
                wfrag(sW2, 0, kc, fh, fl);
                Wshi[(jt * NKC + kc) * 32 + lane] = make_uint4(fh[0], fh[1], fh[2], fh[3]);
                Wslo[(jt * NKC + kc) * 32 + lane] = make_uint4(fl[0], fl[1], fl[2], fl[3]);
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

    // ---- I/O (role-1 warps) ----
    const int io_tid = tid - NT * 32, io_n = NT * 32;
    const long H3 = 3L * H;
    const bool vec_vi = ((ldv & 3) == 0) && (((uintptr_t)vI & 15) == 0);    // then ldv >= roundup4(3H): whole float4s
    const bool vec_y = ((ldy & 3) == 0) && (((uintptr_t)y & 15) == 0);
    auto stage_vi = [&](int t, int slot, int who, int nwho) {          // generic (prologue) path
        if (t < 0 || t >= T) return;
        float *dst = vbuf + slot * BT * VLD;
        for (int e = who; e < BT * (int)H3; e += nwho) {
            const int b = e / (int)H3, c = e - b * (int)H3;
            if (b_base + b < B) dst[b * VLD + c] = __ldg(vI + ((long)t * B + b_base + b) * ldv + c);
        }
    };
    auto stage_vi_slow = [&](int t, int slot, int it) {               // role-1 threads, rows not 16-byte aligned:
        if (t < 0 || t >= T) return;                                   // 4-byte async copies, <= 2 columns per thread and row
        float *dst = vbuf + slot * BT * VLD;
        const float *src = vI + ((long)t * B + b_base) * ldv;
        const int c0 = it, c1 = it + io_n;                             // io_n = 2*HP >= 2*H, so 3H < 2*io_n
#pragma unroll
        for (int b = 0; b < BT; b++) {
            if (b_base + b < B) {
                if (c0 < (int)H3) cp_async4(dst + b * VLD + c0, src + b * ldv + c0);
                if (c1 < (int)H3) cp_async4(dst + b * VLD + c1, src + b * ldv + c1);
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
    auto store_h_slow = [&](int t, int slot, int it) {                // role-1 threads, unaligned rows: one column per
        const float *src = Hf + slot * BT * P;                         // thread and row, coalesced 4-byte stores
        float *dst = y + ((long)t * B + b_base) * ldy;
        if (it < H) {
#pragma unroll
            for (int b = 0; b < BT; b++)
                if (b_base + b < B) dst[(long)b * ldy + it] = src[b * P + it];
        }
    };

    const int tstep = reverse ? -1 : 1;
    int t = reverse ? T - 1 : 0;
    __syncthreads();
    stage_vi(t, 0, tid, NTHREADS);
    stage_vi(t + tstep, 1, tid, NTHREADS);
    __syncthreads();

    // smem offsets of this lane's 4 fragment elements (row j, sequence b): element i sits at compile-time offsets
    // from element 0, so one base register per array serves all four
    const int o_st0 = 2 * t4 * P + 16 * jt + g, o_vi0 = 2 * t4 * VLD + 16 * jt + g, o_hf0 = 2 * t4 * (2 * PW) + 16 * jt + g;
    int o_st[4], o_vi[4], o_hf[4];           // fp32 state, staged vI, fp16 state (in halves)
#pragma unroll
    for (int i = 0; i < 4; i++) {
        o_st[i] = o_st0 + (i & 1) * P + 8 * (i >> 1);
        o_vi[i] = o_vi0 + (i & 1) * VLD + 8 * (i >> 1);
        o_hf[i] = o_hf0 + (i & 1) * (2 * PW) + 8 * (i >> 1);
    }
    __half *Hh_h = reinterpret_cast<__half *>(Hh), *Hl_h = reinterpret_cast<__half *>(Hl);
    __half *RHh_h = reinterpret_cast<__half *>(RHh), *RHl_h = reinterpret_cast<__half *>(RHl);

    // The two roles run separate copies of the scan loop (same barrier sequence) so that neither carries the
    // other's loop state in registers.  The epilogues are written load-all / compute-all / store-all with
    // unconditional stores so that the four elements' MUFU chains overlap.
    if (role == 0) {
        float hreg[4] = {0.f, 0.f, 0.f, 0.f};        // state of this lane's 4 (row, sequence) elements
        for (int s = 0; s < T; s++, t += tstep) {
            const int slot = s & 1;          // Hf[slot] receives h_t, Hf[slot^1] holds h_{t-1}
            const float *vrow = vbuf + (s % 3) * BT * VLD;
            // ---- phase 1: z pre-activations ----
            float pre[4];
            if constexpr (WIDE) matvec_smemA<0, NKC, PW>(Wshi + jt * NKC * 32, Wslo + jt * NKC * 32, lane, Hh, Hl, g, t4, pre);
            else matvec_regA<0, NKC, NWR, PW>(wh, wl, Hh, Hl, g, t4, pre);
            float vz[4], zreg[4];
#pragma unroll
            for (int i = 0; i < 4; i++) vz[i] = vrow[o_vi[i]];
#pragma unroll
            for (int i = 0; i < 4; i++) zreg[i] = sigmoid_fast(pre[i] + vz[i]);
            bar_sync(1, NTHREADS);           // wait for r*h of every row (and: every warp is done reading h_{t-1} halves)
            // ---- phase 2: first part of the k range of the c tile, blend ----
            float cpre[4];
            if constexpr (WIDE) matvec_regA<0, KH0, NWR, PW>(wh, wl, RHh, RHl, g, t4, cpre);
            else matvec_smemA<0, KH0, PW>(Wshi + jt * NKC * 32, Wslo + jt * NKC * 32, lane, RHh, RHl, g, t4, cpre);
            float vc[4];
#pragma unroll
            for (int i = 0; i < 4; i++) vc[i] = vrow[o_vi[i] + 2 * H];
            bar_sync(2 + jt, 64);            // partner's part of the k range
            {
                const float4 px = Cx[jt * 32 + lane];
                cpre[0] += px.x; cpre[1] += px.y; cpre[2] += px.z; cpre[3] += px.w;
            }
            float *hout = Hf + slot * BT * P;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float hbar = tanh_fast(cpre[i] + vc[i]);
                float hn = zreg[i] * hreg[i] + (1.0f - zreg[i]) * hbar;
                hn = t < len[i] ? hn : 0.0f;     // ragged batch: state stays 0 outside the read
                hreg[i] = hn;
            }
            __half hh[4], hl[4];
#pragma unroll
            for (int i = 0; i < 4; i++) split1_f16(hreg[i], hh[i], hl[i]);
#pragma unroll
            for (int i = 0; i < 4; i++) { hout[o_st[i]] = hreg[i]; Hh_h[o_hf[i]] = hh[i]; Hl_h[o_hf[i]] = hl[i]; }
            bar_sync(0, NTHREADS);
        }
    } else {
        // I/O state of this thread for the aligned (128-bit) paths, advanced by one time step per iteration:
        // vI rows b = vq + 2k (k = 0..3), float4 column vc, of step t + 2*tstep; h row yb, float4 column yc, of
        // step t - tstep (io_n = 2*HP threads; 3H/4 < HP and 8 * (HP/4) = io_n)
        const int vq = io_tid / HP, vc = io_tid - vq * HP;
        unsigned vmask = 0;
        if (vec_vi && 4 * vc < (int)H3)
            for (int k = 0; k < BT / 2; k++)
                if (b_base + vq + 2 * k < B) vmask |= 1u << k;
        const float *vptr = vI + ((long)(t + 2 * tstep) * B + b_base + vq) * ldv + 4 * vc;
        const int vdst = vq * VLD + 4 * vc;
        const int yb = io_tid / (HP / 4), yc = io_tid - yb * (HP / 4);
        const int ymode = !(vec_y && 4 * yc < H && b_base + yb < B) ? 0 : (4 * yc + 3 < H ? 2 : 1);   // none / partial / full quad
        float *yptr = y + ((long)(t - tstep) * B + b_base + yb) * ldy + 4 * yc;
        const int ysrc = yb * P + 4 * yc;
        for (int s = 0; s < T; s++, t += tstep) {
            const int slot = s & 1;
            const float *vrow = vbuf + (s % 3) * BT * VLD;
            // ---- phase 1: r pre-activations, r * h_{t-1} ----
            float pre[4];
            if constexpr (WIDE) matvec_smemA<0, NKC, PW>(Wshi + (NT + jt) * NKC * 32, Wslo + (NT + jt) * NKC * 32, lane, Hh, Hl, g, t4, pre);
            else matvec_regA<0, NKC, NWR, PW>(wh, wl, Hh, Hl, g, t4, pre);
            const float *hprev = Hf + (slot ^ 1) * BT * P;
            float vr[4], hp[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { vr[i] = vrow[o_vi[i] + H]; hp[i] = hprev[o_st[i]]; }
            float rh[4];
#pragma unroll
            for (int i = 0; i < 4; i++) rh[i] = sigmoid_fast(pre[i] + vr[i]) * hp[i];
            __half rhh[4], rhl[4];
#pragma unroll
            for (int i = 0; i < 4; i++) split1_f16(rh[i], rhh[i], rhl[i]);
#pragma unroll
            for (int i = 0; i < 4; i++) { RHh_h[o_hf[i]] = rhh[i]; RHl_h[o_hf[i]] = rhl[i]; }
            bar_sync(1, NTHREADS);           // r*h of every row: now every warp consumes it
            // ---- phase 2: the rest of the k range of the c tile ----
            float cpart[4];
            if constexpr (WIDE) matvec_regA<KH0, KH1, NWR, PW>(wh, wl, RHh, RHl, g, t4, cpart);
            else matvec_smemA<KH0, KH1, PW>(Wshi + jt * NKC * 32, Wslo + jt * NKC * 32, lane, RHh, RHl, g, t4, cpart);
            Cx[jt * 32 + lane] = make_float4(cpart[0], cpart[1], cpart[2], cpart[3]);
            bar_arrive(2 + jt, 64);
            // HBM traffic: vI two steps ahead (slot read last in step s-1), h_{t-1} (complete since the last barrier) out
            if (vec_vi) {
                const int tn = t + 2 * tstep;
                if (vmask != 0 && tn >= 0 && tn < T) {
                    float *dst = vbuf + ((s + 2) % 3) * BT * VLD + vdst;
#pragma unroll
                    for (int k = 0; k < BT / 2; k++)
                        if (vmask & (1u << k)) cp_async16(dst + 2 * k * VLD, vptr + (long)(2 * k) * ldv);
                }
                vptr += (long)tstep * B * ldv;
            } else {
                stage_vi_slow(t + 2 * tstep, (s + 2) % 3, io_tid);
            }
            cp_async_commit();
            if (vec_y) {
                if (s > 0 && ymode != 0) {
                    const float *sp = Hf + (slot ^ 1) * BT * P + ysrc;
                    if (ymode == 2) {
                        *reinterpret_cast<float4 *>(yptr) = *reinterpret_cast<const float4 *>(sp);
                    } else {                 // last, partial quad of a row whose width is not a multiple of 4
                        yptr[0] = sp[0];
                        if (4 * yc + 1 < H) yptr[1] = sp[1];
                        if (4 * yc + 2 < H) yptr[2] = sp[2];
                    }
                }
                yptr += (long)tstep * B * ldy;
            } else if (s > 0) {
                store_h_slow(t - tstep, slot ^ 1, io_tid);
            }
            cp_async_wait_1();               // vI of step s+1 (issued one step ago) has landed
            bar_sync(0, NTHREADS);
        }
    }
    t = (reverse ? T - 1 : 0) + T * tstep;
    if (T > 0) store_h(t - tstep, (T - 1) & 1, tid, NTHREADS);           // last step's state
}

template <int HP>
static int launch(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths,
                  int T, int B, int H, int reverse, cudaStream_t st)
{
    constexpr bool WIDE = HP > 96;
    constexpr int P = HP + 4, PW = HP / 2 + 4, VLD = 3 * HP + 4, NT = HP / 16;
    const size_t smem = sizeof(float) * ((size_t)2 * BT * P + (size_t)4 * BT * PW + (size_t)3 * BT * VLD) +
                        (size_t)2 * (WIDE ? 2 : 1) * NT * NT * 32 * 16 + (size_t)NT * 32 * 16;
    auto kern = gru_h16_kernel<HP, WIDE>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    const unsigned grid = (unsigned)ceil_div(B, BT);
    kern<<<grid, HP * 4, smem, st>>>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse);
    SLOIKA_RETURN_LAUNCH_STATUS();
}

// tanh / sigmoid GRUs with H <= 144; SLOIKA_ERR_UNSUPPORTED otherwise (caller falls back to gru_mma.cu / gru.cu).
int dispatch(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths, int T,
             int B, int H, int reverse, int act, int gate_act, cudaStream_t st)
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


// ---------------------------------------------------------------------------------------------------------------
// 144 < H <= 256: the recurrent weights (3 H^2 values as fp16 hi / lo pairs: 786 KB at H = 256) do not fit one SM.
// A thread-block CLUSTER of 4 CTAs keeps them on chip: CTA `rank` owns the units [rank * UC, (rank + 1) * UC) of all
// three gates (A fragments of its 3 * UC weight rows in shared memory, 196 KB at H = 256) and every CTA holds a full
// copy of the state operands h_{t-1} and r * h_{t-1} of the cluster's 8 sequences (packed fp16 hi / lo).  Per time step:
//   phase 1   z and r warps: one 16-row tile each over all of K (mma.sync m16n8k16, three products per term);
//             the r warps write r * h of their units into the CTA's own slice of the r*h operand
//   exchange  cluster barrier, every CTA PULLS the other three slices through distributed shared memory
//             (ld.shared::cluster, 128-bit), CTA barrier
//   phase 2   c warps: candidate tile, blend with z (handed over in shared memory), h_t of the CTA's units -> fp32
//             state, own slice of the h operand
//   exchange  cluster barrier, pull the other slices of h, CTA barrier; the z / r warps store h_t to HBM
// Same arithmetic as gru_h16_kernel (Gru.step, sloika/layers.py:1010-1021).  Replaces the step-by-step scan of
// gru.cu (`gru_stepwise`: four launches per time step) for these sizes.
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint4 ld_dsmem_v4(const void *local_ptr, uint32_t rank) {
    const uint32_t laddr = (uint32_t)__cvta_generic_to_shared(local_ptr);
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(rank));
    uint4 v;
    asm volatile("ld.shared::cluster.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(raddr) : "memory");
    return v;
}

constexpr int CLUSTER = 4;

template <int HP>
__global__ void __launch_bounds__(3 * (HP / CLUSTER / 16) * 32, 1)
gru_cluster_kernel(const float *__restrict__ vI, long ldv, const float *__restrict__ sW, const float *__restrict__ sW2,
                   float *__restrict__ y, long ldy, const int32_t *__restrict__ lengths, int T, int B, int H, int reverse)
{
    constexpr int UC = HP / CLUSTER;             // units per CTA
    constexpr int NTU = UC / 16;                 // 16-row tiles per gate and CTA
    constexpr int NW = 3 * NTU;                  // warps: z tiles, r tiles, candidate tiles
    constexpr int NKC = HP / 16;
    constexpr int PW = HP / 2 + 4;               // pitch (words) of the packed half2 [8][PW] operand arrays
    constexpr int NTHREADS = NW * 32;
    constexpr int SLICE_V4 = UC / 8;             // uint4 per sequence and array in one CTA's operand slice
    extern __shared__ __align__(16) float smem[];
    uint4 *Ahi = reinterpret_cast<uint4 *>(smem);                    // [NW][NKC][32] A fragments, hi
    uint4 *Alo = Ahi + NW * NKC * 32;                                //                             lo
    uint32_t *Hh = reinterpret_cast<uint32_t *>(Alo + NW * NKC * 32);   // [8][PW] h_{t-1}, fp16 hi pairs (all units)
    uint32_t *Hl = Hh + BT * PW;
    uint32_t *RHh = Hl + BT * PW;                                    // [8][PW] r * h_{t-1}
    uint32_t *RHl = RHh + BT * PW;
    float *Hf = reinterpret_cast<float *>(RHl + BT * PW);            // [8][UC] fp32 state of this CTA's units
    float *Zs = Hf + BT * UC;                                        // [8][UC] update gate of this step

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int gate = warp / NTU, jt = warp - gate * NTU;
    const uint32_t rank = cluster_rank();
    const int unit0 = (int)rank * UC;
    const int b_base = (blockIdx.x / CLUSTER) * BT;

    // ---- A fragments of this warp's tile (zero padded, split once) ----
    for (int kc = 0; kc < NKC; kc++) {
        uint32_t fh[4], fl[4];
        const int r0 = unit0 + 16 * jt + g, r1 = r0 + 8, k0 = 16 * kc + 2 * t4, k1 = k0 + 8;
        auto wl = [&](int row, int k) -> float {
            if (row >= H || k >= H) return 0.0f;
            return gate < 2 ? __ldg(sW + (long)(gate * H + row) * H + k) : __ldg(sW2 + (long)row * H + k);
        };
        split2_f16(wl(r0, k0), wl(r0, k0 + 1), fh[0], fl[0]);
        split2_f16(wl(r1, k0), wl(r1, k0 + 1), fh[1], fl[1]);
        split2_f16(wl(r0, k1), wl(r0, k1 + 1), fh[2], fl[2]);
        split2_f16(wl(r1, k1), wl(r1, k1 + 1), fh[3], fl[3]);
        Ahi[(warp * NKC + kc) * 32 + lane] = make_uint4(fh[0], fh[1], fh[2], fh[3]);
        Alo[(warp * NKC + kc) * 32 + lane] = make_uint4(fl[0], fl[1], fl[2], fl[3]);
    }
    for (int e = tid; e < 4 * BT * PW + 2 * BT * UC; e += NTHREADS) reinterpret_cast<uint32_t *>(Hh)[e] = 0u;

    // fragment element i of this lane: local unit ul = 16*jt + g + 8*(i >> 1), sequence b = 2*t4 + (i & 1)
    int len[4], ul[4], sb[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        ul[i] = 16 * jt + g + 8 * (i >> 1);
        sb[i] = 2 * t4 + (i & 1);
        const int bg = b_base + sb[i];
        len[i] = (bg < B && unit0 + ul[i] < H) ? (lengths ? min(lengths[bg], T) : T) : 0;
    }
    const int tstep = reverse ? -1 : 1;
    int t = reverse ? T - 1 : 0;
    // this lane's four projection values of a step (gate block `gate`), clamped to valid addresses
    auto load_vi = [&](int tt, float (&v)[4]) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int bg = b_base + sb[i], u = unit0 + ul[i];
            v[i] = (tt >= 0 && tt < T && bg < B && u < H) ? __ldg(vI + ((long)tt * B + bg) * ldv + (long)gate * H + u) : 0.0f;
        }
    };
    __half *Hh_h = reinterpret_cast<__half *>(Hh), *Hl_h = reinterpret_cast<__half *>(Hl);
    __half *RHh_h = reinterpret_cast<__half *>(RHh), *RHl_h = reinterpret_cast<__half *>(RHl);
    // pull the other CTAs' slices of an operand pair (hi, lo) into this CTA's copies
    auto pull = [&](uint32_t *hi, uint32_t *lo) {
        for (int e = tid; e < (CLUSTER - 1) * BT * 2 * SLICE_V4; e += NTHREADS) {
            const int v4 = e % SLICE_V4, rest = e / SLICE_V4;
            const int arr = rest & 1, b = (rest >> 1) % BT, rr = (rest >> 1) / BT;
            const uint32_t src_rank = (rank + 1 + rr) % CLUSTER;
            uint32_t *base = (arr ? lo : hi) + b * PW + (int)src_rank * (UC / 2) + 4 * v4;
            *reinterpret_cast<uint4 *>(base) = ld_dsmem_v4(base, src_rank);
        }
    };
    float vcur[4], vnext[4];
    load_vi(t, vcur);
    __syncthreads();
    cluster_sync_all();                              // every CTA's operands are zeroed before anyone pulls

    for (int s = 0; s < T; s++, t += tstep) {
        load_vi(t + tstep, vnext);                   // next step's projection values, in flight during this step
        // ---- phase 1 ----
        if (gate < 2) {
            float pre[4];
            matvec_smemA<0, NKC, PW>(Ahi + warp * NKC * 32, Alo + warp * NKC * 32, lane, Hh, Hl, g, t4, pre);
            if (gate == 0) {
#pragma unroll
                for (int i = 0; i < 4; i++) Zs[sb[i] * UC + ul[i]] = sigmoid_fast(pre[i] + vcur[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float rh = sigmoid_fast(pre[i] + vcur[i]) * Hf[sb[i] * UC + ul[i]];
                    __half hh, hl;
                    split1_f16(rh, hh, hl);
                    const int o = sb[i] * (2 * PW) + unit0 + ul[i];
                    RHh_h[o] = hh;
                    RHl_h[o] = hl;
                }
            }
        }
        cluster_sync_all();                          // every CTA's slice of r*h is written
        pull(RHh, RHl);
        __syncthreads();
        // ---- phase 2 ----
        if (gate == 2) {
            float cpre[4];
            matvec_smemA<0, NKC, PW>(Ahi + warp * NKC * 32, Alo + warp * NKC * 32, lane, RHh, RHl, g, t4, cpre);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int o = sb[i] * UC + ul[i];
                const float z = Zs[o], hp = Hf[o];
                const float hbar = tanh_fast(cpre[i] + vcur[i]);
                float hn = z * hp + (1.0f - z) * hbar;
                hn = t < len[i] ? hn : 0.0f;         // ragged batch: state stays 0 outside the read
                Hf[o] = hn;
                __half hh, hl;
                split1_f16(hn, hh, hl);
                const int oh = sb[i] * (2 * PW) + unit0 + ul[i];
                Hh_h[oh] = hh;
                Hl_h[oh] = hl;
            }
        }
        cluster_sync_all();                          // every CTA's slice of h_t is written
        pull(Hh, Hl);
        // h_t of this CTA's units -> HBM (coalesced rows; Hf is next written in phase 2 of the next step, behind a barrier)
        for (int e = tid; e < BT * UC; e += NTHREADS) {
            const int b = e / UC, u = e - b * UC;
            if (b_base + b < B && unit0 + u < H) y[((long)t * B + b_base + b) * ldy + unit0 + u] = Hf[e];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; i++) vcur[i] = vnext[i];
    }
    cluster_sync_all();                              // nobody leaves while a neighbour may still pull from it
}

template <int HP>
static int launch_cluster(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy,
                          const int32_t *lengths, int T, int B, int H, int reverse, cudaStream_t st)
{
    constexpr int UC = HP / CLUSTER, NW = 3 * (UC / 16), NKC = HP / 16, PW = HP / 2 + 4;
    const size_t smem = (size_t)2 * NW * NKC * 32 * 16 + sizeof(float) * ((size_t)4 * BT * PW + (size_t)2 * BT * UC);
    auto kern = gru_cluster_kernel<HP>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(CLUSTER * ceil_div(B, BT)));
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    err = cudaLaunchKernelEx(&cfg, kern, vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse);
    if (err != cudaSuccess) return (int)err;
    SLOIKA_RETURN_LAUNCH_STATUS();
}

// tanh / sigmoid GRUs with 144 < H <= 256 on a 4-CTA cluster; SLOIKA_ERR_UNSUPPORTED otherwise.
int dispatch_cluster(const float *vI, long ldv, const float *sW, const float *sW2, float *y, long ldy, const int32_t *lengths,
                     int T, int B, int H, int reverse, int act, int gate_act, cudaStream_t st)
{
    if (act != SLOIKA_ACT_TANH || gate_act != SLOIKA_ACT_SIGMOID) return SLOIKA_ERR_UNSUPPORTED;
    if (H <= 144 || H > 256) return SLOIKA_ERR_UNSUPPORTED;
    if (H <= 192) return launch_cluster<192>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st);
    return launch_cluster<256>(vI, ldv, sW, sW2, y, ldy, lengths, T, B, H, reverse, st);
}

}  // namespace gru4
}  // namespace sloika
