// Shared device helpers for the sloika_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "sloika_b200.h"

#ifndef __CUDA_ARCH__
#define SLOIKA_HOST_ONLY 1
#endif

namespace sloika {

// --- activations (sloika/activation.py:8, 38-42, 52, 56) ----------------------------------------
// Accurate libdevice functions on purpose (no --use_fast_math): the posteriors must match the fp32
// CPU path to 1e-4 after thousands of recurrent steps, and the elementwise work is a few percent of
// a GRU step.
__device__ __forceinline__ float sigmoid_ref(float x) {
    // Theano's ScalarSigmoid: hard 0 / 1 outside [-88, 15] for float32, else 1/(1+exp(-x)).
    // Selection, never a blend: exp(-x) of a large negative x is +inf and 1/(1+inf) = 0 anyway.
    float y = 1.0f / (1.0f + expf(-x));
    y = x < -88.0f ? 0.0f : y;
    return x > 15.0f ? 1.0f : y;
}

__device__ __forceinline__ float elu_ref(float x) {
    return x > 0.0f ? x : expm1f(x);      // switch(x > 0, x, expm1(x)); expm1f(-large) = -1
}

__device__ __forceinline__ float apply_act(float x, int act) {
    switch (act) {
        case SLOIKA_ACT_TANH:    return tanhf(x);
        case SLOIKA_ACT_SIGMOID: return sigmoid_ref(x);
        case SLOIKA_ACT_ELU:     return elu_ref(x);
        default:                 return x;
    }
}

// Compile-time selected activation: keeps unrolled epilogues branch-free and small (a runtime switch
// inlined 32x blew the instruction cache of the GEMM epilogue).
template <int ACT>
__device__ __forceinline__ float apply_act_t(float x) {
    if constexpr (ACT == SLOIKA_ACT_TANH) return tanhf(x);
    else if constexpr (ACT == SLOIKA_ACT_SIGMOID) return sigmoid_ref(x);
    else if constexpr (ACT == SLOIKA_ACT_ELU) return elu_ref(x);
    else return x;
}

// Raw MUFU approximations (flush-to-zero forms: no denormal fix-up code around them).  Callers guarantee the
// argument range: ex2 inputs <= 0 (results in [0, 1], tiny ones may flush to 0), lg2 inputs >= 1e-10.
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float SLOIKA_LOG2E = 1.4426950408889634f;
constexpr float SLOIKA_LN2 = 0.6931471805599453f;

// Fast variants for the latency-bound recurrence epilogue: MUFU ex2 / rcp based, absolute error ~1e-7
// (relative 2^-21 on the exponential), three orders of magnitude inside the 1e-4 posterior budget.
__device__ __forceinline__ float sigmoid_fast(float x) {
    float y = __fdividef(1.0f, 1.0f + __expf(-x));
    y = x < -88.0f ? 0.0f : y;
    return x > 15.0f ? 1.0f : y;
}
__device__ __forceinline__ float tanh_fast(float x) {
    const float xc = fminf(fmaxf(x, -30.0f), 30.0f);          // keep exp finite; tanh(+-30) == +-1 in fp32
    return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * xc));
}
template <int ACT>
__device__ __forceinline__ float apply_act_fast(float x) {
    if constexpr (ACT == SLOIKA_ACT_TANH) return tanh_fast(x);
    else if constexpr (ACT == SLOIKA_ACT_SIGMOID) return sigmoid_fast(x);
    else if constexpr (ACT == SLOIKA_ACT_ELU) return x > 0.0f ? x : (__expf(x) - 1.0f);
    else return x;
}

__host__ inline bool act_known(int act) { return act >= SLOIKA_ACT_LINEAR && act <= SLOIKA_ACT_ELU; }

__host__ __device__ __forceinline__ long ceil_div(long a, long b) { return (a + b - 1) / b; }

// Packed fp32x2 FMA (FFMA2 on sm_100): the fp32 pipe retires one warp FFMA per 2 cycles per SMSP,
// so the packed form is what reaches the 128 FMA/clk/SM peak.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

}  // namespace sloika

// Host-side launch epilogue: report asynchronous launch errors as positive cudaError_t codes.
#define SLOIKA_RETURN_LAUNCH_STATUS()                  \
    do {                                               \
        cudaError_t err__ = cudaGetLastError();        \
        return err__ == cudaSuccess ? SLOIKA_OK : (int)err__; \
    } while (0)
