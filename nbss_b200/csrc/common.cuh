// common.cuh — status codes and small device helpers shared by all kernels of libnbss_b200.so.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// Status codes returned by every extern "C" entry point (see include/nbss_b200.h):
//   0 ok; < 0 argument / shape error; > 0 cudaError_t of the launch.
#define NBSS_OK 0
#define NBSS_ERR_SHAPE (-1)
#define NBSS_ERR_NULL (-2)
#define NBSS_ERR_UNSUPPORTED (-3)
#define NBSS_ERR_WORKSPACE (-4)

#define NBSS_LAUNCH_CHECK()                       \
    do {                                          \
        cudaError_t e__ = cudaGetLastError();     \
        if (e__ != cudaSuccess) return (int)e__;  \
    } while (0)

// Phase profiling (tools/phase_profile.py; `make prof` builds lib/libnbss_b200_prof.so with -DNBSS_PHASE_PROFILE): thread 0 of
// CTA 0 stamps clock64() at the phase boundaries of its SECOND work item (steady state) into a device array that
// nbss_debug_phases() copies out (one reader per .cu: nbss_debug_phases_<unit>).  Compiled out of the product library.
#ifdef NBSS_PHASE_PROFILE
static __device__ unsigned long long g_nbss_phase[4 * 64];  // one copy per translation unit: up to 4 kernels x 64 stamps
#define NBSS_TICK(kid, idx, iter) \
    do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && (iter) == 1) g_nbss_phase[(kid) * 64 + (idx)] = clock64(); } while (0)
// instantiate once per .cu that stamps phases: copies the unit's stamps to host memory and clears them
#define NBSS_PHASE_READER(fn)                                                                        \
    extern "C" int fn(unsigned long long* host_out) {                                                \
        cudaError_t e = cudaMemcpyFromSymbol(host_out, g_nbss_phase, sizeof(unsigned long long) * 4 * 64); \
        if (e != cudaSuccess) return (int)e;                                                         \
        static unsigned long long zeros[4 * 64];                                                     \
        return (int)cudaMemcpyToSymbol(g_nbss_phase, zeros, sizeof(zeros));                          \
    }
#else
#define NBSS_TICK(kid, idx, iter) do { } while (0)
#define NBSS_PHASE_READER(fn)
#endif

namespace nbss {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// SiLU and its derivative (fp32).  The activation epilogues of the T-ConvFFN kernels are bound by the MUFU pipe (16 results
// per clock per SM): 1 / (1 + exp(-x)) costs two MUFU operations per element (EX2, RCP).  sigmoid(x) = 0.5 + 0.5 tanh(x / 2)
// costs ONE (MUFU.TANH, max relative error 2^-11, i.e. an absolute error <= 2.5e-4 on the sigmoid — the same size as the
// rounding of the result to an fp16 MMA operand, which is where every one of these values goes next).
// -DNBSS_SILU_EXACT selects the two-MUFU form (ex2.approx.ftz + rcp.approx.ftz: ~2 ulp) for A/B accuracy measurements.
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#ifdef NBSS_SILU_EXACT
__device__ __forceinline__ float sigmoidf_(float x) { return rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float silu(float x) { return x * sigmoidf_(x); }
#else
__device__ __forceinline__ float sigmoidf_(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
__device__ __forceinline__ float silu(float x) {
    const float h = 0.5f * x;
    return fmaf(h, tanh_approx(h), h);
}
#endif
__device__ __forceinline__ float silu_grad(float x) {
    float s = sigmoidf_(x);
    return s * (1.f + x * (1.f - s));
}

__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace nbss
