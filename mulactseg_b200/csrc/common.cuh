// Shared helpers for the mulactseg_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mulactseg_b200.h"

namespace mas {

void set_error(const char* fmt, ...);
void count_launches(int n);  // bookkeeping for mas_kernel_launches()

inline int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

#define MAS_CUDA_OK(expr)                                      \
    do {                                                       \
        cudaError_t _e = (expr);                               \
        if (_e != cudaSuccess) return mas::cuda_fail(_e, #expr); \
    } while (0)

#define MAS_LAUNCH_OK(name)                                       \
    do {                                                          \
        cudaError_t _e = cudaGetLastError();                      \
        if (_e != cudaSuccess) return mas::cuda_fail(_e, name);   \
    } while (0)

#define MAS_REQUIRE(cond, code, ...)          \
    do {                                      \
        if (!(cond)) {                        \
            mas::set_error(__VA_ARGS__);      \
            return (code);                    \
        }                                     \
    } while (0)

inline int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;
    }
    return cached;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// order-preserving map float -> uint32 (larger float <-> larger unsigned); -0.0 folded onto +0.0
__device__ __host__ __forceinline__ uint32_t ordered_bits(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } cv; cv.f = f; uint32_t u = cv.u;
#endif
    if ((u << 1) == 0u) u = 0u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

}  // namespace mas
