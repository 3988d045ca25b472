// Shared helpers for the mulactseg_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/mulactseg_b200.h"

namespace mas {

void set_error(const char* fmt, ...);
// losses.cu: the fused loss forward pass; `reduce_group` = false skips the group-loss reduction launch (stage-2 labeller)
int multihot_loss_fwd(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask, const uint32_t* info, int n_img,
                      int channels, int height, int width, int nseg, float temperature, int flags, double* acc,
                      uint64_t* group_max, bool reduce_group, void* stream, const void* tiles = nullptr);
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

// ---- per-device launch state.  One process may drive several GPUs (and several host threads): everything a launcher
// caches (SM count, occupancy, "dynamic shared memory opted in") is keyed by the CURRENT device and held in atomics --
// racing threads compute the same value, so a lost update is harmless.
constexpr int kMaxDevices = 64;

inline int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
    return dev;
}

// one int per device, 0 = "not computed yet"
struct PerDeviceInt {
    std::atomic<int> v[kMaxDevices];
    PerDeviceInt() { for (auto& x : v) x.store(0, std::memory_order_relaxed); }
    int get(int dev) const { return v[dev].load(std::memory_order_relaxed); }
    void set(int dev, int value) { v[dev].store(value, std::memory_order_relaxed); }
};

inline int sm_count() {
    static PerDeviceInt cached;
    const int dev = current_device();
    int n = cached.get(dev);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
        cached.set(dev, n);
    }
    return n;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: opt in once per (kernel, device)
template <typename K>
inline cudaError_t opt_in_smem(K kernel, PerDeviceInt& done, int bytes) {
    const int dev = current_device();
    if (done.get(dev) >= bytes) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.set(dev, bytes);
    return e;
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
