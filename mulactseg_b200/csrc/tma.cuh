// TMA / mbarrier helpers shared by the kernels that stage tiles through shared memory with cp.async.bulk.tensor
// (sm_100a; SASS: UTMALDG / UTMASTG, SYNCS).  The tensor-map encoder is resolved through cudaGetDriverEntryPoint, so the
// library needs no libcuda at link time.
#pragma once

#include <cuda.h>  // CUtensorMap and enums only
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

namespace mas_tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "MAS_TMA_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MAS_TMA_DONE;\n\t"
        "bra MAS_TMA_WAIT;\n\t"
        "MAS_TMA_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (before a TMA store / before re-filling a stage)
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int x, int y, int c, int n) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(x), "r"(y), "r"(c), "r"(n) : "memory");
}
__device__ __forceinline__ void load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int x, int y, int n) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(x), "r"(y), "r"(n) : "memory");
}
// the same with an L2 cache-policy operand (createpolicy): streamed inputs are marked evict-first
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}
__device__ __forceinline__ void load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int x, int y, int c, int n, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(x), "r"(y), "r"(c), "r"(n), "l"(policy) : "memory");
}
__device__ __forceinline__ void load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int x, int y, int n, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(x), "r"(y), "r"(n), "l"(policy) : "memory");
}
// shared -> global; completion is tracked per thread through bulk async-groups
__device__ __forceinline__ void store_4d(const CUtensorMap* tm, uint32_t src, int x, int y, int c, int n) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(tm), "r"(src), "r"(x), "r"(y), "r"(c), "r"(n) : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest `N` groups of this thread have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

// dense (rank)-d map over `dims` (innermost first) with byte strides of the outer dimensions and a box; false if unsupported
inline bool encode(CUtensorMap* map, CUtensorMapDataType type, int rank, const void* base, const cuuint64_t* dims,
                   const cuuint64_t* strides, const cuuint32_t* box) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return fn(map, type, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace mas_tma
