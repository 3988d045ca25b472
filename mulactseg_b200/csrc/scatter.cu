// Operator-level drop-in: the torch_scatter operators the reference calls (scatter sum / mean / max, scatter_max with its
// arg output) as plain segmented reductions over a (outer, n, inner) view of `src`.
//
// Reference call sites (paths relative to the reference checkout): active_selection/my_bvsb.py:73 (mean),
// my_bvsb_banignore.py:43-45 (mean + int64 sum of a one-hot), my_bvsb_predclsbal_pwr*.py:65-69, my_bvsb_clsbal_v2*.py:44-47,
// utils/loss.py:122 and trainer/active_joint_multi_predignore*.py:109 / ..._mclossablation2.py:60 (max),
// trainer/eval_save_cosplbl_prop*.py:178,213 (scatter_max with arg).  The twelve hot-path plugins do NOT go through these
// kernels (their fused passes never materialise the operands); this file serves the reference's OTHER call sites
// (loss ablations etc.), which can switch by importing mulactseg_b200.torch_scatter_compat as torch_scatter.
//
// Semantics of torch_scatter 2.0.9: segments nobody writes hold 0; scatter_max's arg is the FIRST element attaining the
// maximum (CPU behaviour; the CUDA build lets any tying element win) and src.size(dim) for an empty segment.
#include "common.cuh"

#include <algorithm>
#include <type_traits>

namespace {

struct ScatterShape {
    long long outer, n, inner, dim_size;
    int index_inner;      // 1: index has the shape of src; 0: index is (outer, n), shared by the `inner` trailing elements
};

__device__ __forceinline__ long long segment_of(const long long* __restrict__ index, const ScatterShape& s, long long o, long long i, long long k) {
    return s.index_inner ? index[(o * s.n + i) * s.inner + k] : index[o * s.n + i];
}

template <typename T>
__global__ void scatter_sum_kernel(const T* __restrict__ src, const long long* __restrict__ index, ScatterShape s, T* __restrict__ out) {
    const long long total = s.outer * s.n * s.inner;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long k = e % s.inner, i = (e / s.inner) % s.n, o = e / (s.inner * s.n);
        const long long seg = segment_of(index, s, o, i, k);
        if (seg < 0 || seg >= s.dim_size) continue;
        T* dst = out + (o * s.dim_size + seg) * s.inner + k;
        if constexpr (sizeof(T) == 8 && !std::is_floating_point<T>::value)
            atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)src[e]);
        else
            atomicAdd(dst, src[e]);
    }
}

// pass 1: running maximum as an order-preserving key (0 = nobody wrote)
__global__ void scatter_max_key_kernel(const float* __restrict__ src, const long long* __restrict__ index, ScatterShape s,
                                       unsigned int* __restrict__ key) {
    const long long total = s.outer * s.n * s.inner;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long k = e % s.inner, i = (e / s.inner) % s.n, o = e / (s.inner * s.n);
        const long long seg = segment_of(index, s, o, i, k);
        if (seg < 0 || seg >= s.dim_size) continue;
        const unsigned int v = mas::ordered_bits(src[e]);      // > 0 for every float but NaN patterns below -inf
        atomicMax(key + (o * s.dim_size + seg) * s.inner + k, v);
    }
}

// pass 2: first position along `dim` whose value equals the maximum
__global__ void scatter_max_arg_kernel(const float* __restrict__ src, const long long* __restrict__ index, ScatterShape s,
                                       const unsigned int* __restrict__ key, long long* __restrict__ arg) {
    const long long total = s.outer * s.n * s.inner;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long k = e % s.inner, i = (e / s.inner) % s.n, o = e / (s.inner * s.n);
        const long long seg = segment_of(index, s, o, i, k);
        if (seg < 0 || seg >= s.dim_size) continue;
        const long long slot = (o * s.dim_size + seg) * s.inner + k;
        if (mas::ordered_bits(src[e]) == key[slot]) atomicMin(reinterpret_cast<unsigned long long*>(arg + slot), (unsigned long long)i);
    }
}

__global__ void scatter_max_finish_kernel(const unsigned int* __restrict__ key, long long count, float* __restrict__ out) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long long)gridDim.x * blockDim.x) {
        const unsigned int k = key[e];
        out[e] = k == 0u ? 0.f : __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
    }
}

unsigned grid_for(long long total) {
    return (unsigned)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)mas::sm_count() * 16));
}

int check_shape(const char* what, long long outer, long long n, long long inner, long long dim_size) {
    MAS_REQUIRE(outer >= 0 && n >= 0 && inner >= 1 && dim_size >= 0, MAS_E_BADARG, "%s: bad shape", what);
    return 0;
}

}  // namespace

extern "C" int mas_scatter_sum_dev(const void* src, int src_dtype, const int64_t* index, int64_t outer, int64_t n, int64_t inner,
                                   int index_has_inner, int64_t dim_size, void* out, void* stream) {
    MAS_REQUIRE(src && index && out, MAS_E_BADARG, "scatter_sum: null pointer");
    MAS_REQUIRE(src_dtype == MAS_F32 || src_dtype == MAS_SCATTER_I64, MAS_E_BADARG, "scatter_sum: dtype must be float32 or int64");
    if (check_shape("scatter_sum", outer, n, inner, dim_size) != 0) return MAS_E_BADARG;
    const long long total = outer * n * inner;
    if (total == 0) return 0;
    const ScatterShape s = {outer, n, inner, dim_size, index_has_inner ? 1 : 0};
    const long long* idx = reinterpret_cast<const long long*>(index);
    if (src_dtype == MAS_F32)
        scatter_sum_kernel<float><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(src), idx, s, reinterpret_cast<float*>(out));
    else
        scatter_sum_kernel<long long><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(src), idx, s,
                                                                                     reinterpret_cast<long long*>(out));
    mas::count_launches(1);
    MAS_LAUNCH_OK("scatter_sum_kernel");
    return 0;
}

extern "C" int mas_scatter_max_dev(const float* src, const int64_t* index, int64_t outer, int64_t n, int64_t inner, int index_has_inner,
                                   int64_t dim_size, float* out, int64_t* arg, uint32_t* key_workspace, void* stream) {
    MAS_REQUIRE(src && index && out && arg && key_workspace, MAS_E_BADARG, "scatter_max: null pointer");
    if (check_shape("scatter_max", outer, n, inner, dim_size) != 0) return MAS_E_BADARG;
    const long long slots = outer * dim_size * inner, total = outer * n * inner;
    if (slots == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const ScatterShape s = {outer, n, inner, dim_size, index_has_inner ? 1 : 0};
    const long long* idx = reinterpret_cast<const long long*>(index);
    MAS_CUDA_OK(cudaMemsetAsync(key_workspace, 0, (size_t)slots * 4, st));
    if (total > 0) {
        scatter_max_key_kernel<<<grid_for(total), 256, 0, st>>>(src, idx, s, key_workspace);
        scatter_max_arg_kernel<<<grid_for(total), 256, 0, st>>>(src, idx, s, key_workspace, reinterpret_cast<long long*>(arg));
        mas::count_launches(2);
    }
    scatter_max_finish_kernel<<<grid_for(slots), 256, 0, st>>>(key_workspace, slots, out);
    mas::count_launches(1);
    MAS_LAUNCH_OK("scatter_max kernels");
    return 0;
}
