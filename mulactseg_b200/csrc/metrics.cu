// Evaluation counters over label maps: the three per-class counts behind mIoU.
//
// Reference: utils/miou.py:23-38 (MeanIoU._after_step) and :40-54 (_after_step_within_predregion): boolean-filter both
// maps, then 3 * num_classes masked reductions with a `.item()` sync each.  Here: one pass, shared-memory bins per CTA,
// 64-bit global accumulators, no sync.
//   keep(pixel) = target != ignore            (MAS_MIOU_BY_TARGET, _after_step)
//               = output != ignore            (MAS_MIOU_BY_OUTPUT, _after_step_within_predregion)
//   seen[c]     += #kept pixels with target == c
//   correct[c]  += #kept pixels with target == c and output == target
//   positive[c] += #kept pixels with output == c
#include "common.cuh"

#include <algorithm>

namespace {

constexpr int kMaxMiouClasses = 256;

template <typename T>
__global__ void __launch_bounds__(256) miou_counts_kernel(const T* __restrict__ outputs, const T* __restrict__ targets, long long n,
                                                          int C, long long ignore, int by_output,
                                                          unsigned long long* __restrict__ counts) {
    __shared__ unsigned int local[3 * kMaxMiouClasses];
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) local[i] = 0u;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long o = (long long)__ldcs(outputs + i), t = (long long)__ldcs(targets + i);
        if ((by_output ? o : t) == ignore) continue;
        if (t >= 0 && t < C) {
            atomicAdd(&local[t], 1u);
            if (o == t) atomicAdd(&local[C + t], 1u);
        }
        if (o >= 0 && o < C) atomicAdd(&local[2 * C + o], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {
        if (local[i]) atomicAdd(counts + i, (unsigned long long)local[i]);
    }
}

template <typename T>
cudaError_t launch(const void* outputs, const void* targets, long long n, int C, long long ignore, int by_output,
                   unsigned long long* counts, cudaStream_t st) {
    // a CTA's 32-bit bins cannot overflow: it sees at most n / gridDim + 256 elements
    const unsigned blocks = (unsigned)std::max<long long>(std::min<long long>((n + 255) / 256, (long long)mas::sm_count() * 8),
                                                          (n >> 31) + 1);
    miou_counts_kernel<T><<<blocks, 256, 0, st>>>(reinterpret_cast<const T*>(outputs), reinterpret_cast<const T*>(targets), n, C,
                                                  ignore, by_output, counts);
    return cudaGetLastError();
}

}  // namespace

extern "C" int mas_miou_counts_dev(const void* outputs, const void* targets, int labels_dtype, int64_t n, int num_classes,
                                   int64_t ignore_label, int mode, uint64_t* counts, void* stream) {
    MAS_REQUIRE(outputs && targets && counts, MAS_E_BADARG, "miou_counts: null pointer");
    MAS_REQUIRE(n >= 0, MAS_E_BADARG, "miou_counts: negative size");
    MAS_REQUIRE(num_classes >= 1 && num_classes <= kMaxMiouClasses, MAS_E_RANGE, "miou_counts: num_classes outside [1,%d]", kMaxMiouClasses);
    MAS_REQUIRE(mode == MAS_MIOU_BY_TARGET || mode == MAS_MIOU_BY_OUTPUT, MAS_E_BADARG, "miou_counts: bad mode");
    MAS_REQUIRE(labels_dtype == MAS_I32 || labels_dtype == MAS_I64 || labels_dtype == MAS_U8, MAS_E_BADARG, "miou_counts: bad dtype");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* c = reinterpret_cast<unsigned long long*>(counts);
    cudaError_t e;
    if (labels_dtype == MAS_I64) e = launch<long long>(outputs, targets, n, num_classes, ignore_label, mode, c, st);
    else if (labels_dtype == MAS_I32) e = launch<int32_t>(outputs, targets, n, num_classes, ignore_label, mode, c, st);
    else e = launch<uint8_t>(outputs, targets, n, num_classes, ignore_label, mode, c, st);
    mas::count_launches(1);
    if (e != cudaSuccess) return mas::cuda_fail(e, "miou_counts_kernel launch");
    return 0;
}
