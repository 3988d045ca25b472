// Acquisition epilogues over the (image, superpixel, class) tables produced by scorer.cu:
// per-region score / dominant class, pool-wide min/max, dominant-class histogram, normalise / ban / re-weight.
// All of them touch N*S*C' table entries once (<= 0.3 % of the bytes the scorer streams).
#include "common.cuh"

#include <algorithm>

namespace {

// ---------------------------------------------------------------------------------------------
// per-region epilogue: one thread per region, C' consecutive table entries each
__global__ void region_scores_kernel(const float* __restrict__ cls_sum, const int32_t* __restrict__ cls_cnt,
                                     const float* __restrict__ w, long long n_regions, int C,
                                     float* __restrict__ score, int32_t* __restrict__ npix, int32_t* __restrict__ dominant) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_regions) return;
    const float* s = cls_sum + r * C;
    const int32_t* n = cls_cnt + r * C;
    float total = 0.f;
    int count = 0, best = 0, best_n = n[0];
    for (int c = 0; c < C; ++c) {
        const int nc = n[c];
        total += (w ? w[c] : 1.f) * s[c];
        count += nc;
        if (nc > best_n) { best_n = nc; best = c; }
    }
    score[r] = total / (float)max(count, 1);
    if (npix) npix[r] = count;
    if (dominant) dominant[r] = best;
}

__device__ __forceinline__ float ordered_to_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// out_bits[0] = min ordered key over non-zero values, out_bits[1] = max ordered key over all values
__global__ void minmax_nonzero_kernel(const float* __restrict__ v, long long n, uint32_t* out_bits) {
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = v[i];
        const uint32_t k = mas::ordered_bits(x);
        if (x != 0.f) lo = min(lo, k);
        hi = max(hi, k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out_bits, lo);
        atomicMax(out_bits + 1, hi);
    }
}

__global__ void minmax_init_kernel(uint32_t* bits) { bits[0] = 0xffffffffu; bits[1] = 0u; }

__global__ void minmax_decode_kernel(const uint32_t* bits, float* out2) {
    out2[0] = bits[0] == 0xffffffffu ? INFINITY : ordered_to_float(bits[0]);
    out2[1] = bits[1] == 0u ? -INFINITY : ordered_to_float(bits[1]);
}

__global__ void dominant_hist_kernel(const int32_t* __restrict__ dominant, long long n, int C, unsigned long long* hist) {
    __shared__ unsigned int local[MAS_MAX_CLASSES];
    if (threadIdx.x < MAS_MAX_CLASSES) local[threadIdx.x] = 0u;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int d = dominant[i];
        if ((unsigned)d < (unsigned)C) atomicAdd(&local[d], 1u);
    }
    __syncthreads();
    if (threadIdx.x < C && local[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)local[threadIdx.x]);
}

__global__ void finalize_scores_kernel(float* __restrict__ score, const int32_t* __restrict__ dominant, long long n,
                                       const float* __restrict__ minmax, int ban_class,
                                       const float* __restrict__ region_weight) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float u = score[i];
    if (minmax) {
        const float lo = minmax[0];
        u = (u - lo) / (minmax[1] - lo);  // == (u - min) / max(u - min): fp32 subtraction is monotone
    }
    const int d = dominant ? dominant[i] : -1;
    if (ban_class >= 0 && d == ban_class) u = 0.f;
    if (region_weight) u = region_weight[d] * u;
    score[i] = u;
}

// w_c = (coeff * pbar_c + 1)^-2 with pbar = mean over the reference's batches of the per-batch mean probability:
// `cumulated += mean(prob, dim=(0,2,3))` per batch in loader order, fp32, then `/ len(loader)`
// (active_selection/my_bvsb_predclsbal_pwr.py:36-47).  The per-batch means are computed by the whole CTA, kMeanBatches
// at a time, into shared memory; one thread per class then adds them in loader order (the reference's fp32 sequence).
constexpr int kMeanBatches = 256;

__global__ void __launch_bounds__(1024) class_weights_kernel(const double* __restrict__ prob_sum, long long n_img, int C, double pixels,
                                                             int ref_batch, float coeff, float* __restrict__ weight) {
    __shared__ float means[kMeanBatches][MAS_MAX_CLASSES];
    const long long n_batches = (n_img + ref_batch - 1) / ref_batch;
    float cumulated = 0.f;
    for (long long chunk = 0; chunk < n_batches; chunk += kMeanBatches) {
        const int len_chunk = (int)min((long long)kMeanBatches, n_batches - chunk);
        for (int idx = threadIdx.x; idx < len_chunk * C; idx += blockDim.x) {
            const int bi = idx / C, c = idx - bi * C;
            const long long b0 = (chunk + bi) * ref_batch;
            const int len = (int)min((long long)ref_batch, n_img - b0);
            double s = 0.0;
            for (int i = 0; i < len; ++i) s += prob_sum[(b0 + i) * C + c];
            means[bi][c] = (float)(s / ((double)len * pixels));
        }
        __syncthreads();
        if (threadIdx.x < C) {
            for (int bi = 0; bi < len_chunk; ++bi) cumulated += means[bi][threadIdx.x];
        }
        __syncthreads();
    }
    if (threadIdx.x < C) {
        const float pbar = cumulated / (float)n_batches;
        const float x = coeff * pbar + 1.f;
        weight[threadIdx.x] = 1.f / (x * x);
    }
}

}  // namespace

extern "C" int mas_class_weights_dev(const double* prob_sum, int64_t n_img, int channels, int64_t pixels_per_image, int ref_batch,
                                     float coeff, float* weight, void* stream) {
    MAS_REQUIRE(prob_sum && weight, MAS_E_BADARG, "class_weights: null pointer");
    MAS_REQUIRE(n_img > 0 && pixels_per_image > 0 && ref_batch > 0, MAS_E_BADARG, "class_weights: bad size");
    MAS_REQUIRE(channels >= 1 && channels <= MAS_MAX_CLASSES, MAS_E_RANGE, "class_weights: channels out of range");
    class_weights_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(prob_sum, n_img, channels, (double)pixels_per_image,
                                                                         ref_batch, coeff, weight);
    mas::count_launches(1);
    MAS_LAUNCH_OK("class_weights_kernel");
    return 0;
}

extern "C" int mas_region_scores_dev(const float* cls_sum, const int32_t* cls_cnt, const float* class_weight,
                                     int64_t n_regions, int channels, float* score, int32_t* npix, int32_t* dominant,
                                     void* stream) {
    MAS_REQUIRE(cls_sum && cls_cnt && score, MAS_E_BADARG, "region_scores: null pointer");
    MAS_REQUIRE(n_regions >= 0 && channels >= 1, MAS_E_BADARG, "region_scores: bad shape");
    if (n_regions == 0) return 0;
    const int threads = 256;
    const long long blocks = (n_regions + threads - 1) / threads;
    region_scores_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(cls_sum, cls_cnt, class_weight, n_regions,
                                                                                 channels, score, npix, dominant);
    mas::count_launches(1);
    MAS_LAUNCH_OK("region_scores_kernel");
    return 0;
}

extern "C" int mas_minmax_nonzero_dev(const float* values, int64_t n, float* out2, void* stream) {
    MAS_REQUIRE(values && out2 && n >= 0, MAS_E_BADARG, "minmax_nonzero: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    // the two ordered keys live in the caller's output buffer while reducing (same size: 2 x 32 bit)
    uint32_t* bits = reinterpret_cast<uint32_t*>(out2);
    minmax_init_kernel<<<1, 1, 0, st>>>(bits);
    mas::count_launches(n > 0 ? 3 : 2);
    if (n > 0) {
        const int threads = 256;
        const long long blocks = std::min<long long>((n + threads - 1) / threads, (long long)mas::sm_count() * 8);
        minmax_nonzero_kernel<<<(unsigned)blocks, threads, 0, st>>>(values, n, bits);
    }
    minmax_decode_kernel<<<1, 1, 0, st>>>(bits, out2);
    MAS_LAUNCH_OK("minmax_nonzero kernels");
    return 0;
}

extern "C" int mas_dominant_hist_dev(const int32_t* dominant, int64_t n_regions, int channels, int64_t* hist, void* stream) {
    MAS_REQUIRE(dominant && hist && n_regions >= 0, MAS_E_BADARG, "dominant_hist: bad argument");
    MAS_REQUIRE(channels >= 1 && channels <= MAS_MAX_CLASSES, MAS_E_RANGE, "dominant_hist: channels out of range");
    if (n_regions == 0) return 0;
    const int threads = 256;
    const long long blocks = std::min<long long>((n_regions + threads - 1) / threads, (long long)mas::sm_count() * 4);
    dominant_hist_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(dominant, n_regions, channels,
                                                                                 reinterpret_cast<unsigned long long*>(hist));
    mas::count_launches(1);
    MAS_LAUNCH_OK("dominant_hist_kernel");
    return 0;
}

extern "C" int mas_finalize_scores_dev(float* score, const int32_t* dominant, int64_t n_regions, const float* minmax,
                                       int ban_class, const float* region_weight, void* stream) {
    MAS_REQUIRE(score && n_regions >= 0, MAS_E_BADARG, "finalize_scores: bad argument");
    MAS_REQUIRE(dominant || (ban_class < 0 && !region_weight), MAS_E_BADARG, "finalize_scores: dominant required");
    if (n_regions == 0) return 0;
    const int threads = 256;
    const long long blocks = (n_regions + threads - 1) / threads;
    finalize_scores_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(score, dominant, n_regions, minmax,
                                                                                   ban_class, region_weight);
    mas::count_launches(1);
    MAS_LAUNCH_OK("finalize_scores_kernel");
    return 0;
}
