// Acquisition pass: fused best-vs-second-best scorer + segmented reduction keyed by superpixel id.
//
// One pass over the NCHW logits produces, per (image, superpixel, class), the sum of bvsb over the
// pixels whose arg-max class is that class and the pixel count (the arg-max histogram), plus the
// per-image sum of softmax probabilities.  Everything the six reference selectors need
// (active_selection/my_bvsb*.py) follows from these tables -- see include/mulactseg_b200.h.
//
// Mapping (HBM-bound streaming kernel, no tensor cores):
//   * one WARP owns a tile of 32*VEC columns x `rows` rows of one image; lane l owns VEC
//     consecutive columns, so each class plane is read as one fully coalesced 128*VEC-byte row
//     segment per warp (128-bit loads when VEC == 4) and a thread walks DOWN its columns.
//   * the class reduction (top-2, softmax) is a per-thread loop over the C' planes held in registers.
//   * superpixels are spatially compact, so a thread stays inside one superpixel for many rows:
//     it accumulates {sum, count} per class in a PRIVATE shared-memory column (no atomics, no bank
//     conflicts: slot (c, tid) lives at bank pair 2*tid) and flushes the non-empty classes with
//     global reductions (RED) only when its superpixel changes or the tile ends.  Pixels of a row
//     that belong to another superpixel than the thread's current one (boundary straddlers) go to
//     global memory directly.  Adversarial (random) id maps stay correct, just slower.
#include "common.cuh"

#include <algorithm>

namespace {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;

template <typename T, int VEC>
struct VecLoad;

template <>
struct VecLoad<float, 4> {
    static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
};
template <>
struct VecLoad<float, 2> {
    static __device__ __forceinline__ void load(const float* p, float (&o)[2]) {
        const float2 v = __ldcs(reinterpret_cast<const float2*>(p));
        o[0] = v.x; o[1] = v.y;
    }
};
template <>
struct VecLoad<float, 1> {
    static __device__ __forceinline__ void load(const float* p, float (&o)[1]) { o[0] = __ldcs(p); }
};
template <>
struct VecLoad<__nv_bfloat16, 4> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[4]) {
        const uint2 v = __ldcs(reinterpret_cast<const uint2*>(p));
        o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
        o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
    }
};
template <>
struct VecLoad<__nv_bfloat16, 2> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[2]) {
        const uint32_t v = __ldcs(reinterpret_cast<const uint32_t*>(p));
        o[0] = __uint_as_float(v << 16); o[1] = __uint_as_float(v & 0xffff0000u);
    }
};
template <>
struct VecLoad<__nv_bfloat16, 1> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[1]) {
        const unsigned short v = __ldcs(reinterpret_cast<const unsigned short*>(p));
        o[0] = __uint_as_float(((uint32_t)v) << 16);
    }
};

template <int VEC>
__device__ __forceinline__ void load_ids(const int32_t* p, int (&o)[VEC]);
template <>
__device__ __forceinline__ void load_ids<4>(const int32_t* p, int (&o)[4]) {
    const int4 v = __ldcs(reinterpret_cast<const int4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <>
__device__ __forceinline__ void load_ids<2>(const int32_t* p, int (&o)[2]) {
    const int2 v = __ldcs(reinterpret_cast<const int2*>(p));
    o[0] = v.x; o[1] = v.y;
}
template <>
__device__ __forceinline__ void load_ids<1>(const int32_t* p, int (&o)[1]) { o[0] = __ldcs(p); }

struct StatsParams {
    const void* logits;
    const int32_t* ids;
    long long image_stride;  // elements between images of `logits`
    int n_img, C, H, W, S;
    float scale;  // log2(e) / T
    int rows, tiles_x, tiles_y;
    long long total_tiles;
    float* cls_sum;
    int32_t* cls_cnt;
    double* prob_sum;
};

// flush the calling thread's private {sum,count} column into the global tables of region `region`
__device__ __forceinline__ void flush_column(uint2* col, int C, long long region_base, float* cls_sum, int32_t* cls_cnt) {
    for (int c = 0; c < C; ++c) {
        const uint2 slot = col[c * kThreads];
        if (slot.y != 0u) {
            atomicAdd(cls_sum + region_base + c, __uint_as_float(slot.x));
            atomicAdd(cls_cnt + region_base + c, (int)slot.y);
            col[c * kThreads] = make_uint2(0u, 0u);
        }
    }
}

template <int CMAX, bool EXACT, int VEC, bool NEED_PROB, typename T>
__global__ void __launch_bounds__(kThreads) bvsb_stats_kernel(const StatsParams p) {
    extern __shared__ uint2 acc[];  // [C][kThreads]: {float sum bits, int count}
    const int C = EXACT ? CMAX : p.C;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    uint2* col = acc + tid;
    for (int c = 0; c < C; ++c) col[c * kThreads] = make_uint2(0u, 0u);

    const long long tile = (long long)blockIdx.x * kWarps + (tid >> 5);
    if (tile >= p.total_tiles) return;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int img = (int)(tile / tiles_per_img);
    const int rem = (int)(tile - (long long)img * tiles_per_img);
    const int ty = rem / p.tiles_x;
    const int tx = rem - ty * p.tiles_x;
    const int x0 = (tx * 32 + lane) * VEC;
    const int y0 = ty * p.rows;
    const int y1 = min(p.H, y0 + p.rows);
    const size_t P = (size_t)p.H * p.W;
    const T* img_logits = reinterpret_cast<const T*>(p.logits) + (size_t)img * (size_t)p.image_stride;
    const int32_t* img_ids = p.ids + (size_t)img * P;
    const long long img_region = (long long)img * p.S;

    float pacc[CMAX];
    if (NEED_PROB) {
#pragma unroll
        for (int c = 0; c < CMAX; ++c) pacc[c] = 0.f;
    }

    int cur = -1;  // superpixel whose partial sums live in this thread's shared-memory column
    if (x0 < p.W) {
        for (int y = y0; y < y1; ++y) {
            const size_t off = (size_t)y * p.W + x0;
            int id[VEC];
            load_ids<VEC>(img_ids + off, id);
            float v[CMAX][VEC];
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                if (EXACT || c < C) {
                    VecLoad<T, VEC>::load(img_logits + (size_t)c * P + off, v[c]);
                } else {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) v[c][j] = -INFINITY;
                }
            }

            // does this row still touch the current superpixel?  if not, move on to the row's first id
            bool touches = false;
#pragma unroll
            for (int j = 0; j < VEC; ++j) touches |= (id[j] == cur);
            if (!touches) {
                if (cur >= 0) flush_column(col, C, (img_region + cur) * C, p.cls_sum, p.cls_cnt);
                cur = ((unsigned)id[0] < (unsigned)p.S) ? id[0] : -1;
            }

#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                float m1 = v[0][j], m2 = -INFINITY;
                int top1 = 0;
#pragma unroll
                for (int c = 1; c < CMAX; ++c) {
                    const float x = v[c][j];
                    const bool gt = x > m1;          // strict: the first index keeps a tie
                    m2 = fmaxf(m2, gt ? m1 : x);
                    top1 = gt ? c : top1;
                    m1 = gt ? x : m1;
                }
                const float bvsb = mas::ex2_approx((m2 - m1) * p.scale) + 1e-8f;
                if (NEED_PROB) {
                    float e[CMAX];
                    float denom = 0.f;
#pragma unroll
                    for (int c = 0; c < CMAX; ++c) {
                        e[c] = mas::ex2_approx((v[c][j] - m1) * p.scale);
                        denom += e[c];
                    }
                    const float inv = __frcp_rn(denom);
#pragma unroll
                    for (int c = 0; c < CMAX; ++c) pacc[c] = fmaf(e[c], inv, pacc[c]);
                }
                const int s = id[j];
                if (s == cur) {
                    uint2 slot = col[top1 * kThreads];
                    slot.x = __float_as_uint(__uint_as_float(slot.x) + bvsb);
                    slot.y += 1u;
                    col[top1 * kThreads] = slot;
                } else if ((unsigned)s < (unsigned)p.S) {
                    const long long r = (img_region + s) * C + top1;
                    atomicAdd(p.cls_sum + r, bvsb);
                    atomicAdd(p.cls_cnt + r, 1);
                }
            }
        }
        if (cur >= 0) flush_column(col, C, (img_region + cur) * C, p.cls_sum, p.cls_cnt);
    }

    if (NEED_PROB) {
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            float x = pacc[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0 && (EXACT || c < C)) atomicAdd(p.prob_sum + (size_t)img * C + c, (double)x);
        }
    }
}

template <int CMAX, bool EXACT, int VEC, bool NEED_PROB, typename T>
cudaError_t launch_stats(const StatsParams& p, cudaStream_t stream) {
    const long long blocks = (p.total_tiles + kWarps - 1) / kWarps;
    const size_t smem = (size_t)p.C * kThreads * sizeof(uint2);
    bvsb_stats_kernel<CMAX, EXACT, VEC, NEED_PROB, T><<<(unsigned)blocks, kThreads, smem, stream>>>(p);
    mas::count_launches(1);
    return cudaGetLastError();
}

template <int VEC, bool NEED_PROB, typename T>
cudaError_t dispatch_channels(const StatsParams& p, cudaStream_t stream) {
    switch (p.C) {
        case 19: return launch_stats<19, true, VEC, NEED_PROB, T>(p, stream);
        case 20: return launch_stats<20, true, VEC, NEED_PROB, T>(p, stream);
        case 21: return launch_stats<21, true, VEC, NEED_PROB, T>(p, stream);
        case 22: return launch_stats<22, true, VEC, NEED_PROB, T>(p, stream);
        default: break;
    }
    if (p.C <= 8) return launch_stats<8, false, VEC, NEED_PROB, T>(p, stream);
    if (p.C <= 16) return launch_stats<16, false, VEC, NEED_PROB, T>(p, stream);
    if (p.C <= 24) return launch_stats<24, false, VEC, NEED_PROB, T>(p, stream);
    return launch_stats<32, false, VEC, NEED_PROB, T>(p, stream);
}

template <typename T>
cudaError_t dispatch_vec(const StatsParams& p, bool need_prob, int vec, cudaStream_t stream) {
    if (vec == 4) return need_prob ? dispatch_channels<4, true, T>(p, stream) : dispatch_channels<4, false, T>(p, stream);
    return need_prob ? dispatch_channels<1, true, T>(p, stream) : dispatch_channels<1, false, T>(p, stream);
}

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

}  // namespace

extern "C" int mas_bvsb_segment_stats_dev(const void* logits, int logits_dtype, int64_t image_stride, const int32_t* ids,
                                          int n_img, int channels, int height, int width, int nseg,
                                          float temperature, float* cls_sum, int32_t* cls_cnt, double* prob_sum,
                                          void* stream) {
    MAS_REQUIRE(logits && ids && cls_sum && cls_cnt, MAS_E_BADARG, "bvsb_segment_stats: null pointer");
    MAS_REQUIRE(n_img >= 0 && height > 0 && width > 0 && nseg > 0, MAS_E_BADARG, "bvsb_segment_stats: bad shape");
    MAS_REQUIRE(channels >= 2 && channels <= MAS_MAX_CLASSES, MAS_E_RANGE,
                "bvsb_segment_stats: channels=%d outside [2,%d]", channels, MAS_MAX_CLASSES);
    MAS_REQUIRE(temperature > 0.f, MAS_E_BADARG, "bvsb_segment_stats: temperature must be > 0");
    MAS_REQUIRE(logits_dtype == MAS_F32 || logits_dtype == MAS_BF16, MAS_E_BADARG, "bvsb_segment_stats: bad dtype");
    if (n_img == 0) return 0;
    const long long plane = (long long)height * width;
    if (image_stride == 0) image_stride = (long long)channels * plane;
    MAS_REQUIRE(image_stride >= (long long)channels * plane, MAS_E_BADARG, "bvsb_segment_stats: image_stride too small");

    const size_t elt = logits_dtype == MAS_F32 ? 4 : 2;
    // 128-bit (f32) / 64-bit (bf16) row segments need every plane row to start VEC-aligned
    const bool aligned = (width % 4 == 0) && (image_stride % 4 == 0) && (((uintptr_t)logits) % (4 * elt) == 0) && (((uintptr_t)ids) % 16 == 0);
    const int vec = aligned ? 4 : 1;

    StatsParams p;
    p.logits = logits; p.ids = ids; p.image_stride = image_stride;
    p.n_img = n_img; p.C = channels; p.H = height; p.W = width; p.S = nseg;
    p.scale = 1.4426950408889634f / temperature;
    p.tiles_x = (width + 32 * vec - 1) / (32 * vec);
    // rows per warp tile: long walks amortise the flush, but keep >= ~8 warps per SM worth of tiles
    int rows = 32;
    const long long want = (long long)mas::sm_count() * 32;
    while (rows > 4 && (long long)n_img * p.tiles_x * ((height + rows - 1) / rows) < want) rows >>= 1;
    p.rows = rows;
    p.tiles_y = (height + rows - 1) / rows;
    p.total_tiles = (long long)n_img * p.tiles_x * p.tiles_y;
    p.cls_sum = cls_sum; p.cls_cnt = cls_cnt; p.prob_sum = prob_sum;
    MAS_REQUIRE((p.total_tiles + kWarps - 1) / kWarps < 0x7fffffffLL, MAS_E_RANGE, "bvsb_segment_stats: too many tiles");

    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = logits_dtype == MAS_F32 ? dispatch_vec<float>(p, prob_sum != nullptr, vec, st)
                                            : dispatch_vec<__nv_bfloat16>(p, prob_sum != nullptr, vec, st);
    if (e != cudaSuccess) return mas::cuda_fail(e, "bvsb_stats_kernel launch");
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
