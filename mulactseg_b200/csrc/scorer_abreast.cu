// Abreast / low-resolution path of the fused scorer (see scorer.cu for the overall design).
#include "scorer.cuh"

#include <stdlib.h>

#include <algorithm>

using namespace mas_scorer;

namespace {

// ------------------------------------------------------------------------------------------ abreast path
// Rows that are not 16-byte aligned (VOC crops: 513 x 513 floats) cannot use 128-bit loads or TMA, and with an odd plane
// size every class plane has its own alignment phase.  A warp's 128-byte row segment then straddles two lines / five
// sectors, and when the warps of a CTA walk strips that are far apart the straddled sectors are fetched twice.  Here the
// kAbreast warps of a CTA walk ADJACENT strips at the SAME row: what one warp's segment leaves of a sector, its
// neighbour consumes within the same few hundred cycles (L1 / L2 hit).  Work units are (image, group of kAbreast
// strips, y), cut into one contiguous range per CTA.  Two lanes per warp prefetch the lines of the row `ahead` units
// ahead into L1 (no registers held), so the register loads of a row find their data on chip.
//
// The same kernel serves the LOW-RESOLUTION source (LOWRES): the C' planes hold the network head's h_in x w_in logits and
// every full-resolution value is produced on the fly as F.interpolate(mode='bilinear', align_corners=False) would
// (models/segmentation/utils.py:28-34) -- the 16x larger tensor is never written or read.
constexpr int kAbreast = 4;
constexpr int kAheadDefault = 0;     // rows prefetched ahead into L1 (MAS_SCORER_AHEAD overrides): measured on 513x513x22, 0 -> 3.75 TB/s, 2 -> 3.54, 4 -> 3.07 (line-granular prefetch over-fetches)

template <typename T>
__device__ __forceinline__ float load_lowres(const T* p);
template <>
__device__ __forceinline__ float load_lowres<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_lowres<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __uint_as_float(((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
}

template <int VEC>
__device__ __forceinline__ void load_row4(const float* p, float (&o)[VEC]) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = q.x; if (VEC > 1) { o[1 % VEC] = q.y; o[2 % VEC] = q.z; o[3 % VEC] = q.w; }
}
template <int VEC>
__device__ __forceinline__ void load_row4(const __nv_bfloat16* p, float (&o)[VEC]) {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
    o[0] = __uint_as_float(q.x << 16);
    if (VEC > 1) { o[1 % VEC] = __uint_as_float(q.x & 0xffff0000u); o[2 % VEC] = __uint_as_float(q.y << 16); o[3 % VEC] = __uint_as_float(q.y & 0xffff0000u); }
}

// source index and weight of torch's area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=false)
__device__ __forceinline__ void bilinear_tap(int dst, float scale, int size_in, int& i0, int& step, float& lambda1) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = (int)src;
    if (i0 > size_in - 1) i0 = size_in - 1;
    step = (i0 < size_in - 1) ? 1 : 0;
    lambda1 = src - (float)i0;
}

template <int CMAX, bool EXACT, int VEC, bool NEED_PROB, typename T, bool LOWRES>
__global__ void __launch_bounds__(kAbreast * 32) bvsb_stats_abreast_kernel(const StatsParams p) {
    extern __shared__ uint2 acc[];  // [C][kAbreast * 32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Walker<CMAX, EXACT, VEC, NEED_PROB> w;
    w.init(acc + tid, kAbreast * 32, p);
    const int C = w.C;

    long long r0, r1;      // units of this CTA
    mas::warp_range(p.total_rows, (long long)blockIdx.x, (long long)gridDim.x, r0, r1);
    if (r0 >= r1) return;
    const int groups = p.strips;            // strip GROUPS per image (p.strips is repurposed by the launcher)
    mas::Cursor at;
    at.seek(r0, groups, p.H);

    const size_t P = (size_t)p.H * p.W;
    const size_t P_src = LOWRES ? (size_t)p.h_in * p.w_in : P;
    const T* img_logits;
    const int32_t* img_ids;
    auto locate = [&](int img) {
        const int g = seg_of(p, img);
        const size_t local = (size_t)(img - p.seg_first[g]);
        img_logits = reinterpret_cast<const T*>(p.seg_logits[g]) + local * (size_t)p.seg_stride[g];
        img_ids = p.seg_ids[g] + local * P;
    };
    locate(at.img);
    w.img_region = (long long)at.img * p.S;
    int x0 = ((at.strip * kAbreast + warp) * 32 + lane) * VEC;
    // LOWRES: the horizontal taps depend on the column only -- constant while the warp walks down its strip
    const bool quad = LOWRES && p.W == 4 * p.w_in;         // exact x4 up-sampling: three column loads serve four pixels
    int cx[VEC], cstep[VEC];
    float cl1[VEC];
    auto column_taps = [&]() {
        if (LOWRES) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) bilinear_tap(min(x0 + j, p.W - 1), p.rx, p.w_in, cx[j], cstep[j], cl1[j]);
        }
    };
    column_taps();

    // lines of the row `ahead` units further down this strip (same image): lanes 0 and 31 cover the segment's two ends
    auto prefetch_row = [&](int y) {
        if (y >= p.H || x0 >= p.W) return;
        if (VEC == 1 ? (lane == 0 || lane == 31) : ((lane & 7) == 0 || lane == 31)) {
            const size_t off = (size_t)y * p.W + min(x0 + (lane == 31 ? VEC - 1 : 0), p.W - 1);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(img_ids + off));
            if (!LOWRES) {
                const T* q = img_logits + off;
                for (int c = 0; c < C; ++c) { asm volatile("prefetch.global.L1 [%0];" ::"l"(q)); q += P; }
            }
        }
    };
    const int ahead = p.stages;
    for (int k = 1; k <= ahead; ++k) prefetch_row(at.y + k);

    for (long long r = r0; r < r1; ++r) {
        if (ahead > 0) prefetch_row(at.y + ahead);
        if (x0 < p.W) {
            const size_t off = (size_t)at.y * p.W + x0;
            int id[VEC];
            if (VEC == 4) {
                load_ids<VEC>(img_ids + off, id);
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) id[j] = __ldg(img_ids + off + j);
            }
            float v[CMAX][VEC];
            if (LOWRES) {
                int y0i, ystep;
                float l1y;
                bilinear_tap(at.y, p.ry, p.h_in, y0i, ystep, l1y);
                const float l0y = 1.f - l1y;
                const T* row0 = img_logits + (size_t)y0i * p.w_in;
                const size_t down = (size_t)ystep * p.w_in;
                if (VEC == 4 && quad) {
                    // exact x4 (the DeepLab head: 256x512 -> 1024x2048): the thread's four pixels x0 = 4k .. 4k+3 read the
                    // low-resolution columns k-1, k (pixels 0, 1) and k, k+1 (pixels 2, 3).  Three loads per row serve all
                    // four pixels; at the borders the clamped column repeats its neighbour, which is what torch's taps do
                    // (left: src clamps to 0 => weight 0 on the repeated value; right: step 0 => the same element twice).
                    const int k = x0 >> 2;
                    const int ka = max(k - 1, 0) - k, kc = min(k + 1, p.w_in - 1) - k;
                    const T* q0 = row0 + k;
                    const float w1a = cl1[0], w1b = cl1[1], w1c = cl1[2], w1d = cl1[3];
                    const float w0a = 1.f - w1a, w0b = 1.f - w1b, w0c = 1.f - w1c, w0d = 1.f - w1d;
#pragma unroll
                    for (int c = 0; c < CMAX; ++c) {
                        if (EXACT || c < C) {
                            const float a0 = load_lowres<T>(q0 + ka), b0 = load_lowres<T>(q0), c0 = load_lowres<T>(q0 + kc);
                            const float a1 = load_lowres<T>(q0 + down + ka), b1 = load_lowres<T>(q0 + down), c1 = load_lowres<T>(q0 + down + kc);
                            v[c][0] = l0y * (w0a * a0 + w1a * b0) + l1y * (w0a * a1 + w1a * b1);
                            v[c][1 % VEC] = l0y * (w0b * a0 + w1b * b0) + l1y * (w0b * a1 + w1b * b1);
                            v[c][2 % VEC] = l0y * (w0c * b0 + w1c * c0) + l1y * (w0c * b1 + w1c * c1);
                            v[c][3 % VEC] = l0y * (w0d * b0 + w1d * c0) + l1y * (w0d * b1 + w1d * c1);
                        } else {
#pragma unroll
                            for (int j = 0; j < VEC; ++j) v[c][j] = -INFINITY;
                        }
                        q0 += P_src;
                    }
                } else {
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
                    if (EXACT || c < C) {
#pragma unroll
                        for (int j = 0; j < VEC; ++j) {
                            const T* q = row0 + cx[j];
                            const float a = load_lowres<T>(q), b = load_lowres<T>(q + cstep[j]);
                            const float d = load_lowres<T>(q + down), e = load_lowres<T>(q + down + cstep[j]);
                            const float l0x = 1.f - cl1[j];
                            // the expression of upsample_bilinear2d: h0 * (w0 * a + w1 * b) + h1 * (w0 * d + w1 * e)
                            v[c][j] = l0y * (l0x * a + cl1[j] * b) + l1y * (l0x * d + cl1[j] * e);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < VEC; ++j) v[c][j] = -INFINITY;
                    }
                    row0 += P_src;
                }
                }
            } else {
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
                    if (EXACT || c < C) {
                        // cached loads (not the streaming .cs of the other paths): they are meant to hit the lines the
                        // prefetch and the neighbouring warps brought into L1
                        if (VEC == 4) load_row4(img_logits + (size_t)c * P + off, v[c]);
                        else v[c][0] = load_lowres<T>(img_logits + (size_t)c * P + off);
                    } else {
#pragma unroll
                        for (int j = 0; j < VEC; ++j) v[c][j] = -INFINITY;
                    }
                }
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) asm volatile("" : "+f"(v[c][j]));
                }
            }
            w.row(v, id);
        }
        const int img_done = at.img;
        const int step = at.advance(groups, p.H);
        if (step != 0) {
            x0 = ((at.strip * kAbreast + warp) * 32 + lane) * VEC;
            column_taps();
            if (step == 2) {
                w.flush();
                w.flush_prob(p.prob_sum, img_done, lane);
                if (at.img < p.n_img) locate(at.img);
                w.img_region += p.S;
            }
            if (at.img < p.n_img) {
                for (int k = 1; k <= ahead; ++k) prefetch_row(at.y + k);
            }
        }
    }
    w.flush();
    if (at.img < p.n_img && (at.strip != 0 || at.y != 0)) w.flush_prob(p.prob_sum, at.img, lane);
}


template <typename K>
int resident_blocks(K kernel, int threads, size_t smem) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) n = 1;
    return n;
}

template <int CMAX, bool EXACT, int VEC, bool NEED_PROB, typename T, bool LOWRES>
cudaError_t launch_one(StatsParams p, cudaStream_t stream) {
    auto kernel = bvsb_stats_abreast_kernel<CMAX, EXACT, VEC, NEED_PROB, T, LOWRES>;
    constexpr int threads = kAbreast * 32;
    const size_t smem = (size_t)p.C * threads * sizeof(uint2);
    static mas::PerDeviceInt occ;   // per instantiation and device
    const int dev = mas::current_device();
    int per_sm = occ.get(dev);
    if (per_sm == 0) {
        per_sm = resident_blocks(kernel, threads, (size_t)CMAX * threads * sizeof(uint2));
        occ.set(dev, per_sm);
    }
    const int strips = (p.W + 32 * VEC - 1) / (32 * VEC);
    p.strips = (strips + kAbreast - 1) / kAbreast;          // strip GROUPS per image
    const char* ahead = getenv("MAS_SCORER_AHEAD");          // development switch
    p.stages = (ahead && *ahead) ? std::min(std::max(atoi(ahead), 0), 8) : kAheadDefault;
    p.total_rows = (long long)p.n_img * p.strips * p.H;     // work units
    const long long cap = (p.total_rows + 7) / 8;           // never fewer than ~8 units per CTA
    const long long blocks = std::max<long long>(1, std::min<long long>((long long)mas::sm_count() * per_sm, cap));
    kernel<<<(unsigned)blocks, threads, smem, stream>>>(p);
    mas::count_launches(1);
    return cudaGetLastError();
}

template <int CMAX, bool EXACT, bool NEED_PROB, typename T>
cudaError_t launch_variant(const StatsParams& p, int vec, bool lowres, cudaStream_t stream) {
    if (lowres) return vec == 4 ? launch_one<CMAX, EXACT, 4, NEED_PROB, T, true>(p, stream) : launch_one<CMAX, EXACT, 1, NEED_PROB, T, true>(p, stream);
    return vec == 4 ? launch_one<CMAX, EXACT, 4, NEED_PROB, T, false>(p, stream) : launch_one<CMAX, EXACT, 1, NEED_PROB, T, false>(p, stream);
}

template <bool NEED_PROB, typename T>
cudaError_t dispatch_channels(const StatsParams& p, int vec, bool lowres, cudaStream_t stream) {
    switch (p.C) {
        case 19: return launch_variant<19, true, NEED_PROB, T>(p, vec, lowres, stream);
        case 20: return launch_variant<20, true, NEED_PROB, T>(p, vec, lowres, stream);
        case 21: return launch_variant<21, true, NEED_PROB, T>(p, vec, lowres, stream);
        case 22: return launch_variant<22, true, NEED_PROB, T>(p, vec, lowres, stream);
        default: break;
    }
    if (p.C <= 8) return launch_variant<8, false, NEED_PROB, T>(p, vec, lowres, stream);
    if (p.C <= 16) return launch_variant<16, false, NEED_PROB, T>(p, vec, lowres, stream);
    if (p.C <= 24) return launch_variant<24, false, NEED_PROB, T>(p, vec, lowres, stream);
    return launch_variant<32, false, NEED_PROB, T>(p, vec, lowres, stream);
}

}  // namespace

cudaError_t mas_scorer::launch_abreast(const StatsParams& p, int vec, bool lowres, int logits_dtype, cudaStream_t stream) {
    if (logits_dtype == MAS_F32)
        return p.prob_sum ? dispatch_channels<true, float>(p, vec, lowres, stream) : dispatch_channels<false, float>(p, vec, lowres, stream);
    return p.prob_sum ? dispatch_channels<true, __nv_bfloat16>(p, vec, lowres, stream)
                      : dispatch_channels<false, __nv_bfloat16>(p, vec, lowres, stream);
}
