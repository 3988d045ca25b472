// Shared pieces of the fused scorer (scorer.cu: TMA / LDG paths and the C ABI; scorer_abreast.cu: the abreast and
// low-resolution path): launch parameters, vector loads and the per-thread walker.  See scorer.cu for the design notes.
#pragma once

#include "common.cuh"
#include "walk.cuh"

namespace mas_scorer {

constexpr int kLdgThreads = 128;
constexpr int kTmaMaxWarps = 8;
constexpr int kTmaStripPx = 128;  // pixels per strip row on the TMA path (32 lanes x 4)

// A launch covers up to kMaxSeg SEGMENTS: batches of images that live in different allocations (consecutive loader
// batches) but fill consecutive rows of the tables.  Images are numbered 0 .. n_img-1 across the segments.
constexpr int kMaxSeg = MAS_MAX_SEGMENTS;

struct StatsParams {
    const void* seg_logits[kMaxSeg];
    const int32_t* seg_ids[kMaxSeg];
    long long seg_stride[kMaxSeg];   // elements between images of the segment's logits
    int seg_first[kMaxSeg + 1];      // first image of segment g; seg_first[n_seg] = n_img
    int n_seg;
    int n_img, C, H, W, S;
    float scale;             // log2(e) / T
    int strips;              // column strips per image
    long long total_rows;    // n_img * strips * H strip rows
    int stages;              // TMA path: ring depth per warp
    int split_ids;           // TMA path: the id rows use one buffer per warp with its own barrier instead of riding in the stages
    float* cls_sum;
    int32_t* cls_cnt;
    double* prob_sum;
    // low-resolution source (mas_bvsb_segment_stats_lowres_dev): the planes hold h_in x w_in values that are bilinearly
    // interpolated to H x W on the fly (align_corners = False); 0 = the planes are full resolution
    int h_in, w_in;
    float ry, rx;            // h_in / H, w_in / W
};

__device__ __forceinline__ int seg_of(const StatsParams& p, int img) {
    int g = 0;
    while (g + 1 < p.n_seg && img >= p.seg_first[g + 1]) ++g;
    return g;
}

// ------------------------------------------------------------------------------------------ loads
template <typename T, int VEC>
struct VecLoad;

template <>
struct VecLoad<float, 4> {
    // volatile: keeps the C' plane loads of a row back to back (memory-level parallelism) instead of
    // letting the compiler sink each one next to its first use
    static __device__ __forceinline__ void global(const float* p, float (&o)[4]) {
        asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "l"(p));
    }
    static __device__ __forceinline__ void shared(const float* p, float (&o)[4]) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
};
template <>
struct VecLoad<float, 1> {
    static __device__ __forceinline__ void global(const float* p, float (&o)[1]) { o[0] = __ldcs(p); }
};
template <>
struct VecLoad<__nv_bfloat16, 4> {
    static __device__ __forceinline__ void unpack(const uint2 v, float (&o)[4]) {
        o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
        o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void global(const __nv_bfloat16* p, float (&o)[4]) {
        uint2 q;
        asm volatile("ld.global.cs.v2.u32 {%0, %1}, [%2];" : "=r"(q.x), "=r"(q.y) : "l"(p));
        unpack(q, o);
    }
    static __device__ __forceinline__ void shared(const __nv_bfloat16* p, float (&o)[4]) {
        unpack(*reinterpret_cast<const uint2*>(p), o);
    }
};
template <>
struct VecLoad<__nv_bfloat16, 1> {
    static __device__ __forceinline__ void global(const __nv_bfloat16* p, float (&o)[1]) {
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
__device__ __forceinline__ void load_ids<1>(const int32_t* p, int (&o)[1]) { o[0] = __ldcs(p); }

// ------------------------------------------------------------------------------------------ per-thread walker
// State a thread carries down its strip: the superpixel whose partial sums live in its private
// shared-memory column, and (when needed) the running softmax sums of the current image.
template <int CMAX, bool EXACT, int VEC, bool NEED_PROB>
struct Walker {
    uint2* col;       // slot of class c: col[c * col_stride]  {float sum bits, int count}
    int col_stride;   // threads per CTA
    int C, S;
    float scale;
    float* cls_sum;
    int32_t* cls_cnt;
    long long img_region;  // index of region 0 of the current image
    int cur;
    float pacc[NEED_PROB ? CMAX : 1];

    __device__ __forceinline__ void init(uint2* column, int stride, const StatsParams& p) {
        col = column; col_stride = stride;
        C = EXACT ? CMAX : p.C; S = p.S; scale = p.scale;
        cls_sum = p.cls_sum; cls_cnt = p.cls_cnt;
        img_region = 0; cur = -1;
        for (int c = 0; c < C; ++c) col[c * col_stride] = make_uint2(0u, 0u);
        if (NEED_PROB) {
#pragma unroll
            for (int c = 0; c < CMAX; ++c) pacc[c] = 0.f;
        }
    }

    // flush the private {sum,count} column into the global tables of the current superpixel
    __device__ __forceinline__ void flush() {
        if (cur < 0) return;
        const long long base = (img_region + cur) * C;
        for (int c = 0; c < C; ++c) {
            const uint2 slot = col[c * col_stride];
            if (slot.y != 0u) {
                atomicAdd(cls_sum + base + c, __uint_as_float(slot.x));
                atomicAdd(cls_cnt + base + c, (int)slot.y);
                col[c * col_stride] = make_uint2(0u, 0u);
            }
        }
        cur = -1;
    }

    // whole warp: add the running softmax sums of image `img` to prob_sum and restart them
    __device__ __forceinline__ void flush_prob(double* prob_sum, int img, int lane) {
        if (!NEED_PROB) return;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            float x = pacc[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0 && (EXACT || c < C) && x != 0.f) atomicAdd(prob_sum + (size_t)img * C + c, (double)x);
            pacc[c] = 0.f;
        }
    }

    __device__ __forceinline__ void row(float (&v)[CMAX][VEC], const int (&id)[VEC]) {
        // ---- pure arithmetic first, the VEC pixels in lock step (independent chains interleave)
        float m1[VEC], m2[VEC], bvsb[VEC];
        int top1[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) { m1[j] = v[0][j]; m2[j] = -INFINITY; top1[j] = 0; }
#pragma unroll
        for (int c = 1; c < CMAX; ++c) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const float x = v[c][j];
                const bool gt = x > m1[j];          // strict: the first index keeps a tie
                m2[j] = fmaxf(m2[j], gt ? m1[j] : x);
                top1[j] = gt ? c : top1[j];
                m1[j] = gt ? x : m1[j];
            }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) bvsb[j] = mas::ex2_approx((m2[j] - m1[j]) * scale) + 1e-8f;
        if (NEED_PROB) {
            float shift[VEC], den_a[VEC], den_b[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) { shift[j] = -m1[j] * scale; den_a[j] = 0.f; den_b[j] = 0.f; }
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    v[c][j] = mas::ex2_approx(fmaf(v[c][j], scale, shift[j]));   // padded planes hold -inf -> 0
                    if (c & 1) den_b[j] += v[c][j]; else den_a[j] += v[c][j];
                }
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) den_a[j] = mas::rcp_approx(den_a[j] + den_b[j]);
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                float t = pacc[c];
#pragma unroll
                for (int j = 0; j < VEC; ++j) t = fmaf(v[c][j], den_a[j], t);
                pacc[c] = t;
            }
        }
        accumulate(bvsb, top1, id);
    }

    // the segmented accumulation of one row: bvsb[j] / top1[j] of the VEC pixels with ids id[j]
    __device__ __forceinline__ void accumulate(const float (&bvsb)[VEC], const int (&top1)[VEC], const int (&id)[VEC]) {
        // does this row still touch the current superpixel?  if not, move on to the row's first id
        // ids outside [0, S) (crop padding, -1, garbage) become -2: never equal to `cur` (>= -1), never accumulated
        int sid[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) sid[j] = ((unsigned)id[j] < (unsigned)S) ? id[j] : -2;
        bool touches = false;
#pragma unroll
        for (int j = 0; j < VEC; ++j) touches |= (sid[j] == cur);
        if (!touches) {
            flush();
            cur = sid[0] >= 0 ? sid[0] : -1;
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const int s = sid[j];
            if (s == cur) {
                uint2 slot = col[top1[j] * col_stride];
                slot.x = __float_as_uint(__uint_as_float(slot.x) + bvsb[j]);
                slot.y += 1u;
                col[top1[j] * col_stride] = slot;
            } else if (s >= 0) {
                const long long r = (img_region + s) * C + top1[j];
                atomicAdd(cls_sum + r, bvsb[j]);
                atomicAdd(cls_cnt + r, 1);
            }
        }
    }

    // bf16 rows (VEC == 4; q[c] = the lane's four bf16 logits of plane c as two packed pairs): the top-2 / arg-max scan
    // runs on the PACKED pairs -- min / max / compare on bf16x2 are exact, so m1, m2 and the arg-max (first index on ties,
    // strict >) equal the fp32 scan of the widened values at 3 instead of 5 instructions per (class, pixel); only the
    // softmax sums widen the values.
    __device__ __forceinline__ void row_bf16(const uint2 (&q)[CMAX], const int (&id)[VEC]) {
        static_assert(VEC == 4, "packed bf16 rows are 4 pixels wide");
        uint32_t m1p[2] = {q[0].x, q[0].y}, m2p[2] = {0xff80ff80u, 0xff80ff80u};      // -inf pairs
        int top1[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) top1[j] = 0;
#pragma unroll
        for (int c = 1; c < CMAX; ++c) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t x = h == 0 ? q[c].x : q[c].y;
                uint32_t t;
                asm("{\n\t"
                    ".reg .pred lo, hi;\n\t"
                    "setp.gt.bf16x2 lo|hi, %3, %4;\n\t"
                    "selp.s32 %0, %5, %0, lo;\n\t"
                    "selp.s32 %1, %5, %1, hi;\n\t"
                    "min.bf16x2 %2, %4, %3;\n\t"
                    "}"
                    : "+r"(top1[2 * h]), "+r"(top1[2 * h + 1]), "=r"(t)
                    : "r"(x), "r"(m1p[h]), "r"(c));
                asm("max.bf16x2 %0, %0, %1;" : "+r"(m2p[h]) : "r"(t));
                asm("max.bf16x2 %0, %0, %1;" : "+r"(m1p[h]) : "r"(x));
            }
        }
        float m1[VEC], m2[VEC], bvsb[VEC];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            m1[2 * h] = __uint_as_float(m1p[h] << 16); m1[2 * h + 1] = __uint_as_float(m1p[h] & 0xffff0000u);
            m2[2 * h] = __uint_as_float(m2p[h] << 16); m2[2 * h + 1] = __uint_as_float(m2p[h] & 0xffff0000u);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) bvsb[j] = mas::ex2_approx((m2[j] - m1[j]) * scale) + 1e-8f;
        if (NEED_PROB) {
            float shift[VEC], den_a[VEC], den_b[VEC], e[CMAX][VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) { shift[j] = -m1[j] * scale; den_a[j] = 0.f; den_b[j] = 0.f; }
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                const float x[VEC] = {__uint_as_float(q[c].x << 16), __uint_as_float(q[c].x & 0xffff0000u),
                                      __uint_as_float(q[c].y << 16), __uint_as_float(q[c].y & 0xffff0000u)};
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    e[c][j] = mas::ex2_approx(fmaf(x[j], scale, shift[j]));   // padded planes hold -inf -> 0
                    if (c & 1) den_b[j] += e[c][j]; else den_a[j] += e[c][j];
                }
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) den_a[j] = mas::rcp_approx(den_a[j] + den_b[j]);
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                float t = pacc[c];
#pragma unroll
                for (int j = 0; j < VEC; ++j) t = fmaf(e[c][j], den_a[j], t);
                pacc[c] = t;
            }
        }
        accumulate(bvsb, top1, id);
    }
};


// scorer_abreast.cu: launch the abreast kernel (vec = 1 or 4 pixels per thread; lowres = interpolate p.h_in x p.w_in planes)
cudaError_t launch_abreast(const StatsParams& p, int vec, bool lowres, int logits_dtype, cudaStream_t stream);

}  // namespace mas_scorer
