// Acquisition pass: fused best-vs-second-best scorer + segmented reduction keyed by superpixel id.
//
// One pass over the NCHW logits produces, per (image, superpixel, class), the sum of bvsb over the
// pixels whose arg-max class is that class and the pixel count (the arg-max histogram), plus the
// per-image sum of softmax probabilities.  Everything the six reference selectors need
// (active_selection/my_bvsb*.py) follows from these tables -- see include/mulactseg_b200.h.
//
// Work decomposition (HBM-bound streaming kernel, no tensor cores):
//   * the unit of work is a STRIP ROW: 32*VEC consecutive pixels of one image row; lane l of a warp owns
//     VEC consecutive columns, so every class plane is read as one fully coalesced row segment.
//   * strip rows are linearised (image, strip, y) and cut into ONE contiguous range per warp of a
//     single-wave persistent grid (SMs x resident CTAs): every warp gets the same number of rows
//     (+-1) whatever the batch size, and walks DOWN its strip.
//   * the class reduction (top-2, softmax) is a per-thread loop over the C' planes held in registers.
//   * superpixels are spatially compact, so a thread stays inside one superpixel for many rows:
//     it accumulates {sum, count} per class in a PRIVATE shared-memory column (no atomics, no bank
//     conflicts: slot (c, tid) lives at bank pair 2*tid) and flushes the non-empty classes with
//     global reductions (RED) only when its superpixel changes or its range ends.  Pixels of a row
//     that belong to another superpixel than the thread's current one (boundary straddlers) go to
//     global memory directly.  Adversarial (random) id maps stay correct, just slower.
//
// Two data paths feed the same per-row code:
//   * TMA (default when rows are 16-byte aligned): every warp runs its own ring of shared-memory
//     stages; lane 0 issues one cp.async.bulk.tensor box {128 px, 1 row, C' planes} for the logits and
//     one for the ids per strip row, completion is signalled on a per-stage mbarrier.  Bytes in flight
//     do not depend on registers or on the warp being scheduled.
//   * LDG (any shape): 128-bit (VEC = 4) or scalar (VEC = 1) streaming loads straight to registers.
#include "common.cuh"
#include "walk.cuh"

#include <cuda.h>  // CUtensorMap and enums only; the encoder is resolved through cudaGetDriverEntryPoint
#include <stdlib.h>

#include <algorithm>
#include <mutex>

namespace {

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
    float* cls_sum;
    int32_t* cls_cnt;
    double* prob_sum;
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
        // ---- then the segmented accumulation
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
};

// ------------------------------------------------------------------------------------------ LDG path
template <int CMAX, bool EXACT, int VEC, bool NEED_PROB, typename T>
__global__ void __launch_bounds__(kLdgThreads) bvsb_stats_ldg_kernel(const StatsParams p) {
    extern __shared__ uint2 acc[];  // [C][kLdgThreads]
    const int tid = threadIdx.x, lane = tid & 31;
    Walker<CMAX, EXACT, VEC, NEED_PROB> w;
    w.init(acc + tid, kLdgThreads, p);
    const int C = w.C;

    long long r0, r1;
    mas::warp_range(p.total_rows, (long long)blockIdx.x * (kLdgThreads / 32) + (tid >> 5), (long long)gridDim.x * (kLdgThreads / 32), r0, r1);
    if (r0 >= r1) return;
    mas::Cursor at;
    at.seek(r0, p.strips, p.H);

    const size_t P = (size_t)p.H * p.W;
    const T* img_logits;
    const int32_t* img_ids;
    auto locate = [&](int img) {      // pointers of image `img` inside its segment
        const int g = seg_of(p, img);
        const size_t local = (size_t)(img - p.seg_first[g]);
        img_logits = reinterpret_cast<const T*>(p.seg_logits[g]) + local * (size_t)p.seg_stride[g];
        img_ids = p.seg_ids[g] + local * P;
    };
    locate(at.img);
    w.img_region = (long long)at.img * p.S;
    int x0 = (at.strip * 32 + lane) * VEC;

    for (long long r = r0; r < r1; ++r) {
        if (x0 < p.W) {
            const size_t off = (size_t)at.y * p.W + x0;
            int id[VEC];
            load_ids<VEC>(img_ids + off, id);
            float v[CMAX][VEC];
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                if (EXACT || c < C) {
                    VecLoad<T, VEC>::global(img_logits + (size_t)c * P + off, v[c]);
                } else {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) v[c][j] = -INFINITY;
                }
            }
            // all loads of the row are issued before any of them is consumed
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) asm volatile("" : "+f"(v[c][j]));
            }
            w.row(v, id);
        }
        const int img_done = at.img;
        const int step = at.advance(p.strips, p.H);
        if (step != 0) {
            x0 = (at.strip * 32 + lane) * VEC;
            if (step == 2) {
                w.flush();
                w.flush_prob(p.prob_sum, img_done, lane);
                if (at.img < p.n_img) locate(at.img);
                w.img_region += p.S;
            }
        }
    }
    w.flush();
    if (at.img < p.n_img && (at.strip != 0 || at.y != 0)) w.flush_prob(p.prob_sum, at.img, lane);
}

// ------------------------------------------------------------------------------------------ TMA path
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
        "MAS_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MAS_DONE;\n\t"
        "bra MAS_WAIT;\n\t"
        "MAS_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int x, int y, int c, int n,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(dst), "l"(tm), "r"(bar), "r"(x), "r"(y), "r"(c), "r"(n), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int x, int y, int n,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(dst), "l"(tm), "r"(bar), "r"(x), "r"(y), "r"(n), "l"(policy)
        : "memory");
}

// Shared memory of the TMA kernel (dynamic, 128-byte aligned):
//   [warps][stages] stage = { T logits[C][128] ; int32 ids[128] }     filled by TMA
//   [C][threads] uint2                                                private accumulation columns
//   [warps][stages] uint64                                            mbarriers ("stage full")
struct TmaMaps {     // one pair of tensor maps per segment
    CUtensorMap logits[kMaxSeg];
    CUtensorMap ids[kMaxSeg];
};

template <int CMAX, bool EXACT, bool NEED_PROB, typename T>
__global__ void __launch_bounds__(kTmaMaxWarps * 32, 1)
bvsb_stats_tma_kernel(const __grid_constant__ TmaMaps maps, const StatsParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);   // TMA destinations: 128-byte aligned
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int threads = blockDim.x, warps = threads >> 5;
    const int C = EXACT ? CMAX : p.C;
    const int stages = p.stages;
    const uint32_t plane_bytes = kTmaStripPx * sizeof(T);
    const uint32_t stage_bytes = (uint32_t)C * plane_bytes + kTmaStripPx * sizeof(int32_t);
    unsigned char* my_stages = smem + (size_t)warp * stages * stage_bytes;
    uint2* columns = reinterpret_cast<uint2*>(smem + (size_t)warps * stages * stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(columns + (size_t)C * threads) + warp * stages;

    Walker<CMAX, EXACT, 4, NEED_PROB> w;
    w.init(columns + tid, threads, p);

    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(smem_u32(bars + s), 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    long long r0, r1;
    mas::warp_range(p.total_rows, (long long)blockIdx.x * warps + warp, (long long)gridDim.x * warps, r0, r1);
    if (r0 >= r1) return;

    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

    mas::Cursor at, ahead;      // consume / issue positions
    at.seek(r0, p.strips, p.H);
    ahead = at;
    long long issued = r0;
    int ahead_seg = seg_of(p, ahead.img);       // segment of the image being issued (lane 0 only uses it)
    auto issue = [&](int s) {  // lane 0 only
        const uint32_t bar = smem_u32(bars + s);
        const uint32_t dst = smem_u32(my_stages + (size_t)s * stage_bytes);
        const int local = ahead.img - p.seg_first[ahead_seg];
        mbar_expect_tx(bar, stage_bytes);
        tma_load_4d(dst, &maps.logits[ahead_seg], bar, ahead.strip * kTmaStripPx, ahead.y, 0, local, policy);
        tma_load_3d(dst + (uint32_t)C * plane_bytes, &maps.ids[ahead_seg], bar, ahead.strip * kTmaStripPx, ahead.y, local, policy);
        if (ahead.advance(p.strips, p.H) == 2 && ahead_seg + 1 < p.n_seg && ahead.img >= p.seg_first[ahead_seg + 1]) ++ahead_seg;
        ++issued;
    };
    if (lane == 0) {
        for (int s = 0; s < stages && issued < r1; ++s) issue(s);
    }

    w.img_region = (long long)at.img * p.S;
    bool active = (at.strip * kTmaStripPx + lane * 4) < p.W;
    int s = 0;
    uint32_t parity = 0;
    for (long long r = r0; r < r1; ++r) {
        mbar_wait(smem_u32(bars + s), parity);
        if (active) {
            const unsigned char* st = my_stages + (size_t)s * stage_bytes;
            int id[4];
            {
                const int4 q = *reinterpret_cast<const int4*>(st + (size_t)C * plane_bytes + lane * 16);
                id[0] = q.x; id[1] = q.y; id[2] = q.z; id[3] = q.w;
            }
            float v[CMAX][4];
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                if (EXACT || c < C) {
                    VecLoad<T, 4>::shared(reinterpret_cast<const T*>(st + (size_t)c * plane_bytes) + lane * 4, v[c]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[c][j] = -INFINITY;
                }
            }
            w.row(v, id);
        }
        __syncwarp();   // every lane is done with stage s (its loads fed the column updates above)
        if (lane == 0 && issued < r1) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(s);
        }
        if (++s == stages) { s = 0; parity ^= 1u; }

        const int img_done = at.img;
        const int step = at.advance(p.strips, p.H);
        if (step != 0) {
            active = (at.strip * kTmaStripPx + lane * 4) < p.W;
            if (step == 2) {
                w.flush();
                w.flush_prob(p.prob_sum, img_done, lane);
                w.img_region += p.S;
            }
        }
    }
    w.flush();
    if (at.img < p.n_img && (at.strip != 0 || at.y != 0)) w.flush_prob(p.prob_sum, at.img, lane);
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
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

struct LaunchShape {
    int blocks, warps, stages;
    size_t smem;
};

template <typename K>
int resident_blocks(K kernel, int threads, size_t smem) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) n = 1;
    return n;
}

int env_int(const char* name, int fallback) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : fallback;
}

template <int CMAX, bool EXACT, int VEC, bool NEED_PROB, typename T>
cudaError_t launch_ldg(StatsParams p, cudaStream_t stream) {
    auto kernel = bvsb_stats_ldg_kernel<CMAX, EXACT, VEC, NEED_PROB, T>;
    const size_t smem = (size_t)p.C * kLdgThreads * sizeof(uint2);
    static mas::PerDeviceInt occ;   // per instantiation and device; smem differs by at most the generic channel padding
    const int dev = mas::current_device();
    int per_sm = occ.get(dev);
    if (per_sm == 0) {
        per_sm = resident_blocks(kernel, kLdgThreads, (size_t)CMAX * kLdgThreads * sizeof(uint2));
        occ.set(dev, per_sm);
    }
    p.strips = (p.W + 32 * VEC - 1) / (32 * VEC);
    p.total_rows = (long long)p.n_img * p.strips * p.H;
    // one wave of resident CTAs; never fewer than ~8 rows per warp
    const long long cap = (p.total_rows + 8 * (kLdgThreads / 32) - 1) / (8 * (kLdgThreads / 32));
    const long long blocks = std::max<long long>(1, std::min<long long>((long long)mas::sm_count() * per_sm, cap));
    kernel<<<(unsigned)blocks, kLdgThreads, smem, stream>>>(p);
    mas::count_launches(1);
    return cudaGetLastError();
}

template <int CMAX, bool EXACT, bool NEED_PROB, typename T>
cudaError_t launch_tma(StatsParams p, cudaStream_t stream, bool* unsupported) {
    *unsupported = true;
    EncodeTiledFn encode = encode_tiled();
    if (!encode) return cudaSuccess;
    const size_t elt = sizeof(T);
    const uint32_t stage_bytes = (uint32_t)p.C * kTmaStripPx * elt + kTmaStripPx * 4;
    const size_t col_bytes_per_warp = (size_t)p.C * 32 * sizeof(uint2);
    const int dev = mas::current_device();
    int max_smem = 0;
    if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return cudaSuccess;
    int stages = env_int("MAS_SCORER_STAGES", 0);
    int warps = env_int("MAS_SCORER_WARPS", 0);
    if (stages <= 0) stages = (elt == 2) ? 4 : 2;
    stages = std::min(std::max(stages, 2), 8);
    const size_t per_warp = (size_t)stages * stage_bytes + col_bytes_per_warp + (size_t)stages * 8;
    const int fit = (int)(((size_t)max_smem - 128) / per_warp);
    if (warps <= 0) {
        // default: launches of up to ~1 GB run as small CTAs (2 warps, several per SM) so that the CTAs of the next launch
        // on the other lane move in warp-pair by warp-pair as this one drains (+4-5 % at 4 Cityscapes images per launch);
        // bigger launches amortise the hand-over anyway and are a little faster as one 8-warp CTA per SM
        const double launch_bytes = (double)p.n_img * p.H * p.W * ((double)p.C * elt + 4.0);
        warps = launch_bytes <= 1073741824.0 ? 2 : fit;
    }
    if (warps > fit) warps = fit;
    warps = std::min(warps, kTmaMaxWarps);
    if (warps < 1) return cudaSuccess;
    const size_t smem = (size_t)warps * per_warp + 128;

    p.strips = (p.W + kTmaStripPx - 1) / kTmaStripPx;
    p.total_rows = (long long)p.n_img * p.strips * p.H;
    p.stages = stages;

    TmaMaps maps;
    for (int g = 0; g < p.n_seg; ++g) {
        const cuuint64_t n_seg_img = (cuuint64_t)(p.seg_first[g + 1] - p.seg_first[g]);
        {
            const cuuint64_t dims[4] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.C, n_seg_img};
            const cuuint64_t strides[3] = {(cuuint64_t)p.W * elt, (cuuint64_t)p.H * p.W * elt, (cuuint64_t)p.seg_stride[g] * elt};
            const cuuint32_t box[4] = {kTmaStripPx, 1, (cuuint32_t)p.C, 1};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            if (encode(&maps.logits[g], elt == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                       const_cast<void*>(p.seg_logits[g]), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return cudaSuccess;
        }
        {
            const cuuint64_t dims[3] = {(cuuint64_t)p.W, (cuuint64_t)p.H, n_seg_img};
            const cuuint64_t strides[2] = {(cuuint64_t)p.W * 4, (cuuint64_t)p.H * p.W * 4};
            const cuuint32_t box[3] = {kTmaStripPx, 1, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            if (encode(&maps.ids[g], CU_TENSOR_MAP_DATA_TYPE_INT32, 3, const_cast<int32_t*>(p.seg_ids[g]), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return cudaSuccess;
        }
    }
    for (int g = p.n_seg; g < kMaxSeg; ++g) { maps.logits[g] = maps.logits[0]; maps.ids[g] = maps.ids[0]; }
    *unsupported = false;

    auto kernel = bvsb_stats_tma_kernel<CMAX, EXACT, NEED_PROB, T>;
    static mas::PerDeviceInt configured;      // per instantiation AND device (the attribute is a per-device setting)
    {
        cudaError_t e = mas::opt_in_smem(kernel, configured, (int)max_smem);
        if (e != cudaSuccess) return e;
    }
    // one wave of resident CTAs: with fewer warps per CTA several CTAs share an SM, and a CTA of the NEXT launch can move
    // in as soon as one of them retires (finer hand-over between consecutive launches on the two launch lanes).
    // The occupancy query is not free on the launch path: cached per (device, warps) -- smem is a function of warps,
    // stages and C, of which only `warps` varies between launches of one instantiation unless the env switches change.
    static mas::PerDeviceInt occ[kTmaMaxWarps + 1], occ_key[kTmaMaxWarps + 1];
    int per_sm = occ[warps].get(dev);
    if (per_sm == 0 || occ_key[warps].get(dev) != (int)smem) {
        per_sm = resident_blocks(kernel, warps * 32, smem);
        occ[warps].set(dev, per_sm); occ_key[warps].set(dev, (int)smem);
    }
    const long long cap = (p.total_rows + 8 * warps - 1) / (8 * warps);
    const long long blocks = std::max<long long>(1, std::min<long long>((long long)mas::sm_count() * per_sm, cap));
    kernel<<<(unsigned)blocks, warps * 32, smem, stream>>>(maps, p);
    mas::count_launches(1);
    return cudaGetLastError();
}

enum Path { kPathLdg1 = 0, kPathLdg4 = 1, kPathTma = 2 };

template <int CMAX, bool EXACT, bool NEED_PROB, typename T>
cudaError_t launch_path(const StatsParams& p, int path, cudaStream_t stream) {
    if (path == kPathTma) {
        bool unsupported = false;
        cudaError_t e = launch_tma<CMAX, EXACT, NEED_PROB, T>(p, stream, &unsupported);
        if (!unsupported) return e;
        path = kPathLdg4;   // driver without tensor-map support / shared memory too small for one warp
    }
    if (path == kPathLdg4) return launch_ldg<CMAX, EXACT, 4, NEED_PROB, T>(p, stream);
    return launch_ldg<CMAX, EXACT, 1, NEED_PROB, T>(p, stream);
}

template <bool NEED_PROB, typename T>
cudaError_t dispatch_channels(const StatsParams& p, int path, cudaStream_t stream) {
    switch (p.C) {
        case 19: return launch_path<19, true, NEED_PROB, T>(p, path, stream);
        case 20: return launch_path<20, true, NEED_PROB, T>(p, path, stream);
        case 21: return launch_path<21, true, NEED_PROB, T>(p, path, stream);
        case 22: return launch_path<22, true, NEED_PROB, T>(p, path, stream);
        default: break;
    }
    if (p.C <= 8) return launch_path<8, false, NEED_PROB, T>(p, path, stream);
    if (p.C <= 16) return launch_path<16, false, NEED_PROB, T>(p, path, stream);
    if (p.C <= 24) return launch_path<24, false, NEED_PROB, T>(p, path, stream);
    return launch_path<32, false, NEED_PROB, T>(p, path, stream);
}

}  // namespace

extern "C" int mas_bvsb_segment_stats_multi_dev(int n_segments, const void* const* logits, int logits_dtype,
                                                const int64_t* image_strides, const int32_t* const* ids, const int* n_img_per_segment,
                                                int channels, int height, int width, int nseg, float temperature,
                                                float* cls_sum, int32_t* cls_cnt, double* prob_sum, void* stream) {
    MAS_REQUIRE(logits && ids && n_img_per_segment && cls_sum && cls_cnt, MAS_E_BADARG, "bvsb_segment_stats: null pointer");
    MAS_REQUIRE(n_segments >= 0 && n_segments <= kMaxSeg, MAS_E_RANGE, "bvsb_segment_stats: n_segments=%d outside [0,%d]", n_segments, kMaxSeg);
    MAS_REQUIRE(height > 0 && width > 0 && nseg > 0, MAS_E_BADARG, "bvsb_segment_stats: bad shape");
    MAS_REQUIRE(channels >= 2 && channels <= MAS_MAX_CLASSES, MAS_E_RANGE,
                "bvsb_segment_stats: channels=%d outside [2,%d]", channels, MAS_MAX_CLASSES);
    MAS_REQUIRE(temperature > 0.f, MAS_E_BADARG, "bvsb_segment_stats: temperature must be > 0");
    MAS_REQUIRE(logits_dtype == MAS_F32 || logits_dtype == MAS_BF16, MAS_E_BADARG, "bvsb_segment_stats: bad dtype");
    const long long plane = (long long)height * width;
    const size_t elt = logits_dtype == MAS_F32 ? 4 : 2;

    StatsParams p;
    p.n_seg = 0; p.n_img = 0;
    p.seg_first[0] = 0;
    // 128-bit (f32) / 64-bit (bf16) row segments need every plane row to start VEC-aligned;
    // TMA additionally needs 16-byte global strides and base addresses -- for EVERY segment of the launch
    bool vec4 = (width % 4 == 0), tma_ok = ((width * elt) % 16 == 0);
    for (int g = 0; g < n_segments; ++g) {
        MAS_REQUIRE(n_img_per_segment[g] >= 0, MAS_E_BADARG, "bvsb_segment_stats: negative image count");
        if (n_img_per_segment[g] == 0) continue;
        MAS_REQUIRE(logits[g] && ids[g], MAS_E_BADARG, "bvsb_segment_stats: null segment pointer");
        long long stride = image_strides ? image_strides[g] : 0;
        if (stride == 0) stride = (long long)channels * plane;
        MAS_REQUIRE(stride >= (long long)channels * plane, MAS_E_BADARG, "bvsb_segment_stats: image_stride too small");
        const int k = p.n_seg++;
        p.seg_logits[k] = logits[g]; p.seg_ids[k] = ids[g]; p.seg_stride[k] = stride;
        p.n_img += n_img_per_segment[g];
        p.seg_first[k + 1] = p.n_img;
        vec4 = vec4 && (stride % 4 == 0) && (((uintptr_t)logits[g]) % (4 * elt) == 0) && (((uintptr_t)ids[g]) % 16 == 0);
        tma_ok = tma_ok && ((stride * elt) % 16 == 0) && (((uintptr_t)logits[g]) % 16 == 0);
    }
    if (p.n_img == 0) return 0;
    for (int g = p.n_seg; g < kMaxSeg; ++g) {
        p.seg_logits[g] = p.seg_logits[0]; p.seg_ids[g] = p.seg_ids[0]; p.seg_stride[g] = p.seg_stride[0];
        p.seg_first[g + 1] = p.n_img;
    }
    MAS_REQUIRE((long long)p.n_img * ((width + 31) / 32) * height < (1ll << 40), MAS_E_RANGE, "bvsb_segment_stats: too many rows");
    tma_ok = tma_ok && vec4;
    int path = tma_ok ? kPathTma : (vec4 ? kPathLdg4 : kPathLdg1);
    const char* forced = getenv("MAS_SCORER_PATH");   // development switch: "ldg" keeps the register path
    if (forced && forced[0] == 'l' && path == kPathTma) path = kPathLdg4;

    p.C = channels; p.H = height; p.W = width; p.S = nseg;
    p.scale = 1.4426950408889634f / temperature;
    p.strips = 0; p.total_rows = 0; p.stages = 0;
    p.cls_sum = cls_sum; p.cls_cnt = cls_cnt; p.prob_sum = prob_sum;

    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (logits_dtype == MAS_F32)
        e = prob_sum ? dispatch_channels<true, float>(p, path, st) : dispatch_channels<false, float>(p, path, st);
    else
        e = prob_sum ? dispatch_channels<true, __nv_bfloat16>(p, path, st) : dispatch_channels<false, __nv_bfloat16>(p, path, st);
    if (e != cudaSuccess) return mas::cuda_fail(e, "bvsb_stats kernel launch");
    return 0;
}

extern "C" int mas_bvsb_segment_stats_dev(const void* logits, int logits_dtype, int64_t image_stride, const int32_t* ids,
                                          int n_img, int channels, int height, int width, int nseg,
                                          float temperature, float* cls_sum, int32_t* cls_cnt, double* prob_sum,
                                          void* stream) {
    MAS_REQUIRE(logits && ids && cls_sum && cls_cnt, MAS_E_BADARG, "bvsb_segment_stats: null pointer");
    MAS_REQUIRE(n_img >= 0, MAS_E_BADARG, "bvsb_segment_stats: bad shape");
    return mas_bvsb_segment_stats_multi_dev(1, &logits, logits_dtype, &image_stride, &ids, &n_img, channels, height, width, nseg,
                                            temperature, cls_sum, cls_cnt, prob_sum, stream);
}
