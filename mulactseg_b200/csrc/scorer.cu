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
#include "scorer.cuh"

#include <cuda.h>  // CUtensorMap and enums only; the encoder is resolved through cudaGetDriverEntryPoint
#include <stdlib.h>

#include <algorithm>
#include <mutex>

using namespace mas_scorer;

namespace {

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
    constexpr uint32_t kIdBytes = kTmaStripPx * sizeof(int32_t);
    // p.split_ids: the id rows do not ride in the logits stages but in ONE buffer per warp with its own barrier (it is
    // read into registers at the top of a row and refilled at once) -- 512 bytes per warp and stage less, which is what
    // lets an eighth warp fit at C' = 21 / 22 fp32
    const bool split = p.split_ids != 0;
    const uint32_t stage_bytes = (uint32_t)C * plane_bytes + (split ? 0u : kIdBytes);
    unsigned char* my_stages = smem + (size_t)warp * stages * stage_bytes;
    unsigned char* my_ids = smem + (size_t)warps * stages * stage_bytes + (size_t)warp * kIdBytes;      // split mode only
    uint2* columns = reinterpret_cast<uint2*>(smem + (size_t)warps * stages * stage_bytes + (split ? (size_t)warps * kIdBytes : 0));
    uint64_t* bars = reinterpret_cast<uint64_t*>(columns + (size_t)C * threads) + warp * (stages + 1);      // [stages] logits, then ids

    Walker<CMAX, EXACT, 4, NEED_PROB> w;
    w.init(columns + tid, threads, p);

    if (lane == 0) {
        for (int s = 0; s <= stages; ++s) mbar_init(smem_u32(bars + s), 1u);
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
        if (!split) tma_load_3d(dst + (uint32_t)C * plane_bytes, &maps.ids[ahead_seg], bar, ahead.strip * kTmaStripPx, ahead.y, local, policy);
        if (ahead.advance(p.strips, p.H) == 2 && ahead_seg + 1 < p.n_seg && ahead.img >= p.seg_first[ahead_seg + 1]) ++ahead_seg;
        ++issued;
    };
    // split mode: the id row of the next strip row -> the warp's single id buffer
    mas::Cursor ahead_id = at;
    long long issued_id = r0;
    int ahead_id_seg = ahead_seg;
    auto issue_ids = [&]() {  // lane 0 only
        const uint32_t bar = smem_u32(bars + stages);
        mbar_expect_tx(bar, kIdBytes);
        tma_load_3d(smem_u32(my_ids), &maps.ids[ahead_id_seg], bar, ahead_id.strip * kTmaStripPx, ahead_id.y,
                    ahead_id.img - p.seg_first[ahead_id_seg], policy);
        if (ahead_id.advance(p.strips, p.H) == 2 && ahead_id_seg + 1 < p.n_seg && ahead_id.img >= p.seg_first[ahead_id_seg + 1]) ++ahead_id_seg;
        ++issued_id;
    };
    if (lane == 0) {
        if (split) issue_ids();
        for (int s = 0; s < stages && issued < r1; ++s) issue(s);
    }

    w.img_region = (long long)at.img * p.S;
    bool active = (at.strip * kTmaStripPx + lane * 4) < p.W;
    int s = 0;
    uint32_t parity = 0;
    uint32_t parity_id = 0;
    for (long long r = r0; r < r1; ++r) {
        int id[4];
        if (split) {
            // ids first: into registers, and the buffer is refilled with the next row's ids while this row is computed
            mbar_wait(smem_u32(bars + stages), parity_id);
            parity_id ^= 1u;
            const int4 q = *reinterpret_cast<const int4*>(my_ids + lane * 16);
            id[0] = q.x; id[1] = q.y; id[2] = q.z; id[3] = q.w;
            __syncwarp();
            if (lane == 0 && issued_id < r1) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue_ids();
            }
        }
        mbar_wait(smem_u32(bars + s), parity);
        if (active) {
            const unsigned char* st = my_stages + (size_t)s * stage_bytes;
            if (!split) {
                const int4 q = *reinterpret_cast<const int4*>(st + (size_t)C * plane_bytes + lane * 16);
                id[0] = q.x; id[1] = q.y; id[2] = q.z; id[3] = q.w;
            }
            if (sizeof(T) == 2) {
                // bf16: the planes stay packed (two pixels per register) for the top-2 scan, see Walker::row_bf16
                uint2 q[CMAX];
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
                    q[c] = (EXACT || c < C) ? *reinterpret_cast<const uint2*>(st + (size_t)c * plane_bytes + lane * 8)
                                            : make_uint2(0xff80ff80u, 0xff80ff80u);
                }
                w.row_bf16(q, id);
            } else {
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
    uint32_t stage_bytes = (uint32_t)p.C * kTmaStripPx * elt + kTmaStripPx * 4;
    const size_t col_bytes_per_warp = (size_t)p.C * 32 * sizeof(uint2);
    const int dev = mas::current_device();
    int max_smem = 0;
    if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return cudaSuccess;
    int stages = env_int("MAS_SCORER_STAGES", 0);
    int warps = env_int("MAS_SCORER_WARPS", 0);
    if (stages <= 0) stages = (elt == 2) ? 4 : 2;
    stages = std::min(std::max(stages, 2), 8);
    size_t per_warp = (size_t)stages * stage_bytes + col_bytes_per_warp + (size_t)(stages + 1) * 8;
    int fit = (int)(((size_t)max_smem - 128) / per_warp);
    p.split_ids = 0;
    if (fit < kTmaMaxWarps) {
        // one id buffer per warp instead of one per stage: does that buy a warp?  (C' = 21 / 22 fp32: 7 -> 8 warps)
        const uint32_t split_stage = (uint32_t)p.C * kTmaStripPx * elt;
        const size_t split_per_warp = (size_t)stages * split_stage + kTmaStripPx * 4 + col_bytes_per_warp + (size_t)(stages + 1) * 8;
        const int split_fit = (int)(((size_t)max_smem - 128) / split_per_warp);
        if (split_fit > fit && env_int("MAS_SCORER_SPLIT_IDS", 1) != 0) {
            p.split_ids = 1; stage_bytes = split_stage; per_warp = split_per_warp; fit = split_fit;
        }
    }
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

enum Path { kPathLdg1 = 0, kPathLdg4 = 1, kPathTma = 2, kPathAbreast1 = 3, kPathAbreast4 = 4 };

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
    // default: TMA ring when every row is 16-byte aligned, else the abreast kernel (adjacent strips per CTA + L1 prefetch)
    int path = tma_ok ? kPathTma : (vec4 ? kPathAbreast4 : kPathAbreast1);
    const char* forced = getenv("MAS_SCORER_PATH");   // development / test switch: "ldg" = plain register path, "abreast"
    if (forced && forced[0] == 'l') path = vec4 ? kPathLdg4 : kPathLdg1;
    if (forced && forced[0] == 'a') path = vec4 ? kPathAbreast4 : kPathAbreast1;

    p.C = channels; p.H = height; p.W = width; p.S = nseg;
    p.scale = 1.4426950408889634f / temperature;
    p.strips = 0; p.total_rows = 0; p.stages = 0; p.split_ids = 0;
    p.cls_sum = cls_sum; p.cls_cnt = cls_cnt; p.prob_sum = prob_sum;
    p.h_in = 0; p.w_in = 0; p.ry = 1.f; p.rx = 1.f;

    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (path == kPathAbreast1 || path == kPathAbreast4)
        e = launch_abreast(p, path == kPathAbreast4 ? 4 : 1, false, logits_dtype, st);
    else if (logits_dtype == MAS_F32)
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

extern "C" int mas_bvsb_segment_stats_lowres_dev(const void* logits, int logits_dtype, int64_t image_stride, int height_in, int width_in,
                                                 const int32_t* ids, int n_img, int channels, int height, int width, int nseg,
                                                 float temperature, float* cls_sum, int32_t* cls_cnt, double* prob_sum, void* stream) {
    MAS_REQUIRE(logits && ids && cls_sum && cls_cnt, MAS_E_BADARG, "bvsb_segment_stats_lowres: null pointer");
    MAS_REQUIRE(n_img >= 0 && height > 0 && width > 0 && nseg > 0 && height_in > 0 && width_in > 0, MAS_E_BADARG,
                "bvsb_segment_stats_lowres: bad shape");
    MAS_REQUIRE(height_in <= height && width_in <= width, MAS_E_BADARG, "bvsb_segment_stats_lowres: the source must not be larger than the id map");
    MAS_REQUIRE(channels >= 2 && channels <= MAS_MAX_CLASSES, MAS_E_RANGE, "bvsb_segment_stats_lowres: channels=%d outside [2,%d]", channels,
                MAS_MAX_CLASSES);
    MAS_REQUIRE(temperature > 0.f, MAS_E_BADARG, "bvsb_segment_stats_lowres: temperature must be > 0");
    MAS_REQUIRE(logits_dtype == MAS_F32 || logits_dtype == MAS_BF16, MAS_E_BADARG, "bvsb_segment_stats_lowres: bad dtype");
    if (n_img == 0) return 0;
    const long long plane_in = (long long)height_in * width_in;
    if (image_stride == 0) image_stride = (long long)channels * plane_in;
    MAS_REQUIRE(image_stride >= (long long)channels * plane_in, MAS_E_BADARG, "bvsb_segment_stats_lowres: image_stride too small");
    MAS_REQUIRE((long long)n_img * ((width + 31) / 32) * height < (1ll << 40), MAS_E_RANGE, "bvsb_segment_stats_lowres: too many rows");
    StatsParams p;
    p.n_seg = 1; p.n_img = n_img;
    for (int g = 0; g < kMaxSeg; ++g) { p.seg_logits[g] = logits; p.seg_ids[g] = ids; p.seg_stride[g] = image_stride; p.seg_first[g + 1] = n_img; }
    p.seg_first[0] = 0;
    p.C = channels; p.H = height; p.W = width; p.S = nseg;
    p.scale = 1.4426950408889634f / temperature;
    p.strips = 0; p.total_rows = 0; p.stages = 0; p.split_ids = 0;
    p.cls_sum = cls_sum; p.cls_cnt = cls_cnt; p.prob_sum = prob_sum;
    p.h_in = height_in; p.w_in = width_in;
    // torch's area_pixel_compute_scale<float>(input, output, align_corners=false, no scale factor): (float)input / output
    p.ry = (float)height_in / (float)height; p.rx = (float)width_in / (float)width;
    const bool vec4 = (width % 4 == 0) && (((uintptr_t)ids) % 16 == 0);       // only the id map is read with 128-bit loads
    cudaError_t e = launch_abreast(p, vec4 ? 4 : 1, true, logits_dtype, (cudaStream_t)stream);
    if (e != cudaSuccess) return mas::cuda_fail(e, "bvsb_stats lowres kernel launch");
    return 0;
}
