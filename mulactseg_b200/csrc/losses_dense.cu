// Stage-1 losses, DENSE regime: TMA-staged strip walk for batches in which most tiles hold selected pixels.
//
// Same arithmetic as multihot_loss_{fwd,bwd}_kernel (losses.cu; reference: utils/loss.py:81-141, :535-588,
// trainer/active_joint_multi_predignore_lossdecomp.py:16-72), different data path.  The tile walk of losses.cu loads one
// pixel per thread and C' strided planes per row with predicated 32-bit loads -- right when a few percent of the pixels
// are labelled (it never touches the logits of unselected pixels), but issue-bound at ~30 instructions per (class, pixel)
// once most pixels are selected (0.5 of the HBM peak at rho = 1).  Here, like the acquisition scorer (scorer.cu):
//   * the unit of work is a STRIP ROW of 64 pixels (2 per lane; 16 warps per SM); strip rows are linearised (image, strip, y)
//     and cut into one contiguous range per warp of a single-wave persistent grid, so a lane walks DOWN its two columns;
//   * every warp runs its own ring of shared-memory stages; lane 0 issues one cp.async.bulk.tensor box
//     {64 px, 1 row, C' planes} for the logits plus one each for the ids and the mask bytes of the row, completion on a
//     per-stage mbarrier; the planes are then read with vector shared-memory loads at constant offsets (no per-plane
//     address arithmetic, no predicated global loads);
//   * forward: every pixel COLUMN of a lane keeps the running maxima of the superpixel it is walking through as packed
//     64-bit keys in registers, indexed by the RANK of the class inside the superpixel's candidate set (3 slots; further
//     candidates go to global atomicMax directly), flushed with atomicMax when the column enters another superpixel;
//   * backward: the gradient of the row overwrites the logits in the stage and leaves with ONE TMA store (2-stage ring,
//     the drained stage refilled early in the next iteration), so the dense gradient is written as full lines without
//     20 address computations per row either; the arg-max ownership words of a pixel's first three candidates are
//     gathered before the softmax is computed (one latency, not one per candidate).
// Out-of-range boxes are zero-filled on load (mask 0 = unselected) and clipped on store, so ragged right edges need no
// special case.  Requirements: W % 16 == 0 (the mask rows are the narrowest TMA source) and 16-byte aligned bases;
// anything else stays on the tile walk.
#include "losses.cuh"
#include "tma.cuh"
#include "walk.cuh"

#include <stdlib.h>

#include <algorithm>

using namespace mas_loss;
using namespace mas_tma;

namespace {

#ifndef MAS_DENSE_PX
#define MAS_DENSE_PX 2
#endif
constexpr int kPx = MAS_DENSE_PX;       // pixels per lane (2: 64-pixel strip rows, twice the warps per SM; 4: 128-pixel strip rows)
constexpr int kStripPx = 32 * kPx;
#ifdef MAS_DENSE_MAXWARPS
constexpr int kMaxWarps = MAS_DENSE_MAXWARPS;      // comparison builds: fewer warps, more registers per thread
#else
constexpr int kMaxWarps = kPx == 2 ? 16 : 8;
#endif
constexpr int kPrivRanks = 3;           // private running maxima: the first three candidate classes of a superpixel
constexpr uint32_t kFull = 0xffffffffu;

struct DenseMaps {
    CUtensorMap logits, ids, mask, grad;
};

struct DenseShape {
    int strips;               // strips (kStripPx pixels wide) per image row
    long long total_rows;     // n_img * strips * H
    int stages;
    uint32_t stage_bytes, tx_bytes, ids_off, mask_off;      // stage stride (128-byte multiple) / bytes one row's three boxes deliver
    int debug;                // development switches (MAS_LOSS_DENSE_DEBUG): 1 = no ids / mask loads (streaming floor), 2 = no L2 hint
};

__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__device__ __forceinline__ unsigned char* align128(unsigned char* p) { return p + ((128u - (smem_u32(p) & 127u)) & 127u); }

// mask bytes of the lane's pixels (byte j <-> pixel j)
__device__ __forceinline__ uint32_t row_mask(const unsigned char* st, const DenseShape& d, int lane) {
    if (kPx == 4) return *reinterpret_cast<const uint32_t*>(st + d.mask_off + lane * 4);
    return *reinterpret_cast<const unsigned short*>(st + d.mask_off + lane * 2);
}

// ids of the lane's pixels: -1 = not selected by the mask / outside [0, S)
template <typename IdT>
__device__ __forceinline__ void row_ids(const unsigned char* st, const DenseShape& d, int lane, uint32_t mbytes, int S, int (&sid)[kPx]) {
    long long raw[4];
    const unsigned char* src = st + d.ids_off + lane * (kPx * (int)sizeof(IdT));
    if (sizeof(IdT) == 8) {
        const longlong2 a = *reinterpret_cast<const longlong2*>(src);
        raw[0] = a.x; raw[1] = a.y;
        if (kPx == 4) {
            const longlong2 b = *reinterpret_cast<const longlong2*>(src + 16);
            raw[2] = b.x; raw[3] = b.y;
        }
    } else if (kPx == 4) {
        const int4 a = *reinterpret_cast<const int4*>(src);
        raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w;
    } else {
        const int2 a = *reinterpret_cast<const int2*>(src);
        raw[0] = a.x; raw[1] = a.y;
    }
#pragma unroll
    for (int j = 0; j < kPx; ++j) {
        const bool on = ((mbytes >> (8 * j)) & 0xffu) != 0u && (unsigned long long)raw[j] < (unsigned long long)S;
        sid[j] = on ? (int)raw[j] : -1;
    }
}

// one plane of the row: the lane's kPx consecutive values
__device__ __forceinline__ void load_px(const float* src, float (&o)[kPx]) {
    if (kPx == 4) {
        const float4 q = *reinterpret_cast<const float4*>(src);
        o[0] = q.x; o[1] = q.y; o[2 % kPx] = q.z; o[3 % kPx] = q.w;
    } else {
        const float2 q = *reinterpret_cast<const float2*>(src);
        o[0] = q.x; o[1] = q.y;
    }
}
__device__ __forceinline__ void store_px(float* dst, const float (&o)[kPx]) {
    if (kPx == 4) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2 % kPx], o[3 % kPx]);
    else *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
}

// the row's logits -> registers (raw)
template <int CMAX, bool EXACT>
__device__ __forceinline__ void row_logits(const float* sx, int C, float (&v)[CMAX][kPx]) {
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (EXACT || c < C) {
            load_px(sx + c * kStripPx, v[c]);
        } else {
#pragma unroll
            for (int j = 0; j < kPx; ++j) v[c][j] = -INFINITY;
        }
    }
}

// softmax(x / T) numerators (ex2.approx((x - max) * log2e / T), the fast form of losses.cu; KEEP: left in v) and 1 / sum
template <int CMAX, bool KEEP>
__device__ __forceinline__ void row_softmax(float (&v)[CMAX][kPx], float scale, float (&shift)[kPx], float (&inv)[kPx]) {
    float mx[kPx], sa[kPx], sb[kPx];
#pragma unroll
    for (int j = 0; j < kPx; ++j) { mx[j] = v[0][j]; sa[j] = 0.f; sb[j] = 0.f; }
#pragma unroll
    for (int c = 1; c < CMAX; c += 2) {
#pragma unroll
        for (int j = 0; j < kPx; ++j) mx[j] = (c + 1 < CMAX) ? max3(mx[j], v[c][j], v[c + 1 < CMAX ? c + 1 : c][j]) : fmaxf(mx[j], v[c][j]);
    }
#pragma unroll
    for (int j = 0; j < kPx; ++j) shift[j] = -mx[j] * scale;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
#pragma unroll
        for (int j = 0; j < kPx; ++j) {
            const float e = mas::ex2_approx(fmaf(v[c][j], scale, shift[j]));
            if (KEEP) v[c][j] = e;
            if (c & 1) sb[j] += e; else sa[j] += e;
        }
    }
#pragma unroll
    for (int j = 0; j < kPx; ++j) inv[j] = mas::rcp_approx(sa[j] + sb[j]);
}

struct Ring {
    unsigned char* stages;
    uint64_t* bars;
    uint32_t stage_bytes;
    __device__ __forceinline__ unsigned char* stage(int s) const { return stages + (size_t)s * stage_bytes; }
    __device__ __forceinline__ uint32_t bar(int s) const { return smem_u32(bars + s); }
};

// ------------------------------------------------------------------------------------------ forward
template <int CMAX, bool EXACT, typename IdT>
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
multihot_dense_fwd_kernel(const __grid_constant__ DenseMaps maps, const LossParams p, const DenseShape d) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (p.list_mode && p.dense_percent > 0 && !dense_regime(p)) return;      // sparsely selected: the list walk takes it
    unsigned char* smem = align128(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warps = blockDim.x >> 5;
    const int C = EXACT ? CMAX : p.C;
    const int stages = d.stages;
    Ring ring;
    ring.stage_bytes = d.stage_bytes;
    ring.stages = smem + (size_t)warp * stages * d.stage_bytes;
    ring.bars = reinterpret_cast<uint64_t*>(smem + (size_t)warps * stages * d.stage_bytes) + warp * stages;
    const bool do_choice = p.do_choice != 0, do_group = p.do_group != 0;

    float sum_one = 0.f, sum_multi = 0.f, sum_empty = 0.f;
    int n_one = 0, n_multi = 0, n_empty = 0;

    long long r0, r1;
    mas::warp_range(d.total_rows, (long long)blockIdx.x * warps + warp, (long long)gridDim.x * warps, r0, r1);
    if (r0 < r1) {
        if (lane == 0) {
            for (int s = 0; s < stages; ++s) mbar_init(ring.bar(s), 1u);
            fence_barrier_init();
        }
        __syncwarp();

        const uint64_t policy = policy_evict_first();
        mas::Cursor at, ahead;
        at.seek(r0, d.strips, p.H);
        ahead = at;
        long long issued = r0;
        auto issue = [&](int s) {      // lane 0
            const uint32_t bar = ring.bar(s), dst = smem_u32(ring.stage(s));
            if (d.debug & 1) {
                mbar_expect_tx(bar, d.ids_off);
                load_4d(dst, &maps.logits, bar, ahead.strip * kStripPx, ahead.y, 0, ahead.img, policy);
            } else if (d.debug & 2) {
                mbar_expect_tx(bar, d.tx_bytes);
                load_4d(dst, &maps.logits, bar, ahead.strip * kStripPx, ahead.y, 0, ahead.img);
                load_3d(dst + d.ids_off, &maps.ids, bar, ahead.strip * kStripPx, ahead.y, ahead.img);
                load_3d(dst + d.mask_off, &maps.mask, bar, ahead.strip * kStripPx, ahead.y, ahead.img);
            } else {
                mbar_expect_tx(bar, d.tx_bytes);
                load_4d(dst, &maps.logits, bar, ahead.strip * kStripPx, ahead.y, 0, ahead.img, policy);
                load_3d(dst + d.ids_off, &maps.ids, bar, ahead.strip * kStripPx, ahead.y, ahead.img, policy);
                load_3d(dst + d.mask_off, &maps.mask, bar, ahead.strip * kStripPx, ahead.y, ahead.img, policy);
            }
            ahead.advance(d.strips, p.H);
            ++issued;
        };
        if (lane == 0) {
            for (int s = 0; s < stages && issued < r1; ++s) issue(s);
        }

        // running maxima of the superpixel each of the lane's pixel COLUMNS is walking through: one packed (P bits, ~pixel) key
        // per RANK of the class inside the superpixel's candidate set, held in registers (column and rank are compile-time
        // indices below); no pixel of a row needs a global atomic, only a column that enters another superpixel flushes
        int cur[kPx], cur_base[kPx];
        uint32_t cur_bits[kPx];
        unsigned long long best[kPx][kPrivRanks];
#pragma unroll
        for (int j = 0; j < kPx; ++j) {
            cur[j] = -1; cur_base[j] = 0; cur_bits[j] = 0u;
#pragma unroll
            for (int k = 0; k < kPrivRanks; ++k) best[j][k] = 0ull;
        }
        auto flush = [&]() {
#pragma unroll
            for (int j = 0; j < kPx; ++j) {
                if (cur[j] >= 0) {
                    uint32_t b = cur_bits[j];
#pragma unroll
                    for (int k = 0; k < kPrivRanks; ++k) {
                        const int c = b ? __ffs(b) - 1 : 0;
                        b &= b - 1u;
                        if (best[j][k] != 0ull) atomicMax(p.gmax + cur_base[j] + c, best[j][k]);
                        best[j][k] = 0ull;
                    }
                    cur[j] = -1;
                }
            }
        };

        int s = 0;
        uint32_t parity = 0u;
        for (long long r = r0; r < r1; ++r) {
            mbar_wait(ring.bar(s), parity);
            const unsigned char* st = ring.stage(s);
            const uint32_t m4 = (d.debug & 1) ? 0u : row_mask(st, d, lane);
            if (__any_sync(kFull, m4 != 0u)) {
                int sid[kPx];
                row_ids<IdT>(st, d, lane, m4, p.S, sid);
                const int img_row = at.img * p.S;      // 32-bit table indices (the launcher checks n_img * S * C < 2^31)
                const uint32_t* info = p.info + img_row;
                uint32_t inf[kPx];
#pragma unroll
                for (int j = 0; j < kPx; ++j) inf[j] = sid[j] >= 0 ? __ldg(info + sid[j]) : 0u;
                const float* sx = reinterpret_cast<const float*>(st) + lane * kPx;
                float v[CMAX][kPx], shift[kPx], inv[kPx];
                row_logits<CMAX, EXACT>(sx, C, v);
                row_softmax<CMAX, false>(v, p.scale, shift, inv);

                // a pixel counts for the group loss when its region does; a column that left its superpixel flushes first
                bool grp[kPx];
#pragma unroll
                for (int j = 0; j < kPx; ++j) {
                    grp[j] = do_group && (inf[j] & kGroupBit) && (inf[j] & ~kGroupBit) != 0u;      // sid >= 0 follows (inf != 0)
                    if (grp[j] && sid[j] != cur[j]) {
                        if (cur[j] >= 0) {
                            uint32_t b = cur_bits[j];
#pragma unroll
                            for (int k = 0; k < kPrivRanks; ++k) {
                                const int c = b ? __ffs(b) - 1 : 0;
                                b &= b - 1u;
                                if (best[j][k] != 0ull) atomicMax(p.gmax + cur_base[j] + c, best[j][k]);
                                best[j][k] = 0ull;
                            }
                        }
                        cur[j] = sid[j];
                        cur_bits[j] = inf[j] & ~kGroupBit;
                        cur_base[j] = (img_row + sid[j]) * C;
                    }
                }
                const uint32_t pix0 = (uint32_t)at.y * (uint32_t)p.W + (uint32_t)(at.strip * kStripPx + lane * kPx);
                // candidate classes rank by rank, the four pixels side by side (straight-line code: four independent
                // dependency chains per rank instead of one chain per pixel)
                uint32_t rest[kPx];
                float pos[kPx];
#pragma unroll
                for (int j = 0; j < kPx; ++j) {
                    rest[j] = inf[j] & ~kGroupBit;
                    pos[j] = 0.f;
                }
#pragma unroll
                for (int k = 0; k < kPrivRanks; ++k) {
#pragma unroll
                    for (int j = 0; j < kPx; ++j) {
                        const bool has = rest[j] != 0u;
                        const int c = has ? __ffs(rest[j]) - 1 : 0;
                        rest[j] &= rest[j] - 1u;
                        const float pc = mas::ex2_approx(fmaf(sx[c * kStripPx + j], p.scale, shift[j])) * inv[j];
                        pos[j] += has ? pc : 0.f;
                        const unsigned long long key = ((unsigned long long)__float_as_uint(pc) << 32) | (unsigned long long)(~(pix0 + j));
                        if (has && grp[j]) best[j][k] = key > best[j][k] ? key : best[j][k];
                    }
                }
                uint32_t rest_any = 0u;
#pragma unroll
                for (int j = 0; j < kPx; ++j) rest_any |= rest[j];
                if (__any_sync(kFull, rest_any != 0u)) {      // more than kPrivRanks candidate classes (rare)
#pragma unroll
                    for (int j = 0; j < kPx; ++j) {
                        for (uint32_t b = rest[j]; b; b &= b - 1u) {
                            const int c = __ffs(b) - 1;
                            const float pc = mas::ex2_approx(fmaf(sx[c * kStripPx + j], p.scale, shift[j])) * inv[j];
                            pos[j] += pc;
                            if (grp[j])
                                atomicMax(p.gmax + ((img_row + sid[j]) * C + c),
                                          ((unsigned long long)__float_as_uint(pc) << 32) | (unsigned long long)(~(pix0 + j)));
                        }
                    }
                }
                if (do_choice) {
#pragma unroll
                    for (int j = 0; j < kPx; ++j) {
                        const float l = -logf(pos[j] + kEps);
                        const int n = __popc(inf[j] & ~kGroupBit);
                        const bool on = sid[j] >= 0;
                        sum_one += (on && n == 1) ? l : 0.f;   n_one += (on && n == 1) ? 1 : 0;
                        sum_multi += (on && n > 1) ? l : 0.f;  n_multi += (on && n > 1) ? 1 : 0;
                        sum_empty += (on && n == 0) ? l : 0.f; n_empty += (on && n == 0) ? 1 : 0;
                    }
                }
            }
            __syncwarp();      // every lane is done with stage s
            if (lane == 0 && issued < r1) {
                fence_async_shared();
                issue(s);
            }
            if (++s == stages) { s = 0; parity ^= 1u; }
            if (at.advance(d.strips, p.H) == 2) flush();      // ids are per image
        }
        flush();
    }

    if (do_choice) {      // one combine per CTA, in a fixed order
        __shared__ float s_sum[kMaxWarps][3];
        __shared__ int s_cnt[kMaxWarps][3];
        float sv[3] = {sum_one, sum_multi, sum_empty};
        int nv[3] = {n_one, n_multi, n_empty};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sv[k] += __shfl_xor_sync(kFull, sv[k], o);
                nv[k] += __shfl_xor_sync(kFull, nv[k], o);
            }
            if (lane == 0) { s_sum[warp][k] = sv[k]; s_cnt[warp][k] = nv[k]; }
        }
        __syncthreads();
        if (tid < 3) {
            double total = 0.0;
            long long count = 0;
            for (int w = 0; w < warps; ++w) { total += (double)s_sum[w][tid]; count += s_cnt[w][tid]; }
            if (count != 0) {
                atomicAdd(p.acc + 2 * tid, total);
                atomicAdd(p.acc + 2 * tid + 1, (double)count);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ backward
template <int CMAX, bool EXACT, typename IdT>
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
multihot_dense_bwd_kernel(const __grid_constant__ DenseMaps maps, const LossParams p, const DenseShape d) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (p.list_mode && p.dense_percent > 0 && !dense_regime(p)) return;      // sparsely selected: zero sweep + list walk
    unsigned char* smem = align128(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warps = blockDim.x >> 5;
    const int C = EXACT ? CMAX : p.C;
    const int stages = d.stages;      // 3: loading / computing / draining
    Ring ring;
    ring.stage_bytes = d.stage_bytes;
    ring.stages = smem + (size_t)warp * stages * d.stage_bytes;
    ring.bars = reinterpret_cast<uint64_t*>(smem + (size_t)warps * stages * d.stage_bytes) + warp * stages;
    const bool do_choice = p.do_choice != 0, do_group = p.do_group != 0;
    const float w_one = p.coef[0] * p.inv_temp, w_multi = p.coef[1] * p.inv_temp, w_group = p.coef[3] * p.inv_temp;

    long long r0, r1;
    mas::warp_range(d.total_rows, (long long)blockIdx.x * warps + warp, (long long)gridDim.x * warps, r0, r1);
    if (r0 >= r1) return;
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(ring.bar(s), 1u);
        fence_barrier_init();
    }
    __syncwarp();

    const uint64_t policy = policy_evict_first();
    mas::Cursor at, ahead;
    at.seek(r0, d.strips, p.H);
    ahead = at;
    long long issued = r0;
    auto issue = [&](int s) {      // lane 0
        const uint32_t bar = ring.bar(s), dst = smem_u32(ring.stage(s));
        if (d.debug & 1) {
            mbar_expect_tx(bar, d.ids_off);
            load_4d(dst, &maps.logits, bar, ahead.strip * kStripPx, ahead.y, 0, ahead.img, policy);
        } else if (d.debug & 2) {
            mbar_expect_tx(bar, d.tx_bytes);
            load_4d(dst, &maps.logits, bar, ahead.strip * kStripPx, ahead.y, 0, ahead.img);
            load_3d(dst + d.ids_off, &maps.ids, bar, ahead.strip * kStripPx, ahead.y, ahead.img);
            load_3d(dst + d.mask_off, &maps.mask, bar, ahead.strip * kStripPx, ahead.y, ahead.img);
        } else {
            mbar_expect_tx(bar, d.tx_bytes);
            load_4d(dst, &maps.logits, bar, ahead.strip * kStripPx, ahead.y, 0, ahead.img, policy);
            load_3d(dst + d.ids_off, &maps.ids, bar, ahead.strip * kStripPx, ahead.y, ahead.img, policy);
            load_3d(dst + d.mask_off, &maps.mask, bar, ahead.strip * kStripPx, ahead.y, ahead.img, policy);
        }
        ahead.advance(d.strips, p.H);
        ++issued;
    };
    if (lane == 0) {
        for (int s = 0; s + 1 < stages && issued < r1; ++s) issue(s);
    }

    int s = 0;
    uint32_t parity = 0u;
    for (long long r = r0; r < r1; ++r) {
        mbar_wait(ring.bar(s), parity);
        unsigned char* st = ring.stage(s);
        float* sx = reinterpret_cast<float*>(st) + lane * kPx;
        const uint32_t m4 = (d.debug & 1) ? 0u : row_mask(st, d, lane);
        int sid[kPx];
        row_ids<IdT>(st, d, lane, m4, p.S, sid);
        const int img_row = at.img * p.S;      // 32-bit table indices (the launcher checks n_img * S * C < 2^31)
        const uint32_t* info = p.info + img_row;
        uint32_t inf[kPx];
        bool grp[kPx], live[kPx];
        bool any_live = false;
#pragma unroll
        for (int j = 0; j < kPx; ++j) {
            inf[j] = sid[j] >= 0 ? __ldg(info + sid[j]) : 0u;
            const uint32_t bits = inf[j] & ~kGroupBit;
            grp[j] = do_group && (inf[j] & kGroupBit) && bits != 0u;
            live[j] = grp[j] || (do_choice && bits != 0u);            // does any gradient reach this pixel?
            any_live |= live[j];
        }
        // Refill the stage whose gradient row left one iteration ago -- here, not after this row's store: the id / candidate
        // gathers above gave that store time to drain, and the next row's loads now overlap this row's arithmetic
        if (lane == 0 && issued < r1) {
            store_wait_read<0>();
            issue(s == 0 ? stages - 1 : s - 1);
        }
        if (!__any_sync(kFull, any_live)) {
            float z[kPx];
#pragma unroll
            for (int j = 0; j < kPx; ++j) z[j] = 0.f;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                if (EXACT || c < C) store_px(sx + c * kStripPx, z);
            }
        } else {
            // classes of the first kPrivRanks candidates of every pixel (cls < 0: none), and their ownership words: is
            // this pixel the arg-max pixel of (superpixel, class)?  The table entry holds ~pixel in its low word; an
            // empty entry holds 0, which no pixel maps to.  Gathered before the softmax (one latency, not one per rank).
            const uint32_t pix0 = (uint32_t)at.y * (uint32_t)p.W + (uint32_t)(at.strip * kStripPx + lane * kPx);
            int cls[kPx][kPrivRanks];
            uint32_t rest[kPx];
            bool owns[kPx][kPrivRanks];
            bool any_owner = false;
#pragma unroll
            for (int j = 0; j < kPx; ++j) {
                const uint32_t* row = reinterpret_cast<const uint32_t*>(p.gmax) + 2 * ((img_row + (sid[j] >= 0 ? sid[j] : 0)) * C);
                rest[j] = live[j] ? (inf[j] & ~kGroupBit) : 0u;
#pragma unroll
                for (int k = 0; k < kPrivRanks; ++k) {
                    cls[j][k] = rest[j] ? __ffs(rest[j]) - 1 : -1;
                    rest[j] &= rest[j] - 1u;
                    owns[j][k] = cls[j][k] >= 0 && grp[j] && __ldg(row + 2 * cls[j][k]) == ~(pix0 + j);
                    any_owner |= owns[j][k];
                }
            }
            float v[CMAX][kPx], shift[kPx], inv[kPx];
            row_logits<CMAX, EXACT>(sx, C, v);
            row_softmax<CMAX, true>(v, p.scale, shift, inv);
            // d/dx_c = P_c * (a * ([c in row] - pos) + w_group * q_sum) - w_group * [c pooled from this pixel] * q_c,
            // q_c = P_c / (P_c + eps): candidate classes rank by rank, the pixels side by side (straight-line code)
            float s_in[kPx], s_out[kPx], ek[kPx][kPrivRanks], pooled[kPx][kPrivRanks], pos[kPx], q_sum[kPx];
#pragma unroll
            for (int j = 0; j < kPx; ++j) { pos[j] = 0.f; q_sum[j] = 0.f; }
#pragma unroll
            for (int k = 0; k < kPrivRanks; ++k) {
#pragma unroll
                for (int j = 0; j < kPx; ++j) {
                    const int c = cls[j][k] >= 0 ? cls[j][k] : 0;
                    ek[j][k] = mas::ex2_approx(fmaf(sx[c * kStripPx + j], p.scale, shift[j]));
                    pos[j] += cls[j][k] >= 0 ? ek[j][k] * inv[j] : 0.f;
                    pooled[j][k] = 0.f;
                }
            }
            if (__any_sync(kFull, any_owner)) {      // some pixel of the row is the arg-max pixel of a max-pool (under 1 % of the pixels)
#pragma unroll
                for (int k = 0; k < kPrivRanks; ++k) {
#pragma unroll
                    for (int j = 0; j < kPx; ++j) {
                        const float pc = ek[j][k] * inv[j];
                        const float q = owns[j][k] ? __fdividef(pc, pc + kEps) : 0.f;
                        q_sum[j] += q;
                        pooled[j][k] = w_group * q;
                    }
                }
            }
            uint32_t rest_any = 0u;
#pragma unroll
            for (int j = 0; j < kPx; ++j) rest_any |= rest[j];
            const bool many = __any_sync(kFull, rest_any != 0u);
            uint32_t abits[kPx];      // further classes pooled from this pixel (ranks >= kPrivRanks)
#pragma unroll
            for (int j = 0; j < kPx; ++j) abits[j] = 0u;
            if (many) {      // a pixel with more than kPrivRanks candidate classes
#pragma unroll
                for (int j = 0; j < kPx; ++j) {
                    const uint32_t* row = reinterpret_cast<const uint32_t*>(p.gmax) + 2 * ((img_row + (sid[j] >= 0 ? sid[j] : 0)) * C);
                    for (uint32_t b = rest[j]; b; b &= b - 1u) {
                        const int c = __ffs(b) - 1;
                        const float pc = mas::ex2_approx(fmaf(sx[c * kStripPx + j], p.scale, shift[j])) * inv[j];
                        pos[j] += pc;
                        if (grp[j] && __ldg(row + 2 * c) == ~(pix0 + j)) {
                            abits[j] |= 1u << c;
                            q_sum[j] += __fdividef(pc, pc + kEps);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kPx; ++j) {
                const uint32_t bits = inf[j] & ~kGroupBit;
                const float a = (do_choice && live[j]) ? -__fdividef(__popc(bits) == 1 ? w_one : w_multi, pos[j] + kEps) : 0.f;
                const float shared_term = w_group * q_sum[j];      // q_sum = 0 unless the pixel owns a pooled probability
                s_in[j] = live[j] ? (a * (1.f - pos[j]) + shared_term) * inv[j] : 0.f;
                s_out[j] = live[j] ? (shared_term - a * pos[j]) * inv[j] : 0.f;
            }
            // The gradient overwrites the logits in the stage: e_c * s_out for every class (no per-class test), then the
            // entries of the candidate classes patched to e_c * s_in (minus the pooled term where the pixel owns it).
            if (!many) {
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
                    if (EXACT || c < C) {
                        float g[kPx];
#pragma unroll
                        for (int j = 0; j < kPx; ++j) g[j] = v[c][j] * s_out[j];
                        store_px(sx + c * kStripPx, g);
                    }
                }
            } else {
                // candidates beyond the first kPrivRanks are computed in place from the raw logits first, and the
                // sweep over the classes steps around them
#pragma unroll
                for (int j = 0; j < kPx; ++j) {
                    for (uint32_t b = rest[j]; b; b &= b - 1u) {
                        const int c = __ffs(b) - 1;
                        const float e = mas::ex2_approx(fmaf(sx[c * kStripPx + j], p.scale, shift[j]));
                        float g = e * s_in[j];
                        if ((abits[j] >> c) & 1u) {
                            const float pc = e * inv[j];
                            g -= w_group * __fdividef(pc, pc + kEps);
                        }
                        sx[c * kStripPx + j] = g;
                    }
                }
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
                    if (EXACT || c < C) {
                        float g[kPx];
                        load_px(sx + c * kStripPx, g);
#pragma unroll
                        for (int j = 0; j < kPx; ++j) g[j] = ((rest[j] >> c) & 1u) ? g[j] : v[c][j] * s_out[j];
                        store_px(sx + c * kStripPx, g);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kPrivRanks; ++k) {
#pragma unroll
                for (int j = 0; j < kPx; ++j) {
                    if (cls[j][k] >= 0) sx[cls[j][k] * kStripPx + j] = fmaf(ek[j][k], s_in[j], -pooled[j][k]);
                }
            }
        }
        fence_async_shared();      // the stage's new contents -> visible to the TMA store
        __syncwarp();
        if (lane == 0) {
            store_4d(&maps.grad, smem_u32(st), at.strip * kStripPx, at.y, 0, at.img);
            store_commit();
        }
        if (++s == stages) { s = 0; parity ^= 1u; }
        at.advance(d.strips, p.H);
    }
    if (lane == 0) store_wait_all<0>();
}

// ------------------------------------------------------------------------------------------ host side
int env_int(const char* name, int fallback) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : fallback;
}

template <int CMAX, bool EXACT, typename IdT>
cudaError_t launch_dense_one(const LossParams& p, bool backward, cudaStream_t stream, bool* launched) {
    *launched = false;
    const int dev = mas::current_device();
    int max_smem = 0;
    if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return cudaSuccess;
    DenseShape d;
    d.strips = (p.W + kStripPx - 1) / kStripPx;
    d.total_rows = (long long)p.n_img * d.strips * p.H;
    d.stages = backward ? std::min(3, std::max(2, env_int("MAS_LOSS_DENSE_BSTAGES", 2))) : std::min(4, std::max(2, env_int("MAS_LOSS_DENSE_STAGES", 2)));
    d.debug = env_int("MAS_LOSS_DENSE_DEBUG", 0);
    d.ids_off = (uint32_t)p.C * kStripPx * 4;
    d.mask_off = d.ids_off + kStripPx * (uint32_t)sizeof(IdT);
    d.tx_bytes = d.mask_off + kStripPx;
    d.stage_bytes = (d.tx_bytes + 127u) & ~127u;
    const size_t per_warp = (size_t)d.stages * d.stage_bytes + (size_t)d.stages * 8;
    int warps = std::min(kMaxWarps, (int)(((size_t)max_smem - 1024) / per_warp));      // 1 KB: alignment slack + the static combine buffers
    const int want = env_int("MAS_LOSS_DENSE_WARPS", 0);
    if (want > 0) warps = std::min(warps, want);
    if (warps < 1) return cudaSuccess;
    const size_t smem = (size_t)warps * per_warp + 128;

    DenseMaps maps;
    const cuuint64_t W = (cuuint64_t)p.W, H = (cuuint64_t)p.H, N = (cuuint64_t)p.n_img, Cc = (cuuint64_t)p.C;
    {
        const cuuint64_t dims[4] = {W, H, Cc, N};
        const cuuint64_t strides[3] = {W * 4, W * H * 4, W * H * Cc * 4};
        const cuuint32_t box[4] = {kStripPx, 1, (cuuint32_t)p.C, 1};
        if (!encode(&maps.logits, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.logits, dims, strides, box)) return cudaSuccess;
        maps.grad = maps.logits;
        if (backward && !encode(&maps.grad, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.grad, dims, strides, box)) return cudaSuccess;
    }
    {
        const cuuint64_t dims[3] = {W, H, N};
        const cuuint32_t box[3] = {kStripPx, 1, 1};
        const cuuint64_t istr[2] = {W * sizeof(IdT), W * H * sizeof(IdT)};
        if (!encode(&maps.ids, sizeof(IdT) == 8 ? CU_TENSOR_MAP_DATA_TYPE_INT64 : CU_TENSOR_MAP_DATA_TYPE_INT32, 3, p.ids, dims, istr, box))
            return cudaSuccess;
        const cuuint64_t mstr[2] = {W, W * H};
        if (!encode(&maps.mask, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, p.mask, dims, mstr, box)) return cudaSuccess;
    }

    static mas::PerDeviceInt configured[2];
    const long long cap = (d.total_rows + 4 * warps - 1) / (4 * warps);      // never fewer than ~4 rows per warp
    const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>(mas::sm_count(), cap));
    if (backward) {
        auto kernel = multihot_dense_bwd_kernel<CMAX, EXACT, IdT>;
        cudaError_t e = mas::opt_in_smem(kernel, configured[1], (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<blocks, warps * 32, smem, stream>>>(maps, p, d);
    } else {
        auto kernel = multihot_dense_fwd_kernel<CMAX, EXACT, IdT>;
        cudaError_t e = mas::opt_in_smem(kernel, configured[0], (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<blocks, warps * 32, smem, stream>>>(maps, p, d);
    }
    mas::count_launches(1);
    *launched = true;
    return cudaGetLastError();
}

template <typename IdT>
cudaError_t launch_dense_channels(const LossParams& p, bool backward, cudaStream_t stream, bool* launched) {
    switch (p.C) {
        case 19: return launch_dense_one<19, true, IdT>(p, backward, stream, launched);
        case 20: return launch_dense_one<20, true, IdT>(p, backward, stream, launched);
        case 21: return launch_dense_one<21, true, IdT>(p, backward, stream, launched);
        case 22: return launch_dense_one<22, true, IdT>(p, backward, stream, launched);
        default: break;
    }
    if (p.C <= 8) return launch_dense_one<8, false, IdT>(p, backward, stream, launched);
    if (p.C <= 16) return launch_dense_one<16, false, IdT>(p, backward, stream, launched);
    if (p.C <= 24) return launch_dense_one<24, false, IdT>(p, backward, stream, launched);
    return launch_dense_one<31, false, IdT>(p, backward, stream, launched);
}

}  // namespace

cudaError_t mas_loss::launch_dense(const LossParams& p, int ids_dtype, bool backward, cudaStream_t stream, bool* launched) {
    *launched = false;
    // TMA sources: rows of every operand 16-byte aligned (the uint8 mask is the narrowest), bases 16-byte aligned
    const uintptr_t bases = reinterpret_cast<uintptr_t>(p.logits) | reinterpret_cast<uintptr_t>(p.ids) | reinterpret_cast<uintptr_t>(p.mask) |
                            (backward ? reinterpret_cast<uintptr_t>(p.grad) : 0);
    if (p.W % 16 != 0 || (bases & 15) != 0) return cudaSuccess;
    if ((long long)p.H * p.W >= (1ll << 32) || (long long)p.n_img * p.S * p.C >= (1ll << 31)) return cudaSuccess;
    if (ids_dtype == MAS_I64) return launch_dense_channels<long long>(p, backward, stream, launched);
    return launch_dense_channels<int32_t>(p, backward, stream, launched);
}
