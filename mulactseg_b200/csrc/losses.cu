// Stage-1 losses over superpixel ids: multi-hot partial-label (one-hot CE / multi-choice) and MIL
// ("merged positive", per-superpixel x class max-pool) losses, forward and backward.
//
// Reference (paths relative to the reference checkout):
//   utils/loss.py:81-141 GroupMultiLabelCE, :535-588 MultiChoiceCE
//   trainer/active_joint_multi_predignore.py:17-128 MultiChoiceCE_, GroupMultiLabelCE_
//   trainer/active_joint_multi_predignore_mclossablation2.py:17-79 GroupMultiLabelCE_onlymulti
//   trainer/active_joint_multi_predignore_lossdecomp.py:16-72, trainer/active_joint_multi_lossdecomp.py:17-74
//       OnehotCEMultihotChoice
// The reference materialises softmax(x/T) twice, permutes it, boolean-compacts it per image in a Python
// loop and calls torch_scatter.scatter(reduce='max'); here ONE pass over the logits of the masked pixels
// does all of it (and skips the logits of unmasked pixels entirely):
//   P = softmax(x/T) in registers;  row = candidate set of the pixel's superpixel (bit mask);
//   pos = sum_{c in row} P_c;  l = -log(pos + 1e-8) summed into the one-hot / multi-hot / empty bucket;
//   M[s,c] = max P_c over the group-valid pixels of s, kept as a packed 64-bit (P bits, ~pixel) key so
//   that the arg-max pixel (first index on ties, like torch_scatter on CPU) comes with the value.
// Same walking scheme as the scorer (walk.cuh): a thread stays inside one superpixel for many rows and
// keeps that superpixel's running maxima in a private shared-memory column; global atomicMax only when
// the superpixel changes.  Backward recomputes the softmax and writes the dense gradient once.
#include "losses.cuh"

#include <stdlib.h>

#include <algorithm>

using namespace mas_loss;

namespace {

__global__ void multihot_info_kernel(const uint8_t* __restrict__ targets, long long n_regions, int Ct, int C, int group_mode,
                                     uint32_t* __restrict__ info) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_regions) return;
    const uint8_t* t = targets + r * Ct;
    uint32_t bits = 0u;
    int total = 0;
    for (int c = 0; c < Ct; ++c) {
        const int v = t[c];
        total += v;
        if (c < C && v) bits |= 1u << c;
    }
    const bool in_group = group_mode == 0 || total > 1;
    info[r] = bits | (in_group ? kGroupBit : 0u);
}

// ------------------------------------------------------------------------------------------ per-pixel pieces
// Work decomposition of both kernels: a TILE is 32 pixels x kTileRows rows (lane = column, one pixel per thread per
// row); the warps of a persistent grid (resident CTAs x SMs) take tiles round-robin, so neighbouring tiles -- which tend
// to be equally expensive -- land on different warps.  A thread walks DOWN its column, so it stays inside one superpixel
// for many rows (private running maxima, see the header).  No row waits on a chain of dependent round trips to HBM:
//   1. the mask bytes of all kTileRows rows are loaded at once (one 32-byte sector per row and warp); a tile without a
//      selected pixel ends there (forward) or is zero-filled with independent streaming stores (backward);
//   2. rows are software-pipelined three deep: while row r is computed, the C' logits and the candidate word of row
//      r + 1 and the id of row r + 2 are in flight (loads of unselected pixels are predicated off).
constexpr int kTileRows = 16;          // most rows a tile can have (the launcher picks 4, 8 or 16: see launch_one)
constexpr uint32_t kFullWarp = 0xffffffffu;

__device__ __forceinline__ int clamp_id(long long v) { return (v < 0 || v > 0x7fffffffLL) ? -1 : (int)v; }

template <typename IdT>
__device__ __forceinline__ int load_id(const void* ids, size_t i) {
    return clamp_id((long long)__ldcs(reinterpret_cast<const IdT*>(ids) + i));
}

// softmax(x / T) as numerators + normaliser: v[c] <- e_c = exp(x_c/T - max), returns what turns e_c into P_c.
//  ACCURATE: the reference's own operation order (divide by T, subtract the maximum, expf, divide by the sum): the
//            return value is the SUM and P_c = __fdiv_rn(e_c, sum) -- used where the arg-max PIXEL of a max-pool must
//            agree with torch to the last bit (stage-2 prototypes);
//  fast    : ex2.approx((x - max) * log2e / T); the return value is 1 / sum and P_c = e_c * that (<= 4 ulp from the
//            exact form): the losses (1e-5 tolerance).
// Only the few candidate classes of a pixel ever need P_c, so the C' normalisations are not done here.
template <int CMAX, bool ACCURATE>
__device__ __forceinline__ float softmax_numerators(float (&v)[CMAX], float temp, float scale) {
    if (ACCURATE && temp != 1.f) {      // x / 1.0f == x: the stage-2 labeller (T = 1, :140) skips the C' divisions
#pragma unroll
        for (int c = 0; c < CMAX; ++c) v[c] = __fdiv_rn(v[c], temp);   // padded planes hold -inf
    }
    float mx = v[0];
#pragma unroll
    for (int c = 1; c < CMAX; ++c) mx = fmaxf(mx, v[c]);
    float sa = 0.f, sb = 0.f;
    if (ACCURATE) {
#pragma unroll
        for (int c = 0; c < CMAX; ++c) { v[c] = expf(v[c] - mx); sa += v[c]; }
        return sa;
    }
    const float shift = -mx * scale;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        v[c] = mas::ex2_approx(fmaf(v[c], scale, shift));
        if (c & 1) sb += v[c]; else sa += v[c];
    }
    return 1.f / (sa + sb);
}

template <bool ACCURATE>
__device__ __forceinline__ float normalised(float e, float norm) { return ACCURATE ? __fdiv_rn(e, norm) : e * norm; }

__device__ __forceinline__ void prefetch_l1(const void* ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

struct Tile {
    int img, x, y0, rows;
};

__device__ __forceinline__ long long tile_count(const LossParams& p) {
    return (long long)p.n_img * ((p.W + 31) / 32) * ((p.H + p.tile_rows - 1) / p.tile_rows);
}

__device__ __forceinline__ Tile make_tile(const LossParams& p, long long tile) {
    Tile t;
    const int tiles_x = (p.W + 31) / 32, tiles_y = (p.H + p.tile_rows - 1) / p.tile_rows;
    const long long per_img = (long long)tiles_x * tiles_y;
    t.img = (int)(tile / per_img);
    const int rem = (int)(tile - (long long)t.img * per_img);
    const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
    t.x = tx * 32 + (threadIdx.x & 31);
    t.y0 = ty * p.tile_rows;
    t.rows = min(p.H - t.y0, p.tile_rows);
    return t;
}

// Step 1: bit r of the result <-> the pixel of row y0 + r in this thread's column is selected by the mask.
__device__ __forceinline__ uint32_t tile_mask_bits(const LossParams& p, const Tile& t, size_t pix0) {
    const bool active = t.x < p.W;
    uint8_t m[kTileRows];
#pragma unroll
    for (int r = 0; r < kTileRows; ++r) m[r] = (active && r < t.rows) ? __ldcs(p.mask + pix0 + (size_t)r * p.W) : (uint8_t)0;
    uint32_t mbits = 0u;
#pragma unroll
    for (int r = 0; r < kTileRows; ++r) mbits |= (m[r] != 0 ? 1u : 0u) << r;
    return mbits;
}

// pull the mask bytes of a tile this warp will scan later into L1
__device__ __forceinline__ void prefetch_tile_mask(const LossParams& p, const Tile& t, size_t pix0) {
    if (t.x < p.W) {
#pragma unroll
        for (int r = 0; r < kTileRows; ++r) {
            if (r < t.rows) prefetch_l1(p.mask + pix0 + (size_t)r * p.W);
        }
    }
}

constexpr int kPrefetchRows = 4;   // rows ahead of the one being computed whose id + logits are pulled into L1

// Step 2: the three-deep row pipeline of one thread.  After `start`, `next()` hands out rows 0, 1, 2 ... in turn:
// the logits (raw), the id (-1: not selected / out of range) and the candidate word of the thread's pixel.
template <int CMAX, bool EXACT, typename IdT>
struct RowPipe {
    const LossParams& p;
    const float* logits;     // plane 0 of the thread's pixel in row 0 of the tile
    const uint32_t* info;    // candidate words of the image
    size_t pix, P;
    uint32_t mbits;
    int C, r;
    float vn[CMAX];          // row r (in flight)
    uint32_t infn;           // row r
    int idn, idnn;           // rows r, r + 1

    __device__ __forceinline__ RowPipe(const LossParams& params) : p(params) {}

    __device__ __forceinline__ int fetch_id(int row) const {
        if (!((mbits >> row) & 1u)) return -1;
        const int id = load_id<IdT>(p.ids, pix + (size_t)row * p.W);
        return ((unsigned)id < (unsigned)p.S) ? id : -1;
    }
    __device__ __forceinline__ void fetch_row(int row, int id) {
        const bool on = id >= 0;
        infn = on ? __ldg(info + id) : 0u;
        const float* ptr = logits + (size_t)row * p.W;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            vn[c] = (EXACT || c < C) ? (on ? __ldcs(ptr) : 0.f) : -INFINITY;
            ptr += P;       // one 64-bit add per plane
        }
    }
    // L1 prefetch of a row further down (no registers held): the register loads above then hit L1
    __device__ __forceinline__ void prefetch_row(int row) const {
        if (row < kTileRows && ((mbits >> row) & 1u)) {
            prefetch_l1(reinterpret_cast<const IdT*>(p.ids) + pix + (size_t)row * p.W);
            const float* ptr = logits + (size_t)row * p.W;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                if (EXACT || c < C) prefetch_l1(ptr);
                ptr += P;
            }
        }
    }
    __device__ __forceinline__ void start(const Tile& t, int channels, size_t plane, size_t off0, uint32_t mask_bits) {
        C = channels; P = plane; mbits = mask_bits;
        pix = (size_t)t.img * P + off0;
        logits = p.logits + (size_t)t.img * C * P + off0;
        info = p.info + (size_t)t.img * p.S;
        r = 0;
        idn = fetch_id(0);
        idnn = fetch_id(1);
        fetch_row(0, idn);
#pragma unroll 1
        for (int k = 1; k < kPrefetchRows; ++k) prefetch_row(k);
    }
    __device__ __forceinline__ void next(float (&v)[CMAX], int& id, uint32_t& inf) {
#pragma unroll
        for (int c = 0; c < CMAX; ++c) v[c] = vn[c];
        id = idn; inf = infn;
        ++r;
        idn = idnn;
        idnn = (r + 1 < kTileRows) ? fetch_id(r + 1) : -1;
        if (r < kTileRows) fetch_row(r, idn);
        prefetch_row(r + kPrefetchRows - 1);
    }
};

// ------------------------------------------------------------------------------------------ active-tile list
// In training only the labelled superpixels are selected (2-10 % of the pixels after a few acquisition rounds).  A static
// split of ALL tiles over the warps then leaves the step waiting for the unlucky warp that drew several active tiles
// (forward: 0.12 ms for 80 MB of traffic), and the backward pass zero-fills the dense gradient tile by tile (0.22 ms where
// a linear memset takes 0.10).  One scan of the mask (9 MB) builds a bitmap of the 32 px x 8 row tiles that hold a
// selected pixel plus its prefix sums; both passes then deal the ACTIVE tiles round-robin to their warps (rank -> tile by
// binary search + bit select: deterministic, unlike an atomic queue), and a sparsely selected batch gets its gradient
// zeroed by one linear sweep with only the active tiles computed on top.
constexpr int kSparsePercent = 35;      // backward: list mode when fewer than this share of the tiles is active

// tile of rank r in the active list
__device__ __forceinline__ long long list_select(const LossParams& p, int r) {
    int lo = 0, hi = p.n_groups - 1;           // largest g with tile_offsets[g] <= r
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(p.tile_offsets + mid) <= r) lo = mid; else hi = mid - 1;
    }
    const uint32_t word = __ldg(p.tile_words + lo);
    const int k = r - __ldg(p.tile_offsets + lo);
    return (long long)lo * 32 + __fns(word, 0, k + 1);
}

// one lane per tile: does the tile hold a selected pixel?  The last CTA to finish turns the popcounts into prefix sums.
__global__ void __launch_bounds__(256) tile_scan_kernel(const uint8_t* __restrict__ mask, int n_img, int H, int W, int n_groups,
                                                        uint32_t* __restrict__ words, int* __restrict__ offsets,
                                                        unsigned int* __restrict__ ticket) {
    const int lane = threadIdx.x & 31;
    const long long n_tiles = list_tile_count(n_img, H, W);
    const int tiles_x = (W + 31) / 32, tiles_y = (H + kListRows - 1) / kListRows;
    const long long per_img = (long long)tiles_x * tiles_y;
    const size_t P = (size_t)H * W;
    for (long long g = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); g < n_groups; g += (long long)gridDim.x * 8) {
        const long long tile = g * 32 + lane;
        bool any = false;
        if (tile < n_tiles) {
            const int img = (int)(tile / per_img);
            const int rem = (int)(tile - (long long)img * per_img);
            const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
            const int x0 = tx * 32, x1 = min(x0 + 32, W), y0 = ty * kListRows, y1 = min(y0 + kListRows, H);
            const uint8_t* base = mask + (size_t)img * P;
            if (x1 - x0 == 32 && y1 - y0 == kListRows && (W & 15) == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0) {
                // full tile of 16-byte aligned rows: sixteen independent 128-bit loads, OR-reduced
                uint4 q[2 * kListRows];
#pragma unroll
                for (int r = 0; r < kListRows; ++r) {
                    const uint4* row = reinterpret_cast<const uint4*>(base + (size_t)(y0 + r) * W + x0);
                    q[2 * r] = __ldg(row);
                    q[2 * r + 1] = __ldg(row + 1);
                }
                uint32_t acc = 0u;
#pragma unroll
                for (int k = 0; k < 2 * kListRows; ++k) acc |= q[k].x | q[k].y | q[k].z | q[k].w;
                any = acc != 0u;
            } else {
                for (int y = y0; y < y1; ++y) {
                    const uint8_t* row = base + (size_t)y * W;
                    int x = x0;
                    // 32-bit reads over the aligned middle of the 32-byte run
                    for (; x < x1 && ((reinterpret_cast<uintptr_t>(row + x)) & 3); ++x) any |= row[x] != 0;
                    for (; x + 4 <= x1; x += 4) any |= *reinterpret_cast<const uint32_t*>(row + x) != 0u;
                    for (; x < x1; ++x) any |= row[x] != 0;
                }
            }
        }
        const uint32_t word = __ballot_sync(0xffffffffu, any);
        if (lane == 0) { words[g] = word; offsets[g] = __popc(word); }
    }
    __shared__ bool last;
    __shared__ int warp_tot[8];
    __shared__ int carry;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) { carry = 0; *ticket = 0u; }      // the workspace is reusable without another memset
    __syncthreads();
    for (int base = 0; base < n_groups; base += 256) {
        const int i = base + threadIdx.x;
        const int v = i < n_groups ? __ldcg(offsets + i) : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[threadIdx.x >> 5] = x;
        __syncthreads();
        int before = carry;
        for (int wv = 0; wv < (int)(threadIdx.x >> 5); ++wv) before += warp_tot[wv];
        if (i < n_groups) offsets[i] = before + x - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n_groups] = carry;
}

// backward, list mode only: one linear sweep zeroes the dense gradient (the active tiles are then computed on top)
__global__ void __launch_bounds__(256) grad_zero_kernel(float* __restrict__ grad, size_t n, const int* __restrict__ offsets, int n_groups,
                                                        long long n_tiles, int percent) {
    if ((long long)offsets[n_groups] * 100 >= n_tiles * percent) return;      // densely selected: the tile walk fills everything
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    const size_t head = min(n, (size_t)(((16 - (reinterpret_cast<uintptr_t>(grad) & 15)) & 15) / 4));
    const size_t vecs = (n - head) / 4;
    float4* g4 = reinterpret_cast<float4*>(grad + head);
    for (size_t i = tid; i < vecs; i += stride) __stcs(g4 + i, make_float4(0.f, 0.f, 0.f, 0.f));
    for (size_t i = tid; i < head; i += stride) grad[i] = 0.f;
    for (size_t i = head + vecs * 4 + tid; i < n; i += stride) grad[i] = 0.f;
}

// ------------------------------------------------------------------------------------------ forward
template <int CMAX, bool EXACT, typename IdT, bool ACCURATE>
__global__ void __launch_bounds__(kThreads) multihot_loss_fwd_kernel(const LossParams p) {
    // [C][kThreads] u64: running maxima of the thread's current superpixel;  [C][kThreads] f32: softmax numerators of
    // the pixel being processed (registers cannot be indexed by the candidate class; private columns, no conflicts)
    extern __shared__ unsigned long long gcol[];
    if (dense_regime(p)) return;          // densely selected batch: the strip walk of losses_dense.cu takes it
    const int tid = threadIdx.x, lane = tid & 31;
    const int C = EXACT ? CMAX : p.C;
    unsigned long long* col = gcol + tid;
    float* pst = reinterpret_cast<float*>(gcol + (size_t)C * kThreads) + tid;
    const bool do_choice = p.do_choice != 0, do_group = p.do_group != 0;
    const size_t P = (size_t)p.H * p.W;
    float sum_one = 0.f, sum_multi = 0.f, sum_empty = 0.f;
    int n_one = 0, n_multi = 0, n_empty = 0;
    if (do_group) {
        for (int c = 0; c < C; ++c) col[c * kThreads] = 0ull;
    }

    // list mode: `work` = active tiles, dealt round-robin by RANK (every warp gets the same number +-1 of tiles that
    // actually hold selected pixels); otherwise every tile of the batch
    const long long n_work = p.list_mode ? (long long)__ldg(p.tile_offsets + p.n_groups) : tile_count(p);
    const long long warp_stride = (long long)gridDim.x * (kThreads / 32);
    for (long long item = (long long)blockIdx.x * (kThreads / 32) + (tid >> 5); item < n_work; item += warp_stride) {
        const long long tile = p.list_mode ? list_select(p, (int)item) : item;
        const Tile t = make_tile(p, tile);
        const size_t off0 = (size_t)t.y0 * p.W + t.x;
        const uint32_t mbits = tile_mask_bits(p, t, (size_t)t.img * P + off0);
        if (!p.list_mode && item + warp_stride < n_work) {
            const Tile tn = make_tile(p, item + warp_stride);
            prefetch_tile_mask(p, tn, (size_t)tn.img * P + (size_t)tn.y0 * p.W + tn.x);
        }
        const uint32_t any_rows = __reduce_or_sync(kFullWarp, mbits);
        if (any_rows == 0u) continue;                                                         // warp-uniform
        RowPipe<CMAX, EXACT, IdT> pipe(p);
        pipe.start(t, C, P, off0, mbits);
        int cur = -1;
        long long cur_base = 0;   // table offset of superpixel `cur`
#pragma unroll 1
        for (int r = 0; r < t.rows; ++r) {
            float v[CMAX];
            int id;
            uint32_t inf;
            pipe.next(v, id, inf);
            if (!((any_rows >> r) & 1u)) continue;                                            // warp-uniform
            const float norm = softmax_numerators<CMAX, ACCURATE>(v, p.temp, p.scale);
            if (id < 0) continue;
            const uint32_t bits = inf & ~kGroupBit;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                if (EXACT || c < C) pst[c * kThreads] = v[c];
            }
            const bool group = do_group && (inf & kGroupBit) && bits != 0u;
            if (group && id != cur) {
                if (cur >= 0) {
                    for (int c = 0; c < C; ++c) {
                        const unsigned long long e = col[c * kThreads];
                        if (e != 0ull) { atomicMax(p.gmax + cur_base + c, e); col[c * kThreads] = 0ull; }
                    }
                }
                cur = id;
                cur_base = ((long long)t.img * p.S + cur) * C;
            }
            // candidate classes in ascending order: probability mass and running maxima
            const unsigned long long low = (unsigned long long)(~(uint32_t)(off0 + (size_t)r * p.W));
            float pos = 0.f;
            for (uint32_t b = bits; b; b &= b - 1u) {
                const int c = __ffs(b) - 1;
                const float pc = normalised<ACCURATE>(pst[c * kThreads], norm);
                pos += pc;
                if (group) {
                    const unsigned long long key = ((unsigned long long)__float_as_uint(pc) << 32) | low;
                    if (key > col[c * kThreads]) col[c * kThreads] = key;
                }
            }
            if (do_choice) {
                const float l = -logf(pos + kEps);
                const int n = __popc(bits);
                if (n == 1) { sum_one += l; ++n_one; }
                else if (n > 1) { sum_multi += l; ++n_multi; }
                else { sum_empty += l; ++n_empty; }
            }
        }
        if (cur >= 0) {
            for (int c = 0; c < C; ++c) {
                const unsigned long long e = col[c * kThreads];
                if (e != 0ull) { atomicMax(p.gmax + cur_base + c, e); col[c * kThreads] = 0ull; }
            }
        }
    }

    if (do_choice) {      // one combine per CTA of the persistent grid, in a fixed order (run-to-run identical partials)
        __shared__ float s_sum[kThreads / 32][3];
        __shared__ int s_cnt[kThreads / 32][3];
        float s[3] = {sum_one, sum_multi, sum_empty};
        int n[3] = {n_one, n_multi, n_empty};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s[k] += __shfl_xor_sync(kFullWarp, s[k], o);
                n[k] += __shfl_xor_sync(kFullWarp, n[k], o);
            }
            if (lane == 0) { s_sum[tid >> 5][k] = s[k]; s_cnt[tid >> 5][k] = n[k]; }
        }
        __syncthreads();
        if (tid < 3) {
            double total = 0.0;
            long long count = 0;
            for (int w = 0; w < kThreads / 32; ++w) { total += (double)s_sum[w][tid]; count += s_cnt[w][tid]; }
            if (count != 0) {
                atomicAdd(p.acc + 2 * tid, total);
                atomicAdd(p.acc + 2 * tid + 1, (double)count);
            }
        }
    }
}

// acc[6] += sum of -log(M + eps) over labelled (superpixel, class) pairs with M > 0; acc[7] += their number
__global__ void group_loss_reduce_kernel(const unsigned long long* __restrict__ gmax, const uint32_t* __restrict__ info,
                                         long long n_regions, int C, double* acc) {
    float s = 0.f;
    int n = 0;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_regions; r += (long long)gridDim.x * blockDim.x) {
        uint32_t bits = info[r] & ~kGroupBit;
        while (bits) {
            const int c = __ffs(bits) - 1;
            bits &= bits - 1u;
            const uint32_t pb = (uint32_t)(gmax[r * C + c] >> 32);
            if (pb != 0u) { s += -logf(__uint_as_float(pb) + kEps); ++n; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if ((threadIdx.x & 31) == 0 && n != 0) {
        atomicAdd(acc + 6, (double)s);
        atomicAdd(acc + 7, (double)n);
    }
}

// ------------------------------------------------------------------------------------------ backward
template <int CMAX, bool EXACT, typename IdT>
__global__ void __launch_bounds__(kThreads) multihot_loss_bwd_kernel(const LossParams p) {
    extern __shared__ float pst_all[];     // [C][kThreads] softmax numerators of the pixel being processed
    if (dense_regime(p)) return;          // densely selected batch: the strip walk of losses_dense.cu takes it
    const int C = EXACT ? CMAX : p.C;
    float* pst = pst_all + threadIdx.x;
    const bool do_choice = p.do_choice != 0, do_group = p.do_group != 0;
    const float w_one = p.coef[0] * p.inv_temp, w_multi = p.coef[1] * p.inv_temp, w_group = p.coef[3] * p.inv_temp;
    const size_t P = (size_t)p.H * p.W;
    // rows, planes and the gradient base are 16-byte aligned: whole tiles can be zero-filled with 128-bit stores
    const bool vec_ok = (p.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.grad) & 15) == 0);
    // list mode (sparsely selected batch: decided on the device from the number of active tiles, the same test as
    // grad_zero_kernel): the gradient was zeroed by a linear sweep, only the ACTIVE tiles are computed, dealt by rank.
    // Otherwise every tile is walked and the empty ones are zero-filled here.
    LossParams q = p;
    const long long n_dense = tile_count(p);
    if (p.list_mode) {
        const long long n_list_tiles = list_tile_count(p.n_img, p.H, p.W);
        const long long n_active = __ldg(p.tile_offsets + p.n_groups);
        const int pct = p.dense_percent > 0 ? p.dense_percent : kSparsePercent;
        if (n_active * 100 < n_list_tiles * pct) q.tile_rows = kListRows; else q.list_mode = 0;
    }
    const bool listed = q.list_mode != 0;
    const long long n_work = listed ? (long long)__ldg(p.tile_offsets + p.n_groups) : n_dense;
    const long long warp_stride = (long long)gridDim.x * (kThreads / 32);
    for (long long item = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); item < n_work; item += warp_stride) {
        const long long tile = listed ? list_select(q, (int)item) : item;
        const Tile t = make_tile(q, tile);
        const bool active = t.x < p.W;
        const size_t off0 = (size_t)t.y0 * p.W + t.x;
        const uint32_t mbits = tile_mask_bits(q, t, (size_t)t.img * P + off0);
        if (!listed && item + warp_stride < n_work) {
            const Tile tn = make_tile(q, item + warp_stride);
            prefetch_tile_mask(q, tn, (size_t)tn.img * P + (size_t)tn.y0 * p.W + tn.x);
        }
        const uint32_t any_rows = __reduce_or_sync(kFullWarp, mbits);
        float* tgrad = p.grad + (size_t)t.img * C * P + off0;
        if (any_rows == 0u) {
            if (listed) continue;          // (cannot happen for a listed tile; the sweep zeroed it anyway)          // warp-uniform: nothing selected in the tile -> independent streaming zero stores
            if (vec_ok) {
                // 128-bit stores: lane -> (row l / 8 of a group of 4 rows, 4 pixels): 4 full 128-byte lines per instruction
                const int lane = threadIdx.x & 31;
                const int x4 = (t.x - lane) + (lane & 7) * 4;
                if (x4 < p.W) {
                    float* base = p.grad + (size_t)t.img * C * P + (size_t)t.y0 * p.W + x4;
#pragma unroll 1
                    for (int rg = 0; rg * 4 < t.rows; ++rg) {
                        const int r = rg * 4 + (lane >> 3);
                        if (r < t.rows) {
                            float* ptr = base + (size_t)r * p.W;
#pragma unroll
                            for (int c = 0; c < CMAX; ++c) {
                                if (EXACT || c < C) __stcs(reinterpret_cast<float4*>(ptr), make_float4(0.f, 0.f, 0.f, 0.f));
                                ptr += P;
                            }
                        }
                    }
                }
            } else if (active) {
#pragma unroll 2
                for (int r = 0; r < t.rows; ++r) {
                    float* ptr = tgrad + (size_t)r * p.W;
#pragma unroll
                    for (int c = 0; c < CMAX; ++c) {
                        if (EXACT || c < C) __stcs(ptr, 0.f);
                        ptr += P;
                    }
                }
            }
            continue;
        }
        RowPipe<CMAX, EXACT, IdT> pipe(p);
        pipe.start(t, C, P, off0, mbits);
#pragma unroll 1
        for (int r = 0; r < t.rows; ++r) {
            float v[CMAX];
            int id;
            uint32_t inf;
            pipe.next(v, id, inf);
            const uint32_t bits = inf & ~kGroupBit;
            const bool group = do_group && (inf & kGroupBit) && bits != 0u;
            const bool live = id >= 0 && (group || (do_choice && bits != 0u));     // does any gradient reach this pixel?
            const bool any_live = __any_sync(kFullWarp, live);
            if (!active) continue;
            float* gbase = tgrad + (size_t)r * p.W;
            if (!any_live && listed) continue;                    // already zero (linear sweep)
            if (!any_live) {
                float* ptr = gbase;
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
                    if (EXACT || c < C) __stcs(ptr, 0.f);
                    ptr += P;
                }
                continue;
            }
            const float inv = softmax_numerators<CMAX, false>(v, p.temp, p.scale);
            // d/dx_c = P_c * (a * ([c in row] - pos) + w_group * q_sum) - w_group * [c pooled from this pixel] * q_c
            // with q_c = P_c / (P_c + eps).  Hardly any pixel is the arg-max pixel of a max-pool: those entries are patched.
            float s_in = 0.f, s_out = 0.f;
            uint32_t abits = 0u;   // classes whose max-pooled probability comes from this pixel
            if (live) {
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
                    if (EXACT || c < C) pst[c * kThreads] = v[c];
                }
                const uint32_t low = ~(uint32_t)(off0 + (size_t)r * p.W);
                // the table entry holds ~pixel in its low word (an empty entry holds 0, which no pixel maps to)
                const uint32_t* row = reinterpret_cast<const uint32_t*>(p.gmax + ((long long)t.img * p.S + id) * C);
                float pos = 0.f, q_sum = 0.f;
                for (uint32_t b = bits; b; b &= b - 1u) {
                    const int c = __ffs(b) - 1;
                    const float pc = pst[c * kThreads] * inv;
                    pos += pc;
                    if (group && __ldg(row + 2 * c) == low) {
                        abits |= 1u << c;
                        q_sum += __fdividef(pc, pc + kEps);
                    }
                }
                float a = 0.f;
                if (do_choice) a = -(__popc(bits) == 1 ? w_one : w_multi) / (pos + kEps);
                const float shared_term = abits ? w_group * q_sum : 0.f;
                s_in = (a * (1.f - pos) + shared_term) * inv;
                s_out = (shared_term - a * pos) * inv;
            }
            {
                float* ptr = gbase;
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
                    const float gr = v[c] * (((bits >> c) & 1u) ? s_in : s_out);
                    if (EXACT || c < C) __stcs(ptr, live ? gr : 0.f);
                    ptr += P;
                }
            }
            for (uint32_t b = abits; b; b &= b - 1u) {       // rare: this pixel owns a max-pooled probability
                const int c = __ffs(b) - 1;
                const float e = pst[c * kThreads];
                const float pc = e * inv;
                __stcs(gbase + (size_t)c * P, e * s_in - w_group * __fdividef(pc, pc + kEps));
            }
        }
    }
}

// losses[0] one-hot CE            s0 / (1 + n0)
// losses[1] multi-hot, strict     s1 / (1 + n1)                        (active_joint_multi_lossdecomp.py:67-72)
// losses[2] multi-hot := not one  (s1 + s2) / (1 + n1 + n2)            (..._predignore_lossdecomp.py:65-70)
// losses[3] multi-choice          (s0 + s1) / (1 + n0 + n1)            (utils/loss.py:572-588, empty rows dropped)
// losses[4] group / MIL           s3 / (1 + n3)                        (utils/loss.py:131-141)
// fp32 sums divided by fp32 counters, like the reference's `loss / num_valid`
__global__ void loss_finish_kernel(const double* __restrict__ acc, float* __restrict__ losses) {
    if (threadIdx.x != 0) return;
    const float s0 = (float)acc[0], s1 = (float)acc[2], s2 = (float)acc[4], s3 = (float)acc[6];
    const double n0 = acc[1], n1 = acc[3], n2 = acc[5], n3 = acc[7];
    losses[0] = s0 / (float)(1.0 + n0);
    losses[1] = s1 / (float)(1.0 + n1);
    losses[2] = (s1 + s2) / (float)(1.0 + n1 + n2);
    losses[3] = (s0 + s1) / (float)(1.0 + n0 + n1);
    losses[4] = s3 / (float)(1.0 + n3);
    losses[5] = 0.f;
}

// coef[k] = d (sum_j grad[j] * losses[j]) / d (bucket sum k)
__global__ void loss_coef_kernel(const double* __restrict__ acc, const float* __restrict__ grad, float* __restrict__ coef) {
    if (threadIdx.x != 0) return;
    const double n0 = acc[1], n1 = acc[3], n2 = acc[5], n3 = acc[7];
    const float d0 = (float)(1.0 + n0), d1 = (float)(1.0 + n1), d12 = (float)(1.0 + n1 + n2), d01 = (float)(1.0 + n0 + n1);
    coef[0] = grad[0] / d0 + grad[3] / d01;
    coef[1] = grad[1] / d1 + grad[2] / d12 + grad[3] / d01;
    coef[2] = grad[2] / d12;
    coef[3] = grad[4] / (float)(1.0 + n3);
}

// ------------------------------------------------------------------------------------------ launch
template <typename K>
int resident_ctas(K kernel, size_t smem) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, smem) != cudaSuccess || n < 1) n = 1;
    return n;
}

int env_tile_rows() {
    const char* v = getenv("MAS_LOSS_TILE_ROWS");       // development switch
    const int r = (v && *v) ? atoi(v) : 0;
    return (r == 4 || r == 8 || r == 16) ? r : 0;
}

template <int CMAX, bool EXACT, typename IdT>
cudaError_t launch_one(LossParams p, bool backward, bool accurate, cudaStream_t stream) {
    // forward: running maxima (u64) + softmax numerators (f32) per (class, thread); backward: the numerators
    const size_t smem = (size_t)p.C * kThreads * (backward ? sizeof(float) : sizeof(unsigned long long) + sizeof(float));
    // persistent grid: one wave of resident CTAs (per instantiation; the generic channel padding changes smem by < 2x)
    static mas::PerDeviceInt occ[3];     // per instantiation and device
    const int which = backward ? 0 : (accurate ? 1 : 2);
    const int dev = mas::current_device();
    int per_sm_now = occ[which].get(dev);
    if (per_sm_now == 0) {
        const size_t smem_max = (size_t)CMAX * kThreads * (backward ? sizeof(float) : sizeof(unsigned long long) + sizeof(float));
        if (backward) per_sm_now = resident_ctas(multihot_loss_bwd_kernel<CMAX, EXACT, IdT>, smem_max);
        else if (accurate) per_sm_now = resident_ctas(multihot_loss_fwd_kernel<CMAX, EXACT, IdT, true>, smem_max);
        else per_sm_now = resident_ctas(multihot_loss_fwd_kernel<CMAX, EXACT, IdT, false>, smem_max);
        occ[which].set(dev, per_sm_now);
    }
    // Tile height: the cost of a tile is the number of its rows that hold selected pixels, served one after the other by
    // ONE warp, so short tiles spread the selected superpixels over more warps (what matters when few pixels are
    // labelled); tall tiles flush the private maxima less often.  Aim at >= 8 tiles per resident warp.
    const long long warps = (long long)mas::sm_count() * per_sm_now * (kThreads / 32);
    const long long strips = (long long)p.n_img * ((p.W + 31) / 32);
    int rows = kTileRows;
    while (rows > 4 && strips * ((p.H + rows - 1) / rows) < 8 * warps) rows >>= 1;
    if (env_tile_rows()) rows = env_tile_rows();
    p.tile_rows = rows;
    if (p.list_mode && !backward) p.tile_rows = kListRows;      // forward: always the active-tile list (backward decides on the device)
    const long long tiles = strips * ((p.H + rows - 1) / rows);
    const long long want = (tiles + kThreads / 32 - 1) / (kThreads / 32);
    const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)mas::sm_count() * per_sm_now));
    if (backward && p.list_mode) {
        const size_t n = (size_t)p.n_img * p.C * p.H * p.W;
        const unsigned zblocks = (unsigned)std::min<size_t>((n / 4 + 255) / 256 + 1, (size_t)mas::sm_count() * 16);
        grad_zero_kernel<<<zblocks, 256, 0, stream>>>(p.grad, n, p.tile_offsets, p.n_groups, strips * ((p.H + kListRows - 1) / kListRows),
                                                      p.dense_percent > 0 ? p.dense_percent : kSparsePercent);
        mas::count_launches(1);
    }
    if (backward) {
        multihot_loss_bwd_kernel<CMAX, EXACT, IdT><<<blocks, kThreads, smem, stream>>>(p);
    } else {
        if (accurate) multihot_loss_fwd_kernel<CMAX, EXACT, IdT, true><<<blocks, kThreads, smem, stream>>>(p);
        else multihot_loss_fwd_kernel<CMAX, EXACT, IdT, false><<<blocks, kThreads, smem, stream>>>(p);
    }
    mas::count_launches(1);
    return cudaGetLastError();
}

template <typename IdT>
cudaError_t dispatch_channels(const LossParams& p, bool backward, bool accurate, cudaStream_t stream) {
    switch (p.C) {
        case 19: return launch_one<19, true, IdT>(p, backward, accurate, stream);
        case 20: return launch_one<20, true, IdT>(p, backward, accurate, stream);
        case 21: return launch_one<21, true, IdT>(p, backward, accurate, stream);
        case 22: return launch_one<22, true, IdT>(p, backward, accurate, stream);
        default: break;
    }
    if (p.C <= 8) return launch_one<8, false, IdT>(p, backward, accurate, stream);
    if (p.C <= 16) return launch_one<16, false, IdT>(p, backward, accurate, stream);
    if (p.C <= 24) return launch_one<24, false, IdT>(p, backward, accurate, stream);
    return launch_one<31, false, IdT>(p, backward, accurate, stream);
}

// MAS_LOSS_DENSE: unset = the device picks per batch (needs the active-tile list), 0 = never, 1 = always when the shape
// allows (tests, comparisons).  MAS_LOSS_DENSE_PCT: share of active tiles from which the dense kernels take over.
constexpr int kDensePercent = 40;

cudaError_t dispatch(LossParams p, int ids_dtype, bool backward, bool accurate, cudaStream_t stream) {
    const char* mode_env = getenv("MAS_LOSS_DENSE");
    const int mode = (mode_env && *mode_env) ? atoi(mode_env) : -1;
    p.dense_percent = 0;
    if (!accurate && mode != 0 && (p.list_mode || mode == 1)) {
        LossParams d = p;
        if (mode == 1) {
            d.list_mode = 0;      // runs unconditionally
        } else {
            const char* pct_env = getenv("MAS_LOSS_DENSE_PCT");
            d.dense_percent = (pct_env && *pct_env) ? std::min(100, std::max(1, atoi(pct_env))) : kDensePercent;
        }
        bool launched = false;
        cudaError_t e = launch_dense(d, ids_dtype, backward, stream, &launched);
        if (e != cudaSuccess) return e;
        if (launched) {
            if (mode == 1) return cudaSuccess;
            p.dense_percent = d.dense_percent;      // the list kernels below leave densely selected batches alone
        }
    }
    if (ids_dtype == MAS_I64) return dispatch_channels<long long>(p, backward, accurate, stream);
    return dispatch_channels<int32_t>(p, backward, accurate, stream);
}

int check_common(const char* what, const void* logits, const void* ids, int ids_dtype, const uint8_t* mask, const uint32_t* info,
                 int n_img, int channels, int height, int width, int nseg, float temperature, int flags) {
    MAS_REQUIRE(logits && ids && mask && info, MAS_E_BADARG, "%s: null pointer", what);
    MAS_REQUIRE(n_img >= 0 && height > 0 && width > 0 && nseg > 0, MAS_E_BADARG, "%s: bad shape", what);
    MAS_REQUIRE(channels >= 2 && channels <= MAS_MAX_LOSS_CLASSES, MAS_E_RANGE, "%s: channels=%d outside [2,%d]", what, channels,
                MAS_MAX_LOSS_CLASSES);
    MAS_REQUIRE(ids_dtype == MAS_I32 || ids_dtype == MAS_I64, MAS_E_BADARG, "%s: bad ids dtype", what);
    MAS_REQUIRE(temperature > 0.f, MAS_E_BADARG, "%s: temperature must be > 0", what);
    MAS_REQUIRE((flags & ~(MAS_LOSS_CHOICE | MAS_LOSS_GROUP | MAS_LOSS_EXACT_SOFTMAX)) == 0 && (flags & (MAS_LOSS_CHOICE | MAS_LOSS_GROUP)) != 0, MAS_E_BADARG, "%s: bad flags", what);
    return 0;
}

}  // namespace

extern "C" int mas_multihot_info_dev(const uint8_t* targets, int64_t n_regions, int target_channels, int channels, int group_mode,
                                     uint32_t* info, void* stream) {
    MAS_REQUIRE(targets && info, MAS_E_BADARG, "multihot_info: null pointer");
    MAS_REQUIRE(n_regions >= 0 && target_channels >= 1, MAS_E_BADARG, "multihot_info: bad shape");
    MAS_REQUIRE(channels >= 1 && channels <= MAS_MAX_LOSS_CLASSES && channels <= target_channels, MAS_E_RANGE,
                "multihot_info: channels=%d must be in [1,%d] and <= target_channels", channels, MAS_MAX_LOSS_CLASSES);
    MAS_REQUIRE(group_mode == MAS_GROUP_ALL || group_mode == MAS_GROUP_ONLYMULTI, MAS_E_BADARG, "multihot_info: bad group_mode");
    if (n_regions == 0) return 0;
    const int threads = 256;
    multihot_info_kernel<<<(unsigned)((n_regions + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        targets, n_regions, target_channels, channels, group_mode, info);
    mas::count_launches(1);
    MAS_LAUNCH_OK("multihot_info_kernel");
    return 0;
}

namespace {

// workspace of the active-tile list: [ticket u32, pad x3][words: n_groups u32][offsets: n_groups + 1 i32]
struct TileList {
    unsigned int* ticket; uint32_t* words; int* offsets; int n_groups; size_t bytes;
};

TileList carve_tiles(void* base, int n_img, int height, int width) {
    const long long n_tiles = (long long)n_img * ((width + 31) / 32) * ((height + kListRows - 1) / kListRows);
    TileList t;
    t.n_groups = (int)((n_tiles + 31) / 32);
    char* b = reinterpret_cast<char*>(base);
    t.ticket = reinterpret_cast<unsigned int*>(b);
    t.words = reinterpret_cast<uint32_t*>(b ? b + 16 : nullptr);
    t.offsets = reinterpret_cast<int*>(b ? b + 16 + (size_t)t.n_groups * 4 : nullptr);
    t.bytes = 16 + (size_t)t.n_groups * 4 + ((size_t)t.n_groups + 1) * 4;
    return t;
}

void attach_tiles(LossParams& p, const void* tiles) {
    if (!tiles) return;
    const TileList t = carve_tiles(const_cast<void*>(tiles), p.n_img, p.H, p.W);
    p.tile_words = t.words; p.tile_offsets = t.offsets; p.n_groups = t.n_groups; p.list_mode = 1;
}

}  // namespace

extern "C" size_t mas_multihot_tiles_workspace_bytes(int n_img, int height, int width) {
    if (n_img <= 0 || height <= 0 || width <= 0) return 0;
    return carve_tiles(nullptr, n_img, height, width).bytes;
}

extern "C" int mas_multihot_tiles_dev(const uint8_t* mask, int n_img, int height, int width, void* tiles, size_t tiles_bytes, void* stream) {
    MAS_REQUIRE(mask && tiles, MAS_E_BADARG, "multihot_tiles: null pointer");
    MAS_REQUIRE(n_img > 0 && height > 0 && width > 0, MAS_E_BADARG, "multihot_tiles: bad shape");
    MAS_REQUIRE(((uintptr_t)tiles) % 16 == 0, MAS_E_BADARG, "multihot_tiles: workspace must be 16-byte aligned");
    const TileList t = carve_tiles(tiles, n_img, height, width);
    MAS_REQUIRE(tiles_bytes >= t.bytes, MAS_E_WORKSPACE, "multihot_tiles: workspace too small (%zu < %zu)", tiles_bytes, t.bytes);
    MAS_REQUIRE((long long)t.n_groups * 32 < (1ll << 31), MAS_E_RANGE, "multihot_tiles: too many tiles");
    cudaStream_t st = (cudaStream_t)stream;
    MAS_CUDA_OK(cudaMemsetAsync(t.ticket, 0, 16, st));
    const unsigned blocks = (unsigned)std::max(1, std::min((t.n_groups + 7) / 8, mas::sm_count() * 8));
    tile_scan_kernel<<<blocks, 256, 0, st>>>(mask, n_img, height, width, t.n_groups, t.words, t.offsets, t.ticket);
    mas::count_launches(1);
    MAS_LAUNCH_OK("tile_scan_kernel");
    return 0;
}

namespace mas {
// forward pass with the group-loss reduction optional: the stage-2 labeller only needs the packed maxima (arg-max pixels)
int multihot_loss_fwd(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask, const uint32_t* info, int n_img,
                      int channels, int height, int width, int nseg, float temperature, int flags, double* acc,
                      uint64_t* group_max, bool reduce_group, void* stream, const void* tiles);
}  // namespace mas

extern "C" int mas_multihot_loss_fwd_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask,
                                         const uint32_t* info, int n_img, int channels, int height, int width, int nseg,
                                         float temperature, int flags, double* acc, uint64_t* group_max, void* stream) {
    return mas::multihot_loss_fwd(logits, ids, ids_dtype, mask, info, n_img, channels, height, width, nseg, temperature, flags, acc,
                                  group_max, true, stream, nullptr);
}

extern "C" int mas_multihot_loss_fwd_tiles_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask,
                                               const uint32_t* info, const void* tiles, int n_img, int channels, int height, int width,
                                               int nseg, float temperature, int flags, double* acc, uint64_t* group_max, void* stream) {
    return mas::multihot_loss_fwd(logits, ids, ids_dtype, mask, info, n_img, channels, height, width, nseg, temperature, flags, acc,
                                  group_max, true, stream, tiles);
}

int mas::multihot_loss_fwd(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask, const uint32_t* info, int n_img,
                           int channels, int height, int width, int nseg, float temperature, int flags, double* acc,
                           uint64_t* group_max, bool reduce_group, void* stream, const void* tiles) {
    int rc = check_common("multihot_loss_fwd", logits, ids, ids_dtype, mask, info, n_img, channels, height, width, nseg, temperature, flags);
    if (rc != 0) return rc;
    MAS_REQUIRE(acc, MAS_E_BADARG, "multihot_loss_fwd: null acc");
    MAS_REQUIRE(!(flags & MAS_LOSS_GROUP) || group_max, MAS_E_BADARG, "multihot_loss_fwd: group_max required with MAS_LOSS_GROUP");
    if (n_img == 0) return 0;
    LossParams p = {};
    p.logits = logits; p.ids = ids; p.mask = mask; p.info = info;
    p.n_img = n_img; p.C = channels; p.H = height; p.W = width; p.S = nseg;
    p.temp = temperature; p.inv_temp = 1.f / temperature; p.scale = 1.4426950408889634f / temperature;
    p.do_choice = (flags & MAS_LOSS_CHOICE) ? 1 : 0; p.do_group = (flags & MAS_LOSS_GROUP) ? 1 : 0;
    p.acc = acc; p.gmax = reinterpret_cast<unsigned long long*>(group_max);
    attach_tiles(p, tiles);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = dispatch(p, ids_dtype, false, (flags & MAS_LOSS_EXACT_SOFTMAX) != 0, st);
    if (e != cudaSuccess) return mas::cuda_fail(e, "multihot_loss_fwd_kernel launch");
    if (p.do_group && reduce_group) {
        const long long n_regions = (long long)n_img * nseg;
        const int threads = 256;
        const long long blocks = std::min<long long>((n_regions + threads - 1) / threads, (long long)mas::sm_count() * 4);
        group_loss_reduce_kernel<<<(unsigned)blocks, threads, 0, st>>>(p.gmax, info, n_regions, channels, acc);
        mas::count_launches(1);
        MAS_LAUNCH_OK("group_loss_reduce_kernel");
    }
    return 0;
}

extern "C" int mas_multihot_loss_bwd_tiles_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask,
                                               const uint32_t* info, const void* tiles, const uint64_t* group_max, const float* coef,
                                               int n_img, int channels, int height, int width, int nseg, float temperature, int flags,
                                               float* grad_logits, void* stream);

extern "C" int mas_multihot_loss_bwd_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask,
                                         const uint32_t* info, const uint64_t* group_max, const float* coef,
                                         int n_img, int channels, int height, int width, int nseg, float temperature, int flags,
                                         float* grad_logits, void* stream) {
    return mas_multihot_loss_bwd_tiles_dev(logits, ids, ids_dtype, mask, info, nullptr, group_max, coef, n_img, channels, height, width,
                                           nseg, temperature, flags, grad_logits, stream);
}

extern "C" int mas_multihot_loss_bwd_tiles_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask,
                                               const uint32_t* info, const void* tiles, const uint64_t* group_max, const float* coef,
                                               int n_img, int channels, int height, int width, int nseg, float temperature, int flags,
                                               float* grad_logits, void* stream) {
    int rc = check_common("multihot_loss_bwd", logits, ids, ids_dtype, mask, info, n_img, channels, height, width, nseg, temperature, flags);
    if (rc != 0) return rc;
    MAS_REQUIRE(coef && grad_logits, MAS_E_BADARG, "multihot_loss_bwd: null pointer");
    MAS_REQUIRE(!(flags & MAS_LOSS_GROUP) || group_max, MAS_E_BADARG, "multihot_loss_bwd: group_max required with MAS_LOSS_GROUP");
    if (n_img == 0) return 0;
    LossParams p = {};
    p.logits = logits; p.ids = ids; p.mask = mask; p.info = info;
    p.n_img = n_img; p.C = channels; p.H = height; p.W = width; p.S = nseg;
    p.temp = temperature; p.inv_temp = 1.f / temperature; p.scale = 1.4426950408889634f / temperature;
    p.do_choice = (flags & MAS_LOSS_CHOICE) ? 1 : 0; p.do_group = (flags & MAS_LOSS_GROUP) ? 1 : 0;
    p.gmax = const_cast<unsigned long long*>(reinterpret_cast<const unsigned long long*>(group_max));
    p.coef = coef; p.grad = grad_logits;
    attach_tiles(p, tiles);
    cudaError_t e = dispatch(p, ids_dtype, true, false, (cudaStream_t)stream);
    if (e != cudaSuccess) return mas::cuda_fail(e, "multihot_loss_bwd_kernel launch");
    return 0;
}

extern "C" int mas_multihot_loss_finish_dev(const double* acc, float* losses, void* stream) {
    MAS_REQUIRE(acc && losses, MAS_E_BADARG, "multihot_loss_finish: null pointer");
    loss_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc, losses);
    mas::count_launches(1);
    MAS_LAUNCH_OK("loss_finish_kernel");
    return 0;
}

extern "C" int mas_multihot_loss_coef_dev(const double* acc, const float* grad_losses, float* coef, void* stream) {
    MAS_REQUIRE(acc && grad_losses && coef, MAS_E_BADARG, "multihot_loss_coef: null pointer");
    loss_coef_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc, grad_losses, coef);
    mas::count_launches(1);
    MAS_LAUNCH_OK("loss_coef_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------ whole loss step in two calls
// The trainer's step (trainer/active_joint_multi_predignore_lossdecomp.py:101-107) through the entry points above is five
// host calls forward and three backward plus their temporaries -- at a few percent of labelled pixels that host work
// (0.4 ms in Python) exceeds the 0.25 ms the kernels need.  These two entries run the same launches from ONE call each,
// on one caller-provided workspace:
//   [0, 64)    acc      8 doubles (bucket sums / counts)          [64, 96)  losses  6 floats (+ pad)
//   [96, 128)  coef     4 floats (backward)                       then: candidate words, active-tile list, packed maxima
namespace {

struct StepSpace {
    double* acc; float* losses; float* coef; uint32_t* info; void* tiles; uint64_t* gmax; size_t tiles_bytes, zero_from, bytes;
};

size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }

StepSpace carve_step(void* base, int n_img, int channels, int height, int width, int nseg) {
    char* b = reinterpret_cast<char*>(base);
    StepSpace w;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* q = b ? b + off : nullptr; off += up16(bytes); return q; };
    w.acc = reinterpret_cast<double*>(take(64));
    w.losses = reinterpret_cast<float*>(take(32));
    w.coef = reinterpret_cast<float*>(take(32));
    w.info = reinterpret_cast<uint32_t*>(take((size_t)n_img * nseg * 4));
    w.tiles_bytes = mas_multihot_tiles_workspace_bytes(n_img, height, width);
    w.tiles = take(w.tiles_bytes);
    w.zero_from = off;
    w.gmax = reinterpret_cast<uint64_t*>(take((size_t)n_img * nseg * channels * 8));
    w.bytes = off;
    return w;
}

// coef[k] = d (sum_j grad_j * losses[j]) / d (bucket sum k) with the incoming gradients given as six device pointers
// (NULL = that loss does not take part)
struct GradPtrs { const float* g[6]; };

__global__ void loss_coef_ptr_kernel(const double* __restrict__ acc, GradPtrs grads, float* __restrict__ coef) {
    if (threadIdx.x != 0) return;
    float g[6];
    for (int j = 0; j < 6; ++j) g[j] = grads.g[j] ? *grads.g[j] : 0.f;
    const double n0 = acc[1], n1 = acc[3], n2 = acc[5], n3 = acc[7];
    const float d0 = (float)(1.0 + n0), d1 = (float)(1.0 + n1), d12 = (float)(1.0 + n1 + n2), d01 = (float)(1.0 + n0 + n1);
    coef[0] = g[0] / d0 + g[3] / d01;
    coef[1] = g[1] / d1 + g[2] / d12 + g[3] / d01;
    coef[2] = g[2] / d12;
    coef[3] = g[4] / (float)(1.0 + n3);
}

}  // namespace

extern "C" size_t mas_stage1_workspace_bytes(int n_img, int channels, int height, int width, int nseg) {
    if (n_img <= 0 || channels <= 0 || height <= 0 || width <= 0 || nseg <= 0) return 0;
    return carve_step(nullptr, n_img, channels, height, width, nseg).bytes;
}

extern "C" int mas_stage1_loss_fwd_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask, const uint8_t* targets,
                                       int target_channels, int n_img, int channels, int height, int width, int nseg,
                                       float temperature, int group_mode, int flags, void* workspace, size_t workspace_bytes,
                                       void* stream) {
    MAS_REQUIRE(logits && ids && mask && targets && workspace, MAS_E_BADARG, "stage1_loss_fwd: null pointer");
    MAS_REQUIRE(n_img > 0 && height > 0 && width > 0 && nseg > 0, MAS_E_BADARG, "stage1_loss_fwd: bad shape");
    MAS_REQUIRE(((uintptr_t)workspace) % 16 == 0, MAS_E_BADARG, "stage1_loss_fwd: workspace must be 16-byte aligned");
    const StepSpace w = carve_step(workspace, n_img, channels, height, width, nseg);
    MAS_REQUIRE(workspace_bytes >= w.bytes, MAS_E_WORKSPACE, "stage1_loss_fwd: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    cudaStream_t st = (cudaStream_t)stream;
    MAS_CUDA_OK(cudaMemsetAsync(workspace, 0, 128, st));                                            // acc, losses, coef
    if (flags & MAS_LOSS_GROUP) MAS_CUDA_OK(cudaMemsetAsync(w.gmax, 0, w.bytes - w.zero_from, st));
    int rc = mas_multihot_info_dev(targets, (int64_t)n_img * nseg, target_channels, channels, group_mode, w.info, stream);
    if (rc != 0) return rc;
    rc = mas_multihot_tiles_dev(mask, n_img, height, width, w.tiles, w.tiles_bytes, stream);
    if (rc != 0) return rc;
    rc = mas::multihot_loss_fwd(logits, ids, ids_dtype, mask, w.info, n_img, channels, height, width, nseg, temperature, flags, w.acc,
                                (flags & MAS_LOSS_GROUP) ? w.gmax : nullptr, true, stream, w.tiles);
    if (rc != 0) return rc;
    return mas_multihot_loss_finish_dev(w.acc, w.losses, stream);
}

extern "C" int mas_stage1_loss_bwd_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask, int n_img, int channels,
                                       int height, int width, int nseg, float temperature, int flags, const void* workspace,
                                       const float* const* grad_losses, float* grad_logits, void* stream) {
    MAS_REQUIRE(logits && ids && mask && workspace && grad_losses && grad_logits, MAS_E_BADARG, "stage1_loss_bwd: null pointer");
    MAS_REQUIRE(n_img > 0 && height > 0 && width > 0 && nseg > 0, MAS_E_BADARG, "stage1_loss_bwd: bad shape");
    const StepSpace w = carve_step(const_cast<void*>(workspace), n_img, channels, height, width, nseg);
    GradPtrs grads;
    for (int j = 0; j < 6; ++j) grads.g[j] = grad_losses[j];
    loss_coef_ptr_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(w.acc, grads, w.coef);
    mas::count_launches(1);
    MAS_LAUNCH_OK("loss_coef_ptr_kernel");
    return mas_multihot_loss_bwd_tiles_dev(logits, ids, ids_dtype, mask, w.info, w.tiles, (flags & MAS_LOSS_GROUP) ? w.gmax : nullptr, w.coef,
                                           n_img, channels, height, width, nseg, temperature, flags & ~MAS_LOSS_EXACT_SOFTMAX, grad_logits,
                                           stream);
}
