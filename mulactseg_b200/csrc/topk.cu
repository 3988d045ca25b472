// Top-k region selection: 64-bit (score, tie-break) keys, MSB radix select, bitonic sort of the survivors.
//
// Replaces `sorted(scores, reverse=True)` over 6 M Python tuples (active_selection/base.py:37): only the
// prefix that expand_training_set can consume (budget + 1 regions) is ever needed, so the device selects
// the k largest keys exactly (keys are distinct by construction) and sorts just those.
#include "common.cuh"

#include <algorithm>

namespace {

struct SelectState {
    unsigned long long prefix;  // bits decided so far (threshold key once all 8 digits are fixed)
    unsigned long long mask;    // which bits of `prefix` are decided
    long long remaining;        // rank still to locate inside the current bucket (1-based)
    unsigned int out_count;
    unsigned int take_all;      // fewer than k valid keys: everything non-zero is selected
    unsigned int hist[256];
};

__global__ void region_keys_kernel(const float* __restrict__ score, const uint8_t* __restrict__ in_pool,
                                   const int32_t* __restrict__ image_rank, long long n_regions, int nseg,
                                   unsigned long long* __restrict__ keys) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_regions) return;
    unsigned long long key = 0ull;
    if (in_pool[r]) {
        const long long img = r / nseg;
        const unsigned int tie = (unsigned int)image_rank[img] * (unsigned int)nseg + (unsigned int)(r - img * nseg);
        key = ((unsigned long long)mas::ordered_bits(score[r]) << 32) | tie;
    }
    keys[r] = key;
}

__global__ void select_init_kernel(SelectState* st, long long k) {
    const int t = threadIdx.x;
    if (t == 0) {
        st->prefix = 0ull; st->mask = 0ull; st->remaining = k; st->out_count = 0u; st->take_all = 0u;
    }
    if (t < 256) st->hist[t] = 0u;
}

__global__ void select_hist_kernel(const unsigned long long* __restrict__ keys, long long n, SelectState* st, int shift) {
    __shared__ unsigned int local[256];
    local[threadIdx.x & 255] = 0u;
    __syncthreads();
    const unsigned long long prefix = st->prefix, mask = st->mask;
    if (!st->take_all) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            const unsigned long long key = keys[i];
            if (key != 0ull && (key & mask) == prefix) atomicAdd(&local[(unsigned)(key >> shift) & 255u], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < 256 && local[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], local[threadIdx.x]);
}

// one warp: walk the digit histogram from the top, fix the digit holding the `remaining`-th largest key
__global__ void select_pick_kernel(SelectState* st, int shift) {
    if (threadIdx.x == 0 && !st->take_all) {
        long long remaining = st->remaining;
        int digit = -1;
        for (int d = 255; d >= 0; --d) {
            const long long c = st->hist[d];
            if (remaining <= c) { digit = d; break; }
            remaining -= c;
        }
        if (digit < 0) {
            st->take_all = 1u;  // fewer candidates than k
        } else {
            st->prefix |= (unsigned long long)digit << shift;
            st->mask |= 0xffull << shift;
            st->remaining = remaining;
        }
    }
    __syncwarp();
    for (int d = threadIdx.x; d < 256; d += blockDim.x) st->hist[d] = 0u;
}

__global__ void select_compact_kernel(const unsigned long long* __restrict__ keys, long long n, SelectState* st,
                                      unsigned long long* __restrict__ out, long long k) {
    const unsigned long long thr = st->take_all ? 1ull : st->prefix;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        if (key != 0ull && key >= thr) {
            const unsigned int slot = atomicAdd(&st->out_count, 1u);
            if ((long long)slot < k) out[slot] = key;
        }
    }
}

__global__ void select_finish_kernel(const SelectState* st, int32_t* out_count, long long k) {
    *out_count = (int32_t)min((long long)st->out_count, k);
}

// ---------------------------------------------------------------------------------------- fast path: bucket + sort
// Two histogram passes (key bits 63..52, then 51..40 inside the bucket of the k-th largest key) pin down the first 24
// bits of the k-th key: sign, exponent and 15 mantissa bits of its score.  ONE compaction pass then keeps every key from
// that prefix upwards -- k plus the few keys whose scores agree with the k-th to 3e-5 relative -- and the bitonic sort
// that has to run anyway orders them.  If the candidates do not fit the sort buffer (massively tied scores) the caller
// falls back to the exact 8-pass radix select above.
constexpr int kFastBits = 12;
constexpr int kFastBins = 1 << kFastBits;

struct FastState {
    unsigned int hist[2][kFastBins];
    unsigned int ticket[2];
    unsigned int count;            // candidates written (may exceed the capacity: overflow)
    unsigned int take_all;         // fewer non-zero keys than k
    long long remaining;           // rank of the k-th key inside the level-0 bucket (1-based)
    unsigned long long thr;        // smallest key prefix kept
};

// level 0: histogram of key bits 63..52 over all non-zero keys; level 1: bits 51..40 over the keys of the level-0 bucket.
// The last CTA to finish locates the bin holding the wanted rank and extends the threshold prefix.
__global__ void __launch_bounds__(256) fast_hist_kernel(const unsigned long long* __restrict__ keys, long long n, FastState* st,
                                                        long long k, int level) {
    __shared__ unsigned int local[kFastBins];
    __shared__ unsigned int part[256];
    __shared__ bool last;
    if (level == 1 && st->take_all) return;
    const unsigned long long bucket = st->thr >> (64 - kFastBits);     // level 1 only
    for (int i = threadIdx.x; i < kFastBins; i += 256) local[i] = 0u;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const unsigned long long key = keys[i];
        if (key == 0ull) continue;
        if (level == 0) atomicAdd(&local[(unsigned)(key >> (64 - kFastBits))], 1u);
        else if ((key >> (64 - kFastBits)) == bucket) atomicAdd(&local[(unsigned)(key >> (64 - 2 * kFastBits)) & (kFastBins - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kFastBins; i += 256) {
        if (local[i]) atomicAdd(&st->hist[level][i], local[i]);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&st->ticket[level], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    // thread t owns bins [16 t, 16 t + 16); walk from the top
    constexpr int kPer = kFastBins / 256;
    const long long want = level == 0 ? k : st->remaining;
    unsigned int h[kPer];
    unsigned int mine = 0u;
#pragma unroll
    for (int i = 0; i < kPer; ++i) { h[i] = __ldcg(&st->hist[level][threadIdx.x * kPer + i]); mine += h[i]; }
    part[threadIdx.x] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {          // suffix sums over 256 partials
        unsigned int run = 0u;
        for (int t = 255; t >= 0; --t) { const unsigned int v = part[t]; part[t] = run; run += v; }
        if ((long long)run < want) { st->take_all = 1u; st->thr = 1ull; }      // (level 0 only) keep every non-zero key
    }
    __syncthreads();
    long long above = part[threadIdx.x];                  // keys in bins above this thread's range
#pragma unroll
    for (int i = kPer - 1; i >= 0; --i) {
        if (above < want && above + (long long)h[i] >= want) {
            const unsigned long long bin = (unsigned long long)(threadIdx.x * kPer + i);
            if (level == 0) { st->thr = bin << (64 - kFastBits); st->remaining = want - above; }
            else st->thr |= bin << (64 - 2 * kFastBits);
        }
        above += h[i];
    }
}

__global__ void __launch_bounds__(256) fast_compact_kernel(const unsigned long long* __restrict__ keys, long long n, FastState* st,
                                                           unsigned long long* __restrict__ out, long long capacity) {
    const unsigned long long thr = max(st->thr, 1ull);
    const int lane = threadIdx.x & 31;
    const long long padded = (n + 31) & ~31ll;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < padded; i += (long long)gridDim.x * 256) {
        const unsigned long long key = i < n ? keys[i] : 0ull;
        const bool keep = key >= thr;
        const unsigned int votes = __ballot_sync(0xffffffffu, keep);
        if (votes == 0u) continue;
        unsigned int base = 0u;
        if (lane == __ffs(votes) - 1) base = atomicAdd(&st->count, (unsigned int)__popc(votes));
        base = __shfl_sync(0xffffffffu, base, __ffs(votes) - 1);
        const long long slot = (long long)base + __popc(votes & ((1u << lane) - 1u));
        if (keep && slot < capacity) out[slot] = key;
    }
}

__global__ void fast_finish_kernel(const FastState* st, int32_t* out_count, long long k, long long capacity, int clamp,
                                   long long* msg_count) {
    const long long c = st->count;
    const int32_t result = c > capacity ? -1 : (int32_t)(clamp ? min(c, k) : c);
    *out_count = result;
    if (msg_count) *msg_count = (long long)result;      // the count travels in the slot after the candidates
}

// Multi-GPU merge: the all-gathered buffer holds, per rank, `stride - 1` candidate slots followed by that rank's count
// (-1 = its buffer overflowed).  One warp reduces the counts to their minimum and clears the count slots, so that the
// whole buffer can be handed to the bucket select as plain keys (0 = "no key").
__global__ void merge_counts_kernel(unsigned long long* gathered, int world, long long stride, int32_t* worst) {
    long long mine = 0x7fffffffll;
    for (int r = threadIdx.x; r < world; r += 32) {
        unsigned long long* slot = gathered + (long long)r * stride + (stride - 1);
        mine = min(mine, (long long)*slot);
        *slot = 0ull;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
    if (threadIdx.x == 0) *worst = (int32_t)mine;
}

// ---------------------------------------------------------------------------------------- budget cut
// RegionActiveDataset.expand_training_set walks the sorted list and stops AFTER the pick that makes the running label
// cost exceed the budget (strict '>', dataloader/region_active_dataset.py:56-66).  One CTA: gather the cost of every
// ranked region, block-wide inclusive scan with a running carry, first index over the budget.
constexpr int kCutThreads = 1024;
constexpr int kCutPer = 16;        // 16 independent key loads + cost gathers in flight per thread and pass (4: 49 us per round, latency-bound)

__global__ void __launch_bounds__(kCutThreads) prefix_cut_kernel(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ count,
                                                                 const uint8_t* __restrict__ cost_by_tie, long long n_cost,
                                                                 long long budget, int32_t* __restrict__ n_take) {
    __shared__ long long warp_tot[kCutThreads / 32];
    __shared__ long long carry;
    __shared__ int first;
    const int n = *count;
    if (n < 0) { if (threadIdx.x == 0) *n_take = -1; return; }
    if (threadIdx.x == 0) { carry = 0; first = n; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n; base += kCutThreads * kCutPer) {
        const int i0 = base + threadIdx.x * kCutPer;
        int c[kCutPer];
        long long mine = 0;
#pragma unroll
        for (int j = 0; j < kCutPer; ++j) {
            const int i = i0 + j;
            long long tie = i < n ? (long long)(keys[i] & 0xffffffffull) : -1;
            c[j] = (tie >= 0 && tie < n_cost) ? (int)cost_by_tie[tie] : 0;
            mine += c[j];
        }
        long long incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        long long before = carry + incl - mine;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        long long running = before;
#pragma unroll
        for (int j = 0; j < kCutPer; ++j) {
            running += c[j];
            if (i0 + j < n && running > budget) { atomicMin(&first, i0 + j); break; }
        }
        __syncthreads();
        if (threadIdx.x == kCutThreads - 1) carry = before + mine;
        __syncthreads();
        if (first < n) break;
    }
    if (threadIdx.x == 0) *n_take = first < n ? first + 1 : n;
}

// ---------------------------------------------------------------------------------------- bitonic sort (descending)
constexpr int kSortThreads = 1024;
constexpr int kSortTile = 8192;            // keys per CTA: 64 KB of shared memory, 4 comparators per thread and step

__device__ __forceinline__ void cmp_swap_desc(unsigned long long& a, unsigned long long& b, bool descending) {
    if ((a < b) == descending) { const unsigned long long t = a; a = b; b = t; }
}

__global__ void sort_pad_kernel(unsigned long long* keys, long long n, long long n_pad) {
    const long long i = n + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) keys[i] = 0ull;
}

// strides j_from .. 1 of stage k inside the tile held in shared memory
__device__ __forceinline__ void tile_steps(unsigned long long* s, long long tile_base, long long k, int j_from) {
    for (int j = j_from; j > 0; j >>= 1) {
        __syncthreads();
        for (int c = threadIdx.x; c < kSortTile / 2; c += kSortThreads) {
            const int i = 2 * c - (c & (j - 1));
            cmp_swap_desc(s[i], s[i + j], ((tile_base + i) & k) == 0);
        }
    }
    __syncthreads();
}

// sort each tile of kSortTile keys (stages 2 .. kSortTile); the direction of a tile follows the global index
__global__ void __launch_bounds__(kSortThreads) sort_tile_kernel(unsigned long long* keys) {
    extern __shared__ unsigned long long s[];
    const long long base = (long long)blockIdx.x * kSortTile;
    for (int i = threadIdx.x; i < kSortTile; i += kSortThreads) s[i] = keys[base + i];
    for (int k = 2; k <= kSortTile; k <<= 1) tile_steps(s, base, k, k >> 1);
    for (int i = threadIdx.x; i < kSortTile; i += kSortThreads) keys[base + i] = s[i];
}

// finish stage k inside a tile: strides kSortTile / 2 .. 1
__global__ void __launch_bounds__(kSortThreads) sort_tile_merge_kernel(unsigned long long* keys, long long k) {
    extern __shared__ unsigned long long s[];
    const long long base = (long long)blockIdx.x * kSortTile;
    for (int i = threadIdx.x; i < kSortTile; i += kSortThreads) s[i] = keys[base + i];
    tile_steps(s, base, k, kSortTile / 2);
    for (int i = threadIdx.x; i < kSortTile; i += kSortThreads) keys[base + i] = s[i];
}

// one stride (j >= kSortTile) of stage k across tiles
__global__ void sort_global_step_kernel(unsigned long long* keys, long long n_pad, long long k, long long j) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pad / 2) return;
    const long long i = 2 * t - (t & (j - 1));
    unsigned long long a = keys[i], b = keys[i + j];
    const bool desc = (i & k) == 0;
    if ((a < b) == desc) { keys[i] = b; keys[i + j] = a; }
}

// two strides (j and j / 2, both >= kSortTile) of stage k in one pass: every thread owns the 4 keys they connect
__global__ void sort_global_step2_kernel(unsigned long long* keys, long long n_pad, long long k, long long j) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pad / 4) return;
    const long long lo = j >> 1;
    const long long i0 = ((t & ~(lo - 1)) << 2) + (t & (lo - 1));
    unsigned long long a = keys[i0], b = keys[i0 + lo], c = keys[i0 + j], d = keys[i0 + j + lo];
    const bool desc = (i0 & k) == 0;
    cmp_swap_desc(a, c, desc); cmp_swap_desc(b, d, desc);
    cmp_swap_desc(a, b, desc); cmp_swap_desc(c, d, desc);
    keys[i0] = a; keys[i0 + lo] = b; keys[i0 + j] = c; keys[i0 + j + lo] = d;
}

cudaError_t sort_configure() {
    static mas::PerDeviceInt done_tile, done_merge;     // the opt-in is a per-device setting
    const int bytes = kSortTile * (int)sizeof(unsigned long long);
    cudaError_t e = mas::opt_in_smem(sort_tile_kernel, done_tile, bytes);
    if (e == cudaSuccess) e = mas::opt_in_smem(sort_tile_merge_kernel, done_merge, bytes);
    return e;
}

}  // namespace

extern "C" int mas_region_keys_dev(const float* score, const uint8_t* in_pool, const int32_t* image_rank, int64_t n_img,
                                   int nseg, uint64_t* keys, void* stream) {
    MAS_REQUIRE(score && in_pool && image_rank && keys, MAS_E_BADARG, "region_keys: null pointer");
    MAS_REQUIRE(n_img >= 0 && nseg > 0, MAS_E_BADARG, "region_keys: bad shape");
    MAS_REQUIRE((long long)n_img * nseg < (1ll << 32), MAS_E_RANGE, "region_keys: n_img*nseg must fit 32 bits");
    const long long n = (long long)n_img * nseg;
    if (n == 0) return 0;
    const int threads = 256;
    region_keys_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        score, in_pool, image_rank, n, nseg, reinterpret_cast<unsigned long long*>(keys));
    mas::count_launches(1);
    MAS_LAUNCH_OK("region_keys_kernel");
    return 0;
}

constexpr size_t kFastStateOffset = (sizeof(SelectState) + 255) & ~(size_t)255;

extern "C" size_t mas_topk_workspace_bytes(void) { return kFastStateOffset + sizeof(FastState); }

extern "C" int mas_topk_u64_dev(const uint64_t* keys, int64_t n, int64_t k, uint64_t* out, int32_t* out_count, void* workspace,
                                size_t workspace_bytes, void* stream) {
    MAS_REQUIRE(keys && out && out_count && workspace, MAS_E_BADARG, "topk_u64: null pointer");
    MAS_REQUIRE(n >= 0 && k >= 0, MAS_E_BADARG, "topk_u64: negative size");
    MAS_REQUIRE(workspace_bytes >= sizeof(SelectState), MAS_E_WORKSPACE, "topk_u64: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    SelectState* state = reinterpret_cast<SelectState*>(workspace);
    const unsigned long long* kk = reinterpret_cast<const unsigned long long*>(keys);
    select_init_kernel<<<1, 256, 0, st>>>(state, k);
    mas::count_launches((n > 0 && k > 0) ? 2 + 16 + 1 : 2);
    if (n > 0 && k > 0) {
        const int threads = 256;
        const unsigned blocks = (unsigned)std::min<long long>((n + threads - 1) / threads, (long long)mas::sm_count() * 8);
        for (int shift = 56; shift >= 0; shift -= 8) {
            select_hist_kernel<<<blocks, threads, 0, st>>>(kk, n, state, shift);
            select_pick_kernel<<<1, 32, 0, st>>>(state, shift);
        }
        select_compact_kernel<<<blocks, threads, 0, st>>>(kk, n, state, reinterpret_cast<unsigned long long*>(out), k);
    }
    select_finish_kernel<<<1, 1, 0, st>>>(state, out_count, k);
    MAS_LAUNCH_OK("topk_u64 kernels");
    return 0;
}

extern "C" int64_t mas_sort_capacity(int64_t n) {
    long long cap = kSortTile;
    while (cap < n) cap <<= 1;
    return cap;
}

extern "C" int mas_sort_desc_u64_dev(uint64_t* keys, int64_t n, void* stream) {
    MAS_REQUIRE(keys && n >= 0, MAS_E_BADARG, "sort_desc_u64: bad argument");
    MAS_REQUIRE(n <= (1ll << 22), MAS_E_RANGE, "sort_desc_u64: n > 2^22");
    if (n <= 1) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    MAS_CUDA_OK(sort_configure());
    long long n_pad = kSortTile;
    while (n_pad < n) n_pad <<= 1;
    unsigned long long* kk = reinterpret_cast<unsigned long long*>(keys);
    const size_t smem = kSortTile * sizeof(unsigned long long);
    int launches = 1;
    if (n_pad > n) { sort_pad_kernel<<<(unsigned)((n_pad - n + 255) / 256), 256, 0, st>>>(kk, n, n_pad); ++launches; }
    const unsigned tiles = (unsigned)(n_pad / kSortTile);
    sort_tile_kernel<<<tiles, kSortThreads, smem, st>>>(kk);
    for (long long k = 2ll * kSortTile; k <= n_pad; k <<= 1) {
        long long j = k >> 1;
        while (j >= kSortTile) {
            if ((j >> 1) >= kSortTile) {
                sort_global_step2_kernel<<<(unsigned)((n_pad / 4 + 255) / 256), 256, 0, st>>>(kk, n_pad, k, j);
                j >>= 2;
            } else {
                sort_global_step_kernel<<<(unsigned)((n_pad / 2 + 255) / 256), 256, 0, st>>>(kk, n_pad, k, j);
                j >>= 1;
            }
            ++launches;
        }
        sort_tile_merge_kernel<<<tiles, kSortThreads, smem, st>>>(kk, k);
        ++launches;
    }
    mas::count_launches(launches);
    MAS_LAUNCH_OK("sort_desc_u64 kernels");
    return 0;
}

namespace {

// histograms + compaction: out[0 .. count) holds every key from the 24-bit prefix of the k-th largest key upwards
int fast_candidates(const char* what, const uint64_t* keys, int64_t n, int64_t k, uint64_t* out, int64_t capacity, int32_t* out_count,
                    void* workspace, size_t workspace_bytes, void* stream, bool clamp_count, long long* msg_count = nullptr) {
    MAS_REQUIRE(keys && out && out_count && workspace, MAS_E_BADARG, "%s: null pointer", what);
    MAS_REQUIRE(n >= 0 && k >= 0, MAS_E_BADARG, "%s: negative size", what);
    MAS_REQUIRE(workspace_bytes >= mas_topk_workspace_bytes(), MAS_E_WORKSPACE, "%s: workspace too small", what);
    MAS_REQUIRE(capacity >= mas_sort_capacity(k) && capacity == mas_sort_capacity(capacity), MAS_E_BADARG,
                "%s: capacity must be a sort capacity >= mas_sort_capacity(k)", what);
    cudaStream_t st = (cudaStream_t)stream;
    FastState* state = reinterpret_cast<FastState*>(reinterpret_cast<char*>(workspace) + kFastStateOffset);
    MAS_CUDA_OK(cudaMemsetAsync(state, 0, sizeof(FastState), st));
    MAS_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)capacity * sizeof(uint64_t), st));
    if (n > 0 && k > 0) {
        const unsigned long long* kk = reinterpret_cast<const unsigned long long*>(keys);
        const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, (long long)mas::sm_count() * 4);
        fast_hist_kernel<<<blocks, 256, 0, st>>>(kk, n, state, k, 0);
        fast_hist_kernel<<<blocks, 256, 0, st>>>(kk, n, state, k, 1);
        fast_compact_kernel<<<blocks, 256, 0, st>>>(kk, n, state, reinterpret_cast<unsigned long long*>(out), capacity);
        mas::count_launches(3);
    }
    fast_finish_kernel<<<1, 1, 0, st>>>(state, out_count, k, capacity, clamp_count ? 1 : 0, msg_count);
    mas::count_launches(1);
    MAS_LAUNCH_OK(what);
    return 0;
}

}  // namespace

extern "C" int mas_topk_candidates_u64_dev(const uint64_t* keys, int64_t n, int64_t k, uint64_t* out, int64_t capacity,
                                           int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
    return fast_candidates("topk_candidates_u64", keys, n, k, out, capacity, out_count, workspace, workspace_bytes, stream, false);
}

extern "C" int mas_topk_candidates_msg_u64_dev(const uint64_t* keys, int64_t n, int64_t k, uint64_t* msg, int64_t capacity,
                                               int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
    MAS_REQUIRE(msg, MAS_E_BADARG, "topk_candidates_msg_u64: null pointer");
    return fast_candidates("topk_candidates_msg_u64", keys, n, k, msg, capacity, out_count, workspace, workspace_bytes, stream, false,
                           reinterpret_cast<long long*>(msg) + capacity);
}

extern "C" int mas_merge_counts_u64_dev(uint64_t* gathered, int world, int64_t stride, int32_t* worst, void* stream) {
    MAS_REQUIRE(gathered && worst, MAS_E_BADARG, "merge_counts_u64: null pointer");
    MAS_REQUIRE(world >= 1 && stride >= 2, MAS_E_BADARG, "merge_counts_u64: bad shape");
    merge_counts_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(gathered), world, stride, worst);
    mas::count_launches(1);
    MAS_LAUNCH_OK("merge_counts_kernel");
    return 0;
}

extern "C" int mas_topk_sorted_u64_dev(const uint64_t* keys, int64_t n, int64_t k, uint64_t* out, int64_t capacity,
                                       int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
    const int rc = fast_candidates("topk_sorted_u64", keys, n, k, out, capacity, out_count, workspace, workspace_bytes, stream, true);
    if (rc != 0) return rc;
    if (n > 0 && k > 0) return mas_sort_desc_u64_dev(out, capacity, stream);
    return 0;
}

extern "C" int mas_prefix_cut_dev(const uint64_t* sorted_keys, const int32_t* count, const uint8_t* cost_by_tie, int64_t n_cost,
                                  int64_t budget, int32_t* n_take, void* stream) {
    MAS_REQUIRE(sorted_keys && count && cost_by_tie && n_take, MAS_E_BADARG, "prefix_cut: null pointer");
    MAS_REQUIRE(n_cost >= 0 && n_cost <= (1ll << 32), MAS_E_BADARG, "prefix_cut: bad table size");
    prefix_cut_kernel<<<1, kCutThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(sorted_keys), count,
                                                                  cost_by_tie, n_cost, budget, n_take);
    mas::count_launches(1);
    MAS_LAUNCH_OK("prefix_cut_kernel");
    return 0;
}
