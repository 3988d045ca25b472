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

// ---------------------------------------------------------------------------------------- bitonic sort (descending)
constexpr int kSortThreads = 1024;
constexpr int kSortTile = 2 * kSortThreads;

__device__ __forceinline__ void cmp_swap_desc(unsigned long long& a, unsigned long long& b, bool descending) {
    if ((a < b) == descending) { const unsigned long long t = a; a = b; b = t; }
}

__global__ void sort_pad_kernel(unsigned long long* keys, long long n, long long n_pad) {
    const long long i = n + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) keys[i] = 0ull;
}

// sort each tile of kSortTile keys; tile t is sorted descending if bit (t & 1) == 0 else ascending
__global__ void __launch_bounds__(kSortThreads) sort_tile_kernel(unsigned long long* keys) {
    __shared__ unsigned long long s[kSortTile];
    unsigned long long* g = keys + (long long)blockIdx.x * kSortTile;
    s[threadIdx.x] = g[threadIdx.x];
    s[threadIdx.x + kSortThreads] = g[threadIdx.x + kSortThreads];
    for (int k = 2; k <= kSortTile; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            const int i = 2 * threadIdx.x - (threadIdx.x & (j - 1));
            const long long gi = (long long)blockIdx.x * kSortTile + i;
            cmp_swap_desc(s[i], s[i + j], (gi & k) == 0);
        }
    }
    __syncthreads();
    g[threadIdx.x] = s[threadIdx.x];
    g[threadIdx.x + kSortThreads] = s[threadIdx.x + kSortThreads];
}

__global__ void sort_global_step_kernel(unsigned long long* keys, long long n_pad, long long k, long long j) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pad / 2) return;
    const long long i = 2 * t - (t & (j - 1));
    unsigned long long a = keys[i], b = keys[i + j];
    const bool desc = (i & k) == 0;
    if ((a < b) == desc) { keys[i] = b; keys[i + j] = a; }
}

// finish stage k inside a tile: strides kSortThreads .. 1
__global__ void __launch_bounds__(kSortThreads) sort_tile_merge_kernel(unsigned long long* keys, long long k) {
    __shared__ unsigned long long s[kSortTile];
    unsigned long long* g = keys + (long long)blockIdx.x * kSortTile;
    s[threadIdx.x] = g[threadIdx.x];
    s[threadIdx.x + kSortThreads] = g[threadIdx.x + kSortThreads];
    for (int j = kSortThreads; j > 0; j >>= 1) {
        __syncthreads();
        const int i = 2 * threadIdx.x - (threadIdx.x & (j - 1));
        const long long gi = (long long)blockIdx.x * kSortTile + i;
        cmp_swap_desc(s[i], s[i + j], (gi & k) == 0);
    }
    __syncthreads();
    g[threadIdx.x] = s[threadIdx.x];
    g[threadIdx.x + kSortThreads] = s[threadIdx.x + kSortThreads];
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

extern "C" size_t mas_topk_workspace_bytes(void) { return sizeof(SelectState); }

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
    long long n_pad = kSortTile;
    while (n_pad < n) n_pad <<= 1;
    unsigned long long* kk = reinterpret_cast<unsigned long long*>(keys);
    int launches = 1;
    if (n_pad > n) { sort_pad_kernel<<<(unsigned)((n_pad - n + 255) / 256), 256, 0, st>>>(kk, n, n_pad); ++launches; }
    const unsigned tiles = (unsigned)(n_pad / kSortTile);
    sort_tile_kernel<<<tiles, kSortThreads, 0, st>>>(kk);
    for (long long k = 2ll * kSortTile; k <= n_pad; k <<= 1) {
        for (long long j = k >> 1; j >= kSortTile; j >>= 1) {
            sort_global_step_kernel<<<(unsigned)((n_pad / 2 + 255) / 256), 256, 0, st>>>(kk, n_pad, k, j);
            ++launches;
        }
        sort_tile_merge_kernel<<<tiles, kSortThreads, 0, st>>>(kk, k);
        ++launches;
    }
    mas::count_launches(launches);
    MAS_LAUNCH_OK("sort_desc_u64 kernels");
    return 0;
}
