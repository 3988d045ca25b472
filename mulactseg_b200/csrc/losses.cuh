// Shared pieces of the stage-1 loss kernels (losses.cu: tile walk / active-tile list; losses_dense.cu: the TMA-staged
// strip walk for densely selected batches).  See losses.cu for the reference lines and the design notes.
#pragma once

#include "common.cuh"

namespace mas_loss {

constexpr int kThreads = 128;
constexpr uint32_t kGroupBit = 0x80000000u;
constexpr float kEps = 1e-8f;

struct LossParams {
    const float* logits;
    const void* ids;
    const uint8_t* mask;
    const uint32_t* info;
    int n_img, C, H, W, S;
    float temp;      // T
    float inv_temp;  // 1 / T
    float scale;     // log2(e) / T
    int do_choice, do_group;
    int tile_rows;                // rows per tile (<= kTileRows)
    double* acc;                  // fwd: [0..5] one-hot / multi-hot / empty {sum, count}
    unsigned long long* gmax;     // (n_img * S * C) packed maxima
    const float* coef;            // bwd: {w_onehot, w_multihot, w_empty, w_group} = d total / d bucket sum
    float* grad;
    // ACTIVE-TILE LIST (mas_multihot_tiles_dev; NULL = walk every tile): 32 px x kListRows tiles that hold a selected pixel
    const uint32_t* tile_words;   // bit l of word g: tile 32 g + l is active
    const int* tile_offsets;      // exclusive prefix sum of the popcounts; tile_offsets[n_groups] = number of active tiles
    int n_groups;
    int list_mode;                // 1: take tiles from the list (forward always; backward when the image is sparsely selected)
    // share of active tiles (percent) from which the DENSE kernels (losses_dense.cu) take the batch instead of the list
    // walk; both sets of kernels are launched and read the active count on the device, one of them returns at once.
    // 0 = the dense kernels are not in play (shape not eligible / switched off).
    int dense_percent;
};


constexpr int kListRows = 8;            // rows of a tile of the active-tile list (32 px wide)

__device__ __forceinline__ long long list_tile_count(int n_img, int H, int W) {
    return (long long)n_img * ((W + 31) / 32) * ((H + kListRows - 1) / kListRows);
}

// true when the active-tile list says the batch is densely selected (>= p.dense_percent of the tiles hold a selected pixel)
__device__ __forceinline__ bool dense_regime(const LossParams& p) {
    if (p.dense_percent <= 0 || !p.list_mode) return false;
    return (long long)__ldg(p.tile_offsets + p.n_groups) * 100 >= list_tile_count(p.n_img, p.H, p.W) * p.dense_percent;
}

// losses_dense.cu: launch the dense forward / backward kernel when the shape allows it (rows of the mask 16-byte aligned,
// tensor maps available); *launched = false otherwise.  p.dense_percent == 0 and p.list_mode == 0: runs unconditionally.
cudaError_t launch_dense(const LossParams& p, int ids_dtype, bool backward, cudaStream_t stream, bool* launched);

}  // namespace mas_loss
