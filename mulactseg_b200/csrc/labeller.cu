// Stage-2 pseudo-labellers: prototype labeller (P1) and candidate arg-max labeller (P2).
//
// Reference (paths relative to the reference checkout):
//   P1  trainer/eval_save_cosplbl_prop.py:121-314 pseudo_label_generation (prototypes from multi-hot superpixels only)
//       trainer/eval_save_cosplbl_prop_includeonehot.py:121-316 (shipped: every selected superpixel)
//   P2  trainer/eval_within_multihot.py:93-146 top_pseudo_label_generation
//
// What the reference does per image, and what runs here instead:
//   softmax + scatter_max(prob, spx)  -> arg-max pixel per (superpixel, candidate class)        [losses.cu pass, T = 1]
//   prototypes = feat[arg-max pixel]                                                             [proto_gather_kernel]
//   dense mm(protos, valid_feat.T) (nproto x HW') + scatter_max over it, of which only the block of a pixel's OWN
//   superpixel is consumed                      -> per superpixel: sims of its own pixels only   [proto_assign_kernel]
//   per-prototype torch.median / min of the similarities assigned to it                          [same kernel, radix select]
//   per-superpixel skimage binary_dilation on the CPU + torch.unique -> neighbour ids            [spx_adjacency_kernel: bit matrix]
//   python loop over selected superpixels, mm(protos_s, feat[Q].T), ordered overwrite            [proto_propagate_kernel]
// The sequential "later superpixels overwrite earlier ones" (:276-305) becomes: every unselected pixel takes the label
// offered by the LARGEST adjacent selected superpixel id whose threshold test passes; selected pixels keep the label
// of their own superpixel's nearest prototype (:309-310).
//
// Similarities are fp32 FMA chains over the feature channels in a FIXED order (channel 0..F-1), the same in the
// assign and the propagate kernel, so that "threshold < similarity" is evaluated on bit-identical numbers for the
// pixel that defines the threshold.  The contraction is tiny (|protos_s| <= C', typically 1-3, per pixel) and bound by
// reading the (F, H, W) features of the touched superpixels once: no tensor cores.
#include "common.cuh"

#include <stdlib.h>

#include <algorithm>
#include <mutex>

namespace {

constexpr uint32_t kGroupBit = 0x80000000u;
constexpr int kAssignThreads = 256;
constexpr int kGroup = 8;          // prototypes evaluated per pass over a pixel's feature column

struct LabelParams {
    const void* feats;      // (F, H, W) -- or (F, fh_in, fw_in) when the features are low resolution (see FeatSource)
    int fh_in, fw_in;       // low-resolution source size (0 = full resolution)
    float fry, frx;         // fh_in / H, fw_in / W
    const uint8_t* mask;    // (H, W)
    const void* ids;        // (H, W)
    int F, C, H, W, S, P;
    int threshold_min;      // 0: lower median, 1: min
    int only_multihot;
    int overlapped;      // other images of the batch run concurrently on other streams (mas_proto_labeller_batch_dev)
    int32_t* status;
    const uint32_t* info;               // (S)
    const unsigned long long* gmax;     // (S, C)
    const int* offset;                  // (S + 1) CSR offsets
    const int* pixlist;                 // (P) pixels grouped by superpixel
    float* proto;                       // (S, C, F)
    float* own_sim;                     // (P)
    uint8_t* own_cls;                   // (P)
    float* thr;                         // (S, C)
    uint32_t* adj;                      // (S, words)
    uint32_t* svalid;                   // (words) selected superpixels that own prototypes
    uint8_t* touched;                   // (S) 1: the superpixel or one of its 8-neighbours owns prototypes
    int words;
    uint8_t* labels;                    // (H, W)
};

template <typename IdT>
__device__ __forceinline__ int read_id(const void* ids, size_t i, int S) {
    const long long v = (long long)reinterpret_cast<const IdT*>(ids)[i];
    return (v < 0 || v >= S) ? -1 : (int)v;
}

__device__ __forceinline__ float ordered_to_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ------------------------------------------------------------------------------------------ feature source
// Where a pixel's F-dimensional feature column comes from.  Full resolution: element ch of pixel pix = feats[ch * P + pix]
// (fp32, or bf16 widened to fp32).  LOW resolution (mas_proto_labeller_lowres_dev): feats holds the
// network head's (F, fh_in, fw_in) map and every value is F.interpolate(mode='bilinear', align_corners=False)'s
//   h0 * (w0 * a + w1 * b) + h1 * (w0 * d + w1 * e)          (models/segmentation/utils.py:28-34)
// evaluated on the fly from the four taps -- the 16x larger (F, H, W) tensor (2.1 GB per Cityscapes image) never exists;
// the low-resolution map (134 MB) mostly lives in L2.
template <typename FT>
__device__ __forceinline__ float feat_value(const FT* p);
template <>
__device__ __forceinline__ float feat_value<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float feat_value<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __uint_as_float(((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
}

__device__ __forceinline__ void bilinear_tap(int dst, float scale, int size_in, int& i0, int& step, float& lambda1) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;      // torch: area_pixel_compute_source_index(align_corners=false)
    src = src < 0.f ? 0.f : src;
    i0 = (int)src;
    if (i0 > size_in - 1) i0 = size_in - 1;
    step = (i0 < size_in - 1) ? 1 : 0;
    lambda1 = src - (float)i0;
}

template <typename FT, bool LOWRES>
struct FeatColumn {
    const FT* base;      // channel 0 of the pixel (LOWRES: of its top-left tap)
    size_t plane;        // elements between channels
    int right, down;     // LOWRES: offsets of the other taps
    float l0x, l1x, l0y, l1y;

    __device__ __forceinline__ FeatColumn(const LabelParams& p, int pix) {
        if (LOWRES) {
            const int y = pix / p.W, x = pix - y * p.W;
            int y0, ys, x0, xs;
            bilinear_tap(y, p.fry, p.fh_in, y0, ys, l1y);
            bilinear_tap(x, p.frx, p.fw_in, x0, xs, l1x);
            l0y = 1.f - l1y; l0x = 1.f - l1x;
            plane = (size_t)p.fh_in * p.fw_in;
            base = reinterpret_cast<const FT*>(p.feats) + (size_t)y0 * p.fw_in + x0;
            right = xs; down = ys * p.fw_in;
        } else {
            plane = (size_t)p.P;
            base = reinterpret_cast<const FT*>(p.feats) + pix;
            right = 0; down = 0; l0x = l1x = l0y = l1y = 0.f;
        }
    }
    __device__ __forceinline__ float at(int ch) const {
        const FT* q = base + (size_t)ch * plane;
        if (!LOWRES) return feat_value<FT>(q);
        const float a = feat_value<FT>(q), b = feat_value<FT>(q + right);
        const float d = feat_value<FT>(q + down), e = feat_value<FT>(q + down + right);
        return l0y * (l0x * a + l1x * b) + l1y * (l0x * d + l1x * e);
    }
};

// ------------------------------------------------------------------------------------------ P2
template <typename IdT>
__global__ void candidate_argmax_kernel(const float* __restrict__ logits, const void* __restrict__ ids, const uint8_t* __restrict__ mask,
                                        const uint32_t* __restrict__ info, long long n_pix, int P, int C, int S,
                                        uint8_t* __restrict__ labels) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += (long long)gridDim.x * blockDim.x) {
        uint8_t out = 255;
        if (mask[i]) {
            const int img = (int)(i / P);
            const int id = read_id<IdT>(ids, (size_t)i, S);
            if (id >= 0) {
                const uint32_t bits = info[(size_t)img * S + id];
                const float* x = logits + (size_t)img * C * P + (i - (long long)img * P);
                // arg-max over c of logit * target (first index on ties): non-candidates contribute 0, not -inf
                float best = ((bits & 1u) ? x[0] : 0.f);
                int arg = 0;
                for (int c = 1; c < C; ++c) {
                    const float v = ((bits >> c) & 1u) ? x[(size_t)c * P] : 0.f;
                    if (v > best) { best = v; arg = c; }
                }
                out = (uint8_t)arg;
            }
        }
        labels[i] = out;
    }
}

// ------------------------------------------------------------------------------------------ CSR of pixels by superpixel
// lanes holding a run of equal ids aggregate into one atomic; run members keep their order (coalescing later)
__device__ __forceinline__ void run_info(int id, int lane, int& head_lane, int& run_len) {
    const int prev = __shfl_up_sync(0xffffffffu, id, 1);
    const bool head = lane == 0 || id != prev;
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    head_lane = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
    const unsigned above = heads & ~(0xffffffffu >> (31 - head_lane));   // heads after this run's head
    const int next = above ? (__ffs(above) - 1) : 32;
    run_len = next - head_lane;
}

__device__ __forceinline__ void adj_set(uint32_t* adj, int words, int a, int b);

// pixels per superpixel (warp-aggregated runs) and, in the same pass over the id map, the adjacency bit matrix of the
// 3x3 dilation (:260-266): every 8-neighbour pair is seen once from its upper / left pixel
template <typename IdT>
__global__ void spx_count_adjacency_kernel(const void* __restrict__ ids, int H, int W, int S, int* __restrict__ count, int words,
                                           uint32_t* __restrict__ adj) {
    const int lane = threadIdx.x & 31;
    const long long P = (long long)H * W;
    const long long padded = (P + 31) & ~31ll;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += (long long)gridDim.x * blockDim.x) {
        const int id = i < P ? read_id<IdT>(ids, (size_t)i, S) : -1;
        int head_lane, run_len;
        run_info(id, lane, head_lane, run_len);
        if (lane == head_lane && id >= 0) atomicAdd(count + id, run_len);
        if (id < 0) continue;
        const int y = (int)(i / W), x = (int)(i - (long long)y * W);
        const int dx[4] = {1, -1, 0, 1}, dy[4] = {0, 1, 1, 1};      // right, down-left, down, down-right
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xx = x + dx[k], yy = y + dy[k];
            if (xx < 0 || xx >= W || yy >= H) continue;
            const int b = read_id<IdT>(ids, (size_t)yy * W + xx, S);
            if (b < 0 || b == id) continue;
            adj_set(adj, words, id, b);
            adj_set(adj, words, b, id);
        }
    }
}

// exclusive scan of count[0..S) into offset[0..S]; also clears the fill cursors
__global__ void spx_scan_kernel(const int* __restrict__ count, int S, int* __restrict__ offset, int* __restrict__ cursor,
                                const double* __restrict__ acc, int only_multihot, int32_t* __restrict__ status) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    // status: selected pixels whose superpixel has no candidate class (the reference fails on such input, :226)
    if (threadIdx.x == 0) { carry = 0; status[0] = only_multihot ? 0 : (int32_t)acc[5]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < S; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < S ? count[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int before = carry + (warp ? warp_sums[warp - 1] : 0) + x - v;
        if (i < S) { offset[i] = before; cursor[i] = 0; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) offset[S] = carry;
}

template <typename IdT>
__global__ void spx_fill_kernel(const void* __restrict__ ids, int P, int S, const int* __restrict__ offset, int* __restrict__ cursor,
                                int* __restrict__ pixlist) {
    const int lane = threadIdx.x & 31;
    const long long padded = ((long long)P + 31) & ~31ll;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += (long long)gridDim.x * blockDim.x) {
        const int id = i < P ? read_id<IdT>(ids, (size_t)i, S) : -1;
        int head_lane, run_len;
        run_info(id, lane, head_lane, run_len);
        int base = 0;
        if (lane == head_lane && id >= 0) base = atomicAdd(cursor + id, run_len);
        base = __shfl_sync(0xffffffffu, base, head_lane);
        if (id >= 0) pixlist[offset[id] + base + (lane - head_lane)] = (int)i;
    }
}

// ------------------------------------------------------------------------------------------ adjacency (3x3 dilation)
__device__ __forceinline__ void adj_set(uint32_t* adj, int words, int a, int b) {
    uint32_t* w = adj + (size_t)a * words + (b >> 5);
    const uint32_t bit = 1u << (b & 31);
    if (!(*reinterpret_cast<volatile uint32_t*>(w) & bit)) atomicOr(w, bit);
}

// ------------------------------------------------------------------------------------------ prototypes
// one warp per (superpixel, class): copy the feature column of the arg-max-probability pixel
template <typename FT, bool LOWRES>
__global__ void proto_gather_kernel(LabelParams p) {
    const long long entry = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (entry >= (long long)p.S * p.C) return;
    const int s = (int)(entry / p.C), c = (int)(entry - (long long)s * p.C);
    const uint32_t inf = p.info[s];
    if (!(inf & kGroupBit) || !((inf >> c) & 1u)) return;
    const unsigned long long e = p.gmax[entry];
    if (e == 0ull) return;
    const uint32_t pix = ~(uint32_t)e;
    const FeatColumn<FT, LOWRES> col(p, (int)pix);
    for (int ch = lane; ch < p.F; ch += 32) p.proto[entry * p.F + ch] = col.at(ch);
    if (lane == 0) atomicOr(p.svalid + (s >> 5), 1u << (s & 31));
}

__device__ __forceinline__ bool pixel_selected(const LabelParams& p, int pix, uint32_t inf) {
    return (inf & kGroupBit) && p.mask[pix] != 0;
}

// ------------------------------------------------------------------------------------------ similarity engine
// Both prototype kernels evaluate, for one pixel per thread, the inner products of the pixel's feature column with up
// to kGroup prototypes staged in shared memory -- ONE pass over the column for all of them, kFeatBatch channel loads in
// flight per thread (each a coalesced row segment: the pixel list keeps the runs of a superpixel contiguous).
// Every similarity is the same fp32 FMA chain over channels 0..F-1, whichever kernel computes it.
constexpr int kFeatBatch = 16;
constexpr int kSlices = 4;         // CTAs that share the pixels of one superpixel

// sproto layout: [channel][kGroup] so that the kGroup operands of one channel are two 128-bit broadcast reads
// (a 4-prototype instantiation for sparse batches was measured slower: the compiler keeps fewer loads in flight)
// G = 8: all kGroup slots; G = 4: only the first four (most batches hold <= 4 prototypes -- half the FMAs and shared-memory
// reads).  A slot's similarity is the same FMA chain either way.
template <typename FT, bool LOWRES, int G>
__device__ __forceinline__ void dot_some(const LabelParams& p, const float* __restrict__ sproto, int pix, float (&acc)[kGroup]) {
#pragma unroll
    for (int g = 0; g < kGroup; ++g) acc[g] = 0.f;
    const FeatColumn<FT, LOWRES> f(p, pix);
    constexpr int kBatch = LOWRES ? kFeatBatch / 2 : kFeatBatch;     // LOWRES: four taps per value -- same number of loads in flight
    int ch = 0;
    for (; ch + kBatch <= p.F; ch += kBatch) {
        float x[kBatch];
#pragma unroll
        for (int i = 0; i < kBatch; ++i) x[i] = f.at(ch + i);
#pragma unroll
        for (int i = 0; i < kBatch; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(sproto + (size_t)(ch + i) * kGroup);
            acc[0] = fmaf(x[i], a.x, acc[0]); acc[1] = fmaf(x[i], a.y, acc[1]);
            acc[2] = fmaf(x[i], a.z, acc[2]); acc[3] = fmaf(x[i], a.w, acc[3]);
            if (G > 4) {
                const float4 b = *reinterpret_cast<const float4*>(sproto + (size_t)(ch + i) * kGroup + 4);
                acc[4] = fmaf(x[i], b.x, acc[4]); acc[5] = fmaf(x[i], b.y, acc[5]);
                acc[6] = fmaf(x[i], b.z, acc[6]); acc[7] = fmaf(x[i], b.w, acc[7]);
            }
        }
    }
    for (; ch < p.F; ++ch) {
        const float x = f.at(ch);
#pragma unroll
        for (int g = 0; g < G; ++g) acc[g] = fmaf(x, sproto[(size_t)ch * kGroup + g], acc[g]);
    }
}

// n = prototypes staged in sproto (warp-uniform)
template <typename FT, bool LOWRES>
__device__ __forceinline__ void dot_all(const LabelParams& p, const float* __restrict__ sproto, int pix, int n, float (&acc)[kGroup]) {
    if (n <= 4) dot_some<FT, LOWRES, 4>(p, sproto, pix, acc);
    else dot_some<FT, LOWRES, 8>(p, sproto, pix, acc);
}
static_assert(kGroup == 8, "dot_all is written for 8 prototypes per pass");

// One entry of the prototype list a CTA works through: prototype (s, c), in the order the reference visits them.
struct ProtoEntry {
    int s, c;
    int first;      // 1: first prototype of superpixel s (a new neighbour starts here)
};

// copy the prototypes of ent[0..n) (zeros beyond n) into sproto[channel][slot]; all threads
__device__ __forceinline__ void stage_entries(const LabelParams& p, float* sproto, const ProtoEntry* ent, int n) {
    for (int i = threadIdx.x; i < kGroup * p.F; i += blockDim.x) {
        const int g = i / p.F, ch = i - g * p.F;
        sproto[(size_t)ch * kGroup + g] = g < n ? p.proto[((size_t)ent[g].s * p.C + ent[g].c) * p.F + ch] : 0.f;
    }
}

// ------------------------------------------------------------------------------------------ assign
// CTA (s, slice): nearest prototype of s for the selected pixels of s in this slice  (:213-230)
template <typename FT, bool LOWRES>
__global__ void __launch_bounds__(kAssignThreads) proto_assign_kernel(LabelParams p) {
    extern __shared__ __align__(16) float sproto[];          // [F][kGroup]
    __shared__ ProtoEntry ent[kGroup];
    const int s = blockIdx.x;
    if (!((p.svalid[s >> 5] >> (s & 31)) & 1u)) return;
    const uint32_t inf = p.info[s];
    const uint32_t bits = inf & ~kGroupBit;
    const int beg = p.offset[s], end = p.offset[s + 1];
    for (int base = beg + blockIdx.y * kAssignThreads; base < end; base += kSlices * kAssignThreads) {
        const int e = base + threadIdx.x;
        const int pix = e < end ? p.pixlist[e] : -1;
        const bool mine = pix >= 0 && pixel_selected(p, pix, inf);
        float best = -INFINITY;
        int bestc = 255;
        uint32_t rest = bits;
        while (rest) {                                   // uniform over the CTA
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t b = rest;
                for (int g = 0; g < kGroup; ++g) {
                    if (b) { ent[g].s = s; ent[g].c = __ffs(b) - 1; ent[g].first = 0; b &= b - 1u; }
                    else ent[g].c = -1;
                }
            }
            int n = 0;
            for (; n < kGroup && rest; ++n) rest &= rest - 1u;
            __syncthreads();
            stage_entries(p, sproto, ent, n);
            __syncthreads();
            if (mine) {
                float acc[kGroup];
                dot_all<FT, LOWRES>(p, sproto, pix, n, acc);
#pragma unroll
                for (int g = 0; g < kGroup; ++g) {
                    if (g < n && (acc[g] > best || bestc == 255)) { best = acc[g]; bestc = ent[g].c; }
                }
            }
        }
        if (mine) {
            p.own_sim[pix] = best;
            p.own_cls[pix] = (uint8_t)bestc;
            p.labels[pix] = (uint8_t)bestc;
        }
    }
}

// one CTA per selected superpixel: per-prototype threshold = lower median (torch.median) or min of the similarities of
// the pixels assigned to it, 1.0 if none  (:241-255).  The (similarity, class) pairs of the superpixel's selected pixels
// are read once into registers (kHold per thread; larger superpixels re-read them from global memory in every pass);
// each class then takes a count + 4 x 8-bit radix-select passes over them.
constexpr int kHold = 8;

__global__ void __launch_bounds__(kAssignThreads) proto_threshold_kernel(LabelParams p) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int sel_prefix, sel_rank;
    __shared__ float red[kAssignThreads / 32];
    const int s = blockIdx.x;
    if (threadIdx.x < 32) {
        // does s or any superpixel adjacent to it own prototypes?  The propagate CTAs of all other superpixels -- half the
        // image at rho = 0.08 -- then leave after one load instead of scanning the adjacency row in every thread.
        const uint32_t* row = p.adj + (size_t)s * p.words;
        bool any = false;
        for (int w = threadIdx.x; w < p.words; w += 32) any |= ((row[w] | ((w == (s >> 5)) ? (1u << (s & 31)) : 0u)) & p.svalid[w]) != 0u;
        any = __any_sync(0xffffffffu, any);
        if (threadIdx.x == 0) p.touched[s] = any ? 1 : 0;
    }
    if (!((p.svalid[s >> 5] >> (s & 31)) & 1u)) return;
    const uint32_t inf = p.info[s];
    const uint32_t bits = inf & ~kGroupBit;
    const int beg = p.offset[s], end = p.offset[s + 1];
    const bool held = end - beg <= kHold * kAssignThreads;
    float hsim[kHold];
    int hcls[kHold];       // -1: no selected pixel in this slot
    if (held) {
        int pix[kHold];
#pragma unroll
        for (int i = 0; i < kHold; ++i) {
            const int e = beg + i * kAssignThreads + threadIdx.x;
            pix[i] = e < end ? p.pixlist[e] : -1;
        }
#pragma unroll
        for (int i = 0; i < kHold; ++i) {
            const bool on = pix[i] >= 0 && pixel_selected(p, pix[i], inf);
            hcls[i] = on ? (int)p.own_cls[pix[i]] : -1;
            hsim[i] = on ? p.own_sim[pix[i]] : 0.f;
        }
    }
    // visit(sim) for every selected pixel of s assigned to class c
    auto for_each = [&](int c, auto visit) {
        if (held) {
#pragma unroll
            for (int i = 0; i < kHold; ++i) {
                if (hcls[i] == c) visit(hsim[i]);
            }
        } else {
            for (int e = beg + threadIdx.x; e < end; e += blockDim.x) {
                const int pix = p.pixlist[e];
                if (pixel_selected(p, pix, inf) && p.own_cls[pix] == c) visit(p.own_sim[pix]);
            }
        }
    };
    uint32_t b = bits;
    while (b) {
        const int c = __ffs(b) - 1;
        b &= b - 1u;
        float result;
        if (p.threshold_min) {
            float mn = INFINITY;
            for_each(c, [&](float sim) { mn = fminf(mn, sim); });
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mn;
            __syncthreads();
            mn = red[0];
            for (int w = 1; w < kAssignThreads / 32; ++w) mn = fminf(mn, red[w]);
            result = mn == INFINITY ? 1.f : mn;
            __syncthreads();
        } else {
            // count, then 4 x 8-bit radix select of the element of ascending rank (n - 1) / 2
            if (threadIdx.x == 0) { sel_prefix = 0u; sel_rank = 0u; }
            uint32_t mask_bits = 0u;
            bool empty = false;
            for (int pass = -1; pass < 4; ++pass) {
                hist[threadIdx.x] = 0u;     // kAssignThreads == 256
                __syncthreads();
                const uint32_t prefix = sel_prefix;
                const unsigned int rank = sel_rank;         // read by everyone here: the owner lane rewrites it below
                const int shift = pass < 0 ? 0 : 24 - 8 * pass;
                for_each(c, [&](float sim) {
                    if (pass < 0) { atomicAdd(&hist[0], 1u); return; }
                    const uint32_t key = mas::ordered_bits(sim);
                    if ((key & mask_bits) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
                });
                __syncthreads();
                if (pass < 0) {
                    const unsigned int n = hist[0];
                    empty = n == 0u;
                    if (threadIdx.x == 0) sel_rank = n ? (n - 1u) / 2u : 0u;
                } else if (threadIdx.x < 32) {
                    // warp 0: first bin whose cumulative count exceeds the rank (lane l owns bins 8 l .. 8 l + 7)
                    const int lane = threadIdx.x;
                    unsigned int h[8], mine = 0u;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { h[i] = hist[lane * 8 + i]; mine += h[i]; }
                    unsigned int incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned int y = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += y;
                    }
                    const unsigned int owners = __ballot_sync(0xffffffffu, rank < incl);
                    const int owner = owners ? __ffs(owners) - 1 : 31;
                    if (lane == owner) {
                        unsigned int r = rank - (incl - mine), bin = 0u;
#pragma unroll
                        for (int i = 0; i < 7; ++i) {
                            if (bin == (unsigned)i && r >= h[i]) { r -= h[i]; bin = i + 1; }
                        }
                        sel_rank = r;
                        sel_prefix = prefix | (((unsigned)lane * 8u + bin) << shift);
                    }
                }
                __syncthreads();
                if (empty) break;
                if (pass >= 0) mask_bits |= 0xffu << shift;
            }
            result = empty ? 1.f : ordered_to_float(sel_prefix);
            __syncthreads();
        }
        if (threadIdx.x == 0) p.thr[(size_t)s * p.C + c] = result;
    }
}

// ------------------------------------------------------------------------------------------ propagate
// CTA (t, slice): the UNSELECTED pixels of superpixel t in this slice take the label offered by the largest adjacent
// selected superpixel (t itself included) whose test "some threshold < similarity" passes  (:276-305, ascending
// overwrite order == the largest passing id wins).  The prototypes of the neighbours are walked in that order,
// kGroup at a time; a pixel stops at its first passing neighbour.
template <typename FT, bool LOWRES>
__global__ void __launch_bounds__(kAssignThreads) proto_propagate_kernel(LabelParams p) {
    extern __shared__ __align__(16) float sproto[];          // [F][kGroup]
    __shared__ ProtoEntry ent[kGroup];
    __shared__ float ent_thr[kGroup];
    __shared__ int n_ent, cur_word, cur_s;
    __shared__ uint32_t cur_bits, cur_cls;
    const int t = blockIdx.x;
    const int beg = p.offset[t], end = p.offset[t + 1];
    if (beg + (int)blockIdx.y * kAssignThreads >= end) return;
    if (!p.touched[t]) return;                                // no selected neighbour at all (spx_touch_kernel)
    const uint32_t* row = p.adj + (size_t)t * p.words;
    const uint32_t inf_t = p.info[t];

    for (int base = beg + blockIdx.y * kAssignThreads; base < end; base += kSlices * kAssignThreads) {
        const int e = base + threadIdx.x;
        const int pix = e < end ? p.pixlist[e] : -1;
        bool done = pix < 0 || pixel_selected(p, pix, inf_t);   // selected pixels keep their own label
        // state of the neighbour being evaluated (its prototypes may span two batches)
        float best = -INFINITY;
        int bestc = 255;
        bool pass = false;
        __syncthreads();
        if (threadIdx.x == 0) { cur_word = p.words; cur_bits = 0u; cur_cls = 0u; cur_s = -1; }
        while (true) {
            __syncthreads();
            if (threadIdx.x == 0) {
                // next kGroup prototypes: neighbours by descending id, classes ascending
                int n = 0;
                int w = cur_word, sidx = cur_s;
                uint32_t nb = cur_bits, cls = cur_cls;
                while (n < kGroup) {
                    if (cls == 0u) {                     // next neighbour
                        while (nb == 0u && w > 0) {
                            --w;
                            nb = (row[w] | ((w == (t >> 5)) ? (1u << (t & 31)) : 0u)) & p.svalid[w];
                        }
                        if (nb == 0u) break;
                        const int bitpos = 31 - __clz(nb);
                        nb &= ~(1u << bitpos);
                        sidx = w * 32 + bitpos;
                        cls = p.info[sidx] & ~kGroupBit;
                        if (cls == 0u) continue;
                        ent[n].first = 1;
                    } else {
                        ent[n].first = 0;
                    }
                    const int c = __ffs(cls) - 1;
                    cls &= cls - 1u;
                    ent[n].s = sidx; ent[n].c = c;
                    ent_thr[n] = p.thr[(size_t)sidx * p.C + c];
                    ++n;
                }
                n_ent = n; cur_word = w; cur_bits = nb; cur_cls = cls; cur_s = sidx;
            }
            __syncthreads();
            const int n = n_ent;
            if (n == 0) break;
            stage_entries(p, sproto, ent, n);
            const bool all_done = __syncthreads_and(done);          // also orders the staging before the reads
            if (all_done) break;
            if (!done) {
                float acc[kGroup];
                dot_all<FT, LOWRES>(p, sproto, pix, n, acc);
#pragma unroll
                for (int g = 0; g < kGroup; ++g) {
                    if (g < n && !done) {
                        if (ent[g].first) {              // the previous neighbour is complete: did it pass?
                            if (pass) { p.labels[pix] = (uint8_t)bestc; done = true; }
                            best = -INFINITY; bestc = 255; pass = false;
                        }
                        if (!done) {
                            if (acc[g] > best || bestc == 255) { best = acc[g]; bestc = ent[g].c; }
                            pass |= ent_thr[g] < acc[g];
                        }
                    }
                }
            }
        }
        if (!done && pass) p.labels[pix] = (uint8_t)bestc;
    }
}

// ------------------------------------------------------------------------------------------ propagate, tile walk
// Same result as proto_propagate_kernel, other traversal: the CTA owns a SPATIAL tile (64 px x 8 rows; every warp a
// 128-byte-aligned 32-pixel row segment), so each feature line the tile needs is requested once and whole -- the
// per-superpixel CTAs above fetch a line once per superpixel that touches it (1.85x the useful bytes on the bench image;
// 1.42x is the line-granularity floor).  The pixels of a tile belong to a few superpixels: the CTA forms the UNION of their
// selected neighbours, walks its prototypes in the reference's order (neighbours by descending id, classes ascending) in
// batches of kGroup, and every pixel takes part only in the neighbours of ITS superpixel (adjacency bit test when a new
// neighbour starts) -- the per-pixel sequence of (neighbour, prototype) steps is exactly the one the per-superpixel kernel
// walks.
constexpr int kTileW = 64, kTileH = 4, kTileRowsPerWarp = 1;     // 8 warps: 2 across x 4 down, 1 row each

template <typename FT, bool LOWRES, typename IdT>
__global__ void __launch_bounds__(kAssignThreads) proto_propagate_tile_kernel(LabelParams p) {
    extern __shared__ __align__(16) float sproto[];          // [F][kGroup], then tset[words], nset[words]
    uint32_t* tset = reinterpret_cast<uint32_t*>(sproto + (size_t)p.F * kGroup);
    uint32_t* nset = tset + p.words;
    __shared__ ProtoEntry ent[kGroup];
    __shared__ float ent_thr[kGroup];
    __shared__ int n_ent, cur_word, cur_s;
    __shared__ uint32_t cur_bits, cur_cls;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * kTileW + (warp & 1) * 32 + lane;
    const int y0 = blockIdx.y * kTileH + (warp >> 1) * kTileRowsPerWarp;

    for (int w = threadIdx.x; w < 2 * p.words; w += blockDim.x) tset[w] = 0u;      // tset and nset are contiguous
    __syncthreads();
    // this thread's pixels: superpixel id (-1: nothing to do) -- unselected pixels of superpixels with a selected neighbour
    int tid_[kTileRowsPerWarp], pix[kTileRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kTileRowsPerWarp; ++r) {
        const int y = y0 + r;
        tid_[r] = -1; pix[r] = -1;
        if (x < p.W && y < p.H) {
            const int px = y * p.W + x;
            const int t = read_id<IdT>(p.ids, (size_t)px, p.S);
            if (t >= 0 && p.touched[t] && !pixel_selected(p, px, p.info[t])) {
                tid_[r] = t; pix[r] = px;
                const uint32_t bit = 1u << (t & 31);
                if (!(tset[t >> 5] & bit)) atomicOr(&tset[t >> 5], bit);
            }
        }
    }
    bool mine_any = false;
#pragma unroll
    for (int r = 0; r < kTileRowsPerWarp; ++r) mine_any |= tid_[r] >= 0;
    const bool cta_any = __syncthreads_or(mine_any);
    if (!cta_any) return;
    // union of the selected neighbours (each superpixel itself included) of the tile's superpixels
    for (int w = threadIdx.x; w < p.words; w += blockDim.x) {
        uint32_t acc = 0u;
        for (int tw = 0; tw < p.words; ++tw) {
            uint32_t bits = tset[tw];
            while (bits) {
                const int t = tw * 32 + (__ffs(bits) - 1);
                bits &= bits - 1u;
                acc |= p.adj[(size_t)t * p.words + w] | ((w == (t >> 5)) ? (1u << (t & 31)) : 0u);
            }
        }
        nset[w] = acc & p.svalid[w];
    }
    // per-pixel state of the neighbour being evaluated (its prototypes may span two batches)
    float best[kTileRowsPerWarp];
    int bestc[kTileRowsPerWarp];
    bool pass[kTileRowsPerWarp], done[kTileRowsPerWarp], rel[kTileRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kTileRowsPerWarp; ++r) { best[r] = -INFINITY; bestc[r] = 255; pass[r] = false; done[r] = tid_[r] < 0; rel[r] = false; }
    __syncthreads();
    if (threadIdx.x == 0) { cur_word = p.words; cur_bits = 0u; cur_cls = 0u; cur_s = -1; }
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) {
            int n = 0;
            int w = cur_word, sidx = cur_s;
            uint32_t nb = cur_bits, cls = cur_cls;
            while (n < kGroup) {
                if (cls == 0u) {                     // next neighbour of the union, descending id
                    while (nb == 0u && w > 0) { --w; nb = nset[w]; }
                    if (nb == 0u) break;
                    const int bitpos = 31 - __clz(nb);
                    nb &= ~(1u << bitpos);
                    sidx = w * 32 + bitpos;
                    cls = p.info[sidx] & ~kGroupBit;
                    if (cls == 0u) continue;
                    ent[n].first = 1;
                } else {
                    ent[n].first = 0;
                }
                const int c = __ffs(cls) - 1;
                cls &= cls - 1u;
                ent[n].s = sidx; ent[n].c = c;
                ent_thr[n] = p.thr[(size_t)sidx * p.C + c];
                ++n;
            }
            n_ent = n; cur_word = w; cur_bits = nb; cur_cls = cls; cur_s = sidx;
        }
        __syncthreads();
        const int n = n_ent;
        if (n == 0) break;
        stage_entries(p, sproto, ent, n);
        bool mine_done = true;
#pragma unroll
        for (int r = 0; r < kTileRowsPerWarp; ++r) mine_done &= done[r];
        const bool all_done = __syncthreads_and(mine_done);          // also orders the staging before the reads
        if (all_done) break;
#pragma unroll
        for (int r = 0; r < kTileRowsPerWarp; ++r) {
            // does this pixel's superpixel meet any neighbour of the batch?  (a neighbour continued from the previous batch
            // keeps its flag; new ones are tested when they start)
            bool wanted = false;
            if (!done[r]) {
                bool carry = rel[r];
#pragma unroll
                for (int g = 0; g < kGroup; ++g) {
                    if (g < n) {
                        if (ent[g].first) {
                            const int s2 = ent[g].s;
                            carry = s2 == tid_[r] || ((p.adj[(size_t)tid_[r] * p.words + (s2 >> 5)] >> (s2 & 31)) & 1u);
                        }
                        wanted |= carry;
                    }
                }
            }
            if (!__any_sync(0xffffffffu, wanted)) {
                // nobody in the warp row needs this batch: only the neighbour bookkeeping advances
                if (!done[r]) {
#pragma unroll
                    for (int g = 0; g < kGroup; ++g) {
                        if (g < n && ent[g].first && !done[r]) {
                            if (pass[r]) { p.labels[pix[r]] = (uint8_t)bestc[r]; done[r] = true; }
                            best[r] = -INFINITY; bestc[r] = 255; pass[r] = false;
                            const int s2 = ent[g].s;
                            rel[r] = s2 == tid_[r] || ((p.adj[(size_t)tid_[r] * p.words + (s2 >> 5)] >> (s2 & 31)) & 1u);
                        }
                    }
                }
                continue;
            }
            if (!done[r] && wanted) {
                float acc[kGroup];
                dot_all<FT, LOWRES>(p, sproto, pix[r], n, acc);
#pragma unroll
                for (int g = 0; g < kGroup; ++g) {
                    if (g < n && !done[r]) {
                        if (ent[g].first) {              // the previous neighbour is complete: did it pass?
                            if (pass[r]) { p.labels[pix[r]] = (uint8_t)bestc[r]; done[r] = true; }
                            best[r] = -INFINITY; bestc[r] = 255; pass[r] = false;
                            const int s2 = ent[g].s;
                            rel[r] = s2 == tid_[r] || ((p.adj[(size_t)tid_[r] * p.words + (s2 >> 5)] >> (s2 & 31)) & 1u);
                        }
                        if (!done[r] && rel[r]) {
                            if (acc[g] > best[r] || bestc[r] == 255) { best[r] = acc[g]; bestc[r] = ent[g].c; }
                            pass[r] |= ent_thr[g] < acc[g];
                        }
                    }
                }
            } else if (!done[r]) {
#pragma unroll
                for (int g = 0; g < kGroup; ++g) {
                    if (g < n && ent[g].first && !done[r]) {
                        if (pass[r]) { p.labels[pix[r]] = (uint8_t)bestc[r]; done[r] = true; }
                        best[r] = -INFINITY; bestc[r] = 255; pass[r] = false;
                        const int s2 = ent[g].s;
                        rel[r] = s2 == tid_[r] || ((p.adj[(size_t)tid_[r] * p.words + (s2 >> 5)] >> (s2 & 31)) & 1u);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kTileRowsPerWarp; ++r) {
        if (!done[r] && pass[r]) p.labels[pix[r]] = (uint8_t)bestc[r];
    }
}

// one launch instead of two memsets + the candidate-word kernel: zero the accumulating part of the workspace, fill the
// label map with 255 ("unlabeled") and build the candidate word of every superpixel (bit c: class c is a candidate of
// the net's C channels; top bit: the superpixel takes part -- always, or only when multi-hot; same as mas_multihot_info_dev)
__global__ void labeller_init_kernel(uint4* __restrict__ zero, size_t zero_vecs, uint8_t* __restrict__ labels, size_t P,
                                     const uint8_t* __restrict__ targets, int S, int Ct, int C, int only_multihot,
                                     uint32_t* __restrict__ info) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < zero_vecs; i += stride) zero[i] = make_uint4(0u, 0u, 0u, 0u);
    const size_t head = min(P, (size_t)((16 - (reinterpret_cast<uintptr_t>(labels) & 15)) & 15));
    const size_t vecs = (P - head) / 16;
    uint4* lab16 = reinterpret_cast<uint4*>(labels + head);
    for (size_t i = tid; i < vecs; i += stride) lab16[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    for (size_t i = tid; i < head; i += stride) labels[i] = 255;
    for (size_t i = head + vecs * 16 + tid; i < P; i += stride) labels[i] = 255;
    for (size_t r = tid; r < (size_t)S; r += stride) {
        const uint8_t* t = targets + r * Ct;
        uint32_t bits = 0u;
        int total = 0;
        for (int c = 0; c < Ct; ++c) {
            const int v = t[c];
            total += v;
            if (c < C && v) bits |= 1u << c;
        }
        info[r] = bits | ((!only_multihot || total > 1) ? kGroupBit : 0u);
    }
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct Workspace {
    uint32_t* info; unsigned long long* gmax; double* acc; int* count; int* offset; int* cursor; int* pixlist;
    float* proto; float* own_sim; uint8_t* own_cls; float* thr; uint32_t* adj; uint32_t* svalid; uint8_t* touched;
    size_t bytes;
};

Workspace carve(void* base, int F, int C, int H, int W, int S) {
    const size_t P = (size_t)H * W;
    const size_t words = ((size_t)S + 31) / 32;
    char* p = reinterpret_cast<char*>(base);
    size_t off = 0;
    Workspace w;
    auto take = [&](size_t bytes) { char* q = p ? p + off : nullptr; off += align_up(bytes); return q; };
    // zero-initialised block first (one memset): gmax, acc, count, adj, svalid
    w.gmax = reinterpret_cast<unsigned long long*>(take((size_t)S * C * 8));
    w.acc = reinterpret_cast<double*>(take(8 * sizeof(double)));
    w.count = reinterpret_cast<int*>(take((size_t)S * 4));
    w.adj = reinterpret_cast<uint32_t*>(take((size_t)S * words * 4));
    w.svalid = reinterpret_cast<uint32_t*>(take(words * 4));
    const size_t zeroed = off;
    w.info = reinterpret_cast<uint32_t*>(take((size_t)S * 4));
    w.offset = reinterpret_cast<int*>(take(((size_t)S + 1) * 4));
    w.cursor = reinterpret_cast<int*>(take((size_t)S * 4));
    w.pixlist = reinterpret_cast<int*>(take(P * 4));
    w.proto = reinterpret_cast<float*>(take((size_t)S * C * F * 4));
    w.own_sim = reinterpret_cast<float*>(take(P * 4));
    w.own_cls = reinterpret_cast<uint8_t*>(take(P));
    w.thr = reinterpret_cast<float*>(take((size_t)S * C * 4));
    w.touched = reinterpret_cast<uint8_t*>(take((size_t)S));
    w.bytes = off;
    (void)zeroed;
    return w;
}

size_t zeroed_bytes(int C, int S) {
    const size_t words = ((size_t)S + 31) / 32;
    return align_up((size_t)S * C * 8) + align_up(8 * sizeof(double)) + align_up((size_t)S * 4) + align_up((size_t)S * words * 4) +
           align_up(words * 4);
}

template <typename IdT, typename FT, bool LOWRES>
int run_labeller(const LabelParams& p, const Workspace& w, cudaStream_t st) {
    const int threads = 256;
    const unsigned grid_px = (unsigned)std::min<long long>(((long long)p.P + threads - 1) / threads, (long long)mas::sm_count() * 16);
    spx_count_adjacency_kernel<IdT><<<grid_px, threads, 0, st>>>(p.ids, p.H, p.W, p.S, w.count, p.words, w.adj);
    spx_scan_kernel<<<1, 1024, 0, st>>>(w.count, p.S, w.offset, w.cursor, w.acc, p.only_multihot, p.status);
    spx_fill_kernel<IdT><<<grid_px, threads, 0, st>>>(p.ids, p.P, p.S, w.offset, w.cursor, w.pixlist);
    const long long entries = (long long)p.S * p.C;
    proto_gather_kernel<FT, LOWRES><<<(unsigned)((entries * 32 + threads - 1) / threads), threads, 0, st>>>(p);
    const size_t smem = (size_t)kGroup * p.F * sizeof(float);
    const dim3 grid_sp((unsigned)p.S, kSlices);
    proto_assign_kernel<FT, LOWRES><<<grid_sp, kAssignThreads, smem, st>>>(p);
    proto_threshold_kernel<<<p.S, kAssignThreads, 0, st>>>(p);
    {
        // two traversals, same labels: spatial tiles (fewer DRAM bytes, more parallelism on small images; measured on a
        // 375 x 500 image: 0.256 -> 0.223 ms per image) or one CTA per superpixel slice (less per-CTA overhead; 1024 x 2048:
        // 0.57 vs 0.63 ms).  When the images of a batch overlap on several streams the extra parallelism of the tiles buys
        // nothing and their per-tile overhead costs (375 x 500, 8 lanes: 0.076 vs 0.092 ms per image): superpixel walk.
        // MAS_LABELLER_TILE = 0 / 1 forces one.
        const char* v = getenv("MAS_LABELLER_TILE");
        const bool tile = (v && (v[0] == '0' || v[0] == '1')) ? v[0] == '1' : (p.P <= (1 << 19) && !p.overlapped);
        if (tile) {
            const dim3 grid_t((unsigned)((p.W + kTileW - 1) / kTileW), (unsigned)((p.H + kTileH - 1) / kTileH));
            proto_propagate_tile_kernel<FT, LOWRES, IdT><<<grid_t, kAssignThreads, smem + 2 * (size_t)p.words * sizeof(uint32_t), st>>>(p);
        } else {
            proto_propagate_kernel<FT, LOWRES><<<grid_sp, kAssignThreads, smem, st>>>(p);
        }
    }
    mas::count_launches(7);
    MAS_LAUNCH_OK("prototype labeller kernels");
    return 0;
}

}  // namespace

extern "C" int mas_candidate_argmax_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask,
                                        const uint32_t* info, int n_img, int channels, int height, int width, int nseg,
                                        uint8_t* labels, void* stream) {
    MAS_REQUIRE(logits && ids && mask && info && labels, MAS_E_BADARG, "candidate_argmax: null pointer");
    MAS_REQUIRE(n_img >= 0 && height > 0 && width > 0 && nseg > 0, MAS_E_BADARG, "candidate_argmax: bad shape");
    MAS_REQUIRE(channels >= 1 && channels <= MAS_MAX_LOSS_CLASSES, MAS_E_RANGE, "candidate_argmax: channels out of range");
    MAS_REQUIRE(ids_dtype == MAS_I32 || ids_dtype == MAS_I64, MAS_E_BADARG, "candidate_argmax: bad ids dtype");
    if (n_img == 0) return 0;
    const long long P = (long long)height * width;
    MAS_REQUIRE(P < (1ll << 31), MAS_E_RANGE, "candidate_argmax: image too large");
    const long long n_pix = P * n_img;
    const int threads = 256;
    const unsigned blocks = (unsigned)std::min<long long>((n_pix + threads - 1) / threads, (long long)mas::sm_count() * 16);
    if (ids_dtype == MAS_I64)
        candidate_argmax_kernel<long long><<<blocks, threads, 0, (cudaStream_t)stream>>>(logits, ids, mask, info, n_pix, (int)P, channels, nseg, labels);
    else
        candidate_argmax_kernel<int32_t><<<blocks, threads, 0, (cudaStream_t)stream>>>(logits, ids, mask, info, n_pix, (int)P, channels, nseg, labels);
    mas::count_launches(1);
    MAS_LAUNCH_OK("candidate_argmax_kernel");
    return 0;
}

extern "C" size_t mas_proto_labeller_workspace_bytes(int feat_channels, int channels, int height, int width, int nseg) {
    if (feat_channels <= 0 || channels <= 0 || height <= 0 || width <= 0 || nseg <= 0) return 0;
    return carve(nullptr, feat_channels, channels, height, width, nseg).bytes;
}

namespace {

int proto_labeller_impl(const char* what, const void* feats, int feat_dtype, int feat_channels, int feat_height, int feat_width,
                        const float* logits, int channels, const uint8_t* targets, int target_channels, const uint8_t* mask,
                        const void* ids, int ids_dtype, int height, int width, int nseg, int only_multihot, int threshold_mode,
                        uint8_t* labels, int32_t* status, void* workspace, size_t workspace_bytes, void* stream, bool overlapped = false) {
    MAS_REQUIRE(feats && logits && targets && mask && ids && labels && status && workspace, MAS_E_BADARG, "%s: null pointer", what);
    MAS_REQUIRE(feat_channels > 0 && height > 0 && width > 0 && nseg > 0, MAS_E_BADARG, "%s: bad shape", what);
    MAS_REQUIRE(feat_dtype == MAS_F32 || feat_dtype == MAS_BF16, MAS_E_BADARG, "%s: bad feature dtype", what);
    MAS_REQUIRE(feat_height > 0 && feat_width > 0 && feat_height <= height && feat_width <= width, MAS_E_BADARG,
                "%s: the feature map (%d x %d) must not be larger than the image (%d x %d)", what, feat_height, feat_width, height, width);
    MAS_REQUIRE(channels >= 2 && channels <= MAS_MAX_LOSS_CLASSES && channels <= target_channels, MAS_E_RANGE,
                "%s: channels=%d must be in [2,%d] and <= target_channels", what, channels, MAS_MAX_LOSS_CLASSES);
    MAS_REQUIRE(ids_dtype == MAS_I32 || ids_dtype == MAS_I64, MAS_E_BADARG, "%s: bad ids dtype", what);
    MAS_REQUIRE(threshold_mode == MAS_THRESHOLD_MEDIAN || threshold_mode == MAS_THRESHOLD_MIN, MAS_E_BADARG, "%s: bad threshold mode", what);
    MAS_REQUIRE((long long)height * width < (1ll << 31), MAS_E_RANGE, "%s: image too large", what);
    MAS_REQUIRE((size_t)kGroup * feat_channels * sizeof(float) <= 48 * 1024, MAS_E_RANGE, "%s: feat_channels too large", what);
    MAS_REQUIRE(((uintptr_t)workspace) % 256 == 0, MAS_E_BADARG, "%s: workspace must be 256-byte aligned", what);
    const Workspace w = carve(workspace, feat_channels, channels, height, width, nseg);
    MAS_REQUIRE(workspace_bytes >= w.bytes, MAS_E_WORKSPACE, "%s: workspace too small (%zu < %zu)", what, workspace_bytes, w.bytes);
    cudaStream_t st = (cudaStream_t)stream;
    const long long P = (long long)height * width;

    {
        const size_t zero_vecs = zeroed_bytes(channels, nseg) / 16;       // every carved block is 256-byte aligned
        const size_t work = std::max<size_t>(std::max<size_t>(zero_vecs, (size_t)P / 16), (size_t)nseg);
        const unsigned blocks = (unsigned)std::min<size_t>((work + 255) / 256, (size_t)mas::sm_count() * 8);
        labeller_init_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<uint4*>(workspace), zero_vecs, labels, (size_t)P, targets, nseg,
                                                     target_channels, channels, only_multihot, w.info);
        mas::count_launches(1);
        MAS_LAUNCH_OK("labeller_init_kernel");
    }
    // arg-max-probability pixel per (superpixel, candidate class): softmax with T = 1 (:140); the group loss itself is not needed
    int rc = mas::multihot_loss_fwd(logits, ids, ids_dtype, mask, w.info, 1, channels, height, width, nseg, 1.0f,
                                    MAS_LOSS_CHOICE | MAS_LOSS_GROUP | MAS_LOSS_EXACT_SOFTMAX, w.acc, reinterpret_cast<uint64_t*>(w.gmax),
                                    false, stream, nullptr);
    if (rc != 0) return rc;

    LabelParams p = {};
    p.feats = feats; p.mask = mask; p.ids = ids;
    p.F = feat_channels; p.C = channels; p.H = height; p.W = width; p.S = nseg; p.P = (int)P;
    const bool lowres = feat_height != height || feat_width != width;
    p.fh_in = feat_height; p.fw_in = feat_width;
    // torch's area_pixel_compute_scale<float>(input, output, align_corners=false): (float)input / output
    p.fry = (float)feat_height / (float)height; p.frx = (float)feat_width / (float)width;
    p.threshold_min = threshold_mode == MAS_THRESHOLD_MIN;
    p.only_multihot = only_multihot; p.status = status;
    p.info = w.info; p.gmax = w.gmax; p.offset = w.offset; p.pixlist = w.pixlist; p.proto = w.proto;
    p.own_sim = w.own_sim; p.own_cls = w.own_cls; p.thr = w.thr; p.adj = w.adj; p.svalid = w.svalid; p.touched = w.touched;
    p.words = (nseg + 31) / 32;
    p.labels = labels;
    p.overlapped = overlapped ? 1 : 0;
    const bool i64 = ids_dtype == MAS_I64, bf16 = feat_dtype == MAS_BF16;
#define MAS_RUN(IdT)                                                                                                   \
    (bf16 ? (lowres ? run_labeller<IdT, __nv_bfloat16, true>(p, w, st) : run_labeller<IdT, __nv_bfloat16, false>(p, w, st)) \
          : (lowres ? run_labeller<IdT, float, true>(p, w, st) : run_labeller<IdT, float, false>(p, w, st)))
    return i64 ? MAS_RUN(long long) : MAS_RUN(int32_t);
#undef MAS_RUN
}

}  // namespace

extern "C" int mas_proto_labeller_dev(const float* feats, int feat_channels, const float* logits, int channels,
                                      const uint8_t* targets, int target_channels, const uint8_t* mask, const void* ids,
                                      int ids_dtype, int height, int width, int nseg, int only_multihot, int threshold_mode,
                                      uint8_t* labels, int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
    return proto_labeller_impl("proto_labeller", feats, MAS_F32, feat_channels, height, width, logits, channels, targets, target_channels,
                               mask, ids, ids_dtype, height, width, nseg, only_multihot, threshold_mode, labels, status, workspace,
                               workspace_bytes, stream);
}

extern "C" int mas_proto_labeller_src_dev(const void* feats, int feat_dtype, int feat_channels, int feat_height, int feat_width,
                                          const float* logits, int channels, const uint8_t* targets, int target_channels,
                                          const uint8_t* mask, const void* ids, int ids_dtype, int height, int width, int nseg,
                                          int only_multihot, int threshold_mode, uint8_t* labels, int32_t* status, void* workspace,
                                          size_t workspace_bytes, void* stream) {
    return proto_labeller_impl("proto_labeller_src", feats, feat_dtype, feat_channels, feat_height, feat_width, logits, channels, targets,
                               target_channels, mask, ids, ids_dtype, height, width, nseg, only_multihot, threshold_mode, labels, status,
                               workspace, workspace_bytes, stream);
}

// ------------------------------------------------------------------------------------------ a loader batch in one call
// The per-image pipeline is nine short, dependent launches; the images of a batch are independent.  This entry labels a
// whole batch from ONE host call: image i runs on lane_streams[i % n_lanes] with its own slice of the workspace, forked
// from / joined to `stream` with events, so the launch latencies and tails of up to n_lanes images overlap and the host
// pays one call instead of n (a VOC image is ~0.2 ms of mostly latency on its own).
namespace {

constexpr int kMaxLanes = 16;

struct LaneEvents {
    cudaEvent_t fork = nullptr, done[kMaxLanes] = {};
    std::mutex lock;
};

LaneEvents* lane_events_for_current_device() {
    static LaneEvents per_device[mas::kMaxDevices];
    return &per_device[mas::current_device()];
}

}  // namespace

extern "C" int mas_proto_labeller_batch_dev(const void* feats, int feat_dtype, int feat_channels, int feat_height, int feat_width,
                                            const float* logits, int channels, const uint8_t* targets, int target_channels,
                                            const uint8_t* mask, const void* ids, int ids_dtype, int n_img, int height, int width, int nseg,
                                            int only_multihot, int threshold_mode, uint8_t* labels, int32_t* status, void* workspace,
                                            size_t workspace_bytes_per_lane, void* const* lane_streams, int n_lanes, void* stream) {
    MAS_REQUIRE(feats && logits && targets && mask && ids && labels && status && workspace, MAS_E_BADARG, "proto_labeller_batch: null pointer");
    MAS_REQUIRE(n_img >= 0 && n_lanes >= 0 && n_lanes <= kMaxLanes && (n_lanes == 0 || lane_streams), MAS_E_BADARG,
                "proto_labeller_batch: bad image / lane count (at most %d lanes)", kMaxLanes);
    MAS_REQUIRE(feat_dtype == MAS_F32 || feat_dtype == MAS_BF16, MAS_E_BADARG, "proto_labeller_batch: bad feature dtype");
    MAS_REQUIRE(ids_dtype == MAS_I32 || ids_dtype == MAS_I64, MAS_E_BADARG, "proto_labeller_batch: bad ids dtype");
    MAS_REQUIRE(workspace_bytes_per_lane % 256 == 0, MAS_E_BADARG, "proto_labeller_batch: the per-lane workspace size must be a multiple of 256");
    if (n_img == 0) return 0;
    const size_t P = (size_t)height * width;
    const size_t feat_img = (size_t)feat_channels * feat_height * feat_width * (feat_dtype == MAS_F32 ? 4 : 2);
    const size_t id_img = P * (ids_dtype == MAS_I64 ? 8 : 4);
    auto image = [&](int i, void* ws, void* st) {
        return proto_labeller_impl("proto_labeller_batch", (const char*)feats + (size_t)i * feat_img, feat_dtype, feat_channels, feat_height,
                                   feat_width, logits + (size_t)i * channels * P, channels, targets + (size_t)i * nseg * target_channels,
                                   target_channels, mask + (size_t)i * P, (const char*)ids + (size_t)i * id_img, ids_dtype, height, width, nseg,
                                   only_multihot, threshold_mode, labels + (size_t)i * P, status + i, ws, workspace_bytes_per_lane, st,
                                   std::min(n_lanes, n_img) > 1);
    };
    const int lanes = std::min(n_lanes, n_img);
    if (lanes <= 1) {      // nothing to overlap: everything on the caller's stream (or the single lane's)
        for (int i = 0; i < n_img; ++i) {
            const int rc = image(i, workspace, stream);
            if (rc != 0) return rc;
        }
        return 0;
    }
    LaneEvents* ev = lane_events_for_current_device();
    std::lock_guard<std::mutex> guard(ev->lock);
    if (!ev->fork) MAS_CUDA_OK(cudaEventCreateWithFlags(&ev->fork, cudaEventDisableTiming));
    for (int l = 0; l < lanes; ++l) {
        if (!ev->done[l]) MAS_CUDA_OK(cudaEventCreateWithFlags(&ev->done[l], cudaEventDisableTiming));
    }
    cudaStream_t main_stream = (cudaStream_t)stream;
    MAS_CUDA_OK(cudaEventRecord(ev->fork, main_stream));
    for (int l = 0; l < lanes; ++l) MAS_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)lane_streams[l], ev->fork, 0));
    int rc = 0;
    for (int i = 0; i < n_img && rc == 0; ++i) {
        const int l = i % lanes;
        rc = image(i, (char*)workspace + (size_t)l * workspace_bytes_per_lane, lane_streams[l]);
    }
    // join even after an error: the caller's stream must not run ahead of lanes that already hold work
    for (int l = 0; l < lanes; ++l) {
        if (cudaEventRecord(ev->done[l], (cudaStream_t)lane_streams[l]) == cudaSuccess) cudaStreamWaitEvent(main_stream, ev->done[l], 0);
    }
    return rc;
}

