// Offline multi-hot label generation: per-superpixel class histogram of the
// ground-truth map, with the optional boundary trimming.
//
// Reference: dataloader/region_cityscapes_tensor.py:33-84 (__getitem__), driven by tools/label_assignment_tensor.py:50-67:
//   bdry = binary_dilation(find_boundaries(superpixel, mode='thick'), ones(k, k));  trimmed = superpixel with bdry -> nseg
//   for p in preserving_labels:  mask = (trimmed == p) if any else (superpixel == p)
//       u, c = np.unique(target[mask]);  classes present -> multi-hot row, 255 -> last column;  size = mask.sum()
// i.e. a python loop of np.unique over <= 2048 ids per image with CPU morphology; here one pass builds the two
// (superpixel x class) histograms (trimmed / untrimmed) and a second tiny kernel picks per superpixel.
//   find_boundaries(mode='thick', connectivity 1) (scikit-image 0.19.2, actsegmul.yml:106): a pixel is a boundary when the
//   maximum and minimum id over its in-image 4-neighbourhood (itself included) differ; binary_dilation with ones(k,k)
//   (scipy.ndimage semantics, centre k // 2, outside = False): trimmed(y,x) iff a boundary pixel lies at
//   (y - dy, x - dx) for some dy, dx in [-(k//2), k-1-k//2].
#include "common.cuh"

#include <algorithm>

namespace {

template <typename IdT>
__device__ __forceinline__ long long raw_id(const void* ids, size_t i) { return (long long)reinterpret_cast<const IdT*>(ids)[i]; }

template <typename IdT>
__device__ __forceinline__ bool is_boundary(const void* ids, int y, int x, int H, int W) {
    const long long c = raw_id<IdT>(ids, (size_t)y * W + x);
    long long lo = c, hi = c;
    if (x > 0) { const long long v = raw_id<IdT>(ids, (size_t)y * W + x - 1); lo = min(lo, v); hi = max(hi, v); }
    if (x + 1 < W) { const long long v = raw_id<IdT>(ids, (size_t)y * W + x + 1); lo = min(lo, v); hi = max(hi, v); }
    if (y > 0) { const long long v = raw_id<IdT>(ids, (size_t)(y - 1) * W + x); lo = min(lo, v); hi = max(hi, v); }
    if (y + 1 < H) { const long long v = raw_id<IdT>(ids, (size_t)(y + 1) * W + x); lo = min(lo, v); hi = max(hi, v); }
    return lo != hi;
}

// hist[(variant * S + s) * (C + 1) + bin] += 1;  variant 0 = untrimmed, 1 = trimmed;  bin = class, C for label 255
template <typename IdT>
__global__ void label_hist_kernel(const void* __restrict__ ids, const uint8_t* __restrict__ target, int H, int W, int S, int C,
                                  int trim_k, int32_t* __restrict__ hist) {
    const long long P = (long long)H * W;
    const int bins = C + 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        const long long s = raw_id<IdT>(ids, (size_t)i);
        if (s < 0 || s >= S) continue;
        const int t = target[i];
        const int bin = t == 255 ? C : t;
        if (bin > C) continue;                       // not a train id: the reference would index out of range
        atomicAdd(hist + (size_t)s * bins + bin, 1);
        if (trim_k > 0) {
            const int y = (int)(i / W), x = (int)(i - (long long)y * W);
            const int c = trim_k / 2;
            bool trimmed = false;
            for (int dy = -c; dy <= trim_k - 1 - c && !trimmed; ++dy) {
                const int yy = y - dy;
                if (yy < 0 || yy >= H) continue;
                for (int dx = -c; dx <= trim_k - 1 - c; ++dx) {
                    const int xx = x - dx;
                    if (xx < 0 || xx >= W) continue;
                    if (is_boundary<IdT>(ids, yy, xx, H, W)) { trimmed = true; break; }
                }
            }
            if (!trimmed) atomicAdd(hist + ((size_t)S + s) * bins + bin, 1);
        }
    }
}

// one thread per superpixel: choose the trimmed histogram unless it is empty, emit the multi-hot row and the size
__global__ void label_rows_kernel(const int32_t* __restrict__ hist, const uint8_t* __restrict__ keep, int S, int C, int trim_k,
                                  uint8_t* __restrict__ multi_hot, int32_t* __restrict__ size) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int bins = C + 1;
    uint8_t* row = multi_hot + (size_t)s * bins;
    if (!keep[s]) {
        for (int b = 0; b < bins; ++b) row[b] = 0;
        size[s] = -1;
        return;
    }
    const int32_t* h = hist + (size_t)s * bins;
    if (trim_k > 0) {
        const int32_t* ht = hist + ((size_t)S + s) * bins;
        int n = 0;
        for (int b = 0; b < bins; ++b) n += ht[b];
        if (n > 0) h = ht;                            // "prevent disappearing because of the boundary" (:58)
    }
    int n = 0;
    for (int b = 0; b < bins; ++b) { row[b] = h[b] > 0 ? 1 : 0; n += h[b]; }
    size[s] = n;
}

// ---------------------------------------------------------------------------------------- dominant label assignment
// dataloader/region_dataset.py:201-240 (RegionCityscapesDominantAll.__getitem__; tools/label_assignment_dominant.py):
//   for p in preserving_labels:  mask = (superpixel == p) & (target != 255);  u, c = np.unique(target[mask]);
//       target[mask] = u[c.argmax()]            -- the smallest class wins a tie (np.unique sorts, argmax takes the first)
// i.e. the untrimmed histogram above, an arg-max per superpixel and one relabelling pass.
__global__ void dominant_pick_kernel(const int32_t* __restrict__ hist, const uint8_t* __restrict__ keep, int S, int C,
                                     uint8_t* __restrict__ dom) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int best = 0, arg = 255;
    if (keep[s]) {
        const int32_t* h = hist + (size_t)s * (C + 1);
        for (int c = 0; c < C; ++c) {
            if (h[c] > best) { best = h[c]; arg = c; }
        }
    }
    dom[s] = (uint8_t)arg;
}

template <typename IdT>
__global__ void dominant_relabel_kernel(const void* __restrict__ ids, const uint8_t* __restrict__ target, long long P, int S, int C,
                                        const uint8_t* __restrict__ dom, uint8_t* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        const uint8_t t = target[i];
        uint8_t o = t;
        if (t < C) {                                  // ignore (255) and anything that is not a train id stay as they are
            const long long s = raw_id<IdT>(ids, (size_t)i);
            if (s >= 0 && s < S && dom[s] != 255) o = dom[s];
        }
        out[i] = o;
    }
}

}  // namespace

extern "C" size_t mas_multihot_labels_workspace_bytes(int nseg, int num_classes) {
    if (nseg <= 0 || num_classes <= 0) return 0;
    return (size_t)2 * nseg * (num_classes + 1) * sizeof(int32_t);
}

extern "C" int mas_multihot_labels_dev(const void* ids, int ids_dtype, const uint8_t* target, const uint8_t* keep,
                                       int height, int width, int nseg, int num_classes, int trim_kernel_size,
                                       uint8_t* multi_hot, int32_t* size, void* workspace, size_t workspace_bytes, void* stream) {
    MAS_REQUIRE(ids && target && keep && multi_hot && size && workspace, MAS_E_BADARG, "multihot_labels: null pointer");
    MAS_REQUIRE(height > 0 && width > 0 && nseg > 0 && num_classes > 0 && num_classes < 255, MAS_E_BADARG, "multihot_labels: bad shape");
    MAS_REQUIRE(trim_kernel_size >= 0 && trim_kernel_size <= 15, MAS_E_RANGE, "multihot_labels: trim_kernel_size outside [0,15]");
    MAS_REQUIRE(ids_dtype == MAS_I32 || ids_dtype == MAS_I64, MAS_E_BADARG, "multihot_labels: bad ids dtype");
    const size_t need = mas_multihot_labels_workspace_bytes(nseg, num_classes);
    MAS_REQUIRE(workspace_bytes >= need, MAS_E_WORKSPACE, "multihot_labels: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    MAS_CUDA_OK(cudaMemsetAsync(workspace, 0, need, st));
    int32_t* hist = reinterpret_cast<int32_t*>(workspace);
    const long long P = (long long)height * width;
    const int threads = 256;
    const unsigned blocks = (unsigned)std::min<long long>((P + threads - 1) / threads, (long long)mas::sm_count() * 16);
    if (ids_dtype == MAS_I64)
        label_hist_kernel<long long><<<blocks, threads, 0, st>>>(ids, target, height, width, nseg, num_classes, trim_kernel_size, hist);
    else
        label_hist_kernel<int32_t><<<blocks, threads, 0, st>>>(ids, target, height, width, nseg, num_classes, trim_kernel_size, hist);
    label_rows_kernel<<<(nseg + threads - 1) / threads, threads, 0, st>>>(hist, keep, nseg, num_classes, trim_kernel_size, multi_hot, size);
    mas::count_launches(2);
    MAS_LAUNCH_OK("multihot_labels kernels");
    return 0;
}

extern "C" size_t mas_dominant_labels_workspace_bytes(int nseg, int num_classes) {
    if (nseg <= 0 || num_classes <= 0) return 0;
    return (((size_t)nseg * (num_classes + 1) * sizeof(int32_t) + 255) & ~(size_t)255) + (size_t)nseg;
}

extern "C" int mas_dominant_labels_dev(const void* ids, int ids_dtype, const uint8_t* target, const uint8_t* keep, int height,
                                       int width, int nseg, int num_classes, uint8_t* out, void* workspace, size_t workspace_bytes,
                                       void* stream) {
    MAS_REQUIRE(ids && target && keep && out && workspace, MAS_E_BADARG, "dominant_labels: null pointer");
    MAS_REQUIRE(height > 0 && width > 0 && nseg > 0 && num_classes > 0 && num_classes < 255, MAS_E_BADARG, "dominant_labels: bad shape");
    MAS_REQUIRE(ids_dtype == MAS_I32 || ids_dtype == MAS_I64, MAS_E_BADARG, "dominant_labels: bad ids dtype");
    MAS_REQUIRE(workspace_bytes >= mas_dominant_labels_workspace_bytes(nseg, num_classes), MAS_E_WORKSPACE, "dominant_labels: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t hist_bytes = (size_t)nseg * (num_classes + 1) * sizeof(int32_t);
    MAS_CUDA_OK(cudaMemsetAsync(workspace, 0, hist_bytes, st));
    int32_t* hist = reinterpret_cast<int32_t*>(workspace);
    uint8_t* dom = reinterpret_cast<uint8_t*>(workspace) + ((hist_bytes + 255) & ~(size_t)255);
    const long long P = (long long)height * width;
    const int threads = 256;
    const unsigned blocks = (unsigned)std::min<long long>((P + threads - 1) / threads, (long long)mas::sm_count() * 16);
    if (ids_dtype == MAS_I64)
        label_hist_kernel<long long><<<blocks, threads, 0, st>>>(ids, target, height, width, nseg, num_classes, 0, hist);
    else
        label_hist_kernel<int32_t><<<blocks, threads, 0, st>>>(ids, target, height, width, nseg, num_classes, 0, hist);
    dominant_pick_kernel<<<(nseg + threads - 1) / threads, threads, 0, st>>>(hist, keep, nseg, num_classes, dom);
    if (ids_dtype == MAS_I64)
        dominant_relabel_kernel<long long><<<blocks, threads, 0, st>>>(ids, target, P, nseg, num_classes, dom, out);
    else
        dominant_relabel_kernel<int32_t><<<blocks, threads, 0, st>>>(ids, target, P, nseg, num_classes, dom, out);
    mas::count_launches(3);
    MAS_LAUNCH_OK("dominant_labels kernels");
    return 0;
}
