// Host-buffer entry points: the end-to-end path of the C ABI (copies inside the call).
//
// mas_acquisition_host streams logits + ids from (ideally pinned) host memory through two staging
// buffers: the H2D copy of chunk i+1 runs on the copy stream while the stats kernel of chunk i runs on
// the compute stream.  The (image, superpixel, class) tables stay on the device; only the per-region
// results travel back.
#include "common.cuh"

#include <math.h>
#include <string.h>
#include <map>
#include <mutex>
#include <vector>

namespace {

// Streams, events and grow-only device buffers of the host-buffer entry points, cached per device for the life of
// the process: cudaMalloc / cudaFree of the 1.4 GB staging area cost more than the scoring itself.  One call at a
// time per device (the mutex is held for the whole call).
struct HostContext {
    std::mutex lock;
    cudaStream_t copy = nullptr, compute = nullptr;
    cudaEvent_t filled[2] = {nullptr, nullptr}, drained[2] = {nullptr, nullptr};
    struct Slot { void* ptr = nullptr; size_t bytes = 0; };
    std::vector<Slot> slots;
    size_t next = 0;

    cudaError_t init() {
        if (compute) return cudaSuccess;
        cudaError_t e;
        if ((e = cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking)) != cudaSuccess) return e;
        for (int i = 0; i < 2; ++i) {
            if ((e = cudaEventCreateWithFlags(&filled[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&drained[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    void begin() { next = 0; }
    // the i-th buffer requested in a call is the i-th slot: calls of the same shape never reallocate
    template <typename T>
    cudaError_t alloc(T** out, size_t bytes) {
        if (bytes == 0) bytes = 1;
        if (next == slots.size()) slots.push_back(Slot());
        Slot& s = slots[next++];
        if (s.bytes < bytes) {
            if (s.ptr) cudaFree(s.ptr);
            s.ptr = nullptr; s.bytes = 0;
            cudaError_t e = cudaMalloc(&s.ptr, bytes);
            if (e != cudaSuccess) return e;
            s.bytes = bytes;
        }
        *out = reinterpret_cast<T*>(s.ptr);
        return cudaSuccess;
    }
};

// Error paths of the host entry points return while H2D copies from the CALLER's buffers may still be in flight: drain both
// streams before the caller gets its buffers back (the success path ends with its own synchronize and disarms the guard).
struct DrainOnExit {
    HostContext& d;
    bool armed = true;
    explicit DrainOnExit(HostContext& ctx) : d(ctx) {}
    ~DrainOnExit() {
        if (!armed) return;
        if (d.copy) cudaStreamSynchronize(d.copy);
        if (d.compute) cudaStreamSynchronize(d.compute);
    }
};

HostContext* context_for_current_device() {
    static std::mutex table_lock;
    static std::map<int, HostContext*> table;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> g(table_lock);
    auto it = table.find(dev);
    if (it == table.end()) it = table.emplace(dev, new HostContext()).first;   // lives until process exit
    return it->second;
}

}  // namespace

extern "C" int mas_acquisition_host(const void* logits, int logits_dtype, const int32_t* ids, int n_img, int channels,
                                    int height, int width, int nseg, float temperature, int weighting, float coeff,
                                    int ref_batch, int normalise, int ban_class, int clsbal, int chunk_img, float* score,
                                    int32_t* dominant, double* prob_sum) {
    MAS_REQUIRE(logits && ids && score, MAS_E_BADARG, "acquisition_host: null pointer");
    MAS_REQUIRE(n_img > 0 && channels >= 2 && channels <= MAS_MAX_CLASSES && height > 0 && width > 0 && nseg > 0,
                MAS_E_BADARG, "acquisition_host: bad shape");
    MAS_REQUIRE(logits_dtype == MAS_F32 || logits_dtype == MAS_BF16, MAS_E_BADARG, "acquisition_host: bad dtype");
    MAS_REQUIRE(weighting == MAS_WEIGHT_NONE || weighting == MAS_WEIGHT_PREDCLSBAL, MAS_E_BADARG, "acquisition_host: bad weighting");
    if (chunk_img <= 0) chunk_img = 4;
    if (chunk_img > n_img) chunk_img = n_img;
    if (ref_batch <= 0) ref_batch = 1;
    const bool need_prob = prob_sum != nullptr || weighting == MAS_WEIGHT_PREDCLSBAL;
    const bool need_dom = dominant != nullptr || ban_class >= 0 || clsbal;

    const size_t P = (size_t)height * width;
    const size_t elt = logits_dtype == MAS_F32 ? 4 : 2;
    const size_t img_logit_bytes = (size_t)channels * P * elt;
    const size_t img_id_bytes = P * sizeof(int32_t);
    const long long n_regions = (long long)n_img * nseg;
    const size_t table = (size_t)n_regions * channels;

    HostContext* ctx = context_for_current_device();
    MAS_REQUIRE(ctx, MAS_E_BADARG, "acquisition_host: no CUDA device");
    std::lock_guard<std::mutex> guard(ctx->lock);
    HostContext& d = *ctx;
    MAS_CUDA_OK(d.init());
    DrainOnExit drain(d);
    d.begin();
    char* stage_logits[2];
    int32_t* stage_ids[2];
    for (int i = 0; i < 2; ++i) {
        MAS_CUDA_OK(d.alloc(&stage_logits[i], img_logit_bytes * chunk_img));
        MAS_CUDA_OK(d.alloc(&stage_ids[i], img_id_bytes * chunk_img));
    }
    float *cls_sum, *d_score, *d_w = nullptr, *d_minmax = nullptr, *d_rw = nullptr;
    int32_t *cls_cnt, *d_dom = nullptr;
    double* d_prob = nullptr;
    int64_t* d_hist = nullptr;
    MAS_CUDA_OK(d.alloc(&cls_sum, table * sizeof(float)));
    MAS_CUDA_OK(d.alloc(&cls_cnt, table * sizeof(int32_t)));
    MAS_CUDA_OK(d.alloc(&d_score, (size_t)n_regions * sizeof(float)));
    if (need_dom) MAS_CUDA_OK(d.alloc(&d_dom, (size_t)n_regions * sizeof(int32_t)));
    if (need_prob) MAS_CUDA_OK(d.alloc(&d_prob, (size_t)n_img * channels * sizeof(double)));
    MAS_CUDA_OK(cudaMemsetAsync(cls_sum, 0, table * sizeof(float), d.compute));
    MAS_CUDA_OK(cudaMemsetAsync(cls_cnt, 0, table * sizeof(int32_t), d.compute));
    if (need_prob) MAS_CUDA_OK(cudaMemsetAsync(d_prob, 0, (size_t)n_img * channels * sizeof(double), d.compute));

    int chunk_index = 0;
    for (int first = 0; first < n_img; first += chunk_img, ++chunk_index) {
        const int b = chunk_index & 1;
        const int count = (n_img - first < chunk_img) ? (n_img - first) : chunk_img;
        if (chunk_index >= 2) MAS_CUDA_OK(cudaStreamWaitEvent(d.copy, d.drained[b], 0));
        MAS_CUDA_OK(cudaMemcpyAsync(stage_logits[b], (const char*)logits + (size_t)first * img_logit_bytes,
                                    img_logit_bytes * count, cudaMemcpyHostToDevice, d.copy));
        MAS_CUDA_OK(cudaMemcpyAsync(stage_ids[b], ids + (size_t)first * P, img_id_bytes * count, cudaMemcpyHostToDevice, d.copy));
        MAS_CUDA_OK(cudaEventRecord(d.filled[b], d.copy));
        MAS_CUDA_OK(cudaStreamWaitEvent(d.compute, d.filled[b], 0));
        int rc = mas_bvsb_segment_stats_dev(stage_logits[b], logits_dtype, 0, stage_ids[b], count, channels, height, width, nseg,
                                            temperature, cls_sum + (size_t)first * nseg * channels,
                                            cls_cnt + (size_t)first * nseg * channels,
                                            need_prob ? d_prob + (size_t)first * channels : nullptr, d.compute);
        if (rc != 0) return rc;
        MAS_CUDA_OK(cudaEventRecord(d.drained[b], d.compute));
    }

    std::vector<double> h_prob;
    if (need_prob) {
        h_prob.resize((size_t)n_img * channels);
        MAS_CUDA_OK(cudaMemcpyAsync(h_prob.data(), d_prob, h_prob.size() * sizeof(double), cudaMemcpyDeviceToHost, d.compute));
        MAS_CUDA_OK(cudaStreamSynchronize(d.compute));
        if (prob_sum) memcpy(prob_sum, h_prob.data(), h_prob.size() * sizeof(double));
    }
    if (weighting == MAS_WEIGHT_PREDCLSBAL) {
        // mean over reference batches of the per-batch mean probability (my_bvsb_predclsbal_pwr.py:41-46)
        std::vector<float> acc(channels, 0.f), w(channels);
        int n_batches = 0;
        for (int first = 0; first < n_img; first += ref_batch, ++n_batches) {
            const int count = (n_img - first < ref_batch) ? (n_img - first) : ref_batch;
            for (int c = 0; c < channels; ++c) {
                double s = 0.0;
                for (int i = 0; i < count; ++i) s += h_prob[(size_t)(first + i) * channels + c];
                acc[c] += (float)(s / ((double)count * (double)P));
            }
        }
        for (int c = 0; c < channels; ++c) {
            const float base = coeff * (acc[c] / (float)n_batches) + 1.f;
            w[c] = 1.f / (base * base);
        }
        MAS_CUDA_OK(d.alloc(&d_w, channels * sizeof(float)));
        MAS_CUDA_OK(cudaMemcpyAsync(d_w, w.data(), channels * sizeof(float), cudaMemcpyHostToDevice, d.compute));
        MAS_CUDA_OK(cudaStreamSynchronize(d.compute));  // `w` leaves scope below
    }
    int rc = mas_region_scores_dev(cls_sum, cls_cnt, d_w, n_regions, channels, d_score, nullptr, d_dom, d.compute);
    if (rc != 0) return rc;
    if (normalise) {
        MAS_CUDA_OK(d.alloc(&d_minmax, 2 * sizeof(float)));
        rc = mas_minmax_nonzero_dev(d_score, n_regions, d_minmax, d.compute);
        if (rc != 0) return rc;
    }
    if (clsbal) {
        MAS_CUDA_OK(d.alloc(&d_hist, channels * sizeof(int64_t)));
        MAS_CUDA_OK(cudaMemsetAsync(d_hist, 0, channels * sizeof(int64_t), d.compute));
        rc = mas_dominant_hist_dev(d_dom, n_regions, channels, d_hist, d.compute);
        if (rc != 0) return rc;
        std::vector<int64_t> hist(channels);
        MAS_CUDA_OK(cudaMemcpyAsync(hist.data(), d_hist, channels * sizeof(int64_t), cudaMemcpyDeviceToHost, d.compute));
        MAS_CUDA_OK(cudaStreamSynchronize(d.compute));
        std::vector<float> rw(channels);
        for (int c = 0; c < channels; ++c) rw[c] = expf(-((float)hist[c] / (float)n_regions));
        MAS_CUDA_OK(d.alloc(&d_rw, channels * sizeof(float)));
        MAS_CUDA_OK(cudaMemcpyAsync(d_rw, rw.data(), channels * sizeof(float), cudaMemcpyHostToDevice, d.compute));
        MAS_CUDA_OK(cudaStreamSynchronize(d.compute));
    }
    if (normalise || ban_class >= 0 || clsbal) {
        rc = mas_finalize_scores_dev(d_score, d_dom, n_regions, d_minmax, ban_class, d_rw, d.compute);
        if (rc != 0) return rc;
    }
    MAS_CUDA_OK(cudaMemcpyAsync(score, d_score, (size_t)n_regions * sizeof(float), cudaMemcpyDeviceToHost, d.compute));
    if (dominant) MAS_CUDA_OK(cudaMemcpyAsync(dominant, d_dom, (size_t)n_regions * sizeof(int32_t), cudaMemcpyDeviceToHost, d.compute));
    MAS_CUDA_OK(cudaStreamSynchronize(d.compute));
    drain.armed = false;    // every copy was issued on / waited for by the compute stream
    return 0;
}

extern "C" int mas_select_topk_host(const float* score, const uint8_t* in_pool, const int32_t* image_rank, int64_t n_img,
                                    int nseg, int64_t k, uint64_t* out_keys, int32_t* out_count) {
    MAS_REQUIRE(score && in_pool && image_rank && out_keys && out_count, MAS_E_BADARG, "select_topk_host: null pointer");
    MAS_REQUIRE(n_img > 0 && nseg > 0 && k > 0, MAS_E_BADARG, "select_topk_host: bad size");
    const long long n = (long long)n_img * nseg;
    const long long cap = mas_sort_capacity(k);
    HostContext* ctx = context_for_current_device();
    MAS_REQUIRE(ctx, MAS_E_BADARG, "select_topk_host: no CUDA device");
    std::lock_guard<std::mutex> guard(ctx->lock);
    HostContext& d = *ctx;
    MAS_CUDA_OK(d.init());
    DrainOnExit drain(d);
    d.begin();
    d.next = 16;   // slots 16.. : keeps the acquisition buffers of the same process untouched
    while (d.slots.size() < 16) d.slots.push_back(HostContext::Slot());
    float* d_score; uint8_t* d_pool; int32_t* d_rank; uint64_t *d_keys, *d_out; int32_t* d_count; void* d_ws;
    MAS_CUDA_OK(d.alloc(&d_score, n * sizeof(float)));
    MAS_CUDA_OK(d.alloc(&d_pool, n));
    MAS_CUDA_OK(d.alloc(&d_rank, n_img * sizeof(int32_t)));
    MAS_CUDA_OK(d.alloc(&d_keys, n * sizeof(uint64_t)));
    MAS_CUDA_OK(d.alloc(&d_out, cap * sizeof(uint64_t)));
    MAS_CUDA_OK(d.alloc(&d_count, sizeof(int32_t)));
    MAS_CUDA_OK(d.alloc(reinterpret_cast<char**>(&d_ws), mas_topk_workspace_bytes()));
    MAS_CUDA_OK(cudaMemcpyAsync(d_score, score, n * sizeof(float), cudaMemcpyHostToDevice, d.compute));
    MAS_CUDA_OK(cudaMemcpyAsync(d_pool, in_pool, n, cudaMemcpyHostToDevice, d.compute));
    MAS_CUDA_OK(cudaMemcpyAsync(d_rank, image_rank, n_img * sizeof(int32_t), cudaMemcpyHostToDevice, d.compute));
    MAS_CUDA_OK(cudaMemsetAsync(d_out, 0, cap * sizeof(uint64_t), d.compute));
    int rc = mas_region_keys_dev(d_score, d_pool, d_rank, n_img, nseg, d_keys, d.compute);
    if (rc != 0) return rc;
    rc = mas_topk_u64_dev(d_keys, n, k, d_out, d_count, d_ws, mas_topk_workspace_bytes(), d.compute);
    if (rc != 0) return rc;
    rc = mas_sort_desc_u64_dev(d_out, k, d.compute);
    if (rc != 0) return rc;
    MAS_CUDA_OK(cudaMemcpyAsync(out_keys, d_out, k * sizeof(uint64_t), cudaMemcpyDeviceToHost, d.compute));
    MAS_CUDA_OK(cudaMemcpyAsync(out_count, d_count, sizeof(int32_t), cudaMemcpyDeviceToHost, d.compute));
    MAS_CUDA_OK(cudaStreamSynchronize(d.compute));
    drain.armed = false;
    return 0;
}
