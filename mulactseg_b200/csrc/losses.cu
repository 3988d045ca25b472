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
#include "common.cuh"
#include "walk.cuh"

#include <algorithm>

namespace {

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
    int strips;
    long long total_rows;
    int do_choice, do_group;
    double* acc;                  // fwd: [0..5] one-hot / multi-hot / empty {sum, count}
    unsigned long long* gmax;     // (n_img * S * C) packed maxima
    const float* coef;            // bwd: {w_onehot, w_multihot, w_empty, w_group} = d total / d bucket sum
    float* grad;
};

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

// ------------------------------------------------------------------------------------------ loads
template <int VEC>
__device__ __forceinline__ uint32_t load_mask(const uint8_t* p);   // byte j of the result = mask of pixel j
template <>
__device__ __forceinline__ uint32_t load_mask<4>(const uint8_t* p) { return __ldcs(reinterpret_cast<const uint32_t*>(p)); }
template <>
__device__ __forceinline__ uint32_t load_mask<1>(const uint8_t* p) { return __ldcs(p); }

__device__ __forceinline__ int clamp_id(long long v) { return (v < 0 || v > 0x7fffffffLL) ? -1 : (int)v; }

template <typename IdT, int VEC>
struct IdLoad;
template <>
struct IdLoad<int32_t, 4> {
    static __device__ __forceinline__ void load(const int32_t* p, int (&o)[4]) {
        const int4 v = __ldcs(reinterpret_cast<const int4*>(p));
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
};
template <>
struct IdLoad<int32_t, 1> {
    static __device__ __forceinline__ void load(const int32_t* p, int (&o)[1]) { o[0] = __ldcs(p); }
};
template <>
struct IdLoad<long long, 4> {
    static __device__ __forceinline__ void load(const long long* p, int (&o)[4]) {
        const longlong2 a = __ldcs(reinterpret_cast<const longlong2*>(p));
        const longlong2 b = __ldcs(reinterpret_cast<const longlong2*>(p) + 1);
        o[0] = clamp_id(a.x); o[1] = clamp_id(a.y); o[2] = clamp_id(b.x); o[3] = clamp_id(b.y);
    }
};
template <>
struct IdLoad<long long, 1> {
    static __device__ __forceinline__ void load(const long long* p, int (&o)[1]) { o[0] = clamp_id(__ldcs(p)); }
};

template <int VEC>
__device__ __forceinline__ void load_logits(const float* p, float (&o)[VEC]);
template <>
__device__ __forceinline__ void load_logits<4>(const float* p, float (&o)[4]) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <>
__device__ __forceinline__ void load_logits<1>(const float* p, float (&o)[1]) { o[0] = __ldcs(p); }

template <int VEC>
__device__ __forceinline__ void store_grad(float* p, const float (&o)[VEC]);
template <>
__device__ __forceinline__ void store_grad<4>(float* p, const float (&o)[4]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(o[0], o[1], o[2], o[3]));
}
template <>
__device__ __forceinline__ void store_grad<1>(float* p, const float (&o)[1]) { __stcs(p, o[0]); }

// softmax(x / T) of pixel j in place: v[c][j] <- P_c.  Same operations as the reference's F.softmax(inputs / T):
// divide by T, subtract the maximum, accurate expf, divide by the sum -- the arg-max pixel of a max-pool and the
// -log of small probabilities are sensitive to the last bits, and only selected pixels pay for it.
template <int CMAX, int VEC>
__device__ __forceinline__ void softmax_inplace(float (&v)[CMAX][VEC], int j, float temp) {
#pragma unroll
    for (int c = 0; c < CMAX; ++c) v[c][j] = __fdiv_rn(v[c][j], temp);   // padded planes hold -inf
    float mx = v[0][j];
#pragma unroll
    for (int c = 1; c < CMAX; ++c) mx = fmaxf(mx, v[c][j]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        v[c][j] = expf(v[c][j] - mx);
        sum += v[c][j];
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c) v[c][j] = __fdiv_rn(v[c][j], sum);
}

// ------------------------------------------------------------------------------------------ forward
template <int CMAX, bool EXACT, int VEC, typename IdT>
__global__ void __launch_bounds__(kThreads) multihot_loss_fwd_kernel(const LossParams p) {
    extern __shared__ unsigned long long gcol[];   // [C][kThreads] running maxima of the thread's current superpixel
    const int tid = threadIdx.x, lane = tid & 31;
    const int C = EXACT ? CMAX : p.C;
    unsigned long long* col = gcol + tid;
    for (int c = 0; c < C; ++c) col[c * kThreads] = 0ull;

    long long r0, r1;
    mas::warp_range(p.total_rows, (long long)blockIdx.x * (kThreads / 32) + (tid >> 5), (long long)gridDim.x * (kThreads / 32), r0, r1);

    float sum_one = 0.f, sum_multi = 0.f, sum_empty = 0.f;
    int n_one = 0, n_multi = 0, n_empty = 0;
    int cur = -1;
    long long cur_base = 0;   // table offset of superpixel `cur`

    auto flush = [&]() {
        if (cur < 0) return;
        for (int c = 0; c < C; ++c) {
            const unsigned long long e = col[c * kThreads];
            if (e != 0ull) {
                atomicMax(p.gmax + cur_base + c, e);
                col[c * kThreads] = 0ull;
            }
        }
        cur = -1;
    };

    if (r0 < r1) {
        mas::Cursor at;
        at.seek(r0, p.strips, p.H);
        const size_t P = (size_t)p.H * p.W;
        for (long long r = r0; r < r1; ++r) {
            const int x0 = (at.strip * 32 + lane) * VEC;
            if (x0 < p.W) {
                const size_t off = (size_t)at.y * p.W + x0;
                const size_t pix0 = (size_t)at.img * P + off;
                const uint32_t m = load_mask<VEC>(p.mask + pix0);
                if (m != 0u) {
                    int id[VEC];
                    IdLoad<IdT, VEC>::load(reinterpret_cast<const IdT*>(p.ids) + pix0, id);
                    uint32_t inf[VEC];
                    bool any_valid = false, any_group = false, touches = false;
                    int first_group = -1;
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        const bool valid = ((m >> (8 * j)) & 0xffu) != 0u && (unsigned)id[j] < (unsigned)p.S;
                        inf[j] = 0u;
                        if (valid) {
                            inf[j] = __ldg(p.info + (size_t)at.img * p.S + id[j]);
                            any_valid = true;
                            if (p.do_group && (inf[j] & kGroupBit)) {
                                any_group = true;
                                touches |= (id[j] == cur);
                                if (first_group < 0) first_group = id[j];
                            }
                        } else {
                            id[j] = -1;
                        }
                    }
                    if (any_valid) {
                        float v[CMAX][VEC];
                        const float* base = p.logits + (size_t)at.img * C * P + off;
#pragma unroll
                        for (int c = 0; c < CMAX; ++c) {
                            if (EXACT || c < C) {
                                load_logits<VEC>(base + (size_t)c * P, v[c]);
                            } else {
#pragma unroll
                                for (int j = 0; j < VEC; ++j) v[c][j] = -INFINITY;
                            }
                        }
                        if (any_group && !touches) {
                            flush();
                            cur = first_group;
                            cur_base = ((long long)at.img * p.S + cur) * C;
                        }
#pragma unroll
                        for (int j = 0; j < VEC; ++j) {
                            if (id[j] < 0) continue;
                            softmax_inplace<CMAX, VEC>(v, j, p.temp);
                            const uint32_t bits = inf[j] & ~kGroupBit;
                            if (p.do_choice) {
                                float pos = 0.f;
#pragma unroll
                                for (int c = 0; c < CMAX; ++c) pos += ((bits >> c) & 1u) ? v[c][j] : 0.f;
                                const float l = -logf(pos + kEps);
                                const int n = __popc(bits);
                                if (n == 1) { sum_one += l; ++n_one; }
                                else if (n > 1) { sum_multi += l; ++n_multi; }
                                else { sum_empty += l; ++n_empty; }
                            }
                            if (p.do_group && (inf[j] & kGroupBit)) {
                                const unsigned long long low = (unsigned long long)(~(uint32_t)(off + j));
                                const bool own = id[j] == cur;
                                const long long tbase = ((long long)at.img * p.S + id[j]) * C;
#pragma unroll
                                for (int c = 0; c < CMAX; ++c) {
                                    if ((bits >> c) & 1u) {
                                        const unsigned long long key = ((unsigned long long)__float_as_uint(v[c][j]) << 32) | low;
                                        if (own) {
                                            const unsigned long long old = col[c * kThreads];
                                            if (key > old) col[c * kThreads] = key;
                                        } else {
                                            atomicMax(p.gmax + tbase + c, key);
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
            }
            const int step = at.advance(p.strips, p.H);
            if (step == 2) flush();
        }
        flush();
    }

    if (p.do_choice) {
        float s[3] = {sum_one, sum_multi, sum_empty};
        int n[3] = {n_one, n_multi, n_empty};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
                n[k] += __shfl_xor_sync(0xffffffffu, n[k], o);
            }
            if (lane == 0 && n[k] != 0) {
                atomicAdd(p.acc + 2 * k, (double)s[k]);
                atomicAdd(p.acc + 2 * k + 1, (double)n[k]);
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
template <int CMAX, bool EXACT, int VEC, typename IdT>
__global__ void __launch_bounds__(kThreads) multihot_loss_bwd_kernel(const LossParams p) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int C = EXACT ? CMAX : p.C;
    long long r0, r1;
    mas::warp_range(p.total_rows, (long long)blockIdx.x * (kThreads / 32) + (tid >> 5), (long long)gridDim.x * (kThreads / 32), r0, r1);
    if (r0 >= r1) return;
    const float w_one = p.coef[0] * p.inv_temp, w_multi = p.coef[1] * p.inv_temp, w_group = p.coef[3] * p.inv_temp;
    mas::Cursor at;
    at.seek(r0, p.strips, p.H);
    const size_t P = (size_t)p.H * p.W;
    for (long long r = r0; r < r1; ++r, at.advance(p.strips, p.H)) {
        const int x0 = (at.strip * 32 + lane) * VEC;
        if (x0 >= p.W) continue;
        const size_t off = (size_t)at.y * p.W + x0;
        const size_t pix0 = (size_t)at.img * P + off;
        float* gbase = p.grad + (size_t)at.img * C * P + off;
        const uint32_t m = load_mask<VEC>(p.mask + pix0);
        int id[VEC];
        uint32_t inf[VEC];
        bool any_valid = false;
        if (m != 0u) {
            IdLoad<IdT, VEC>::load(reinterpret_cast<const IdT*>(p.ids) + pix0, id);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const bool valid = ((m >> (8 * j)) & 0xffu) != 0u && (unsigned)id[j] < (unsigned)p.S;
                inf[j] = valid ? __ldg(p.info + (size_t)at.img * p.S + id[j]) : 0u;
                if (!valid) id[j] = -1;
                any_valid |= valid;
            }
        }
        if (!any_valid) {
            float z[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) z[j] = 0.f;
            for (int c = 0; c < C; ++c) store_grad<VEC>(gbase + (size_t)c * P, z);
            continue;
        }
        float v[CMAX][VEC];
        const float* base = p.logits + (size_t)at.img * C * P + off;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            if (EXACT || c < C) {
                load_logits<VEC>(base + (size_t)c * P, v[c]);
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) v[c][j] = -INFINITY;
            }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            if (id[j] < 0) {
#pragma unroll
                for (int c = 0; c < CMAX; ++c) v[c][j] = 0.f;
                continue;
            }
            softmax_inplace<CMAX, VEC>(v, j, p.temp);
            const uint32_t bits = inf[j] & ~kGroupBit;
            float pos = 0.f;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) pos += ((bits >> c) & 1u) ? v[c][j] : 0.f;
            float a = 0.f;
            if (p.do_choice) {
                const int n = __popc(bits);
                const float wb = n == 1 ? w_one : (n > 1 ? w_multi : 0.f);
                a = -wb / (pos + kEps);
            }
            uint32_t abits = 0u;   // classes whose max-pooled probability comes from this pixel
            float q_sum = 0.f;
            if (p.do_group && (inf[j] & kGroupBit)) {
                const uint32_t low = ~(uint32_t)(off + j);
                const unsigned long long* row = p.gmax + ((long long)at.img * p.S + id[j]) * C;
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
                    if ((bits >> c) & 1u) {
                        const unsigned long long e = __ldg(row + c);
                        if ((uint32_t)e == low && (e >> 32) != 0ull) {
                            abits |= 1u << c;
                            q_sum += v[c][j] / (v[c][j] + kEps);
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                const float pc = v[c][j];
                float g = a * pc * ((((bits >> c) & 1u) ? 1.f : 0.f) - pos);
                if (abits) g -= w_group * ((((abits >> c) & 1u) ? pc / (pc + kEps) : 0.f) - pc * q_sum);
                v[c][j] = g;
            }
        }
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            if (EXACT || c < C) store_grad<VEC>(gbase + (size_t)c * P, v[c]);
        }
    }
}

// ------------------------------------------------------------------------------------------ launch
template <typename K>
int resident_blocks(K kernel, size_t smem) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, smem) != cudaSuccess || n < 1) n = 1;
    return n;
}

template <int CMAX, bool EXACT, int VEC, typename IdT>
cudaError_t launch_one(LossParams p, bool backward, cudaStream_t stream) {
    p.strips = (p.W + 32 * VEC - 1) / (32 * VEC);
    p.total_rows = (long long)p.n_img * p.strips * p.H;
    const long long cap = (p.total_rows + 8 * (kThreads / 32) - 1) / (8 * (kThreads / 32));
    if (backward) {
        auto kernel = multihot_loss_bwd_kernel<CMAX, EXACT, VEC, IdT>;
        static int per_sm = 0;
        if (per_sm == 0) per_sm = resident_blocks(kernel, 0);
        const long long blocks = std::max<long long>(1, std::min<long long>((long long)mas::sm_count() * per_sm, cap));
        kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(p);
    } else {
        auto kernel = multihot_loss_fwd_kernel<CMAX, EXACT, VEC, IdT>;
        const size_t smem = (size_t)p.C * kThreads * sizeof(unsigned long long);
        static int per_sm = 0;
        if (per_sm == 0) per_sm = resident_blocks(kernel, (size_t)CMAX * kThreads * sizeof(unsigned long long));
        const long long blocks = std::max<long long>(1, std::min<long long>((long long)mas::sm_count() * per_sm, cap));
        kernel<<<(unsigned)blocks, kThreads, smem, stream>>>(p);
    }
    mas::count_launches(1);
    return cudaGetLastError();
}

template <int VEC, typename IdT>
cudaError_t dispatch_channels(const LossParams& p, bool backward, cudaStream_t stream) {
    switch (p.C) {
        case 19: return launch_one<19, true, VEC, IdT>(p, backward, stream);
        case 20: return launch_one<20, true, VEC, IdT>(p, backward, stream);
        case 21: return launch_one<21, true, VEC, IdT>(p, backward, stream);
        case 22: return launch_one<22, true, VEC, IdT>(p, backward, stream);
        default: break;
    }
    if (p.C <= 8) return launch_one<8, false, VEC, IdT>(p, backward, stream);
    if (p.C <= 16) return launch_one<16, false, VEC, IdT>(p, backward, stream);
    if (p.C <= 24) return launch_one<24, false, VEC, IdT>(p, backward, stream);
    return launch_one<31, false, VEC, IdT>(p, backward, stream);
}

cudaError_t dispatch(const LossParams& p, int ids_dtype, bool backward, cudaStream_t stream) {
    const size_t id_bytes = ids_dtype == MAS_I64 ? 8 : 4;
    const bool vec4 = (p.W % 4 == 0) && (((uintptr_t)p.logits) % 16 == 0) && (((uintptr_t)p.ids) % (4 * id_bytes) == 0) &&
                      (((uintptr_t)p.mask) % 4 == 0) && (!backward || ((uintptr_t)p.grad) % 16 == 0);
    if (ids_dtype == MAS_I64)
        return vec4 ? dispatch_channels<4, long long>(p, backward, stream) : dispatch_channels<1, long long>(p, backward, stream);
    return vec4 ? dispatch_channels<4, int32_t>(p, backward, stream) : dispatch_channels<1, int32_t>(p, backward, stream);
}

int check_common(const char* what, const void* logits, const void* ids, int ids_dtype, const uint8_t* mask, const uint32_t* info,
                 int n_img, int channels, int height, int width, int nseg, float temperature, int flags) {
    MAS_REQUIRE(logits && ids && mask && info, MAS_E_BADARG, "%s: null pointer", what);
    MAS_REQUIRE(n_img >= 0 && height > 0 && width > 0 && nseg > 0, MAS_E_BADARG, "%s: bad shape", what);
    MAS_REQUIRE(channels >= 2 && channels <= MAS_MAX_LOSS_CLASSES, MAS_E_RANGE, "%s: channels=%d outside [2,%d]", what, channels,
                MAS_MAX_LOSS_CLASSES);
    MAS_REQUIRE(ids_dtype == MAS_I32 || ids_dtype == MAS_I64, MAS_E_BADARG, "%s: bad ids dtype", what);
    MAS_REQUIRE(temperature > 0.f, MAS_E_BADARG, "%s: temperature must be > 0", what);
    MAS_REQUIRE((flags & ~(MAS_LOSS_CHOICE | MAS_LOSS_GROUP)) == 0 && flags != 0, MAS_E_BADARG, "%s: bad flags", what);
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

extern "C" int mas_multihot_loss_fwd_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask,
                                         const uint32_t* info, int n_img, int channels, int height, int width, int nseg,
                                         float temperature, int flags, double* acc, uint64_t* group_max, void* stream) {
    int rc = check_common("multihot_loss_fwd", logits, ids, ids_dtype, mask, info, n_img, channels, height, width, nseg, temperature, flags);
    if (rc != 0) return rc;
    MAS_REQUIRE(acc, MAS_E_BADARG, "multihot_loss_fwd: null acc");
    MAS_REQUIRE(!(flags & MAS_LOSS_GROUP) || group_max, MAS_E_BADARG, "multihot_loss_fwd: group_max required with MAS_LOSS_GROUP");
    if (n_img == 0) return 0;
    LossParams p = {};
    p.logits = logits; p.ids = ids; p.mask = mask; p.info = info;
    p.n_img = n_img; p.C = channels; p.H = height; p.W = width; p.S = nseg;
    p.temp = temperature; p.inv_temp = 1.f / temperature;
    p.do_choice = (flags & MAS_LOSS_CHOICE) ? 1 : 0; p.do_group = (flags & MAS_LOSS_GROUP) ? 1 : 0;
    p.acc = acc; p.gmax = reinterpret_cast<unsigned long long*>(group_max);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = dispatch(p, ids_dtype, false, st);
    if (e != cudaSuccess) return mas::cuda_fail(e, "multihot_loss_fwd_kernel launch");
    if (p.do_group) {
        const long long n_regions = (long long)n_img * nseg;
        const int threads = 256;
        const long long blocks = std::min<long long>((n_regions + threads - 1) / threads, (long long)mas::sm_count() * 4);
        group_loss_reduce_kernel<<<(unsigned)blocks, threads, 0, st>>>(p.gmax, info, n_regions, channels, acc);
        mas::count_launches(1);
        MAS_LAUNCH_OK("group_loss_reduce_kernel");
    }
    return 0;
}

extern "C" int mas_multihot_loss_bwd_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask,
                                         const uint32_t* info, const uint64_t* group_max, const float* coef,
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
    p.temp = temperature; p.inv_temp = 1.f / temperature;
    p.do_choice = (flags & MAS_LOSS_CHOICE) ? 1 : 0; p.do_group = (flags & MAS_LOSS_GROUP) ? 1 : 0;
    p.gmax = const_cast<unsigned long long*>(reinterpret_cast<const unsigned long long*>(group_max));
    p.coef = coef; p.grad = grad_logits;
    cudaError_t e = dispatch(p, ids_dtype, true, (cudaStream_t)stream);
    if (e != cudaSuccess) return mas::cuda_fail(e, "multihot_loss_bwd_kernel launch");
    return 0;
}
