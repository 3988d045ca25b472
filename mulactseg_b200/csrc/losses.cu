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
    float scale;     // log2(e) / T
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

// ------------------------------------------------------------------------------------------ per-pixel pieces
// Work decomposition of both kernels: one WARP per tile of 32 pixels x kTileRows rows (lane = column), one
// pixel per thread per row, many more tiles than resident warps: the hardware scheduler balances the very uneven
// cost of rows (an unselected row is one 32-byte mask load; a selected pixel is a C'-wide softmax).  A thread walks
// DOWN its column, so it stays inside one superpixel for many rows (private running maxima, see the header).
constexpr int kTileRows = 16;

__device__ __forceinline__ int clamp_id(long long v) { return (v < 0 || v > 0x7fffffffLL) ? -1 : (int)v; }

template <typename IdT>
__device__ __forceinline__ int load_id(const void* ids, size_t i) {
    return clamp_id((long long)__ldcs(reinterpret_cast<const IdT*>(ids) + i));
}

// softmax(x / T) in place: v[c] <- P_c.
//  ACCURATE: the reference's own operation order (divide by T, subtract the maximum, expf, divide by the sum) -- used
//            where the arg-max PIXEL of a max-pool must agree with torch to the last bit (stage-2 prototypes);
//  fast    : ex2.approx((x - max) * log2e / T) and one reciprocal (<= 4 ulp): the losses (1e-5 tolerance).
template <int CMAX, bool ACCURATE>
__device__ __forceinline__ void softmax_inplace(float (&v)[CMAX], float temp, float scale) {
    if (ACCURATE) {
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
#pragma unroll
        for (int c = 0; c < CMAX; ++c) v[c] = __fdiv_rn(v[c], sa);
    } else {
        const float shift = -mx * scale;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            v[c] = mas::ex2_approx(fmaf(v[c], scale, shift));
            if (c & 1) sb += v[c]; else sa += v[c];
        }
        const float inv = 1.f / (sa + sb);
#pragma unroll
        for (int c = 0; c < CMAX; ++c) v[c] *= inv;
    }
}

struct Tile {
    int img, x, y0, y1;
    bool in_range;
};

__device__ __forceinline__ Tile my_tile(const LossParams& p) {
    Tile t;
    const long long tile = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int tiles_x = (p.W + 31) / 32, tiles_y = (p.H + kTileRows - 1) / kTileRows;
    const long long per_img = (long long)tiles_x * tiles_y;
    t.in_range = tile < per_img * p.n_img;
    t.img = (int)(tile / per_img);
    const int rem = (int)(tile - (long long)t.img * per_img);
    const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
    t.x = tx * 32 + (threadIdx.x & 31);
    t.y0 = ty * kTileRows;
    t.y1 = min(p.H, t.y0 + kTileRows);
    return t;
}

// ------------------------------------------------------------------------------------------ forward
template <int CMAX, bool EXACT, typename IdT, bool ACCURATE>
__global__ void __launch_bounds__(kThreads) multihot_loss_fwd_kernel(const LossParams p) {
    extern __shared__ unsigned long long gcol[];   // [C][kThreads] running maxima of the thread's current superpixel
    const int tid = threadIdx.x, lane = tid & 31;
    const int C = EXACT ? CMAX : p.C;
    unsigned long long* col = gcol + tid;
    const Tile t = my_tile(p);
    float sum_one = 0.f, sum_multi = 0.f, sum_empty = 0.f;
    int n_one = 0, n_multi = 0, n_empty = 0;

    if (t.in_range) {     // warp-uniform
        if (p.do_group) {
            for (int c = 0; c < C; ++c) col[c * kThreads] = 0ull;
        }
        int cur = -1;
        long long cur_base = 0;   // table offset of superpixel `cur`
        const size_t P = (size_t)p.H * p.W;
        const bool active = t.x < p.W;
        for (int y = t.y0; y < t.y1; ++y) {
            const size_t off = (size_t)y * p.W + t.x;
            const size_t pix = (size_t)t.img * P + off;
            const bool m = active && __ldcs(p.mask + pix) != 0;
            if (!__any_sync(0xffffffffu, m)) continue;          // whole row unselected: nothing else is read
            if (!m) continue;
            const int id = load_id<IdT>(p.ids, pix);
            if ((unsigned)id >= (unsigned)p.S) continue;
            const uint32_t inf = __ldg(p.info + (size_t)t.img * p.S + id);
            const uint32_t bits = inf & ~kGroupBit;
            const bool group = p.do_group && (inf & kGroupBit) && bits != 0u;
            if (!p.do_choice && !group) continue;
            float v[CMAX];
            const float* base = p.logits + (size_t)t.img * C * P + off;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) v[c] = (EXACT || c < C) ? __ldcs(base + (size_t)c * P) : -INFINITY;
            softmax_inplace<CMAX, ACCURATE>(v, p.temp, p.scale);
            if (p.do_choice) {
                float pos = 0.f;
#pragma unroll
                for (int c = 0; c < CMAX; ++c) pos += ((bits >> c) & 1u) ? v[c] : 0.f;
                const float l = -logf(pos + kEps);
                const int n = __popc(bits);
                if (n == 1) { sum_one += l; ++n_one; }
                else if (n > 1) { sum_multi += l; ++n_multi; }
                else { sum_empty += l; ++n_empty; }
            }
            if (group) {
                if (id != cur) {
                    if (cur >= 0) {
                        for (int c = 0; c < C; ++c) {
                            const unsigned long long e = col[c * kThreads];
                            if (e != 0ull) { atomicMax(p.gmax + cur_base + c, e); col[c * kThreads] = 0ull; }
                        }
                    }
                    cur = id;
                    cur_base = ((long long)t.img * p.S + id) * C;
                }
                const unsigned long long low = (unsigned long long)(~(uint32_t)off);
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
                    if ((bits >> c) & 1u) {
                        const unsigned long long key = ((unsigned long long)__float_as_uint(v[c]) << 32) | low;
                        if (key > col[c * kThreads]) col[c * kThreads] = key;
                    }
                }
            }
        }
        if (cur >= 0) {
            for (int c = 0; c < C; ++c) {
                const unsigned long long e = col[c * kThreads];
                if (e != 0ull) atomicMax(p.gmax + cur_base + c, e);
            }
        }
    }

    if (p.do_choice) {
        __shared__ float s_sum[3];
        __shared__ int s_cnt[3];
        if (tid < 3) { s_sum[tid] = 0.f; s_cnt[tid] = 0; }
        __syncthreads();
        float s[3] = {sum_one, sum_multi, sum_empty};
        int n[3] = {n_one, n_multi, n_empty};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
                n[k] += __shfl_xor_sync(0xffffffffu, n[k], o);
            }
            if (lane == 0 && n[k] != 0) { atomicAdd(&s_sum[k], s[k]); atomicAdd(&s_cnt[k], n[k]); }
        }
        __syncthreads();
        if (tid < 3 && s_cnt[tid] != 0) {
            atomicAdd(p.acc + 2 * tid, (double)s_sum[tid]);
            atomicAdd(p.acc + 2 * tid + 1, (double)s_cnt[tid]);
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
    const int C = EXACT ? CMAX : p.C;
    const Tile t = my_tile(p);
    if (!t.in_range || t.x >= p.W) return;
    const float w_one = p.coef[0] * p.inv_temp, w_multi = p.coef[1] * p.inv_temp, w_group = p.coef[3] * p.inv_temp;
    const size_t P = (size_t)p.H * p.W;
    for (int y = t.y0; y < t.y1; ++y) {
        const size_t off = (size_t)y * p.W + t.x;
        const size_t pix = (size_t)t.img * P + off;
        float* gbase = p.grad + (size_t)t.img * C * P + off;
        int id = -1;
        uint32_t inf = 0u;
        if (__ldcs(p.mask + pix) != 0) {
            id = load_id<IdT>(p.ids, pix);
            if ((unsigned)id < (unsigned)p.S) inf = __ldg(p.info + (size_t)t.img * p.S + id); else id = -1;
        }
        const uint32_t bits = inf & ~kGroupBit;
        const bool group = p.do_group && (inf & kGroupBit) && bits != 0u;
        if (id < 0 || (!group && !(p.do_choice && bits != 0u))) {      // no gradient reaches this pixel
            for (int c = 0; c < C; ++c) __stcs(gbase + (size_t)c * P, 0.f);
            continue;
        }
        float v[CMAX];
        const float* base = p.logits + (size_t)t.img * C * P + off;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) v[c] = (EXACT || c < C) ? __ldcs(base + (size_t)c * P) : -INFINITY;
        softmax_inplace<CMAX, false>(v, p.temp, p.scale);
        float pos = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) pos += ((bits >> c) & 1u) ? v[c] : 0.f;
        float a = 0.f;
        if (p.do_choice) {
            const int n = __popc(bits);
            a = -(n == 1 ? w_one : w_multi) / (pos + kEps);
        }
        uint32_t abits = 0u;   // classes whose max-pooled probability comes from this pixel
        float q_sum = 0.f;
        if (group) {
            const uint32_t low = ~(uint32_t)off;
            const unsigned long long* row = p.gmax + ((long long)t.img * p.S + id) * C;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                if ((bits >> c) & 1u) {
                    const unsigned long long e = __ldg(row + c);
                    if ((uint32_t)e == low && (e >> 32) != 0ull) {
                        abits |= 1u << c;
                        q_sum += v[c] / (v[c] + kEps);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            const float pc = v[c];
            float g = a * pc * ((((bits >> c) & 1u) ? 1.f : 0.f) - pos);
            if (abits) g -= w_group * ((((abits >> c) & 1u) ? pc / (pc + kEps) : 0.f) - pc * q_sum);
            if (EXACT || c < C) __stcs(gbase + (size_t)c * P, g);
        }
    }
}

// ------------------------------------------------------------------------------------------ launch
template <int CMAX, bool EXACT, typename IdT>
cudaError_t launch_one(const LossParams& p, bool backward, bool accurate, cudaStream_t stream) {
    const long long tiles = (long long)p.n_img * ((p.W + 31) / 32) * ((p.H + kTileRows - 1) / kTileRows);
    const long long blocks = (tiles + kThreads / 32 - 1) / (kThreads / 32);
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    if (backward) {
        multihot_loss_bwd_kernel<CMAX, EXACT, IdT><<<(unsigned)blocks, kThreads, 0, stream>>>(p);
    } else {
        const size_t smem = p.do_group ? (size_t)p.C * kThreads * sizeof(unsigned long long) : 0;
        if (accurate) multihot_loss_fwd_kernel<CMAX, EXACT, IdT, true><<<(unsigned)blocks, kThreads, smem, stream>>>(p);
        else multihot_loss_fwd_kernel<CMAX, EXACT, IdT, false><<<(unsigned)blocks, kThreads, smem, stream>>>(p);
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

cudaError_t dispatch(const LossParams& p, int ids_dtype, bool backward, bool accurate, cudaStream_t stream) {
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
    p.temp = temperature; p.inv_temp = 1.f / temperature; p.scale = 1.4426950408889634f / temperature;
    p.do_choice = (flags & MAS_LOSS_CHOICE) ? 1 : 0; p.do_group = (flags & MAS_LOSS_GROUP) ? 1 : 0;
    p.acc = acc; p.gmax = reinterpret_cast<unsigned long long*>(group_max);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = dispatch(p, ids_dtype, false, (flags & MAS_LOSS_EXACT_SOFTMAX) != 0, st);
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
    p.temp = temperature; p.inv_temp = 1.f / temperature; p.scale = 1.4426950408889634f / temperature;
    p.do_choice = (flags & MAS_LOSS_CHOICE) ? 1 : 0; p.do_group = (flags & MAS_LOSS_GROUP) ? 1 : 0;
    p.gmax = const_cast<unsigned long long*>(reinterpret_cast<const unsigned long long*>(group_max));
    p.coef = coef; p.grad = grad_logits;
    cudaError_t e = dispatch(p, ids_dtype, true, false, (cudaStream_t)stream);
    if (e != cudaSuccess) return mas::cuda_fail(e, "multihot_loss_bwd_kernel launch");
    return 0;
}
