/*
 * mulactseg_b200 -- C ABI of the B200-native superpixel-segmented scoring engine.
 *
 * This is the drop-in boundary for the hot path of sehyun03/MulActSeg named in
 * BASELINE.json `north_star`.  The reference is pure Python: what it binds at
 * this level is the third-party `torch_scatter` operator set plus the torch ops
 * around it.  Each entry point below replaces one such operator chain; the
 * comment above it cites the reference call sites (paths relative to the
 * reference checkout).  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types.
 *  - `*_dev` entry points take DEVICE pointers and enqueue on `stream`
 *    (a `cudaStream_t` passed as void*; NULL = legacy default stream) without
 *    synchronising.  `*_host` entry points take HOST pointers, do their own
 *    host<->device copies and return after the result is in host memory.
 *  - return value: 0 on success, a negative MAS_E_* code for argument errors,
 *    a positive value = the cudaError_t that was raised.  `mas_last_error()`
 *    returns a thread-local human-readable message for the last failure.
 *  - there is NO CPU fallback: every compute entry point fails with a CUDA
 *    error when no sm_100 device is present.
 *  - images are NCHW-contiguous: logits[(img*C + c)*H*W + y*W + x].
 */
#ifndef MULACTSEG_B200_H
#define MULACTSEG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAS_ABI_VERSION 1

#define MAS_E_BADARG (-1)   /* null pointer / non-positive size / unsupported combination */
#define MAS_E_RANGE (-2)    /* value out of the supported range (e.g. channels > MAS_MAX_CLASSES) */
#define MAS_E_WORKSPACE (-3) /* workspace too small */

#define MAS_MAX_CLASSES 32

/* element type of the logits / feature tensors */
#define MAS_F32 0
#define MAS_BF16 1
/* element type of superpixel id maps (the reference hands int64 maps over: ext_transforms.py:389,406) */
#define MAS_I32 0
#define MAS_I64 1
#define MAS_U8 2      /* label maps only (mas_miou_counts_dev) */

int mas_abi_version(void);
const char* mas_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench bookkeeping) */
int64_t mas_kernel_launches(void);

/* ------------------------------------------------------------------ acquisition
 *
 * mas_bvsb_segment_stats_dev -- ONE pass over the logits of `n_img` images.
 * Replaces, per batch (reference file:line):
 *   F.softmax(preds/T) + torch.topk(prob,2) + ratio + 1e-8   active_selection/my_bvsb.py:19-27
 *   torch_scatter.scatter(bvsb, spx, reduce='mean')           my_bvsb.py:73 (and the five sibling selectors)
 *   F.one_hot(top1) + scatter(..., reduce='sum')              my_bvsb_banignore.py:44-45, my_bvsb_predclsbal_pwr*.py:68-69
 *   torch.softmax(preds/T).mean(dim=(0,2,3))  (pass 1)        my_bvsb_predclsbal_pwr*.py:41-43
 *
 * For every pixel: (l1,top1),(l2) = best / second-best LOGIT (first index wins a tie),
 * bvsb = exp((l2-l1)/T) + 1e-8  (== p2/p1 of the softmax), and
 *   cls_sum[img][s][c] += bvsb   where s = ids[pixel], c = top1       (float32, atomics)
 *   cls_cnt[img][s][c] += 1                                           (int32, exact)
 *   prob_sum[img][c]   += softmax(l/T)[c]   for every c               (float64; skipped if NULL)
 * Pixels whose id is outside [0, nseg) are ignored.  Outputs are ACCUMULATED
 * into: the caller zeroes them (cudaMemsetAsync) before the first call.
 * Region sums / counts / histograms / class means all follow from these tables
 * (see mas_region_scores_dev), so the six selectors share this single pass.
 * `image_stride` = elements between consecutive images of `logits` (0 = channels*height*width); a
 * larger stride lets a channel-sliced view such as preds[:, :-1] (my_bvsb.py:65-66) be read in place.
 */
int mas_bvsb_segment_stats_dev(const void* logits, int logits_dtype, int64_t image_stride, const int32_t* ids,
                               int n_img, int channels, int height, int width, int nseg,
                               float temperature,
                               float* cls_sum, int32_t* cls_cnt, double* prob_sum,
                               void* stream);

/* mas_bvsb_segment_stats_multi_dev -- the same pass over up to MAS_MAX_SEGMENTS batches that live in different
 * allocations (consecutive DataLoader batches) in ONE launch: segment g holds n_img_per_segment[g] images at logits[g]
 * / ids[g] (image stride image_strides[g], 0 = dense; image_strides may be NULL), and the segments fill consecutive image
 * rows of cls_sum / cls_cnt / prob_sum (sum of n_img_per_segment rows).  Small batches (VOC-sized images) are
 * launch-latency-bound one by one; grouped they stream like one big batch.
 */
#define MAS_MAX_SEGMENTS 32
int mas_bvsb_segment_stats_multi_dev(int n_segments, const void* const* logits, int logits_dtype, const int64_t* image_strides,
                                     const int32_t* const* ids, const int* n_img_per_segment, int channels, int height,
                                     int width, int nseg, float temperature, float* cls_sum, int32_t* cls_cnt,
                                     double* prob_sum, void* stream);

/* mas_bvsb_segment_stats_lowres_dev -- the same pass fed with the network head's LOW-RESOLUTION logits
 * (n_img, channels, height_in, width_in): every full-resolution logit is produced on the fly exactly as the model's final
 * F.interpolate(x, size=(height, width), mode='bilinear', align_corners=False) would (models/segmentation/utils.py:28-34,
 * deeplabv3.py:115-129; source index = (in/out) * (dst + 0.5) - 0.5 clamped at 0, taps i0 and min(i0 + 1, in - 1)), so
 * the 16x larger up-sampled tensor is never written nor read.  Opt-in: it changes what the
 * caller hands over (the head's output instead of net(images)); results equal mas_bvsb_segment_stats_dev on the
 * interpolated tensor up to fp32 rounding of the interpolation.  ids stay full resolution (height, width).
 */
int mas_bvsb_segment_stats_lowres_dev(const void* logits, int logits_dtype, int64_t image_stride, int height_in, int width_in,
                                      const int32_t* ids, int n_img, int channels, int height, int width, int nseg,
                                      float temperature, float* cls_sum, int32_t* cls_cnt, double* prob_sum, void* stream);

/* mas_class_weights_dev -- w_c = (coeff * pbar_c + 1)^-2, pbar = the reference's `cumulated_pred_prob / len(loader)`
 * (active_selection/my_bvsb_predclsbal_pwr.py:33-47): per reference batch of `ref_batch` images (the last may be short)
 * the mean probability of class c = sum of prob_sum rows / (images * pixels_per_image), accumulated over the batches in
 * loader order in fp32.  prob_sum (n_img, channels) f64 from mas_bvsb_segment_stats_dev (all ranks' rows in pool order),
 * weight: channels DEVICE floats -- the `class_weight` argument of mas_region_scores_dev.
 */
int mas_class_weights_dev(const double* prob_sum, int64_t n_img, int channels, int64_t pixels_per_image, int ref_batch,
                          float coeff, float* weight, void* stream);

/* mas_region_scores_dev -- per-region epilogue over the tables above.
 *   n[r]        = sum_c cls_cnt[r][c]
 *   score[r]    = (sum_c w[c] * cls_sum[r][c]) / max(n[r], 1)     (w == NULL -> all ones)
 *   dominant[r] = first arg-max_c cls_cnt[r][c]                    (0 for an empty region)
 * Replaces the divide of scatter(reduce='mean') (my_bvsb.py:73), the cls_weight[top1] weighting
 * (my_bvsb_predclsbal_pwr.py:59-65, re-associated) and hist.argmax(dim=1) (my_bvsb_banignore.py:60).
 * `n_regions` = n_img * nseg.  `npix` and `dominant` may be NULL.
 */
int mas_region_scores_dev(const float* cls_sum, const int32_t* cls_cnt, const float* class_weight,
                          int64_t n_regions, int channels,
                          float* score, int32_t* npix, int32_t* dominant, void* stream);

/* mas_minmax_nonzero_dev -- out[0] = min over entries != 0, out[1] = max over all entries
 * (my_bvsb.py:80-81).  out[0] = +inf / out[1] = -inf when there is nothing to reduce. */
int mas_minmax_nonzero_dev(const float* values, int64_t n, float* out2, void* stream);

/* mas_dominant_hist_dev -- hist[c] += #regions with dominant == c (int64, `channels` bins;
 * my_bvsb_clsbal_v2.py:64-66).  Accumulates: zero `hist` first. */
int mas_dominant_hist_dev(const int32_t* dominant, int64_t n_regions, int channels, int64_t* hist, void* stream);

/* mas_finalize_scores_dev -- in place, in the reference's order of operations:
 *   if (minmax)  u = (u - minmax[0]) / (minmax[1] - minmax[0])   my_bvsb.py:80-81
 *                (minmax = DEVICE pointer to {min over non-zero, max}; NULL = no normalisation)
 *   if (ban_class >= 0 && dominant == ban_class) u = 0        my_bvsb_banignore.py:59-61
 *   if (region_weight) u = region_weight[dominant] * u        my_bvsb_clsbal_v2.py:67-70
 */
int mas_finalize_scores_dev(float* score, const int32_t* dominant, int64_t n_regions,
                            const float* minmax, int ban_class,
                            const float* region_weight, void* stream);

/* ------------------------------------------------------------------ top-k region selection
 *
 * Replaces `sorted(scores, reverse=True)` over (score, path, id) tuples
 * (active_selection/base.py:37) restricted to the prefix that
 * RegionActiveDataset.expand_training_set can consume
 * (dataloader/region_active_dataset.py:31-73).
 *
 * mas_region_keys_dev builds one 64-bit key per region:
 *   key = ordered_bits(score) << 32 | (image_rank[img] * nseg + id)
 * where image_rank is the rank of the image's joined path string in ascending
 * string order, so that descending key order == the reference's descending
 * tuple order including ties.  Regions with in_pool == 0 get key 0 (never
 * selected).  -0.0 is canonicalised to +0.0 (Python compares them equal).
 *
 * mas_topk_u64_dev writes the `k` largest non-zero keys of `keys[0..n)` to
 * `out` (UNSORTED) and the number written to *out_count (min(k, #non-zero)).
 * `workspace` needs mas_topk_workspace_bytes() bytes.  Keys must be distinct
 * (they are, by construction) for the count to be exact.
 */
int mas_region_keys_dev(const float* score, const uint8_t* in_pool, const int32_t* image_rank,
                        int64_t n_img, int nseg, uint64_t* keys, void* stream);
size_t mas_topk_workspace_bytes(void);
int mas_topk_u64_dev(const uint64_t* keys, int64_t n, int64_t k, uint64_t* out, int32_t* out_count,
                     void* workspace, size_t workspace_bytes, void* stream);
/* mas_sort_desc_u64_dev -- in-place descending sort of n keys (n <= 2^22).  The buffer must have
 * room for mas_sort_capacity(n) keys (n rounded up to a power of two, >= 2048): the tail is used as
 * padding and holds zeros afterwards. */
int64_t mas_sort_capacity(int64_t n);
int mas_sort_desc_u64_dev(uint64_t* keys, int64_t n, void* stream);

/* mas_topk_sorted_u64_dev -- the same result as mas_topk_u64_dev followed by mas_sort_desc_u64_dev in 3 + sort launches:
 * two histogram passes (key bits 63..52, then 51..40) locate the 24-bit prefix of the k-th largest key, one compaction
 * pass keeps every key from that prefix upwards, the sort orders them.  `out` has `capacity` slots (a sort capacity >=
 * mas_sort_capacity(k)) and is fully overwritten: out[0 .. count) = the count = min(k, #non-zero keys) largest keys in
 * descending order.  *out_count = -1 when the candidates did not fit `capacity` (massively tied scores): call the
 * exact two-step path instead.  workspace: mas_topk_workspace_bytes().
 */
int mas_topk_sorted_u64_dev(const uint64_t* keys, int64_t n, int64_t k, uint64_t* out, int64_t capacity, int32_t* out_count,
                            void* workspace, size_t workspace_bytes, void* stream);

/* mas_topk_candidates_u64_dev -- the first half of mas_topk_sorted_u64_dev (histograms + compaction, no sort): a SUPERSET of
 * the k largest keys, unordered, in out[0 .. count); remaining slots are 0 ("no key").  *out_count = count (>= min(k,
 * #non-zero keys), <= capacity) or -1 on overflow.  Used per GPU before the all-gather of the multi-GPU merge: the merged
 * candidates go through mas_topk_sorted_u64_dev once.
 */
int mas_topk_candidates_u64_dev(const uint64_t* keys, int64_t n, int64_t k, uint64_t* out, int64_t capacity,
                                int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream);

/* mas_topk_candidates_msg_u64_dev -- mas_topk_candidates_u64_dev writing the MESSAGE one rank contributes to the multi-GPU
 * merge (north_star: "global top-k merges per-GPU candidate lists with an NCCL all-gather"; replaces the host-side
 * sorted() over the whole pool, active_selection/base.py:37): msg has capacity + 1 slots, msg[0 .. capacity) = the
 * candidates (0 = "no key"), msg[capacity] = the count as int64 (-1 on overflow).  One all_gather_into_tensor of these
 * messages is the only collective of the merge.
 */
int mas_topk_candidates_msg_u64_dev(const uint64_t* keys, int64_t n, int64_t k, uint64_t* msg, int64_t capacity,
                                    int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream);

/* mas_merge_counts_u64_dev -- on the gathered messages (world x stride slots, stride = capacity + 1): *worst = the
 * smallest per-rank count (-1 if any rank overflowed) and the count slots are cleared, so the whole buffer can go through
 * mas_topk_sorted_u64_dev as plain keys.
 */
int mas_merge_counts_u64_dev(uint64_t* gathered, int world, int64_t stride, int32_t* worst, void* stream);

/* mas_prefix_cut_dev -- how much of the ranked list RegionActiveDataset.expand_training_set consumes
 * (dataloader/region_active_dataset.py:56-66): walk sorted_keys[0 .. *count), add cost_by_tie[key & 0xffffffff] (the label
 * cost of the region: 1, or its multi-hot class count with --fair_counting --or_labeling) and stop AFTER the pick that
 * makes the running cost exceed `budget` (strict '>').  *n_take = that prefix length (the whole list if the budget is
 * never exceeded; -1 if *count is -1).  cost_by_tie: n_cost uint8, indexed like the low key word (image rank * nseg + id).
 */
int mas_prefix_cut_dev(const uint64_t* sorted_keys, const int32_t* count, const uint8_t* cost_by_tie, int64_t n_cost,
                       int64_t budget, int32_t* n_take, void* stream);

/* ------------------------------------------------------------------ host-buffer entries (end-to-end)
 *
 * mas_acquisition_host -- the whole scoring pass of one selector with HOST buffers.  Streams
 * `n_img` images of logits + ids to the device in chunks of `chunk_img` images (double-buffered:
 * the copy of chunk i+1 overlaps the kernel of chunk i), accumulates the tables on the device,
 * then runs the selector epilogue and copies back
 *   score    (n_img*nseg f32)
 *   dominant (n_img*nseg i32, may be NULL)
 *   prob_sum (n_img*channels f64, may be NULL; forced on for MAS_WEIGHT_PREDCLSBAL).
 * `weighting`: MAS_WEIGHT_NONE, or MAS_WEIGHT_PREDCLSBAL = w_c = (coeff * pbar_c + 1)^-2 with pbar the
 * mean over reference batches (`ref_batch` images each, the last one short) of the per-batch mean
 * softmax probability (my_bvsb_predclsbal_pwr.py:36-47).
 * `normalise` != 0: (u - min_nonzero) / (max - min_nonzero) (my_bvsb.py:80-81).
 * `ban_class` >= 0: zero regions whose dominant arg-max class is ban_class.
 * `clsbal` != 0: multiply by exp(-freq[dominant]) (my_bvsb_clsbal_v2.py:64-70).
 * Host buffers should be pinned (cudaHostAlloc / cudaHostRegister) for the copies to overlap.
 */
#define MAS_WEIGHT_NONE 0
#define MAS_WEIGHT_PREDCLSBAL 1
int mas_acquisition_host(const void* logits, int logits_dtype, const int32_t* ids,
                         int n_img, int channels, int height, int width, int nseg,
                         float temperature, int weighting, float coeff, int ref_batch,
                         int normalise, int ban_class, int clsbal, int chunk_img,
                         float* score, int32_t* dominant, double* prob_sum);

/* mas_select_topk_host -- host scores -> the k best (sorted descending) region keys, see
 * mas_region_keys_dev for the key layout.  out_keys needs k slots; *out_count <= k. */
int mas_select_topk_host(const float* score, const uint8_t* in_pool, const int32_t* image_rank,
                         int64_t n_img, int nseg, int64_t k, uint64_t* out_keys, int32_t* out_count);

/* ------------------------------------------------------------------ stage-1 losses (forward + backward)
 *
 * Replaces, in ONE pass over the logits of the masked pixels (reference file:line):
 *   F.softmax(inputs/T) + permute + per-image boolean compaction        utils/loss.py:99-118, :550-568
 *   torch_scatter.scatter(valid_output, valid_spx, reduce='max')        utils/loss.py:122,
 *        trainer/active_joint_multi_predignore.py:109, ..._mclossablation2.py:60          (group / MIL loss)
 *   targets[spx] gather, (prob * trg).sum(1), -log(. + 1e-8)            utils/loss.py:572-581,
 *        trainer/active_joint_multi_predignore_lossdecomp.py:52-70, active_joint_multi_lossdecomp.py:37-56
 *
 * mas_multihot_info_dev -- per (image, superpixel) candidate word from the multi-hot targets
 * (n_regions x target_channels, uint8):
 *   bits 0..channels-1 : targets[r][c] != 0      (channels < target_channels reproduces targets[..., :-1])
 *   bit 31             : the region takes part in the group loss: MAS_GROUP_ALL always,
 *                        MAS_GROUP_ONLYMULTI iff sum_c targets[r][c] over ALL target_channels > 1
 *                        (is_trg_multi, ..._mclossablation2.py:35)
 */
#define MAS_MAX_LOSS_CLASSES 31
#define MAS_GROUP_ALL 0
#define MAS_GROUP_ONLYMULTI 1
int mas_multihot_info_dev(const uint8_t* targets, int64_t n_regions, int target_channels, int channels, int group_mode,
                          uint32_t* info, void* stream);

/* mas_multihot_loss_fwd_dev -- logits (n_img, channels, H, W) f32 contiguous, ids (n_img, H, W) int32|int64,
 * mask (n_img, H, W) uint8 (non-zero = selected pixel), info from mas_multihot_info_dev.  A pixel counts when
 * its mask is set and 0 <= id < nseg (the crop-padding id == nseg never counts).  With P = softmax(x / T),
 * row = candidate set of the pixel's superpixel, pos = sum_{c in row} P_c, l = -log(pos + 1e-8):
 *   MAS_LOSS_CHOICE: acc[0] += sum l, acc[1] += #pixels with |row| == 1      (one-hot bucket)
 *                    acc[2], acc[3]: |row| > 1 (multi-hot bucket);  acc[4], acc[5]: |row| == 0
 *   MAS_LOSS_GROUP : group_max[(img*nseg + s)*channels + c] = max over the masked pixels of group regions of
 *                    (bits(P_c) << 32 | ~pixel) for c in row  -- value and first arg-max pixel of the
 *                    max-pool; then acc[6] += sum -log(M + 1e-8) over entries with M > 0, acc[7] += their number.
 * `acc` (8 doubles) is ACCUMULATED into and `group_max` must be zero on entry (cudaMemsetAsync both).
 * The reference's losses follow as  sum / (1 + count)  per bucket (counters start at 1).
 */
#define MAS_LOSS_CHOICE 1
#define MAS_LOSS_GROUP 2
/* softmax in the reference's own operation order (x / T, max-subtract, expf, divide by the sum) instead of the fast
 * ex2 / reciprocal form (<= 4 ulp apart): for callers that need the arg-max PIXEL of the max-pool to agree with torch
 * on near-ties (the stage-2 prototype labeller).  Forward only. */
#define MAS_LOSS_EXACT_SOFTMAX 4
int mas_multihot_loss_fwd_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask, const uint32_t* info,
                              int n_img, int channels, int height, int width, int nseg, float temperature, int flags,
                              double* acc, uint64_t* group_max, void* stream);

/* mas_multihot_loss_bwd_dev -- dense gradient of  coef[0]*acc[0] + coef[1]*acc[2] + coef[2]*acc[4] + coef[3]*acc[6]
 * with respect to the logits (coef = 4 DEVICE floats; the caller folds the 1/(1+count) factors and the loss
 * coefficients in).  The group term reaches only the arg-max pixel of each counted (superpixel, class), like the
 * autograd of torch_scatter's max.  grad_logits (n_img, channels, H, W) f32 is fully written (zeros elsewhere).
 */
int mas_multihot_loss_bwd_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask, const uint32_t* info,
                              const uint64_t* group_max, const float* coef, int n_img, int channels, int height, int width,
                              int nseg, float temperature, int flags, float* grad_logits, void* stream);

/* ACTIVE-TILE LIST of a mask.  In training only the labelled superpixels are selected (spmasks, 2-10 % of the pixels
 * after a few acquisition rounds; trainer/active_joint_multi_predignore_lossdecomp.py:101-104 hands the same masks to both
 * criteria).  mas_multihot_tiles_dev scans the mask once (tiles of 32 px x 8 rows) and writes a bitmap of the tiles that
 * hold a selected pixel plus prefix sums into `tiles` (mas_multihot_tiles_workspace_bytes() bytes, 16-byte aligned).  The
 * *_tiles_dev twins of the two passes take that list:
 *   forward : only the active tiles are visited, dealt round-robin BY RANK to the warps (balanced whatever the layout of
 *             the labelled regions; deterministic -- no atomic queue);
 *   backward: a sparsely selected batch gets its dense gradient zeroed by one linear sweep and only the active tiles
 *             computed on top;
 *   both    : when at least 40 % of the tiles are active and the shape allows (W % 16 == 0, 16-byte aligned bases, fast
 *             softmax) the batch is taken by the DENSE kernels instead (csrc/losses_dense.cu: TMA-staged 64-pixel strip
 *             rows, the gradient leaves by TMA store).  The choice is made ON THE DEVICE from the active-tile count: both
 *             kernel sets are launched and one returns at once, so no call ever synchronises.  Other shapes: below 35 %
 *             the list walk, otherwise the tile walk of the plain entry points.
 * Results are those of the plain entry points (tiles == NULL falls back to them) up to the order of the fp32 bucket sums.
 */
size_t mas_multihot_tiles_workspace_bytes(int n_img, int height, int width);
int mas_multihot_tiles_dev(const uint8_t* mask, int n_img, int height, int width, void* tiles, size_t tiles_bytes, void* stream);
int mas_multihot_loss_fwd_tiles_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask, const uint32_t* info,
                                    const void* tiles, int n_img, int channels, int height, int width, int nseg, float temperature,
                                    int flags, double* acc, uint64_t* group_max, void* stream);
int mas_multihot_loss_bwd_tiles_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask, const uint32_t* info,
                                    const void* tiles, const uint64_t* group_max, const float* coef, int n_img, int channels,
                                    int height, int width, int nseg, float temperature, int flags, float* grad_logits, void* stream);

/* mas_multihot_loss_finish_dev -- the reference's normalisations (counters start at 1; fp32 sum / fp32 count, like
 * `loss / num_valid`) of the bucket sums in `acc`, in one launch.  losses = 6 DEVICE floats:
 *   [0] one-hot CE                  acc[0] / (1 + acc[1])                      ..._predignore_lossdecomp.py:69, ..._lossdecomp.py:71
 *   [1] multi-hot (row-sum > 1)     acc[2] / (1 + acc[3])                      trainer/active_joint_multi_lossdecomp.py:67-72
 *   [2] multi-hot (not one-hot)     (acc[2] + acc[4]) / (1 + acc[3] + acc[5])  trainer/active_joint_multi_predignore_lossdecomp.py:65-70
 *   [3] multi-choice CE             (acc[0] + acc[2]) / (1 + acc[1] + acc[3])  utils/loss.py:572-588 (empty rows dropped)
 *   [4] group / MIL loss            acc[6] / (1 + acc[7])                      utils/loss.py:131-141
 *   [5] 0
 * mas_multihot_loss_coef_dev -- its transpose for the backward pass: coef[k] = d(sum_j grad_losses[j] * losses[j]) / d acc[2k]
 * (4 DEVICE floats, the `coef` argument of mas_multihot_loss_bwd_dev).
 */
int mas_multihot_loss_finish_dev(const double* acc, float* losses, void* stream);
int mas_multihot_loss_coef_dev(const double* acc, const float* grad_losses, float* coef, void* stream);

/* mas_stage1_loss_fwd_dev / mas_stage1_loss_bwd_dev -- the trainer's whole loss step
 * (trainer/active_joint_multi_predignore_lossdecomp.py:101-107: group loss + decomposed partial-label loss on the same
 * tensors) in ONE call per direction: candidate words (mas_multihot_info_dev), active-tile list, fused forward, group
 * reduction and the six normalised losses of mas_multihot_loss_finish_dev; then coefficients, zero sweep and the fused
 * backward.  Same launches as the separate entry points, an eighth of the host work (at a few percent of labelled pixels
 * the host side of the step, not its kernels, is the bottleneck).  `workspace`: mas_stage1_workspace_bytes() bytes,
 * 16-byte aligned, kept by the caller from forward to backward; bytes [0,64) = acc (8 doubles), [64,88) = the six losses.
 * targets (n_img, nseg, target_channels) uint8, group_mode as mas_multihot_info_dev, flags as mas_multihot_loss_fwd_dev.
 * grad_losses: HOST array of six DEVICE pointers to the incoming gradients of the six losses (NULL = not used).
 */
size_t mas_stage1_workspace_bytes(int n_img, int channels, int height, int width, int nseg);
int mas_stage1_loss_fwd_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask, const uint8_t* targets,
                            int target_channels, int n_img, int channels, int height, int width, int nseg, float temperature,
                            int group_mode, int flags, void* workspace, size_t workspace_bytes, void* stream);
int mas_stage1_loss_bwd_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask, int n_img, int channels,
                            int height, int width, int nseg, float temperature, int flags, const void* workspace,
                            const float* const* grad_losses, float* grad_logits, void* stream);

/* ------------------------------------------------------------------ operator level: torch_scatter
 *
 * The reference delegates every segmented reduction to torch_scatter 2.0.9 (actsegmul.yml:99; call sites: SURVEY.md 2b).
 * The twelve hot-path plugins never materialise those operands (fused passes above); these two entries serve the
 * reference's OTHER call sites through mulactseg_b200/torch_scatter_compat.py (`scatter`, `scatter_max`, same
 * signatures).  `src` is viewed as (outer, n, inner) with the reduced dimension in the middle; index is int64, either of
 * src's shape (index_has_inner = 1) or (outer, n) shared by the `inner` trailing elements (the broadcast torch_scatter
 * applies to e.g. a (B, HW) index against a (B, HW, C') one-hot, my_bvsb_banignore.py:44-45).  Indices outside
 * [0, dim_size) are ignored.
 *   mas_scatter_sum_dev: out (outer, dim_size, inner) += src   (float32 or int64; ACCUMULATES: zero `out` first).
 *   mas_scatter_max_dev: out = per-segment maximum, 0 for a segment nobody writes; arg = FIRST position along the reduced
 *     dimension attaining it (torch_scatter's CPU tie rule), pre-filled by the caller with n (the "empty" marker,
 *     utils/loss.py:202-204).  key_workspace: outer * dim_size * inner uint32.
 */
#define MAS_SCATTER_I64 3
int mas_scatter_sum_dev(const void* src, int src_dtype, const int64_t* index, int64_t outer, int64_t n, int64_t inner,
                        int index_has_inner, int64_t dim_size, void* out, void* stream);
int mas_scatter_max_dev(const float* src, const int64_t* index, int64_t outer, int64_t n, int64_t inner, int index_has_inner,
                        int64_t dim_size, float* out, int64_t* arg, uint32_t* key_workspace, void* stream);

/* ------------------------------------------------------------------ stage-2 pseudo-labellers
 *
 * mas_candidate_argmax_dev -- trainer/eval_within_multihot.py:93-146 top_pseudo_label_generation:
 *   labels[pixel] = first arg-max over c of logit[c] * targets[spx[pixel]][c] on selected pixels (mask set,
 *   0 <= id < nseg), 255 elsewhere.  Non-candidate classes contribute 0 (not -inf), exactly like the reference's
 *   multiply-then-max.  logits (n_img, channels, H, W) f32, info from mas_multihot_info_dev, labels (n_img, H, W) u8.
 */
int mas_candidate_argmax_dev(const float* logits, const void* ids, int ids_dtype, const uint8_t* mask, const uint32_t* info,
                             int n_img, int channels, int height, int width, int nseg, uint8_t* labels, void* stream);

/* mas_proto_labeller_dev -- ONE image of ActiveTrainer.pseudo_label_generation
 * (trainer/eval_save_cosplbl_prop.py:121-314 with only_multihot = 1; the shipped
 * trainer/eval_save_cosplbl_prop_includeonehot.py:121-316 with only_multihot = 0):
 *   selected pixel  = mask set, 0 <= id < nseg (and, only_multihot: its superpixel has > 1 candidate class);
 *   prototype(s,c)  = feats[:, first arg-max over the selected pixels of s of softmax(logits)[c]]  for c in targets[s];
 *   selected pixels : label = class of the prototype of their OWN superpixel with the largest inner product (first on ties);
 *   threshold(s,c)  = lower median (MAS_THRESHOLD_MEDIAN, torch.median) or minimum (MAS_THRESHOLD_MIN) of those largest
 *                     inner products over the pixels assigned to prototype (s,c); 1.0 if none;
 *   other pixels    : among the selected superpixels s whose 3x3-dilated mask reaches the pixel's superpixel (s itself
 *                     included), the LARGEST s with  threshold(s,k) < <prototype(s,k), feat[pixel]>  for some k labels the
 *                     pixel with the class of its best prototype (== the reference's ascending overwrite order, :276-305);
 *   everything else : 255.
 * feats (feat_channels, H, W) f32, logits (channels, H, W) f32, targets (nseg, target_channels) u8 (the first `channels`
 * columns are the candidate sets), labels (H, W) u8.  status[0] (device int32) = number of selected pixels whose
 * superpixel has no candidate class -- the reference raises on such input (:226); callers should too.
 * `workspace`: 256-byte aligned device buffer of mas_proto_labeller_workspace_bytes() bytes.
 */
#define MAS_THRESHOLD_MEDIAN 0
#define MAS_THRESHOLD_MIN 1
size_t mas_proto_labeller_workspace_bytes(int feat_channels, int channels, int height, int width, int nseg);
int mas_proto_labeller_dev(const float* feats, int feat_channels, const float* logits, int channels,
                           const uint8_t* targets, int target_channels, const uint8_t* mask, const void* ids, int ids_dtype,
                           int height, int width, int nseg, int only_multihot, int threshold_mode,
                           uint8_t* labels, int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* mas_proto_labeller_src_dev -- mas_proto_labeller_dev with a described FEATURE SOURCE (north_star: "stream bf16/fp32 ...
 * features"):
 *   feat_dtype MAS_F32 | MAS_BF16 (bf16 features are widened to fp32; every similarity is still an fp32 FMA chain);
 *   (feat_height, feat_width) == (height, width): full-resolution features as in the reference
 *   (trainer/eval_save_cosplbl_prop.py:55-60, feats = the x4 up-sampled head features);
 *   smaller: the network head's LOW-RESOLUTION (F, feat_height, feat_width) map -- every feature value is evaluated on
 *   the fly as F.interpolate(feat, size=(height, width), mode='bilinear', align_corners=False) would produce it
 *   (models/segmentation/utils.py:28-34, deeplabv3.py:122), so the 2.1 GB up-sampled tensor of a Cityscapes image is
 *   never written nor read.  Labels equal mas_proto_labeller_dev on the interpolated tensor wherever similarities are
 *   not within fp32 rounding of each other / of a threshold.  Everything else as mas_proto_labeller_dev (logits stay
 *   full resolution: 8 % of the bytes).
 */
int mas_proto_labeller_src_dev(const void* feats, int feat_dtype, int feat_channels, int feat_height, int feat_width,
                               const float* logits, int channels, const uint8_t* targets, int target_channels,
                               const uint8_t* mask, const void* ids, int ids_dtype, int height, int width, int nseg,
                               int only_multihot, int threshold_mode, uint8_t* labels, int32_t* status, void* workspace,
                               size_t workspace_bytes, void* stream);

/* mas_proto_labeller_batch_dev -- mas_proto_labeller_src_dev over a loader batch, as the reference's loop calls
 * pseudo_label_generation(labels, feats, inputs, targets, spmasks, superpixels) with batched tensors
 * (trainer/eval_save_cosplbl_prop_includeonehot.py:121-130): feats (n_img, F, fh, fw), logits (n_img, C, H, W), targets
 * (n_img, nseg, Ct), mask / ids / labels (n_img, H, W), status (n_img) int32 ZEROED by the caller, all contiguous.
 * Image i runs on lane_streams[i % n_lanes] (n_lanes <= 16 caller-owned streams) with the workspace slice
 * [lane * workspace_bytes_per_lane, ...) -- workspace_bytes_per_lane >= mas_proto_labeller_workspace_bytes(), a multiple
 * of 256 -- forked from and joined to `stream`; n_lanes <= 1 runs the images one after the other on `stream`.  Same
 * labels as n_img calls of mas_proto_labeller_src_dev.
 */
int mas_proto_labeller_batch_dev(const void* feats, int feat_dtype, int feat_channels, int feat_height, int feat_width,
                                 const float* logits, int channels, const uint8_t* targets, int target_channels,
                                 const uint8_t* mask, const void* ids, int ids_dtype, int n_img, int height, int width, int nseg,
                                 int only_multihot, int threshold_mode, uint8_t* labels, int32_t* status, void* workspace,
                                 size_t workspace_bytes_per_lane, void* const* lane_streams, int n_lanes, void* stream);

/* ------------------------------------------------------------------ offline multi-hot label generation
 *
 * mas_multihot_labels_dev -- ONE image of RegionCityscapesTensor.__getitem__ (dataloader/region_cityscapes_tensor.py:33-84,
 * driven by tools/label_assignment_tensor.py:50-67): per-superpixel class histogram of the ground-truth train-id map.
 *   ids (H, W) int32|int64, target (H, W) uint8 train ids (255 = ignore), keep (nseg) uint8 = 1 for the ids listed in
 *   the region dict ("preserving_labels").
 *   multi_hot[s][c] = 1 iff class c occurs in superpixel s, multi_hot[s][num_classes] = 1 iff label 255 occurs;
 *   size[s] = number of pixels counted; rows of ids with keep == 0 are zero with size -1.
 *   trim_kernel_size k > 0 (--trim_multihot_boundary): pixels within the k x k dilation of the superpixel boundaries
 *   (skimage find_boundaries mode='thick' + binary_dilation(ones(k,k))) are left out, unless that empties the superpixel.
 * multi_hot (nseg, num_classes + 1) uint8, size (nseg) int32; workspace: mas_multihot_labels_workspace_bytes() bytes.
 */
size_t mas_multihot_labels_workspace_bytes(int nseg, int num_classes);
int mas_multihot_labels_dev(const void* ids, int ids_dtype, const uint8_t* target, const uint8_t* keep,
                            int height, int width, int nseg, int num_classes, int trim_kernel_size,
                            uint8_t* multi_hot, int32_t* size, void* workspace, size_t workspace_bytes, void* stream);

/* mas_dominant_labels_dev -- ONE image of RegionCityscapesDominantAll.__getitem__ (dataloader/region_dataset.py:201-240;
 * tools/label_assignment_dominant.py): inside every superpixel listed in the region dict (keep == 1) the pixels that are
 * not 'ignore' take the most frequent train id among them (the smallest id on a tie: np.unique order + argmax); label 255
 * and the pixels of other superpixels are copied.  ids (H, W) int32|int64, target / out (H, W) uint8.
 */
size_t mas_dominant_labels_workspace_bytes(int nseg, int num_classes);
int mas_dominant_labels_dev(const void* ids, int ids_dtype, const uint8_t* target, const uint8_t* keep, int height, int width,
                            int nseg, int num_classes, uint8_t* out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ evaluation counters
 *
 * mas_miou_counts_dev -- utils/miou.py:23-54 (MeanIoU._after_step / _after_step_within_predregion) in one pass:
 *   a pixel is kept when target != ignore_label (MAS_MIOU_BY_TARGET) or output != ignore_label (MAS_MIOU_BY_OUTPUT);
 *   counts[c] += kept pixels with target == c                          (total_seen)
 *   counts[num_classes + c]   += ... with target == c == output        (total_correct)
 *   counts[2*num_classes + c] += kept pixels with output == c          (total_positive)
 * outputs / targets: n labels of the same dtype (MAS_I64 | MAS_I32 | MAS_U8); counts: 3 * num_classes uint64, ACCUMULATED.
 */
#define MAS_MIOU_BY_TARGET 0
#define MAS_MIOU_BY_OUTPUT 1
int mas_miou_counts_dev(const void* outputs, const void* targets, int labels_dtype, int64_t n, int num_classes,
                        int64_t ignore_label, int mode, uint64_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MULACTSEG_B200_H */
