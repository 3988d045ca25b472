"""Tensor-level wrappers over the C ABI: torch only provides device memory and the stream.

Every function validates device / dtype / contiguity (``RuntimeError`` like the
operators it replaces), passes raw pointers plus the current CUDA stream to the
library and never synchronises.  No function here has a CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor) -> int:
    """Raw handle of torch's current stream on the tensor's device (no Stream object is built: this sits on the
    launch path of every kernel)."""
    return torch._C._cuda_getCurrentRawStream(t.device.index)


class _on:
    """``with _on(t):`` makes the tensor's device current for the C-ABI call; free when it already is."""
    __slots__ = ("guard",)

    def __init__(self, t: torch.Tensor):
        self.guard = None if t.device.index == torch.cuda.current_device() else torch.cuda.device(t.device)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *exc):
        if self.guard is not None:
            self.guard.__exit__(*exc)
        return False


def _want(t: torch.Tensor, name: str, dtype=None, dims=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    if dtype is not None and t.dtype not in (dtype if isinstance(dtype, tuple) else (dtype,)):
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if dims is not None and t.dim() != dims:
        raise RuntimeError(f"{name}: expected {dims} dims, got shape {tuple(t.shape)}")


def bvsb_segment_stats(logits: torch.Tensor, spx: torch.Tensor, nseg: int, temperature: float,
                       cls_sum: torch.Tensor, cls_cnt: torch.Tensor, prob_sum: Optional[torch.Tensor]) -> None:
    """Accumulate per-(image, superpixel, class) bvsb sums / counts (+ per-image softmax sums).

    logits (B,C,H,W) f32|bf16, spx (B,H,W) i32, cls_sum/cls_cnt (B,nseg,C) f32/i32, prob_sum (B,C) f64 or None.
    Outputs are accumulated into (zero them first).  See ``mas_bvsb_segment_stats_dev``.
    """
    if not isinstance(logits, torch.Tensor) or not logits.is_cuda:
        raise RuntimeError("logits: expected a CUDA tensor (there is no CPU path)")
    if logits.dim() != 4 or logits.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"logits: expected (B,C,H,W) float32/bfloat16, got {tuple(logits.shape)} {logits.dtype}")
    _want(spx, "spx", torch.int32, 3)
    b, c, h, w = logits.shape
    # NCHW with dense planes; the image stride may exceed C*H*W (channel-sliced view, e.g. preds[:, :-1])
    if b * c * h * w > 0 and (logits.stride(3) != 1 or logits.stride(2) != w or logits.stride(1) != h * w
                              or (b > 1 and logits.stride(0) < c * h * w)):
        raise RuntimeError("logits: expected NCHW layout with contiguous planes")
    image_stride = logits.stride(0) if b > 1 else c * h * w
    if tuple(spx.shape) != (b, h, w):
        raise RuntimeError(f"spx shape {tuple(spx.shape)} does not match logits {tuple(logits.shape)}")
    _want(cls_sum, "cls_sum", torch.float32)
    _want(cls_cnt, "cls_cnt", torch.int32)
    if cls_sum.numel() != b * nseg * c or cls_cnt.numel() != b * nseg * c:
        raise RuntimeError("cls_sum / cls_cnt must hold B*nseg*C elements")
    if prob_sum is not None:
        _want(prob_sum, "prob_sum", torch.float64)
        if prob_sum.numel() != b * c:
            raise RuntimeError("prob_sum must hold B*C elements")
    if b == 0:
        return
    with _on(logits):
        _lib.call("mas_bvsb_segment_stats_dev", logits.data_ptr(),
                  _lib.MAS_F32 if logits.dtype == torch.float32 else _lib.MAS_BF16, int(image_stride), spx.data_ptr(),
                  b, c, h, w, int(nseg), float(temperature), cls_sum.data_ptr(), cls_cnt.data_ptr(),
                  _ptr(prob_sum), _stream(logits))


def bvsb_segment_stats_multi(batches, nseg: int, temperature: float, cls_sum: torch.Tensor, cls_cnt: torch.Tensor,
                             prob_sum: Optional[torch.Tensor], stream: Optional[int] = None) -> None:
    """``bvsb_segment_stats`` over several (logits, spx) batches in ONE launch (``mas_bvsb_segment_stats_multi_dev``):
    the batches live in different allocations but fill consecutive image rows of the tables.  All batches share
    dtype, C, H, W.  ``stream``: raw CUDA stream handle (default: torch's current stream)."""
    import ctypes
    if not batches:
        return
    if len(batches) > _lib.MAS_MAX_SEGMENTS:
        raise RuntimeError(f"at most {_lib.MAS_MAX_SEGMENTS} batches per launch")
    first = batches[0][0]
    if not isinstance(first, torch.Tensor) or not first.is_cuda:
        raise RuntimeError("logits: expected a CUDA tensor (there is no CPU path)")
    _, c, h, w = first.shape
    ptrs, strides, ids, counts, total = [], [], [], [], 0
    for logits, spx in batches:
        if logits.dim() != 4 or logits.dtype != first.dtype or logits.dtype not in (torch.float32, torch.bfloat16) \
                or tuple(logits.shape[1:]) != (c, h, w) or logits.device != first.device:
            raise RuntimeError("logits: every batch must be (B,C,H,W) float32/bfloat16 with the same C, H, W, dtype and device")
        _want(spx, "spx", torch.int32, 3)
        b = logits.shape[0]
        if b * c * h * w > 0 and (logits.stride(3) != 1 or logits.stride(2) != w or logits.stride(1) != h * w
                                  or (b > 1 and logits.stride(0) < c * h * w)):
            raise RuntimeError("logits: expected NCHW layout with contiguous planes")
        if tuple(spx.shape) != (b, h, w):
            raise RuntimeError(f"spx shape {tuple(spx.shape)} does not match logits {tuple(logits.shape)}")
        ptrs.append(logits.data_ptr()); ids.append(spx.data_ptr()); counts.append(b)
        strides.append(logits.stride(0) if b > 1 else c * h * w)
        total += b
    _want(cls_sum, "cls_sum", torch.float32)
    _want(cls_cnt, "cls_cnt", torch.int32)
    if cls_sum.numel() != total * nseg * c or cls_cnt.numel() != total * nseg * c:
        raise RuntimeError("cls_sum / cls_cnt must hold (sum of B)*nseg*C elements")
    if prob_sum is not None:
        _want(prob_sum, "prob_sum", torch.float64)
        if prob_sum.numel() != total * c:
            raise RuntimeError("prob_sum must hold (sum of B)*C elements")
    if total == 0:
        return
    n = len(batches)
    with _on(first):
        _lib.call("mas_bvsb_segment_stats_multi_dev", n, (ctypes.c_void_p * n)(*ptrs),
                  _lib.MAS_F32 if first.dtype == torch.float32 else _lib.MAS_BF16, (ctypes.c_int64 * n)(*strides),
                  (ctypes.c_void_p * n)(*ids), (ctypes.c_int * n)(*counts), c, h, w, int(nseg), float(temperature),
                  cls_sum.data_ptr(), cls_cnt.data_ptr(), _ptr(prob_sum), _stream(first) if stream is None else int(stream))


def bvsb_segment_stats_lowres(logits_lo: torch.Tensor, spx: torch.Tensor, nseg: int, temperature: float,
                              cls_sum: torch.Tensor, cls_cnt: torch.Tensor, prob_sum: Optional[torch.Tensor],
                              stream: Optional[int] = None) -> None:
    """``bvsb_segment_stats`` fed with the head's low-resolution logits (B,C,h,w): the bilinear up-sampling to the id
    map's (H,W) (``F.interpolate(..., align_corners=False)``) happens inside the kernel
    (``mas_bvsb_segment_stats_lowres_dev``)."""
    if not isinstance(logits_lo, torch.Tensor) or not logits_lo.is_cuda:
        raise RuntimeError("logits: expected a CUDA tensor (there is no CPU path)")
    if logits_lo.dim() != 4 or logits_lo.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"logits: expected (B,C,h,w) float32/bfloat16, got {tuple(logits_lo.shape)} {logits_lo.dtype}")
    _want(spx, "spx", torch.int32, 3)
    b, c, h_in, w_in = logits_lo.shape
    if b * c * h_in * w_in > 0 and (logits_lo.stride(3) != 1 or logits_lo.stride(2) != w_in or logits_lo.stride(1) != h_in * w_in
                                    or (b > 1 and logits_lo.stride(0) < c * h_in * w_in)):
        raise RuntimeError("logits: expected NCHW layout with contiguous planes")
    image_stride = logits_lo.stride(0) if b > 1 else c * h_in * w_in
    if spx.shape[0] != b or spx.shape[1] < h_in or spx.shape[2] < w_in:
        raise RuntimeError(f"spx shape {tuple(spx.shape)} does not match low-resolution logits {tuple(logits_lo.shape)}")
    h, w = spx.shape[1:]
    _want(cls_sum, "cls_sum", torch.float32)
    _want(cls_cnt, "cls_cnt", torch.int32)
    if cls_sum.numel() != b * nseg * c or cls_cnt.numel() != b * nseg * c:
        raise RuntimeError("cls_sum / cls_cnt must hold B*nseg*C elements")
    if prob_sum is not None:
        _want(prob_sum, "prob_sum", torch.float64)
        if prob_sum.numel() != b * c:
            raise RuntimeError("prob_sum must hold B*C elements")
    if b == 0:
        return
    with _on(logits_lo):
        _lib.call("mas_bvsb_segment_stats_lowres_dev", logits_lo.data_ptr(),
                  _lib.MAS_F32 if logits_lo.dtype == torch.float32 else _lib.MAS_BF16, int(image_stride), h_in, w_in, spx.data_ptr(),
                  b, c, h, w, int(nseg), float(temperature), cls_sum.data_ptr(), cls_cnt.data_ptr(), _ptr(prob_sum),
                  _stream(logits_lo) if stream is None else int(stream))


def class_weights(prob_sum: torch.Tensor, pixels_per_image: int, ref_batch: int, coeff: float) -> torch.Tensor:
    """(N,C) f64 per-image probability sums in pool order -> (C,) f32 class weights (``mas_class_weights_dev``)."""
    _want(prob_sum, "prob_sum", torch.float64, 2)
    n, c = prob_sum.shape
    weight = torch.empty(c, dtype=torch.float32, device=prob_sum.device)
    with _on(prob_sum):
        _lib.call("mas_class_weights_dev", prob_sum.data_ptr(), n, c, int(pixels_per_image), max(int(ref_batch), 1), float(coeff),
                  weight.data_ptr(), _stream(prob_sum))
    return weight


def prefix_cut(sorted_keys: torch.Tensor, count: torch.Tensor, cost_by_tie: torch.Tensor, budget: int) -> torch.Tensor:
    """Device int32: length of the prefix of the ranked list whose running cost first exceeds ``budget``
    (``mas_prefix_cut_dev``)."""
    _want(sorted_keys, "sorted_keys", torch.int64, 1)
    _want(count, "count", torch.int32, 1)
    _want(cost_by_tie, "cost_by_tie", torch.uint8, 1)
    n_take = torch.empty(1, dtype=torch.int32, device=sorted_keys.device)
    with _on(sorted_keys):
        _lib.call("mas_prefix_cut_dev", sorted_keys.data_ptr(), count.data_ptr(), cost_by_tie.data_ptr(), cost_by_tie.numel(),
                  int(budget), n_take.data_ptr(), _stream(sorted_keys))
    return n_take


def region_scores(cls_sum: torch.Tensor, cls_cnt: torch.Tensor, class_weight: Optional[torch.Tensor] = None,
                  want_npix: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor]:
    """(score, npix | None, dominant), each shaped like cls_sum without its last dim."""
    _want(cls_sum, "cls_sum", torch.float32)
    _want(cls_cnt, "cls_cnt", torch.int32)
    c = cls_sum.shape[-1]
    shape = cls_sum.shape[:-1]
    n = cls_sum.numel() // c
    if class_weight is not None:
        _want(class_weight, "class_weight", torch.float32)
        if class_weight.numel() != c:
            raise RuntimeError("class_weight must have C elements")
    score = torch.empty(shape, dtype=torch.float32, device=cls_sum.device)
    dominant = torch.empty(shape, dtype=torch.int32, device=cls_sum.device)
    npix = torch.empty(shape, dtype=torch.int32, device=cls_sum.device) if want_npix else None
    with _on(cls_sum):
        _lib.call("mas_region_scores_dev", cls_sum.data_ptr(), cls_cnt.data_ptr(), _ptr(class_weight), n, c,
                  score.data_ptr(), _ptr(npix), dominant.data_ptr(), _stream(cls_sum))
    return score, npix, dominant


def minmax_nonzero(values: torch.Tensor) -> torch.Tensor:
    """Device tensor [min over non-zero entries, max over all entries]."""
    _want(values, "values", torch.float32)
    out = torch.empty(2, dtype=torch.float32, device=values.device)
    with _on(values):
        _lib.call("mas_minmax_nonzero_dev", values.data_ptr(), values.numel(), out.data_ptr(), _stream(values))
    return out


def dominant_hist(dominant: torch.Tensor, channels: int) -> torch.Tensor:
    _want(dominant, "dominant", torch.int32)
    hist = torch.zeros(channels, dtype=torch.int64, device=dominant.device)
    with _on(dominant):
        _lib.call("mas_dominant_hist_dev", dominant.data_ptr(), dominant.numel(), int(channels), hist.data_ptr(),
                  _stream(dominant))
    return hist


def finalize_scores(score: torch.Tensor, dominant: Optional[torch.Tensor], minmax: Optional[torch.Tensor] = None,
                    ban_class: int = -1, region_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In-place normalise / ban / re-weight (the reference's order of operations)."""
    _want(score, "score", torch.float32)
    if dominant is not None:
        _want(dominant, "dominant", torch.int32)
    if minmax is not None:
        _want(minmax, "minmax", torch.float32)
    if region_weight is not None:
        _want(region_weight, "region_weight", torch.float32)
    with _on(score):
        _lib.call("mas_finalize_scores_dev", score.data_ptr(), _ptr(dominant), score.numel(), _ptr(minmax),
                  int(ban_class), _ptr(region_weight), _stream(score))
    return score


def region_keys(score: torch.Tensor, in_pool: torch.Tensor, image_rank: torch.Tensor) -> torch.Tensor:
    """(N,S) scores -> (N*S,) uint64-as-int64 keys; 0 for regions outside the pool."""
    _want(score, "score", torch.float32, 2)
    _want(in_pool, "in_pool", torch.uint8, 2)
    _want(image_rank, "image_rank", torch.int32, 1)
    n, s = score.shape
    if tuple(in_pool.shape) != (n, s) or image_rank.numel() != n:
        raise RuntimeError("in_pool / image_rank shape mismatch")
    keys = torch.empty(n * s, dtype=torch.int64, device=score.device)
    with _on(score):
        _lib.call("mas_region_keys_dev", score.data_ptr(), in_pool.data_ptr(), image_rank.data_ptr(), n, s,
                  keys.data_ptr(), _stream(score))
    return keys


def sort_capacity(n: int) -> int:
    return int(_lib.load().mas_sort_capacity(int(n)))


def topk_sorted(keys: torch.Tensor, k: int, sort: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fast path of ``topk_keys``: bucket histograms + compaction (+ sort) -- ``mas_topk_sorted_u64_dev`` /
    ``mas_topk_candidates_u64_dev``.  ``sort=True``: the k largest keys, descending, count = min(k, #keys);
    ``sort=False``: an unordered superset of them, count = its size.  The device count is -1 when the candidates
    overflowed the buffer (``sort_capacity(k)`` slots); the caller then uses ``topk_keys``."""
    _want(keys, "keys", torch.int64, 1)
    k = int(k)
    cap = sort_capacity(max(k, 1))
    out = torch.empty(cap, dtype=torch.int64, device=keys.device)
    count = torch.empty(1, dtype=torch.int32, device=keys.device)
    ws_bytes = int(_lib.load().mas_topk_workspace_bytes())
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=keys.device)
    with _on(keys):
        _lib.call("mas_topk_sorted_u64_dev" if sort else "mas_topk_candidates_u64_dev", keys.data_ptr(), keys.numel(), k,
                  out.data_ptr(), cap, count.data_ptr(), ws.data_ptr(), ws_bytes, _stream(keys))
    return out, count


def topk_candidates_msg(keys: torch.Tensor, k: int) -> torch.Tensor:
    """This rank's message for the multi-GPU top-k merge (``mas_topk_candidates_msg_u64_dev``): ``sort_capacity(k) + 1``
    int64 slots = an unordered superset of the k largest keys (0 = no key) followed by their count (-1 = overflow)."""
    _want(keys, "keys", torch.int64, 1)
    k = int(k)
    cap = sort_capacity(max(k, 1))
    msg = torch.empty(cap + 1, dtype=torch.int64, device=keys.device)
    count = torch.empty(1, dtype=torch.int32, device=keys.device)
    ws_bytes = int(_lib.load().mas_topk_workspace_bytes())
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=keys.device)
    with _on(keys):
        _lib.call("mas_topk_candidates_msg_u64_dev", keys.data_ptr(), keys.numel(), k, msg.data_ptr(), cap, count.data_ptr(),
                  ws.data_ptr(), ws_bytes, _stream(keys))
    return msg


def merge_counts(gathered: torch.Tensor, worst: torch.Tensor) -> None:
    """(world, cap + 1) gathered messages: worst[0] <- the smallest per-rank count (-1 if any rank overflowed); the count
    slots are cleared in place so that ``gathered.view(-1)`` is a plain key list (``mas_merge_counts_u64_dev``)."""
    _want(gathered, "gathered", torch.int64, 2)
    _want(worst, "worst", torch.int32, 1)
    with _on(gathered):
        _lib.call("mas_merge_counts_u64_dev", gathered.data_ptr(), gathered.shape[0], gathered.shape[1], worst.data_ptr(),
                  _stream(gathered))


def topk_keys(keys: torch.Tensor, k: int, sort: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """The k largest non-zero keys (sorted descending if ``sort``) and a device int32 count.

    The returned buffer has ``sort_capacity(k)`` slots; slots past the count hold zeros when sorted.
    """
    _want(keys, "keys", torch.int64, 1)
    k = int(k)
    cap = sort_capacity(max(k, 1))
    out = torch.zeros(cap, dtype=torch.int64, device=keys.device)
    count = torch.zeros(1, dtype=torch.int32, device=keys.device)
    ws_bytes = int(_lib.load().mas_topk_workspace_bytes())
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=keys.device)
    with _on(keys):
        _lib.call("mas_topk_u64_dev", keys.data_ptr(), keys.numel(), k, out.data_ptr(), count.data_ptr(),
                  ws.data_ptr(), ws_bytes, _stream(keys))
        if sort and k > 1:
            _lib.call("mas_sort_desc_u64_dev", out.data_ptr(), k, _stream(keys))
    return out, count


# ------------------------------------------------------------------------------------------------ stage-1 losses
def _ids_dtype(spx: torch.Tensor) -> int:
    if spx.dtype == torch.int64:
        return _lib.MAS_I64
    if spx.dtype == torch.int32:
        return _lib.MAS_I32
    raise RuntimeError(f"superpixels: expected int32 or int64, got {spx.dtype}")


def multihot_info(targets: torch.Tensor, channels: int, group_mode: int) -> torch.Tensor:
    """(..., Ct) uint8 multi-hot targets -> (...) candidate words (int32 storage, see ``mas_multihot_info_dev``)."""
    _want(targets, "targets", torch.uint8)
    ct = targets.shape[-1]
    info = torch.empty(targets.shape[:-1], dtype=torch.int32, device=targets.device)
    with _on(targets):
        _lib.call("mas_multihot_info_dev", targets.data_ptr(), info.numel(), int(ct), int(channels), int(group_mode),
                  info.data_ptr(), _stream(targets))
    return info


_tile_bytes = {}


def multihot_tiles(mask: torch.Tensor) -> torch.Tensor:
    """(N,H,W) bool|uint8 mask -> the active-tile list both loss passes take (``mas_multihot_tiles_dev``): which 32 px x 8 row
    tiles hold a selected pixel, with prefix sums.  One scan of the mask; reused by the forward and the backward pass."""
    _want(mask, "spmasks", (torch.bool, torch.uint8), 3)
    n, h, w = mask.shape
    need = _tile_bytes.get((n, h, w))
    if need is None:
        need = _tile_bytes[(n, h, w)] = int(_lib.load().mas_multihot_tiles_workspace_bytes(n, h, w))
    tiles = torch.empty(need, dtype=torch.uint8, device=mask.device)
    if n * h * w:
        with _on(mask):
            _lib.call("mas_multihot_tiles_dev", mask.data_ptr(), n, h, w, tiles.data_ptr(), need, _stream(mask))
    return tiles


def _loss_args(logits, spx, mask, info, nseg):
    _want(logits, "inputs", torch.float32, 4)
    _want(spx, "superpixels", (torch.int32, torch.int64), 3)
    _want(mask, "spmasks", (torch.bool, torch.uint8), 3)
    _want(info, "info", torch.int32, 2)
    n, c, h, w = logits.shape
    if tuple(spx.shape) != (n, h, w) or tuple(mask.shape) != (n, h, w):
        raise RuntimeError(f"superpixels {tuple(spx.shape)} / spmasks {tuple(mask.shape)} do not match inputs {tuple(logits.shape)}")
    if tuple(info.shape) != (n, nseg):
        raise RuntimeError(f"targets: expected {(n, nseg)} regions, got {tuple(info.shape)}")
    return n, c, h, w


def multihot_loss_forward(logits: torch.Tensor, spx: torch.Tensor, mask: torch.Tensor, info: torch.Tensor, nseg: int,
                          temperature: float, flags: int, tiles: Optional[torch.Tensor] = None):
    """-> (acc (8,) f64 bucket sums / counts, group_max (N,nseg,C) i64 packed maxima or None).
    ``tiles``: the active-tile list of ``mask`` (``multihot_tiles``) -- only tiles with selected pixels are visited."""
    n, c, h, w = _loss_args(logits, spx, mask, info, nseg)
    # one zero-fill for both: 8 accumulators (viewed as f64) followed by the max-pool table
    want_group = bool(flags & _lib.MAS_LOSS_GROUP)
    buf = torch.zeros(8 + (n * nseg * c if want_group else 0), dtype=torch.int64, device=logits.device)
    acc = buf[:8].view(torch.float64)
    gmax = buf[8:].view(n, nseg, c) if want_group else None
    with _on(logits):
        _lib.call("mas_multihot_loss_fwd_tiles_dev", logits.data_ptr(), spx.data_ptr(), _ids_dtype(spx), mask.data_ptr(),
                  info.data_ptr(), _ptr(tiles), n, c, h, w, int(nseg), float(temperature), int(flags), acc.data_ptr(), _ptr(gmax),
                  _stream(logits))
    return acc, gmax


def multihot_loss_backward(logits: torch.Tensor, spx: torch.Tensor, mask: torch.Tensor, info: torch.Tensor,
                           gmax: Optional[torch.Tensor], coef: torch.Tensor, nseg: int, temperature: float,
                           flags: int, tiles: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Dense d(sum_k coef[k] * bucket_sum[k]) / d logits; coef = 4 device floats.  ``tiles`` as in the forward pass: a
    sparsely selected batch gets its gradient zeroed by one linear sweep and only the active tiles computed."""
    n, c, h, w = _loss_args(logits, spx, mask, info, nseg)
    _want(coef, "coef", torch.float32, 1)
    if coef.numel() != 4:
        raise RuntimeError("coef must hold 4 floats")
    grad = torch.empty_like(logits)
    with _on(logits):
        _lib.call("mas_multihot_loss_bwd_tiles_dev", logits.data_ptr(), spx.data_ptr(), _ids_dtype(spx), mask.data_ptr(),
                  info.data_ptr(), _ptr(tiles), _ptr(gmax), coef.data_ptr(), n, c, h, w, int(nseg), float(temperature), int(flags),
                  grad.data_ptr(), _stream(logits))
    return grad


_step_bytes = {}


def stage1_forward(logits: torch.Tensor, spx: torch.Tensor, mask: torch.Tensor, targets: torch.Tensor, temperature: float,
                   group_mode: int, flags: int):
    """The whole forward side of the loss step in ONE library call (``mas_stage1_loss_fwd_dev``): candidate words,
    active-tile list, fused pass, group reduction, normalised losses.
    -> (workspace (opaque, needed by ``stage1_backward``), acc (8,) f64 view, losses (6,) f32 view)."""
    _want(logits, "inputs", torch.float32, 4)
    _want(spx, "superpixels", (torch.int32, torch.int64), 3)
    _want(mask, "spmasks", (torch.bool, torch.uint8), 3)
    _want(targets, "targets", torch.uint8, 3)
    n, c, h, w = logits.shape
    nseg, ct = targets.shape[1], targets.shape[2]
    if tuple(spx.shape) != (n, h, w) or tuple(mask.shape) != (n, h, w) or targets.shape[0] != n:
        raise RuntimeError(f"superpixels {tuple(spx.shape)} / spmasks {tuple(mask.shape)} / targets {tuple(targets.shape)} "
                           f"do not match inputs {tuple(logits.shape)}")
    key = (n, c, h, w, nseg)
    need = _step_bytes.get(key)
    if need is None:
        need = _step_bytes[key] = int(_lib.load().mas_stage1_workspace_bytes(n, c, h, w, nseg))
    ws = torch.empty(need, dtype=torch.uint8, device=logits.device)
    with _on(logits):
        _lib.call("mas_stage1_loss_fwd_dev", logits.data_ptr(), spx.data_ptr(), _ids_dtype(spx), mask.data_ptr(), targets.data_ptr(), ct,
                  n, c, h, w, nseg, float(temperature), int(group_mode), int(flags), ws.data_ptr(), need, _stream(logits))
    return ws, ws[:64].view(torch.float64), ws[64:88].view(torch.float32)


def stage1_backward(logits: torch.Tensor, spx: torch.Tensor, mask: torch.Tensor, ws: torch.Tensor, nseg: int, temperature: float,
                    flags: int, grads) -> torch.Tensor:
    """Dense gradient of sum_j grads[j] * losses[j] (``mas_stage1_loss_bwd_dev``); ``grads``: six 0-dim fp32 CUDA tensors or None."""
    import ctypes
    n, c, h, w = logits.shape
    keep = [None if g is None else (g if g.dtype == torch.float32 else g.float()) for g in grads]
    ptrs = (ctypes.c_void_p * 6)(*[None if g is None else g.data_ptr() for g in keep])
    grad = torch.empty_like(logits)
    with _on(logits):
        _lib.call("mas_stage1_loss_bwd_dev", logits.data_ptr(), spx.data_ptr(), _ids_dtype(spx), mask.data_ptr(), n, c, h, w, int(nseg),
                  float(temperature), int(flags), ws.data_ptr(), ptrs, grad.data_ptr(), _stream(logits))
    return grad


def multihot_loss_finish(acc: torch.Tensor) -> torch.Tensor:
    """(8,) f64 bucket sums / counts -> (6,) f32 normalised losses (see ``mas_multihot_loss_finish_dev``)."""
    _want(acc, "acc", torch.float64, 1)
    losses = torch.empty(6, dtype=torch.float32, device=acc.device)
    with _on(acc):
        _lib.call("mas_multihot_loss_finish_dev", acc.data_ptr(), losses.data_ptr(), _stream(acc))
    return losses


def multihot_loss_coef(acc: torch.Tensor, grad_losses: torch.Tensor) -> torch.Tensor:
    """Gradient of the 6 normalised losses folded back onto the 4 bucket sums (device floats)."""
    _want(acc, "acc", torch.float64, 1)
    _want(grad_losses, "grad_losses", torch.float32, 1)
    if grad_losses.numel() != 6:
        raise RuntimeError("grad_losses must hold 6 floats")
    coef = torch.empty(4, dtype=torch.float32, device=acc.device)
    with _on(acc):
        _lib.call("mas_multihot_loss_coef_dev", acc.data_ptr(), grad_losses.data_ptr(), coef.data_ptr(), _stream(acc))
    return coef


# ------------------------------------------------------------------------------------------------ stage-2 labellers
def candidate_argmax(logits: torch.Tensor, spx: torch.Tensor, mask: torch.Tensor, info: torch.Tensor, nseg: int) -> torch.Tensor:
    """(N,H,W) uint8 labels: arg-max of logit * multi-hot row on selected pixels, 255 elsewhere."""
    n, c, h, w = _loss_args(logits, spx, mask, info, nseg)
    labels = torch.empty((n, h, w), dtype=torch.uint8, device=logits.device)
    with _on(logits):
        _lib.call("mas_candidate_argmax_dev", logits.data_ptr(), spx.data_ptr(), _ids_dtype(spx), mask.data_ptr(),
                  info.data_ptr(), n, c, h, w, int(nseg), labels.data_ptr(), _stream(logits))
    return labels


_workspaces = {}
_workspace_bytes = {}


def _labeller_workspace(device, fch, c, h, w, nseg) -> torch.Tensor:
    """Scratch of the prototype labeller, cached per (device, stream): calls on different streams never share (or
    regrow) a buffer that kernels queued on another stream still use."""
    shape = (fch, c, h, w, nseg)
    need = _workspace_bytes.get(shape)
    if need is None:
        need = _workspace_bytes[shape] = int(_lib.load().mas_proto_labeller_workspace_bytes(fch, c, h, w, nseg))
    index = device.index if device.index is not None else torch.cuda.current_device()
    key = (index, torch._C._cuda_getCurrentRawStream(index))
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty(need, dtype=torch.uint8, device=device)   # caching-allocator blocks are >= 512-byte aligned
        _workspaces[key] = buf
    return buf


def proto_labeller(feats: torch.Tensor, logits: torch.Tensor, targets: torch.Tensor, mask: torch.Tensor, spx: torch.Tensor,
                   only_multihot: bool, threshold: str, labels: Optional[torch.Tensor] = None,
                   status: Optional[torch.Tensor] = None):
    """One image: feats (F,H,W) f32|bf16 -- or the head's LOW-RESOLUTION (F,h,w) map, interpolated inside the kernels --,
    logits (C,H,W) f32, targets (S,Ct) u8, mask (H,W) bool, spx (H,W) i32|i64
    -> (labels (H,W) uint8, status (1,) int32 on the device).  See ``mas_proto_labeller_src_dev``.
    ``labels`` / ``status``: optional preallocated outputs (status must be zero) -- the batched caller allocates them on
    its own stream and runs the images on side streams."""
    _want(feats, "feats", (torch.float32, torch.bfloat16), 3)
    _want(logits, "inputs", torch.float32, 3)
    _want(targets, "targets", torch.uint8, 2)
    _want(mask, "spmasks", (torch.bool, torch.uint8), 2)
    _want(spx, "superpixels", (torch.int32, torch.int64), 2)
    fch, fh, fw = feats.shape
    c, h, w = logits.shape
    nseg, ct = targets.shape
    if tuple(mask.shape) != (h, w) or tuple(spx.shape) != (h, w):
        raise RuntimeError("inputs / spmasks / superpixels spatial shapes differ")
    if fh > h or fw > w:
        raise RuntimeError(f"feats ({fh}x{fw}) larger than the image ({h}x{w})")
    if threshold not in ("median", "min"):
        raise NotImplementedError(f"cosprop_threshold_method={threshold!r}")
    if labels is None:
        labels = torch.empty((h, w), dtype=torch.uint8, device=feats.device)
    if status is None:
        status = torch.zeros(1, dtype=torch.int32, device=feats.device)
    if labels.dtype != torch.uint8 or tuple(labels.shape) != (h, w) or not labels.is_contiguous() or status.dtype != torch.int32:
        raise RuntimeError("labels must be a contiguous (H,W) uint8 tensor and status an int32 tensor")
    ws = _labeller_workspace(feats.device, fch, c, h, w, nseg)
    with _on(feats):
        _lib.call("mas_proto_labeller_src_dev", feats.data_ptr(), _lib.MAS_F32 if feats.dtype == torch.float32 else _lib.MAS_BF16,
                  fch, fh, fw, logits.data_ptr(), c, targets.data_ptr(), ct, mask.data_ptr(), spx.data_ptr(), _ids_dtype(spx), h, w,
                  nseg, int(bool(only_multihot)), _lib.MAS_THRESHOLD_MEDIAN if threshold == "median" else _lib.MAS_THRESHOLD_MIN,
                  labels.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), _stream(feats))
    return labels, status


_batch_workspaces = {}


def proto_labeller_batch(feats: torch.Tensor, logits: torch.Tensor, targets: torch.Tensor, mask: torch.Tensor, spx: torch.Tensor,
                         only_multihot: bool, threshold: str, lanes):
    """A loader batch in ONE library call (``mas_proto_labeller_batch_dev``): feats (N,F,fh,fw) f32|bf16, logits (N,C,H,W) f32,
    targets (N,S,Ct) u8, mask (N,H,W) bool, spx (N,H,W) i32|i64, ``lanes``: side streams (torch.cuda.Stream) the images are
    dealt to round-robin -> (labels (N,H,W) uint8, status (N,) int32 on the device)."""
    import ctypes
    _want(feats, "feats", (torch.float32, torch.bfloat16), 4)
    _want(logits, "inputs", torch.float32, 4)
    _want(targets, "targets", torch.uint8, 3)
    _want(mask, "spmasks", (torch.bool, torch.uint8), 3)
    _want(spx, "superpixels", (torch.int32, torch.int64), 3)
    n, fch, fh, fw = feats.shape
    _, c, h, w = logits.shape
    nseg, ct = targets.shape[1], targets.shape[2]
    if logits.shape[0] != n or targets.shape[0] != n or tuple(mask.shape) != (n, h, w) or tuple(spx.shape) != (n, h, w):
        raise RuntimeError("feats / inputs / targets / spmasks / superpixels batch or spatial shapes differ")
    if fh > h or fw > w:
        raise RuntimeError(f"feats ({fh}x{fw}) larger than the image ({h}x{w})")
    if threshold not in ("median", "min"):
        raise NotImplementedError(f"cosprop_threshold_method={threshold!r}")
    device = feats.device
    labels = torch.empty((n, h, w), dtype=torch.uint8, device=device)
    status = torch.zeros(n, dtype=torch.int32, device=device)
    if n == 0:
        return labels, status
    shape = (fch, c, h, w, nseg)
    need = _workspace_bytes.get(shape)
    if need is None:
        need = _workspace_bytes[shape] = int(_lib.load().mas_proto_labeller_workspace_bytes(fch, c, h, w, nseg))
    per_lane = (need + 255) // 256 * 256
    n_lanes = max(1, min(len(lanes), n))
    index = device.index if device.index is not None else torch.cuda.current_device()
    key = (index, torch._C._cuda_getCurrentRawStream(index))
    ws = _batch_workspaces.get(key)
    if ws is None or ws.numel() < per_lane * n_lanes:
        ws = _batch_workspaces[key] = torch.empty(per_lane * n_lanes, dtype=torch.uint8, device=device)
    handles = (ctypes.c_void_p * n_lanes)(*[s.cuda_stream for s in lanes[:n_lanes]])
    with _on(feats):
        _lib.call("mas_proto_labeller_batch_dev", feats.data_ptr(), _lib.MAS_F32 if feats.dtype == torch.float32 else _lib.MAS_BF16,
                  fch, fh, fw, logits.data_ptr(), c, targets.data_ptr(), ct, mask.data_ptr(), spx.data_ptr(), _ids_dtype(spx), n, h, w,
                  nseg, int(bool(only_multihot)), _lib.MAS_THRESHOLD_MEDIAN if threshold == "median" else _lib.MAS_THRESHOLD_MIN,
                  labels.data_ptr(), status.data_ptr(), ws.data_ptr(), per_lane, handles, n_lanes, _stream(feats))
    return labels, status


# ------------------------------------------------------------------------------------------------ offline label generation
def multihot_labels(spx: torch.Tensor, target: torch.Tensor, keep: torch.Tensor, nseg: int, num_classes: int,
                    trim_kernel_size: int = 0):
    """One image: spx (H,W) i32|i64, target (H,W) u8 train ids (255 = ignore), keep (nseg,) u8
    -> (multi_hot (nseg, num_classes+1) u8, size (nseg,) i32).  See ``mas_multihot_labels_dev``."""
    _want(spx, "superpixel", (torch.int32, torch.int64), 2)
    _want(target, "target", torch.uint8, 2)
    _want(keep, "keep", torch.uint8, 1)
    h, w = spx.shape
    if tuple(target.shape) != (h, w) or keep.numel() != nseg:
        raise RuntimeError("target / keep shapes do not match the superpixel map")
    multi_hot = torch.empty((nseg, num_classes + 1), dtype=torch.uint8, device=spx.device)
    size = torch.empty((nseg,), dtype=torch.int32, device=spx.device)
    need = int(_lib.load().mas_multihot_labels_workspace_bytes(int(nseg), int(num_classes)))
    ws = torch.empty(need, dtype=torch.uint8, device=spx.device)
    with _on(spx):
        _lib.call("mas_multihot_labels_dev", spx.data_ptr(), _ids_dtype(spx), target.data_ptr(), keep.data_ptr(), h, w,
                  int(nseg), int(num_classes), int(trim_kernel_size), multi_hot.data_ptr(), size.data_ptr(), ws.data_ptr(),
                  need, _stream(spx))
    return multi_hot, size


def dominant_labels(spx: torch.Tensor, target: torch.Tensor, keep: torch.Tensor, nseg: int, num_classes: int) -> torch.Tensor:
    """One image: spx (H,W) i32|i64, target (H,W) u8 train ids (255 = ignore), keep (nseg,) u8 -> relabelled (H,W) u8.
    See ``mas_dominant_labels_dev``."""
    _want(spx, "superpixel", (torch.int32, torch.int64), 2)
    _want(target, "target", torch.uint8, 2)
    _want(keep, "keep", torch.uint8, 1)
    if tuple(target.shape) != tuple(spx.shape) or keep.numel() != nseg:
        raise RuntimeError("dominant_labels: shape mismatch")
    h, w = spx.shape
    out = torch.empty_like(target)
    need = int(_lib.load().mas_dominant_labels_workspace_bytes(int(nseg), int(num_classes)))
    ws = torch.empty(need, dtype=torch.uint8, device=spx.device)
    with _on(spx):
        _lib.call("mas_dominant_labels_dev", spx.data_ptr(), _ids_dtype(spx), target.data_ptr(), keep.data_ptr(), h, w,
                  int(nseg), int(num_classes), out.data_ptr(), ws.data_ptr(), need, _stream(spx))
    return out


_LABEL_DTYPES = {torch.int64: _lib.MAS_I64, torch.int32: _lib.MAS_I32, torch.uint8: _lib.MAS_U8}


def miou_counts(outputs: torch.Tensor, targets: torch.Tensor, num_classes: int, ignore_label: int, by_output: bool,
                counts: torch.Tensor) -> None:
    """Accumulate [seen | correct | positive] (3*num_classes int64, device) over two label maps of the same shape."""
    if not (isinstance(outputs, torch.Tensor) and outputs.is_cuda and isinstance(targets, torch.Tensor) and targets.is_cuda):
        raise RuntimeError("miou_counts: expected CUDA tensors (there is no CPU path)")
    if outputs.shape != targets.shape:
        raise RuntimeError(f"miou_counts: outputs {tuple(outputs.shape)} and targets {tuple(targets.shape)} differ")
    if outputs.dtype != targets.dtype or outputs.dtype not in _LABEL_DTYPES:
        outputs, targets = outputs.long(), targets.long()
    outputs, targets = outputs.contiguous(), targets.contiguous()
    _want(counts, "counts", torch.int64, 1)
    if counts.numel() != 3 * num_classes:
        raise RuntimeError("miou_counts: counts must hold 3 * num_classes int64")
    with _on(outputs):
        _lib.call("mas_miou_counts_dev", outputs.data_ptr(), targets.data_ptr(), _LABEL_DTYPES[outputs.dtype], outputs.numel(),
                  int(num_classes), int(ignore_label), _lib.MAS_MIOU_BY_OUTPUT if by_output else _lib.MAS_MIOU_BY_TARGET,
                  counts.data_ptr(), _stream(outputs))
