"""Device-side acquisition engine shared by the six ``my_bvsb*`` selectors.

One streaming pass over the logits fills three tables (``RegionStats``); every
selector variant is then a cheap epilogue over them (``finalize``).  The
re-association that makes the reference's two-pass ``predclsbal`` selector a
single pass:  score_s = mean_{pix in s}(bvsb * w[top1]) = sum_c w_c * B[s,c] / n_s
with B[s,c] = sum of bvsb over the pixels of s whose arg-max class is c
(reference: active_selection/my_bvsb_predclsbal_pwr.py:50-65).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from . import dist as mdist
from . import ops


@dataclass(frozen=True)
class SelectorSpec:
    """What distinguishes the reference selector modules (active_selection/<name>.py)."""
    normalise: bool        # min/max normalisation over the whole pool   (my_bvsb.py:79-81)
    ban_ignore: bool       # zero regions dominated by the last channel   (my_bvsb_banignore.py:59-61)
    weighting: str         # "none" | "predclsbal" (pixel-wise power weights) | "clsbal" (region-wise exp(-freq))
    slice_ignore: bool     # my_bvsb drops the ignore channel of predignore nets (my_bvsb.py:65-66)


SELECTORS = {
    "my_bvsb": SelectorSpec(True, False, "none", True),
    "my_bvsb_banignore": SelectorSpec(True, True, "none", False),
    "my_bvsb_predclsbal_pwr": SelectorSpec(False, False, "predclsbal", False),
    "my_bvsb_predclsbal_pwr_banignore": SelectorSpec(False, True, "predclsbal", False),
    "my_bvsb_clsbal_v2": SelectorSpec(True, False, "clsbal", False),
    "my_bvsb_clsbal_v2_banignore": SelectorSpec(True, True, "clsbal", False),
}


class RegionStats:
    """(image, superpixel, class) tables of one pool shard, resident in HBM.

    cls_sum (n,S,C) f32 : sum of bvsb over pixels with arg-max class c
    cls_cnt (n,S,C) i32 : pixel count = arg-max histogram (exact)
    prob_sum (n,C)  f64 : sum over pixels of softmax(l/T)   (only when ``need_prob``)

    Consecutive ``add_batch`` calls touch disjoint table rows and only add, so they carry no mutual dependency:
    * launches alternate over ``lanes`` side streams (forked from / joined to the caller's stream with events).  The scorer
      is a single-wave grid, so the CTAs of launch i+1 move onto SMs as the CTAs of launch i retire -- ramp and tail of the
      ~100 us launches overlap instead of adding a launch gap each.  ``lanes`` = 1 keeps everything on the caller's stream.
    * batches are GROUPED: ``add_batch`` only queues the batch (holding its tensors); a launch covers up to
      ``MAS_MAX_SEGMENTS`` queued batches and goes out once ``group_bytes`` of logits are waiting (or at ``join``).  One
      launch per loader batch is launch-latency-bound for small images (VOC: 66 MB per batch of 4) and leaves ramp / tail
      on the table for big ones; grouped, every launch streams ~1.5 GB.  ``group_bytes`` = 0 launches every batch at once.
    """

    def __init__(self, n_img: int, nseg: int, channels: int, device, need_prob: bool, lanes: int = 2,
                 group_bytes: int = 1536 << 20):
        self.n_img, self.nseg, self.channels = int(n_img), int(nseg), int(channels)
        self.lanes = [torch.cuda.Stream(device=device) for _ in range(lanes)] if lanes > 1 else []
        self.group_bytes = int(group_bytes)
        self._turn, self._dirty = 0, False
        self._queue, self._queued_bytes, self._queue_first, self._queue_temp, self._queue_n = [], 0, 0, None, 0
        self.launches = 0
        self._cls_sum = torch.zeros((n_img, nseg, channels), dtype=torch.float32, device=device)
        self._cls_cnt = torch.zeros((n_img, nseg, channels), dtype=torch.int32, device=device)
        self._prob_sum = torch.zeros((n_img, channels), dtype=torch.float64, device=device) if need_prob else None
        self.pixels_per_image: Optional[int] = None

    # reading a table orders the caller's stream after the launches still running on the side streams
    @property
    def cls_sum(self) -> torch.Tensor:
        self.join()
        return self._cls_sum

    @property
    def cls_cnt(self) -> torch.Tensor:
        self.join()
        return self._cls_cnt

    @property
    def prob_sum(self) -> Optional[torch.Tensor]:
        self.join()
        return self._prob_sum

    def join(self):
        """Launch what is still queued and make the caller's current stream wait for every outstanding launch."""
        self.flush()
        if self._dirty:
            main = torch.cuda.current_stream(self._cls_sum.device)
            for lane in self.lanes:
                main.wait_stream(lane)
            self._dirty = False

    def zero_(self):
        self.join()
        self._cls_sum.zero_()
        self._cls_cnt.zero_()
        if self._prob_sum is not None:
            self._prob_sum.zero_()

    def flush(self):
        """One launch over the queued batches (consecutive image rows of the tables)."""
        if not self._queue:
            return
        queue, first, temperature, n = self._queue, self._queue_first, self._queue_temp, self._queue_n
        self._queue, self._queued_bytes, self._queue_n = [], 0, 0
        tables = (self._cls_sum[first:first + n], self._cls_cnt[first:first + n],
                  None if self._prob_sum is None else self._prob_sum[first:first + n])
        self.launches += 1
        if not self.lanes:
            ops.bvsb_segment_stats_multi(queue, self.nseg, temperature, *tables)
            return
        lane = self.lanes[self._turn % len(self.lanes)]
        self._turn += 1
        lane.wait_stream(torch.cuda.current_stream(self._cls_sum.device))   # inputs (and the zeroed tables) are ready
        ops.bvsb_segment_stats_multi(queue, self.nseg, temperature, *tables, stream=lane.cuda_stream)
        for x, ids in queue:
            x.record_stream(lane)     # the caching allocator must not recycle them before the lane is done
            ids.record_stream(lane)
        self._dirty = True

    def add_batch(self, first_img: int, logits: torch.Tensor, spx: torch.Tensor, temperature: float) -> None:
        """Fold images [first_img, first_img + B) into the tables (asynchronous; the launch may be deferred until enough
        batches are queued -- ``join`` / ``finalize`` / reading a table flushes)."""
        shape = logits.shape
        b = shape[0]
        if first_img < 0 or first_img + b > self.n_img:
            raise RuntimeError(f"batch [{first_img},{first_img + b}) outside the shard of {self.n_img} images")
        if shape[1] != self.channels:
            raise RuntimeError(f"expected {self.channels} channels, got {shape[1]}")
        if not logits.is_cuda or not spx.is_cuda:
            raise RuntimeError("add_batch: expected CUDA tensors (there is no CPU path)")
        if not temperature > 0.0:
            raise RuntimeError("add_batch: temperature must be > 0")
        # the launch may be deferred: everything the kernel entry would reject is rejected here, at the call that caused it
        if not 2 <= shape[1] <= _lib.MAS_MAX_CLASSES:
            raise RuntimeError(f"add_batch: channels={shape[1]} outside [2,{_lib.MAS_MAX_CLASSES}]")
        if logits.dim() != 4 or logits.dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError(f"logits: expected (B,C,H,W) float32/bfloat16, got {tuple(shape)} {logits.dtype}")
        if tuple(spx.shape) != (b, shape[2], shape[3]):
            raise RuntimeError(f"spx shape {tuple(spx.shape)} does not match logits {tuple(shape)}")
        if b * shape[1] > 0 and (logits.stride(3) != 1 or logits.stride(2) != shape[3] or logits.stride(1) != shape[2] * shape[3]
                                 or (b > 1 and logits.stride(0) < shape[1] * shape[2] * shape[3])):
            raise RuntimeError("logits: expected NCHW layout with contiguous planes")
        if spx.dtype != torch.int32:
            spx = spx.to(torch.int32)
        if not spx.is_contiguous():
            spx = spx.contiguous()
        if self._prob_sum is not None and self.pixels_per_image not in (None, shape[2] * shape[3]):
            # the class weights divide every image's probability sum by ONE pixel count (mas_class_weights_dev); the
            # reference's per-batch torch.mean would follow a shape change, so refuse instead of weighting wrongly
            raise RuntimeError(f"add_batch: images of {shape[2]}x{shape[3]} after images of {self.pixels_per_image} pixels: "
                               "the predclsbal selectors need one image size per pool")
        self.pixels_per_image = shape[2] * shape[3]
        queue = self._queue
        if queue:
            head = queue[0][0]
            if first_img != self._queue_first + self._queue_n or temperature != self._queue_temp \
                    or head.shape[1:] != shape[1:] or head.dtype != logits.dtype:
                self.flush()
                queue = self._queue
        if not queue:
            self._queue_first, self._queue_temp = first_img, temperature
        queue.append((logits, spx))
        self._queue_n += b
        self._queued_bytes += logits.numel() * logits.element_size()
        if self._queued_bytes >= self.group_bytes or len(queue) >= _lib.MAS_MAX_SEGMENTS:
            self.flush()


def _add_batch_lowres(self, first_img: int, logits_lo: torch.Tensor, spx: torch.Tensor, temperature: float) -> None:
    """Fold images [first_img, first_img + B) into the tables from the network head's LOW-RESOLUTION logits
    (B,C',h,w): the final ``F.interpolate(..., size=spx.shape[1:], mode='bilinear', align_corners=False)`` of the model
    (models/segmentation/utils.py:28-34) is evaluated inside the kernel, so the 16x larger tensor never exists
    (opt-in -- the caller hands over the head's output instead of ``net(images)``)."""
    b = logits_lo.shape[0]
    if first_img < 0 or first_img + b > self.n_img:
        raise RuntimeError(f"batch [{first_img},{first_img + b}) outside the shard of {self.n_img} images")
    if logits_lo.shape[1] != self.channels:
        raise RuntimeError(f"expected {self.channels} channels, got {logits_lo.shape[1]}")
    if spx.dtype != torch.int32:
        spx = spx.to(torch.int32)
    spx = spx.contiguous()
    pixels = spx.shape[1] * spx.shape[2]
    if self._prob_sum is not None and self.pixels_per_image not in (None, pixels):
        raise RuntimeError("add_batch_lowres: the predclsbal selectors need one image size per pool")
    self.pixels_per_image = pixels
    self.flush()                       # keeps the table rows in launch order with queued full-resolution batches
    tables = (self._cls_sum[first_img:first_img + b], self._cls_cnt[first_img:first_img + b],
              None if self._prob_sum is None else self._prob_sum[first_img:first_img + b])
    self.launches += 1
    if not self.lanes:
        ops.bvsb_segment_stats_lowres(logits_lo, spx, self.nseg, temperature, *tables)
        return
    lane = self.lanes[self._turn % len(self.lanes)]
    self._turn += 1
    lane.wait_stream(torch.cuda.current_stream(self._cls_sum.device))
    ops.bvsb_segment_stats_lowres(logits_lo, spx, self.nseg, temperature, *tables, stream=lane.cuda_stream)
    logits_lo.record_stream(lane)
    spx.record_stream(lane)
    self._dirty = True


RegionStats.add_batch_lowres = _add_batch_lowres


def finalize(stats: RegionStats, spec: SelectorSpec, coeff: float = 0.0, ref_batch: int = 1, group=None, shard_counts=None):
    """Selector epilogue -> (scores (n,S) f32, dominant (n,S) i32) on the device.

    With a process group the pool-wide quantities (class means, min/max, dominant-class histogram) are
    exchanged over it; each rank keeps the scores of its own shard.  ``shard_counts`` = images per rank when known
    (``dist.shard_sizes``): the gather of the class-probability sums then needs no size exchange / host sync.
    """
    stats.join()
    weight = None
    if spec.weighting == "predclsbal":
        if stats.prob_sum is None:
            raise RuntimeError("predclsbal weighting needs RegionStats(need_prob=True)")
        prob_all = mdist.all_gather_rows(stats.prob_sum, group, shard_counts)
        weight = ops.class_weights(prob_all.contiguous(), stats.pixels_per_image, ref_batch, coeff)
    score, _, dominant = ops.region_scores(stats.cls_sum, stats.cls_cnt, weight)
    minmax = None
    if spec.normalise:
        minmax = mdist.all_reduce_minmax(ops.minmax_nonzero(score), group)
    region_weight = None
    if spec.weighting == "clsbal":
        hist = mdist.all_reduce_sum(ops.dominant_hist(dominant, stats.channels), group)
        freq = hist / hist.sum()          # int64 / int64 -> float32, as my_bvsb_clsbal_v2.py:66
        region_weight = torch.exp(-freq).to(torch.float32).contiguous()
    if minmax is not None or spec.ban_ignore or region_weight is not None:
        ops.finalize_scores(score, dominant, minmax, stats.channels - 1 if spec.ban_ignore else -1, region_weight)
    return score, dominant
