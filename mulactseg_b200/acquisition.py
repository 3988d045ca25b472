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

    Consecutive ``add_batch`` launches touch disjoint table rows and only add, so they carry no mutual
    dependency: with ``lanes`` > 1 they alternate over that many side streams (forked from / joined to the
    caller's stream with events).  The scorer is a single-wave grid holding a whole SM per CTA, so the CTAs of
    launch i+1 move onto SMs as the CTAs of launch i retire -- ramp and tail of the ~100 us launches overlap
    instead of adding a launch gap each.  ``lanes`` = 1 keeps everything on the caller's stream.
    """

    def __init__(self, n_img: int, nseg: int, channels: int, device, need_prob: bool, lanes: int = 2):
        self.n_img, self.nseg, self.channels = int(n_img), int(nseg), int(channels)
        self.lanes = [torch.cuda.Stream(device=device) for _ in range(lanes)] if lanes > 1 else []
        self._turn, self._dirty = 0, False
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
        """Make the caller's current stream wait for every outstanding ``add_batch`` launch."""
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

    def add_batch(self, first_img: int, logits: torch.Tensor, spx: torch.Tensor, temperature: float) -> None:
        """Fold images [first_img, first_img + B) into the tables (asynchronous)."""
        b = logits.shape[0]
        if first_img < 0 or first_img + b > self.n_img:
            raise RuntimeError(f"batch [{first_img},{first_img + b}) outside the shard of {self.n_img} images")
        if logits.shape[1] != self.channels:
            raise RuntimeError(f"expected {self.channels} channels, got {logits.shape[1]}")
        if not logits.is_cuda or not spx.is_cuda:
            raise RuntimeError("add_batch: expected CUDA tensors (there is no CPU path)")
        if spx.dtype != torch.int32:
            spx = spx.to(torch.int32)
        self.pixels_per_image = logits.shape[2] * logits.shape[3]
        spx = spx.contiguous()
        tables = (self._cls_sum[first_img:first_img + b], self._cls_cnt[first_img:first_img + b],
                  None if self._prob_sum is None else self._prob_sum[first_img:first_img + b])
        if not self.lanes:
            ops.bvsb_segment_stats(logits, spx, self.nseg, temperature, *tables)
            return
        lane = self.lanes[self._turn % len(self.lanes)]
        self._turn += 1
        lane.wait_stream(torch.cuda.current_stream(logits.device))   # inputs (and the zeroed tables) are ready
        with torch.cuda.stream(lane):
            ops.bvsb_segment_stats(logits, spx, self.nseg, temperature, *tables)
        logits.record_stream(lane)    # the caching allocator must not recycle them before the lane is done
        spx.record_stream(lane)
        self._dirty = True


def predicted_class_weights(prob_sum_all: torch.Tensor, pixels_per_image: int, ref_batch: int, coeff: float) -> torch.Tensor:
    """w_c = (coeff * pbar_c + 1)^-2 with pbar = mean over REFERENCE batches of the per-batch mean
    probability (my_bvsb_predclsbal_pwr.py:36-47: ``cumulated += mean(prob, dim=(0,2,3))`` per batch,
    divided by ``len(loader)``).  ``prob_sum_all`` (N,C) f64 holds per-image sums in pool order."""
    n, c = prob_sum_all.shape
    ref_batch = max(int(ref_batch), 1)
    n_batches = (n + ref_batch - 1) // ref_batch
    batch_of = torch.arange(n, device=prob_sum_all.device) // ref_batch
    batch_sum = torch.zeros((n_batches, c), dtype=torch.float64, device=prob_sum_all.device)
    batch_sum.index_add_(0, batch_of, prob_sum_all)
    batch_len = torch.bincount(batch_of, minlength=n_batches).to(torch.float64)
    batch_mean = (batch_sum / (batch_len * float(pixels_per_image)).unsqueeze(1)).to(torch.float32)
    pbar = batch_mean.sum(dim=0) / float(n_batches)
    return (float(coeff) * pbar + 1.0) ** (-2)


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
        if prob_all.is_cuda:
            weight = ops.class_weights(prob_all.contiguous(), stats.pixels_per_image, ref_batch, coeff)
        else:       # gloo / CPU tensors in the host-logic tests
            weight = predicted_class_weights(prob_all, stats.pixels_per_image, ref_batch, coeff).contiguous()
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
