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
    """

    def __init__(self, n_img: int, nseg: int, channels: int, device, need_prob: bool):
        self.n_img, self.nseg, self.channels = int(n_img), int(nseg), int(channels)
        self.cls_sum = torch.zeros((n_img, nseg, channels), dtype=torch.float32, device=device)
        self.cls_cnt = torch.zeros((n_img, nseg, channels), dtype=torch.int32, device=device)
        self.prob_sum = torch.zeros((n_img, channels), dtype=torch.float64, device=device) if need_prob else None
        self.pixels_per_image: Optional[int] = None

    def zero_(self):
        self.cls_sum.zero_()
        self.cls_cnt.zero_()
        if self.prob_sum is not None:
            self.prob_sum.zero_()

    def add_batch(self, first_img: int, logits: torch.Tensor, spx: torch.Tensor, temperature: float) -> None:
        """Fold images [first_img, first_img + B) into the tables (asynchronous)."""
        b = logits.shape[0]
        if first_img < 0 or first_img + b > self.n_img:
            raise RuntimeError(f"batch [{first_img},{first_img + b}) outside the shard of {self.n_img} images")
        if logits.shape[1] != self.channels:
            raise RuntimeError(f"expected {self.channels} channels, got {logits.shape[1]}")
        if spx.dtype != torch.int32:
            spx = spx.to(torch.int32)
        self.pixels_per_image = logits.shape[2] * logits.shape[3]
        ops.bvsb_segment_stats(logits, spx.contiguous(), self.nseg, temperature,
                               self.cls_sum[first_img:first_img + b], self.cls_cnt[first_img:first_img + b],
                               None if self.prob_sum is None else self.prob_sum[first_img:first_img + b])


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


def finalize(stats: RegionStats, spec: SelectorSpec, coeff: float = 0.0, ref_batch: int = 1, group=None):
    """Selector epilogue -> (scores (n,S) f32, dominant (n,S) i32) on the device.

    With a process group the pool-wide quantities (class means, min/max, dominant-class histogram) are
    exchanged over it; each rank keeps the scores of its own shard.
    """
    weight = None
    if spec.weighting == "predclsbal":
        if stats.prob_sum is None:
            raise RuntimeError("predclsbal weighting needs RegionStats(need_prob=True)")
        prob_all = mdist.all_gather_rows(stats.prob_sum, group)
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
