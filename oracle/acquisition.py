"""CPU oracle: BvSB acquisition selectors and region selection.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Torch-CPU restatement of the
reference's acquisition pass, op for op (softmax -> topk -> ratio ->
segment-mean via scatter_add; one_hot -> scatter sum for the histogram), so that
it is both the parity checker and a fair stand-in for the reference's CPU cost.

A "pool" here is a list of batches ``(preds (B,C',H,W) f32, spx (B,H,W) i64)``
exactly as the reference's loader would hand them over (the last batch may be
short -- that matters for the mean-of-batch-means in the ``predclsbal`` pass 1).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple

import torch
import torch.nn.functional as F

from .scatter_ref import scatter

Batch = Tuple[torch.Tensor, torch.Tensor]


def softmax_bvsb(preds: torch.Tensor, temperature: float):
    """active_selection/my_bvsb.py:19-27 -- second-best / best softmax prob + 1e-8, and the top-1 class."""
    prob = F.softmax(preds / temperature, dim=1)
    val, idx = torch.topk(prob, 2, dim=1)
    ratio = val[:, 1] / val[:, 0]
    ratio += 1e-8
    return ratio, idx[:, 0]


def _segment_mean(values: torch.Tensor, spx: torch.Tensor, nseg: int) -> torch.Tensor:
    """my_bvsb.py:70-73 -- (B,H,W) values -> (B,nseg) mean, 0 for absent ids."""
    b = values.shape[0]
    return scatter(values.reshape(b, -1), spx.reshape(b, -1), dim=1, reduce="mean", dim_size=nseg)


def _segment_hist(top1: torch.Tensor, spx: torch.Tensor, nseg: int, nbins: int) -> torch.Tensor:
    """my_bvsb_banignore.py:44-45 -- per-superpixel histogram of the arg-max class, int64 (B,nseg,nbins)."""
    b = top1.shape[0]
    onehot = F.one_hot(top1.reshape(b, -1), num_classes=nbins)
    return scatter(onehot, spx.reshape(b, -1), dim=1, reduce="sum", dim_size=nseg)


def _minmax_normalise(flat: torch.Tensor) -> torch.Tensor:
    """my_bvsb.py:79-81 -- subtract the min over non-zero entries, divide by the max."""
    flat = flat - flat[flat != 0].min()
    return flat / flat.max()


def scores_my_bvsb(pool: Iterable[Batch], nseg: int, temperature: float, predignore: bool) -> torch.Tensor:
    """active_selection/my_bvsb.py:50-87 (the ignore channel is sliced off for predignore nets, :65-66)."""
    rows = []
    for preds, spx in pool:
        if predignore:
            preds = preds[:, :-1]
        bvsb, _ = softmax_bvsb(preds, temperature)
        rows.append(_segment_mean(bvsb, spx, nseg))
    u = torch.cat(rows, dim=0).view(-1)
    return _minmax_normalise(u).view(-1, nseg)


def scores_my_bvsb_banignore(pool: Iterable[Batch], nseg: int, temperature: float) -> torch.Tensor:
    """active_selection/my_bvsb_banignore.py:19-67 -- all C+1 channels kept; ignore-dominant regions -> 0."""
    rows, hists = [], []
    for preds, spx in pool:
        bvsb, top1 = softmax_bvsb(preds, temperature)
        rows.append(_segment_mean(bvsb, spx, nseg))
        hists.append(_segment_hist(top1, spx, nseg, preds.shape[1]))
    u = _minmax_normalise(torch.cat(rows, dim=0).view(-1))
    hist = torch.cat(hists, dim=0)
    hist = hist.view(-1, hist.shape[-1])
    u[hist.argmax(dim=1) == hist.shape[1] - 1] = 0
    return u.view(-1, nseg)


def predicted_class_weights(pool: Iterable[Batch], temperature: float, coeff: float) -> torch.Tensor:
    """my_bvsb_predclsbal_pwr.py:36-47 -- mean of per-BATCH mean softmax probs -> (coeff*p+1)^-2."""
    acc, nb = None, 0
    for preds, _ in pool:
        p = torch.softmax(preds / temperature, dim=1).mean(dim=(0, 2, 3))
        acc = p if acc is None else acc + p
        nb += 1
    return (coeff * (acc / nb) + 1) ** (-2)


def scores_predclsbal_pwr(pool: Sequence[Batch], nseg: int, temperature: float, coeff: float,
                          ban_ignore: bool) -> torch.Tensor:
    """my_bvsb_predclsbal_pwr.py:23-88 / my_bvsb_predclsbal_pwr_banignore.py:22-91.

    Two passes over the pool: class weights from the pool-wide mean softmax,
    then segment-mean of bvsb * w[top1].  No min/max normalisation.  The
    ``_banignore`` twin zeroes regions whose dominant arg-max class is the last
    (ignore) channel (:78-84).
    """
    w = predicted_class_weights(pool, temperature, coeff)
    rows, hists = [], []
    for preds, spx in pool:
        bvsb, top1 = softmax_bvsb(preds, temperature)
        rows.append(_segment_mean(bvsb * w[top1], spx, nseg))
        hists.append(_segment_hist(top1, spx, nseg, preds.shape[1]))
    u = torch.cat(rows, dim=0).view(-1)
    if ban_ignore:
        hist = torch.cat(hists, dim=0)
        hist = hist.view(-1, hist.shape[-1])
        u[hist.argmax(dim=1) == hist.shape[1] - 1] = 0
    return u.view(-1, nseg)


def scores_clsbal_v2(pool: Iterable[Batch], nseg: int, temperature: float, ban_ignore: bool) -> torch.Tensor:
    """my_bvsb_clsbal_v2.py:21-73 / my_bvsb_clsbal_v2_banignore.py:21-76 -- exp(-freq[dominant]) * normalised bvsb."""
    rows, hists = [], []
    for preds, spx in pool:
        bvsb, top1 = softmax_bvsb(preds, temperature)
        rows.append(_segment_mean(bvsb, spx, nseg))
        hists.append(_segment_hist(top1, spx, nseg, preds.shape[1]))
    u = _minmax_normalise(torch.cat(rows, dim=0).view(-1))
    hist = torch.cat(hists, dim=0)
    nbins = hist.shape[-1]
    dominant = hist.view(-1, nbins).argmax(dim=1)
    if ban_ignore:
        u[dominant == nbins - 1] = 0
    dom_oh = F.one_hot(dominant, num_classes=nbins)
    freq = dom_oh.sum(dim=0) / dom_oh.sum()
    return (torch.exp(-freq)[dominant] * u).view(-1, nseg)


def region_histograms(pool: Iterable[Batch], nseg: int, temperature: float) -> torch.Tensor:
    """The int64 (N,nseg,C') arg-max histogram on its own (bit-exact parity target).

    The arg-max is taken on the softmax PROBABILITIES like the reference does
    (my_bvsb.py:20-21), so it is only defined where the top-2 probabilities are
    distinct floats; fixtures avoid exact ties.
    """
    return torch.cat([_segment_hist(softmax_bvsb(p, temperature)[1], s, nseg, p.shape[1]) for p, s in pool], dim=0)


def score_list(im_idx: Sequence[Sequence[str]], suppix: Dict[str, List[int]],
               scores_tensor: torch.Tensor) -> List[Tuple[float, str, int]]:
    """my_bvsb.py:29-48 -- (score, 'img,lbl,spx', id) for every id still in the pool."""
    out = []
    for k, key in enumerate(im_idx):
        ids = suppix[key[2]]
        vals = scores_tensor[k][ids].tolist()
        path = ",".join(key)
        out.extend((v, path, i) for v, i in zip(vals, ids))
    return out


def rank_regions(scores: List[Tuple[float, str, int]]) -> List[Tuple[float, str, int]]:
    """active_selection/base.py:37 -- descending tuple order (score, then path, then id)."""
    return sorted(scores, reverse=True)


def expand_training_set(ranked: List[Tuple[float, str, int]], budget: int,
                        label_im_idx: List[List[str]], label_suppix: Dict[str, List[int]],
                        pool_im_idx: List[List[str]], pool_suppix: Dict[str, List[int]],
                        region_cost=None) -> int:
    """dataloader/region_active_dataset.py:16-73 -- walk the ranked list moving ids pool -> label.

    ``region_cost(spx_path, id)`` is the multi-hot class count used with
    ``--fair_counting --or_labeling`` (:56-61); ``None`` = one unit per region.
    Stops AFTER the pick that makes the cumulative cost exceed ``budget`` (strict
    ``>``, :66).  Mutates the four containers in place like the reference and
    returns the length of the selected prefix (what ``*_selection_XX.pkl`` holds).
    """
    spent = 0
    for n, (_, joined, sid) in enumerate(ranked):
        key = joined.split(",")
        spx_path = key[2]
        if key not in label_im_idx:
            label_im_idx.append(key)
            label_suppix[spx_path] = [sid]
        else:
            label_suppix[spx_path].append(sid)
        pool_suppix[spx_path].remove(sid)
        if not pool_suppix[spx_path]:
            pool_suppix.pop(spx_path)
            pool_im_idx.remove(key)
        spent += 1 if region_cost is None else int(region_cost(spx_path, sid))
        if spent > budget:
            return n + 1
    return len(ranked)
