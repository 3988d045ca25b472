"""CPU oracle: multi-hot partial-label and MIL (merged-positive) losses.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Torch-CPU restatement, with
autograd, of the stage-1 losses the shipped recipes use.  Shapes follow the
reference: ``inputs (N,C',H,W) f32``, ``targets (N,S,Ct) u8``, ``superpixels
(N,H,W) i64`` (crop padding carries id == S and is always masked out),
``spmasks (N,H,W) bool``.

Variants and the reference classes they follow:

group loss ``variant=``
  ``"base"``       utils/loss.py:81-141  GroupMultiLabelCE  (targets[..., :-1])
  ``"predignore"`` trainer/active_joint_multi_predignore.py:74-128  GroupMultiLabelCE_
  ``"onlymulti"``  trainer/active_joint_multi_predignore_mclossablation2.py:17-79
multi-choice loss ``variant=``
  ``"base"``       utils/loss.py:535-588  MultiChoiceCE  (targets[..., :-1], empty rows dropped)
  ``"predignore"`` trainer/active_joint_multi_predignore.py:17-73  MultiChoiceCE_
one-hot / multi-hot decomposition
  trainer/active_joint_multi_predignore_lossdecomp.py:16-72 (``strict_multihot=False``:
  multi-hot := not one-hot) and trainer/active_joint_multi_lossdecomp.py:17-74
  (``strict_multihot=True``: multi-hot := row-sum > 1).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .scatter_ref import scatter

EPS = 1e-8


def _pixel_major(inputs: torch.Tensor, temperature: float) -> torch.Tensor:
    n, c = inputs.shape[:2]
    return F.softmax(inputs / temperature, dim=1).permute(0, 2, 3, 1).reshape(n, -1, c)


def group_multilabel_ce(inputs, targets, superpixels, spmasks, nseg: int, temperature: float = 1.0,
                        variant: str = "onlymulti", reduction: str = "mean"):
    """Per-(superpixel, class) max-pooled probability, -log over the labelled classes.

    ``num_valid`` starts at 1 and counts only NON-ZERO pooled values
    (``nonzero`` at loss.py:133 / predignore.py:120 / mclossablation2.py:71).
    """
    n = inputs.shape[0]
    prob = _pixel_major(inputs, temperature)
    spx = superpixels.reshape(n, -1, 1)
    mask = spmasks.reshape(n, -1)
    trg = targets[..., :-1] if variant == "base" else targets
    has_label = torch.any(trg, dim=2).bool()
    is_multi = targets.sum(dim=2) > 1
    loss, count = 0, 1
    for i in range(n):
        keep = mask[i]
        if not torch.any(keep):
            continue
        if variant == "onlymulti":
            inner = is_multi[i][spx[i].squeeze(1)[mask[i]]]
            keep = mask[i].clone()
            keep[mask[i]] = inner
            if not torch.any(keep):
                continue
        pooled = scatter(prob[i][keep], spx[i][keep], dim=0, reduce="max", dim_size=nseg)
        picked = pooled[has_label[i]] * trg[i][has_label[i]]
        vals = picked[picked.nonzero(as_tuple=True)]
        count += vals.shape[0]
        loss = loss + (-torch.log(vals + EPS)).sum()
    if reduction == "mean":
        return loss / count
    return loss, count


def multi_choice_ce(inputs, targets, superpixels, spmasks, temperature: float = 1.0,
                    variant: str = "predignore"):
    """-log of the probability mass on the candidate set, averaged over labelled masked pixels (+1)."""
    n = inputs.shape[0]
    prob = _pixel_major(inputs, temperature)
    spx = superpixels.reshape(n, -1)
    mask = spmasks.reshape(n, -1)
    trg = targets[..., :-1] if variant == "base" else targets
    loss, count = 0, 1
    for i in range(n):
        keep = mask[i]
        if not torch.any(keep):
            continue
        rows = trg[i][spx[i][keep]]
        nonempty = torch.any(rows, dim=1).bool()
        pos = (prob[i][keep][nonempty] * rows[nonempty]).sum(dim=1)
        count += pos.shape[0]
        loss = loss + (-torch.log(pos + EPS)).sum()
    return loss / count


def onehot_ce_multihot_choice(inputs, targets, superpixels, spmasks, temperature: float = 1.0,
                              strict_multihot: bool = False):
    """(CE over one-hot regions, multi-choice over multi-hot regions); both counters start at 1."""
    n = inputs.shape[0]
    prob = _pixel_major(inputs, temperature)
    spx = superpixels.reshape(n, -1)
    mask = spmasks.reshape(n, -1)
    oh_loss, oh_n, mh_loss, mh_n = 0, 1, 0, 1
    for i in range(n):
        keep = mask[i]
        if not torch.any(keep):
            continue
        rows = targets[i][spx[i][keep]]
        pos = (prob[i][keep] * rows).sum(dim=1)
        ncand = rows.sum(dim=1)
        one = ncand == 1
        multi = (ncand > 1) if strict_multihot else torch.logical_not(one)
        if torch.any(one):
            oh_loss = oh_loss + (-torch.log(pos[one] + EPS)).sum()
            oh_n += int(one.sum())
        if torch.any(multi):
            if not strict_multihot:
                assert torch.all(multi == (ncand > 1))  # predignore_lossdecomp.py:67
            mh_loss = mh_loss + (-torch.log(pos[multi] + EPS)).sum()
            mh_n += int(multi.sum())
    return oh_loss / oh_n, mh_loss / mh_n


def stage1_total(inputs, targets, superpixels, spmasks, nseg, t_group, t_multi,
                 coeff=16.0, coeff_mc=8.0, coeff_gm=1.0, strict_multihot=False):
    """trainer/active_joint_multi_predignore_lossdecomp.py:101-104 -- the scalar the trainer back-propagates."""
    group = group_multilabel_ce(inputs, targets, superpixels, spmasks, nseg, t_group, "onlymulti")
    ce, mc = onehot_ce_multihot_choice(inputs, targets, superpixels, spmasks, t_multi, strict_multihot)
    return coeff * ce + coeff_mc * mc + coeff_gm * group, (ce, mc, group)
