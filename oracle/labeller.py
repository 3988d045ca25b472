"""CPU oracle: stage-2 prototype pseudo-labeller.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Restates, per image,
``ActiveTrainer.pseudo_label_generation`` of
``trainer/eval_save_cosplbl_prop.py:121-314`` (``only_multihot=True``: prototypes
only from multi-hot superpixels, :167-170) and of the shipped
``trainer/eval_save_cosplbl_prop_includeonehot.py:121-316``
(``only_multihot=False``, :172), plus the candidate-arg-max labeller
``top_pseudo_label_generation`` of ``trainer/eval_within_multihot.py:93-146``.

The reference computes a dense (nproto x HW') similarity matrix and a
(nvspx x HW') scatter_max over it, but only ever consumes the block of a
pixel's own superpixel (:213-230) and, in the propagation step, the blocks of
the 3x3-adjacent superpixels (:260-305); this restatement computes just those
blocks, superpixel by superpixel (python loops -> small cases only).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from scipy import ndimage


def neighbour_ids(spx_map: np.ndarray, sid: int) -> np.ndarray:
    """eval_save_cosplbl_prop.py:260-266 -- ids met inside the 3x3-dilated mask of ``sid`` (itself included)."""
    own = spx_map == sid
    ys, xs = np.nonzero(own)
    if ys.size == 0:
        return np.empty(0, dtype=spx_map.dtype)
    # the dilation only reaches one pixel beyond the superpixel's bounding box: work on that window (same result as
    # dilating the full-size mask, without a pass over the whole image per superpixel)
    y0, y1 = max(int(ys.min()) - 1, 0), min(int(ys.max()) + 2, spx_map.shape[0])
    x0, x1 = max(int(xs.min()) - 1, 0), min(int(xs.max()) + 2, spx_map.shape[1])
    grown = ndimage.binary_dilation(own[y0:y1, x0:x1], structure=np.ones((3, 3), np.uint8))
    return np.unique(spx_map[y0:y1, x0:x1][grown])


def pseudo_label_image(feat: torch.Tensor, logits: torch.Tensor, target: torch.Tensor,
                       spmask: torch.Tensor, spx: torch.Tensor, only_multihot: bool = False,
                       threshold: str = "median") -> torch.Tensor:
    """One image.  feat (Ch,H,W) f32, logits (C',H,W) f32, target (S,C') u8, spmask (H,W) bool,
    spx (H,W) i64 -> (H,W) i64 labels, 255 where nothing was assigned."""
    ch, h, w = feat.shape
    c = logits.shape[0]
    prob = F.softmax(logits, dim=0).reshape(c, -1).T          # (HW, C'), T = 1 (:140)
    fpix = feat.reshape(ch, -1).T                              # (HW, Ch)
    ids = spx.reshape(-1)
    out = torch.full((h * w,), 255, dtype=torch.long)
    valid = spmask.reshape(-1).clone()
    if not torch.any(valid):
        return out.view(h, w)
    if only_multihot:
        valid &= (target.sum(dim=1) > 1)[ids.clamp(max=target.shape[0] - 1)] & (ids < target.shape[0])
        if not torch.any(valid):
            return out.view(h, w)
    vpix = valid.nonzero().squeeze(1)                          # ascending pixel index
    vids = ids[vpix]
    spx_np = spx.numpy()

    own_label = torch.empty(vpix.shape[0], dtype=torch.long)
    protos, proto_cls, thresholds, order = {}, {}, {}, []
    for s in torch.unique(vids).tolist():                      # ascending superpixel id
        where = (vids == s).nonzero().squeeze(1)
        pix = vpix[where]
        cls = target[s].nonzero().squeeze(1)                   # ascending candidate classes
        if cls.numel() == 0:
            raise ValueError(f"valid superpixel {s} has no candidate class (reference raises at :226)")
        p = prob[pix][:, cls]                                  # (n_s, K)
        best = p.max(dim=0).values
        first = torch.stack([(p[:, k] == best[k]).nonzero()[0, 0] for k in range(cls.numel())])
        proto = fpix[pix[first]]                               # (K, Ch) feature at the arg-max-prob pixel
        sim = proto @ fpix[pix].T                              # (K, n_s)
        smax = sim.max(dim=0).values
        # first prototype index attaining the maximum, per pixel (vectorised `(sim[:, j] == smax[j]).nonzero()[0, 0]`)
        rows = torch.arange(cls.numel()).view(-1, 1).expand_as(sim)
        kstar = torch.where(sim == smax[None, :], rows, torch.full_like(rows, cls.numel())).min(dim=0).values
        own_label[where] = cls[kstar]
        thr = torch.ones(cls.numel())
        for k in range(cls.numel()):
            mine = smax[kstar == k]
            if mine.numel():
                thr[k] = torch.median(mine) if threshold == "median" else torch.min(mine)
        protos[s], proto_cls[s], thresholds[s] = proto, cls, thr
        order.append(s)

    for s in order:                                            # later superpixels overwrite earlier (:276-305)
        near = torch.from_numpy(neighbour_ids(spx_np, s))
        q = torch.isin(ids, near).nonzero().squeeze(1)
        sim = protos[s] @ fpix[q].T                            # (K, |Q|)
        k = sim.argmax(dim=0)
        ok = torch.any(thresholds[s][:, None] < sim, dim=0)
        out[q[ok]] = proto_cls[s][k[ok]]

    out[vpix] = own_label                                      # (:309-310)
    return out.view(h, w)


def pseudo_label_generation(feats, inputs, targets, spmasks, superpixels, only_multihot=False,
                            threshold="median") -> torch.Tensor:
    """Batch wrapper with the reference's argument order minus ``labels`` (used only for its shape)."""
    return torch.stack([
        pseudo_label_image(feats[i], inputs[i], targets[i], spmasks[i], superpixels[i], only_multihot, threshold)
        for i in range(inputs.shape[0])])


def top_pseudo_label_generation(inputs, targets, spmasks, superpixels) -> torch.Tensor:
    """eval_within_multihot.py:93-146 -- arg-max of (logit * multi-hot row) on masked pixels, 255 elsewhere."""
    n, c, h, w = inputs.shape
    out = torch.full((n, h * w), 255, dtype=torch.long)
    for i in range(n):
        keep = spmasks[i].reshape(-1)
        if not torch.any(keep):
            continue
        x = inputs[i].reshape(c, -1).T[keep]
        rows = targets[i][superpixels[i].reshape(-1)[keep]]
        out[i, keep.nonzero().squeeze(1)] = (x * rows).max(dim=1)[1]
    return out.view(n, h, w)
