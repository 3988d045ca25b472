"""Seeded synthetic inputs shaped like the reference's tensors.

Used by tests, ``bench.py`` and ``oracle/gen_golden.py``.  Everything takes a
``torch.Generator`` (or a seed) and a device so that the big benchmark shapes
can be produced directly in HBM.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch


def _gen(seed_or_gen, device="cpu") -> torch.Generator:
    if isinstance(seed_or_gen, torch.Generator):
        return seed_or_gen
    g = torch.Generator(device=device)
    g.manual_seed(int(seed_or_gen))
    return g


def grid_shape(nseg: int, h: int, w: int) -> Tuple[int, int]:
    """Rows x cols of a superpixel grid with rows*cols == nseg and cells as square as possible."""
    best = (1, nseg)
    target = math.sqrt(nseg * h / w)
    for gy in range(1, nseg + 1):
        if nseg % gy == 0 and abs(gy - target) < abs(best[0] - target):
            best = (gy, nseg // gy)
    return best


def superpixel_map(n: int, h: int, w: int, nseg: int, kind: str = "jitter", seed=0,
                   device="cpu", drop_ids: int = 0, dtype=torch.int64) -> torch.Tensor:
    """(n,h,w) superpixel ids in [0,nseg).

    ``grid``   : compact cells like SEEDS output;
    ``jitter`` : grid whose cell boundaries wander by a few pixels per row/column;
    ``random`` : i.i.d. ids per pixel (adversarial, correctness only).
    ``drop_ids`` ids per image are merged into a neighbour so that they are absent
    (the real Cityscapes dict has images with a missing id).
    """
    g = _gen(seed, device)
    if kind == "random":
        spx = torch.randint(0, nseg, (n, h, w), generator=g, device=device)
    else:
        gy, gx = grid_shape(nseg, h, w)
        yy = torch.arange(h, device=device).view(1, h, 1).expand(n, h, w).float()
        xx = torch.arange(w, device=device).view(1, 1, w).expand(n, h, w).float()
        if kind == "jitter":
            amp_y, amp_x = min(4.0, h / gy / 4), min(4.0, w / gx / 4)
            phase = torch.rand((n, 4), generator=g, device=device) * 6.283
            yy = yy + amp_y * torch.sin(xx * (6.283 / max(w / gx, 2.0)) * 0.37 + phase[:, 0].view(n, 1, 1))
            xx = xx + amp_x * torch.sin(yy * (6.283 / max(h / gy, 2.0)) * 0.41 + phase[:, 1].view(n, 1, 1))
        cy = (yy * gy / h).floor().clamp_(0, gy - 1).long()
        cx = (xx * gx / w).floor().clamp_(0, gx - 1).long()
        spx = cy * gx + cx
    for _ in range(drop_ids):
        victim = torch.randint(1, nseg, (n,), generator=g, device=device).view(n, 1, 1)
        spx = torch.where(spx == victim, victim - 1, spx)
    return spx.to(dtype)


def logits(n: int, c: int, h: int, w: int, kind: str = "cosine", seed=0, device="cpu",
           dtype=torch.float32, coherent: int = 0) -> torch.Tensor:
    """(n,c,h,w) logits.

    ``cosine`` mimics the reference's cosine-classifier head (values in (-1,1), used with T=0.1);
    ``normal`` is N(0,1) (used with T=1).  ``coherent=k`` draws the noise at 1/k resolution and
    up-samples it bilinearly like the model's x4 up-sampling (``models/segmentation/utils.py:28-34``).
    """
    g = _gen(seed, device)
    if coherent and coherent > 1:
        z = torch.randn((n, c, max(h // coherent, 2), max(w // coherent, 2)), generator=g, device=device)
        z = torch.nn.functional.interpolate(z, size=(h, w), mode="bilinear", align_corners=False)
    else:
        z = torch.randn((n, c, h, w), generator=g, device=device)
    if kind == "cosine":
        z = torch.tanh(z) * 0.9
    return z.to(dtype).contiguous()


def multihot_targets(n: int, nseg: int, c: int, seed=0, device="cpu", p_extra: float = 0.08,
                     p_ignore: float = 0.05, ignore_channel: bool = True) -> torch.Tensor:
    """(n,nseg,c) uint8: one dominant class per region + Bernoulli extras (+ ignore channel)."""
    g = _gen(seed, device)
    ncls = c - 1 if ignore_channel else c
    dom = torch.randint(0, ncls, (n, nseg), generator=g, device=device)
    t = (torch.rand((n, nseg, c), generator=g, device=device) < p_extra)
    if ignore_channel:
        t[..., -1] = torch.rand((n, nseg), generator=g, device=device) < p_ignore
    t.scatter_(2, dom.unsqueeze(-1), True)
    return t.to(torch.uint8)


def region_mask(spx: torch.Tensor, nseg: int, rho: float, seed=0) -> torch.Tensor:
    """(n,h,w) bool: superpixel-level Bernoulli(rho) selection; ids >= nseg (crop padding) are never selected."""
    g = _gen(seed, spx.device)
    n = spx.shape[0]
    chosen = torch.rand((n, nseg + 1), generator=g, device=spx.device) < rho
    chosen[:, nseg] = False
    return torch.gather(chosen, 1, spx.reshape(n, -1).clamp(max=nseg).long()).view(spx.shape)


def pad_border(spx: torch.Tensor, nseg: int, pad: int) -> torch.Tensor:
    """Mimic the crop padding of the train transform: a border carrying id == nseg."""
    if pad <= 0:
        return spx
    out = spx.clone()
    out[:, :pad, :] = nseg
    out[:, :, -pad:] = nseg
    return out


def features(n: int, ch: int, h: int, w: int, seed=0, device="cpu", stride: int = 4) -> torch.Tensor:
    """(n,ch,h,w): L2-normalised low-res features bilinearly up-sampled (deeplabv3.py:122, utils.py:32)."""
    g = _gen(seed, device)
    low = torch.randn((n, ch, max(h // stride, 2), max(w // stride, 2)), generator=g, device=device)
    low = torch.nn.functional.normalize(low, dim=1)
    return torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False).contiguous()


def pool_lists(n: int, nseg: int, spx: Optional[torch.Tensor] = None, labelled_frac: float = 0.0, seed=0):
    """``im_idx`` / ``suppix`` like the reference pool dataset (region_active_dataset.py:82-91)."""
    g = _gen(seed)
    im_idx = [[f"img{i:05d}.png", f"lbl{i:05d}.png", f"spx{i:05d}.png"] for i in range(n)]
    suppix = {}
    for i in range(n):
        ids = torch.unique(spx[i]).tolist() if spx is not None else list(range(nseg))
        ids = [int(s) for s in ids if s < nseg]
        if labelled_frac > 0:
            keep = torch.rand(len(ids), generator=g) >= labelled_frac
            ids = [s for s, k in zip(ids, keep.tolist()) if k]
        suppix[im_idx[i][2]] = ids
    return im_idx, suppix
