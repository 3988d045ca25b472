"""Offline multi-hot label generation on the GPU.

Drop-in for the per-image body of the reference's ``RegionCityscapesTensor.__getitem__``
(``dataloader/region_cityscapes_tensor.py:33-84``) and for the accumulation loop of
``tools/label_assignment_tensor.py:50-67`` that writes ``multi_hot_cls.npy`` / ``sp_size.npy``.
"""
from __future__ import annotations

from typing import Iterable, Sequence, Tuple

import numpy as np
import torch

from . import ops


def superpixel_info(target: torch.Tensor, superpixel: torch.Tensor, preserving_labels: Sequence[int], nseg: int,
                    num_classes: int, trim_multihot_boundary: bool = False, trim_kernel_size: int = 3,
                    device="cuda") -> Tuple[torch.Tensor, torch.Tensor]:
    """``sample['superpixel_info']`` of the reference: (superpixel_cls (nseg, C+1) uint8, superpixel_size (nseg,) int32),
    both on the CPU like the reference's.  ``target``: encoded train ids (255 = ignore), ``superpixel``: id map."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("mulactseg_b200.label_assignment needs a CUDA device (there is no CPU path)")
    spx = superpixel.to(dev)
    if spx.dtype not in (torch.int32, torch.int64):
        spx = spx.long()
    keep = torch.zeros(nseg, dtype=torch.uint8)
    if len(preserving_labels):
        keep[torch.as_tensor(list(preserving_labels), dtype=torch.long)] = 1
    cls, size = ops.multihot_labels(spx.contiguous(), target.to(dev, torch.uint8).contiguous(), keep.to(dev), nseg, num_classes,
                                    trim_kernel_size if trim_multihot_boundary else 0)
    return cls.cpu(), size.cpu()


def assign_all(samples: Iterable[Tuple[torch.Tensor, torch.Tensor, Sequence[int]]], nseg: int, num_classes: int,
               trim_multihot_boundary: bool = False, trim_kernel_size: int = 3, device="cuda"):
    """tools/label_assignment_tensor.py:52-60: stack per-image results -> (multi_hot_cls (N,nseg,C+1) uint8,
    multi_hot_size (N,nseg) int64) ready for ``np.save``.  ``samples`` yields (target, superpixel, preserving_labels)."""
    cls_all, size_all = [], []
    for target, superpixel, ids in samples:
        c, s = superpixel_info(target, superpixel, ids, nseg, num_classes, trim_multihot_boundary, trim_kernel_size, device)
        cls_all.append(c.numpy())
        size_all.append(s.numpy().astype(np.int64))
    return np.stack(cls_all).astype("uint8"), np.stack(size_all).astype("int")


def dominant_target(target: torch.Tensor, superpixel: torch.Tensor, preserving_labels: Sequence[int], nseg: int,
                    num_classes: int, device="cuda") -> torch.Tensor:
    """The "dominant label assignment" block of ``RegionCityscapesDominantAll.__getitem__``
    (``dataloader/region_dataset.py:221-233``): (H,W) uint8 target in which every listed superpixel carries its most
    frequent non-ignore train id (ignore pixels stay 255); returned on the CPU like the reference's."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("mulactseg_b200.label_assignment needs a CUDA device (there is no CPU path)")
    spx = torch.as_tensor(superpixel).to(dev)
    if spx.dtype not in (torch.int32, torch.int64):
        spx = spx.long()
    keep = torch.zeros(nseg, dtype=torch.uint8)
    ids = [int(p) for p in preserving_labels if 0 <= int(p) < nseg]
    if ids:
        keep[torch.as_tensor(ids, dtype=torch.long)] = 1
    out = ops.dominant_labels(spx.contiguous(), torch.as_tensor(target).to(dev, torch.uint8).contiguous(), keep.to(dev), nseg,
                              num_classes)
    return out.cpu()
