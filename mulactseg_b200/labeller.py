"""Drop-in stage-2 pseudo-label generation (same signatures as the reference's ``ActiveTrainer`` methods).

Reference (file:line relative to the reference checkout):
  ``pseudo_label_generation(labels, feats, inputs, targets, spmasks, superpixels) -> LongTensor (N,H,W)``
      trainer/eval_save_cosplbl_prop.py:121-314                (``only_multihot=True``)
      trainer/eval_save_cosplbl_prop_includeonehot.py:121-316  (``only_multihot=False``; the shipped recipe)
  ``top_pseudo_label_generation(labels, inputs, targets, spmasks, superpixels)``
      trainer/eval_within_multihot.py:93-146
``ProtoLabellerMixin`` / ``TopLabellerMixin`` give a trainer class these methods with the reference's argument
lists; ``args.cosprop_threshold_method`` ('median' | 'min') and ``args.nseg`` are read like the reference does.
The VOC multi-scale variant (``..._includeonehot_voc_ms.py``) calls the same function on fused features
(``trainer/eval_save_cosplbl_prop_includeonehot_voc_ms.fuse_multiscale``).
``feats`` may be fp32 or bf16, and either the reference's full-resolution (N,F,H,W) map or -- opt-in -- the head's
low-resolution (N,F,h,w) map, which the kernels interpolate on the fly.
"""
from __future__ import annotations

import os

import torch

from . import _lib, ops


def _prep(targets, spmasks, superpixels):
    trg = (targets if targets.dtype == torch.uint8 else targets.to(torch.uint8)).contiguous()
    mask = (spmasks if spmasks.dtype in (torch.bool, torch.uint8) else spmasks.bool()).contiguous()
    spx = (superpixels if superpixels.dtype in (torch.int32, torch.int64) else superpixels.long()).contiguous()
    return trg, mask, spx


_LANES = max(1, min(16, int(os.environ.get("MAS_LABELLER_LANES", "8"))))      # images of a batch in flight at once (side streams)
_side_streams = {}


def _lanes(device, n):
    """Side streams for the images of a batch.  The per-image pipeline is nine short, dependent launches (a VOC image:
    ~0.2 ms of mostly launch latency and tails), and the images are independent, so up to ``_LANES`` of them run
    concurrently (each on its own workspace slice); the labels are the ones the reference's sequential loop produces."""
    index = device.index if device.index is not None else torch.cuda.current_device()
    lanes = _side_streams.get(index)
    if lanes is None:
        lanes = _side_streams[index] = [torch.cuda.Stream(device=device) for _ in range(_LANES)]
    return lanes[:min(n, _LANES)]


def pseudo_label_generation(labels, feats, inputs, targets, spmasks, superpixels, only_multihot: bool = False,
                            threshold: str = "median", check: bool = True) -> torch.Tensor:
    """(N,H,W) int64 pseudo labels, 255 = unlabeled.  ``labels`` is only used for its shape, as in the reference.
    ``check`` (one sync per call, the caller reads the result back anyway): raise like the reference (:226) when a
    selected superpixel has no candidate class."""
    if not feats.is_cuda:
        raise RuntimeError("mulactseg_b200 labellers need CUDA tensors (there is no CPU path)")
    n = inputs.shape[0]
    trg, mask, spx = _prep(targets, spmasks, superpixels)
    # fp32 or bf16 features are read as they are; full resolution (the reference's x4 up-sampled map) or the head's
    # low-resolution map (interpolated inside the kernels, ``mas_proto_labeller_src_dev``)
    feats = feats.contiguous() if feats.dtype in (torch.float32, torch.bfloat16) else feats.contiguous().float()
    inputs = inputs.contiguous().float()
    # one library call per loader batch: the images are dealt to side streams inside (mas_proto_labeller_batch_dev)
    lab, stats = ops.proto_labeller_batch(feats, inputs, trg, mask, spx, only_multihot, threshold, _lanes(feats.device, n))
    out = lab.long()
    if check:
        bad = int(stats.sum())
        if bad:
            raise RuntimeError(f"{bad} selected pixels belong to superpixels without any candidate class "
                               "(the reference fails on such input, eval_save_cosplbl_prop.py:226)")
    return out


def top_pseudo_label_generation(labels, inputs, targets, spmasks, superpixels) -> torch.Tensor:
    """eval_within_multihot.py:93-146: arg-max of (logit * multi-hot row) on selected pixels, 255 elsewhere."""
    if not inputs.is_cuda:
        raise RuntimeError("mulactseg_b200 labellers need CUDA tensors (there is no CPU path)")
    trg, mask, spx = _prep(targets, spmasks, superpixels)
    info = ops.multihot_info(trg, inputs.shape[1], _lib.MAS_GROUP_ALL)
    return ops.candidate_argmax(inputs.contiguous().float(), spx, mask, info, trg.shape[1]).long()


class ProtoLabellerMixin:
    """``ActiveTrainer.pseudo_label_generation`` of eval_save_cosplbl_prop*.py for a trainer with ``self.args``."""
    only_multihot = False

    def pseudo_label_generation(self, labels, feats, inputs, targets, spmasks, superpixels):
        return pseudo_label_generation(labels, feats, inputs, targets, spmasks, superpixels, self.only_multihot,
                                       getattr(self.args, "cosprop_threshold_method", "median"))


class TopLabellerMixin:
    def top_pseudo_label_generation(self, labels, inputs, targets, spmasks, superpixels):
        return top_pseudo_label_generation(labels, inputs, targets, spmasks, superpixels)
