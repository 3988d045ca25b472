"""Hot-path part of the reference's ``trainer/eval_save_cosplbl_prop_includeonehot_voc_ms.py`` (VOC multi-scale stage 2).

``pseudo_label_generation`` (:152-354) is the same function as in ``..._includeonehot.py`` (the reference files differ
only in ``inference()``, :56-79, and a debugging try/except at :263-266), so the same kernels serve it.  What the
multi-scale ``inference()`` does to the tensors BEFORE calling it is reproduced by ``fuse_multiscale`` with torch ops
(it is part of the caller, like the network forward): the second half of the scale list is flipped back, every scale is
resized to the image size (bilinear), the scales are averaged and the features re-normalised along the channels.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
import torch.nn.functional as F

from ..labeller import ProtoLabellerMixin


def fuse_multiscale(feat_list: Sequence[torch.Tensor], output_list: Sequence[torch.Tensor],
                    im_size: Tuple[int, int]) -> Tuple[torch.Tensor, torch.Tensor]:
    """(1,F,H,W) features and (1,C',H,W) logits from per-scale ``feat_forward`` results (each (1,·,h_k,w_k)).

    Reference :59-79.  Entries with index > (n - 1) // 2 come from horizontally flipped images and are flipped back;
    ``tF.resize(..., BILINEAR)`` on tensors is ``interpolate(mode='bilinear', align_corners=False)`` without
    anti-aliasing in the pinned torchvision 0.12 (actsegmul.yml)."""
    n = len(feat_list)
    feats, outs = [], []
    for k, (feat, out) in enumerate(zip(feat_list, output_list)):
        if (n - 1) // 2 < k:
            feat, out = torch.flip(feat, dims=[-1]), torch.flip(out, dims=[-1])
        feats.append(F.interpolate(feat, size=tuple(im_size), mode="bilinear", align_corners=False)[0])
        outs.append(F.interpolate(out, size=tuple(im_size), mode="bilinear", align_corners=False)[0])
    feats = F.normalize(torch.stack(feats).mean(dim=0), dim=0)
    return feats[None], torch.stack(outs).mean(dim=0)[None]


class LabellerMixin(ProtoLabellerMixin):
    only_multihot = False
    fuse_multiscale = staticmethod(fuse_multiscale)


from ._bind import bind  # noqa: E402

# the reference's own trainer with the hot-path methods replaced (None when the reference checkout is not importable)
ActiveTrainer = bind("eval_save_cosplbl_prop_includeonehot_voc_ms", LabellerMixin, "trainer/eval_save_cosplbl_prop_includeonehot_voc_ms.py with the fused pseudo_label_generation (:152-354).")
