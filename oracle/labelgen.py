"""CPU oracle: offline multi-hot label generation.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Restates ``RegionCityscapesTensor.__getitem__``
(``dataloader/region_cityscapes_tensor.py:33-84``; driver ``tools/label_assignment_tensor.py:50-67``) in numpy.

Third-party arithmetic absent from the reference tree and from this image: ``skimage.segmentation.find_boundaries``
and ``skimage.morphology.binary_dilation`` (scikit-image 0.19.2, ``actsegmul.yml:106``).  Published algorithm restated:
``find_boundaries(mode='thick', connectivity=1)`` = grey dilation != grey erosion with the 4-neighbour cross
(reflecting borders); ``binary_dilation(img, ones(k, k))`` = ``scipy.ndimage.binary_dilation`` (outside = False).
No reference test pins either ("parity unpinned" at that boundary); the golden vectors in
``tests/golden/labelgen.npz`` were produced by the unmodified reference class over these restatements.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage


def find_boundaries_thick(labels: np.ndarray) -> np.ndarray:
    fp = ndimage.generate_binary_structure(labels.ndim, 1)
    return ndimage.grey_dilation(labels, footprint=fp) != ndimage.grey_erosion(labels, footprint=fp)


def multi_hot_labels(target: np.ndarray, superpixel: np.ndarray, preserving_labels, nseg: int, num_classes: int,
                     trim_kernel_size: int = 0):
    """-> (superpixel_cls (nseg, num_classes + 1) uint8, superpixel_size (nseg,) int32, -1 for ids not preserved)."""
    cls_out = np.zeros((nseg, num_classes + 1), dtype=np.uint8)
    size_out = np.full((nseg,), -1, dtype=np.int32)
    flat_t = target.reshape(-1)
    flat_s = superpixel.reshape(-1)
    if trim_kernel_size:
        bdry = ndimage.binary_dilation(find_boundaries_thick(superpixel), structure=np.ones((trim_kernel_size,) * 2, np.uint8))
        trimmed = np.where(bdry, nseg, superpixel).reshape(-1)
    for p in preserving_labels:
        if trim_kernel_size:
            mask = trimmed == p
            if not mask.any():
                mask = flat_s == p
        else:
            mask = flat_s == p
        u = np.unique(flat_t[mask])
        valid = u[u != 255]
        cls_out[p, valid] = 1
        if (u == 255).any():
            cls_out[p, -1] = 1
        size_out[p] = int(mask.sum())
    return cls_out, size_out
