"""Import shims that let the UNMODIFIED reference run in the build container.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Used only by
``oracle/gen_golden.py`` (fixture generation) -- the GPU box has no
``/root/reference`` and nothing at test/bench run time touches this module's
``install()``.

What is shimmed and why (reference file:line):
* ``torch_scatter``  -> ``oracle.scatter_ref`` (package absent from the image;
  ``active_selection/my_bvsb.py:3``, ``utils/loss.py:6``).
* ``imageio``        -> stub; every ``dataloader/*`` module calls
  ``imageio.plugins.freeimage.download()`` at import time, which needs network
  (``dataloader/region_dataset.py:11-12``).
* ``skimage``        -> stub mapping ``binary_dilation(img, fp)`` to
  ``scipy.ndimage.binary_dilation(img, structure=fp)`` and ``find_boundaries``
  to a numpy equivalent (``trainer/eval_save_cosplbl_prop.py:8-9``).
* ``collections.Iterable`` alias (py3.8-ism, ``dataloader/ext_transforms.py:568``).
* ``torch.Tensor.cuda`` -> identity when no GPU is present (hard ``.cuda()`` at
  ``trainer/eval_save_cosplbl_prop.py:266``).
"""
from __future__ import annotations

import collections
import collections.abc
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MULACTSEG_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "active_selection"))


def install() -> None:
    import numpy as np
    import torch

    from . import scatter_ref

    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")

    sys.modules.setdefault("torch_scatter", scatter_ref)

    if "imageio" not in sys.modules:
        imageio = types.ModuleType("imageio")
        plugins = types.ModuleType("imageio.plugins")
        freeimage = types.ModuleType("imageio.plugins.freeimage")
        freeimage.download = lambda *a, **k: None
        plugins.freeimage = freeimage
        imageio.plugins = plugins
        imageio.imread = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("imageio stub"))
        sys.modules["imageio"] = imageio
        sys.modules["imageio.plugins"] = plugins
        sys.modules["imageio.plugins.freeimage"] = freeimage

    if "skimage" not in sys.modules:
        from scipy import ndimage

        skimage = types.ModuleType("skimage")
        morphology = types.ModuleType("skimage.morphology")
        segmentation = types.ModuleType("skimage.segmentation")

        def binary_dilation(image, footprint=None, out=None):
            return ndimage.binary_dilation(np.asarray(image).astype(bool), structure=footprint)

        def find_boundaries(label_img, connectivity=1, mode="thick", background=0):
            lab = np.asarray(label_img)
            fp = ndimage.generate_binary_structure(lab.ndim, connectivity)
            return ndimage.grey_dilation(lab, footprint=fp) != ndimage.grey_erosion(lab, footprint=fp)

        def mark_boundaries(image, label_img, *a, **k):
            return np.asarray(image, dtype=float) / 255.0

        morphology.binary_dilation = binary_dilation
        segmentation.find_boundaries = find_boundaries
        segmentation.mark_boundaries = mark_boundaries
        skimage.morphology = morphology
        skimage.segmentation = segmentation
        sys.modules["skimage"] = skimage
        sys.modules["skimage.morphology"] = morphology
        sys.modules["skimage.segmentation"] = segmentation

    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
