"""Binding of the hot-path mixins to the reference's own ``ActiveTrainer`` classes.

The reference loads a trainer as ``importlib.import_module("trainer." + args.method).ActiveTrainer(args, logger,
selection_iter)`` (``train_AL.py:32,38``, ``eval_AL.py:31``).  When the reference checkout is importable (its root on
``sys.path``, so that ``trainer.<name>`` resolves to the reference's module) every ``mulactseg_b200.trainer.<name>``
defines ``ActiveTrainer = (mixin, reference ActiveTrainer)``: the training / evaluation loops stay the reference's, the
criteria / pseudo-label generators are the fused kernels.  A one-line module in the reference tree then switches a recipe:

    # trainer/b200_active_joint_multi_predignore_lossdecomp.py        (--method b200_active_joint_multi_predignore_lossdecomp)
    from mulactseg_b200.trainer.active_joint_multi_predignore_lossdecomp import ActiveTrainer  # noqa: F401

Without the reference on the path the modules still import and export their mixins (``ActiveTrainer`` is then ``None``).
"""
from __future__ import annotations

import importlib
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_trainer(name: str):
    """The reference's ``trainer.<name>.ActiveTrainer`` or None when that package cannot be imported here."""
    try:
        mod = importlib.import_module("trainer." + name)
    except Exception:          # no reference checkout, or one of its own imports (torch_scatter, imageio, wandb ...) is missing
        return None
    path = os.path.dirname(os.path.abspath(getattr(mod, "__file__", "") or ""))
    if path == _HERE:          # resolved to this package under another name: not the reference
        return None
    return getattr(mod, "ActiveTrainer", None)


def bind(name: str, mixin, doc: str):
    """``class ActiveTrainer(mixin, reference ActiveTrainer)`` or None."""
    ref = reference_trainer(name)
    if ref is None:
        return None
    return type("ActiveTrainer", (mixin, ref), {"__doc__": doc, "__module__": mixin.__module__})
