"""CPU tier: the ``mulactseg_b200.trainer.<name>`` modules define ``ActiveTrainer`` on top of the reference's own trainer
when the reference checkout is importable, and a one-line module in the reference's ``trainer`` package is then loaded
exactly like ``train_AL.py:32`` / ``eval_AL.py:31`` load any trainer (``importlib.import_module("trainer." + method)``).
Needs ``/root/reference`` (build container only; skipped on the GPU box) -- the binding itself is host logic."""
import importlib
import sys
import types

import pytest

from oracle import ref_shims

NAMES = ["active_joint_multi_lossdecomp", "active_joint_multi_predignore_lossdecomp", "eval_save_cosplbl_prop",
         "eval_save_cosplbl_prop_includeonehot", "eval_save_cosplbl_prop_includeonehot_voc_ms", "eval_within_multihot"]


def test_modules_import_without_the_reference():
    for name in NAMES:
        mod = importlib.import_module(f"mulactseg_b200.trainer.{name}")
        assert hasattr(mod, "ActiveTrainer")
        assert hasattr(mod, "CriterionMixin") or hasattr(mod, "LabellerMixin")


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference checkout not present")
@pytest.mark.parametrize("name", NAMES)
def test_active_trainer_is_the_reference_trainer_with_the_fused_hot_path(name, tmp_path):
    ref_shims.install()                       # reference root on sys.path + stand-ins for torch_scatter / imageio / skimage
    ours = importlib.reload(importlib.import_module(f"mulactseg_b200.trainer.{name}"))
    ref = importlib.import_module(f"trainer.{name}")
    assert ours.ActiveTrainer is not None and issubclass(ours.ActiveTrainer, ref.ActiveTrainer)

    # the reference-side stub, found through the reference's own plugin mechanism
    import trainer as ref_pkg
    (tmp_path / f"b200_{name}.py").write_text(f"from mulactseg_b200.trainer.{name} import ActiveTrainer  # noqa: F401\n")
    ref_pkg.__path__.append(str(tmp_path))
    try:
        Trainer = importlib.import_module("trainer.{}".format(f"b200_{name}".lower()))      # train_AL.py:32
        cls = Trainer.ActiveTrainer
        assert cls is ours.ActiveTrainer
        obj = cls.__new__(cls)                # BaseTrainer.__init__ needs cuda:0, data and wandb: not part of the binding
        obj.args = types.SimpleNamespace(nseg=32, group_ce_temp=0.1, multi_ce_temp=0.1, cosprop_threshold_method="median")
        obj.num_classes = 19
        if hasattr(ours, "CriterionMixin"):
            from mulactseg_b200 import losses
            obj.get_criterion()
            assert isinstance(obj.group_multi_loss, losses.GroupMultiLabelCE_onlymulti)
            assert isinstance(obj.multi_pos_loss, losses.OnehotCEMultihotChoice)
            # the loop stays the reference's: the mixin's train_impl only wraps it to raise the last deferred partition check
            assert super(ours.CriterionMixin, obj).train_impl.__func__ is ref.ActiveTrainer.train_impl
        else:
            mixin = ours.LabellerMixin
            method = "top_pseudo_label_generation" if name == "eval_within_multihot" else "pseudo_label_generation"
            assert getattr(cls, method) is getattr(mixin, method)
            assert cls.inference is ref.ActiveTrainer.inference            # the loop stays the reference's
    finally:
        ref_pkg.__path__.remove(str(tmp_path))
        sys.modules.pop(f"trainer.b200_{name}", None)
