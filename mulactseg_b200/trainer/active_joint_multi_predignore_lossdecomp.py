"""Hot-path part of the reference's ``trainer/active_joint_multi_predignore_lossdecomp.py``: ``get_criterion``
(:76-81) with the fused criteria; ``train_impl`` (:83-117) calls them unchanged."""
from ..losses import GroupMultiLabelCE_onlymulti, OnehotCEMultihotChoice, stage1_criterion  # noqa: F401


class CriterionMixin:
    def get_criterion(self):
        self.group_multi_loss, self.multi_pos_loss = stage1_criterion(self.args, self.num_classes, voc=False)

    def train_impl(self, total_itrs, val_period):
        """The reference's loop (trainer/active.py:73 / active_joint_multi*.py) unchanged; afterwards the partition assertion
        of the LAST step is raised (it is checked one call late so that no training step waits for the GPU)."""
        try:
            return super().train_impl(total_itrs, val_period)
        finally:
            check = getattr(getattr(self, "multi_pos_loss", None), "check_partition", None)
            if check is not None:
                check()


from ._bind import bind  # noqa: E402

# the reference's own trainer with the hot-path methods replaced (None when the reference checkout is not importable)
ActiveTrainer = bind("active_joint_multi_predignore_lossdecomp", CriterionMixin, "trainer/active_joint_multi_predignore_lossdecomp.py:74 with the fused criteria installed by get_criterion (:76-81).")
