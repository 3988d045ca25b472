"""Hot-path part of the reference's ``trainer/eval_within_multihot.py``: ``top_pseudo_label_generation`` (:93-146)."""
from ..labeller import TopLabellerMixin as LabellerMixin


from ._bind import bind  # noqa: E402

# the reference's own trainer with the hot-path methods replaced (None when the reference checkout is not importable)
ActiveTrainer = bind("eval_within_multihot", LabellerMixin, "trainer/eval_within_multihot.py with the fused top_pseudo_label_generation (:93-146).")
