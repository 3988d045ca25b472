"""Hot-path part of the reference's ``trainer/eval_within_multihot.py``: ``top_pseudo_label_generation`` (:93-146)."""
from ..labeller import TopLabellerMixin as LabellerMixin  # noqa: F401
