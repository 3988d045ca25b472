"""Hot-path part of the reference's ``trainer/eval_save_cosplbl_prop.py``: ``pseudo_label_generation`` (:121-314),
prototypes from multi-hot superpixels only."""
from ..labeller import ProtoLabellerMixin


class LabellerMixin(ProtoLabellerMixin):
    only_multihot = True


from ._bind import bind  # noqa: E402

# the reference's own trainer with the hot-path methods replaced (None when the reference checkout is not importable)
ActiveTrainer = bind("eval_save_cosplbl_prop", LabellerMixin, "trainer/eval_save_cosplbl_prop.py:20 with the fused pseudo_label_generation (:121-314).")
