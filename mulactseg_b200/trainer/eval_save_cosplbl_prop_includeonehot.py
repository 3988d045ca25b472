"""Hot-path part of the reference's ``trainer/eval_save_cosplbl_prop_includeonehot.py`` (the shipped stage-2 recipe):
``pseudo_label_generation`` (:121-316), prototypes from every selected superpixel."""
from ..labeller import ProtoLabellerMixin


class LabellerMixin(ProtoLabellerMixin):
    only_multihot = False


from ._bind import bind  # noqa: E402

# the reference's own trainer with the hot-path methods replaced (None when the reference checkout is not importable)
ActiveTrainer = bind("eval_save_cosplbl_prop_includeonehot", LabellerMixin, "trainer/eval_save_cosplbl_prop_includeonehot.py:20 with the fused pseudo_label_generation (:121-316).")
