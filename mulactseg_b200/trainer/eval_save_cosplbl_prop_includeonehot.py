"""Hot-path part of the reference's ``trainer/eval_save_cosplbl_prop_includeonehot.py`` (the shipped stage-2 recipe):
``pseudo_label_generation`` (:121-316), prototypes from every selected superpixel."""
from ..labeller import ProtoLabellerMixin


class LabellerMixin(ProtoLabellerMixin):
    only_multihot = False
