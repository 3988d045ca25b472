"""Hot-path part of the reference's ``trainer/eval_save_cosplbl_prop.py``: ``pseudo_label_generation`` (:121-314),
prototypes from multi-hot superpixels only."""
from ..labeller import ProtoLabellerMixin


class LabellerMixin(ProtoLabellerMixin):
    only_multihot = True
