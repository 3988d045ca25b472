"""ctypes binding of ``libmulactseg_b200.so`` (the C ABI of ``include/mulactseg_b200.h``).

There is no CPU fallback anywhere in this package: if the shared library has not
been built, or a compute entry point is called without an sm_100 device, the
call raises.  ``python -m mulactseg_b200.build`` (or ``__graft_entry__.build()``)
produces the library in-tree under ``mulactseg_b200/lib/``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MULACTSEG_B200_LIB") or os.path.join(_HERE, "lib", "libmulactseg_b200.so")   # override: comparison builds

MAS_F32, MAS_BF16 = 0, 1
MAS_I32, MAS_I64, MAS_U8 = 0, 1, 2
MAS_SCATTER_I64 = 3
MAS_MIOU_BY_TARGET, MAS_MIOU_BY_OUTPUT = 0, 1
MAS_GROUP_ALL, MAS_GROUP_ONLYMULTI = 0, 1
MAS_LOSS_CHOICE, MAS_LOSS_GROUP, MAS_LOSS_EXACT_SOFTMAX = 1, 2, 4
MAS_MAX_CLASSES = 32
MAS_MAX_LOSS_CLASSES = 31
MAS_MAX_SEGMENTS = 32
MAS_THRESHOLD_MEDIAN, MAS_THRESHOLD_MIN = 0, 1

# name -> (restype, argtypes); mirrors include/mulactseg_b200.h one to one
SIGNATURES = {
    "mas_abi_version": (c_int, []),
    "mas_last_error": (c_char_p, []),
    "mas_kernel_launches": (c_int64, []),
    "mas_bvsb_segment_stats_dev": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                           c_void_p, c_void_p, c_void_p, c_void_p]),
    "mas_bvsb_segment_stats_lowres_dev": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                                  c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mas_class_weights_dev": (c_int, [c_void_p, c_int64, c_int, c_int64, c_int, c_float, c_void_p, c_void_p]),
    "mas_prefix_cut_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "mas_bvsb_segment_stats_multi_dev": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                                 c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mas_region_scores_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mas_minmax_nonzero_dev": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "mas_dominant_hist_dev": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "mas_finalize_scores_dev": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "mas_region_keys_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "mas_topk_workspace_bytes": (c_size_t, []),
    "mas_topk_u64_dev": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mas_topk_sorted_u64_dev": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mas_topk_candidates_u64_dev": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mas_topk_candidates_msg_u64_dev": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mas_merge_counts_u64_dev": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "mas_sort_capacity": (c_int64, [c_int64]),
    "mas_sort_desc_u64_dev": (c_int, [c_void_p, c_int64, c_void_p]),
    "mas_acquisition_host": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_float,
                                     c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mas_select_topk_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p]),
    "mas_scatter_sum_dev": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int64, c_int64, c_int, c_int64, c_void_p, c_void_p]),
    "mas_scatter_max_dev": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mas_multihot_info_dev": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mas_multihot_loss_fwd_dev": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                          c_float, c_int, c_void_p, c_void_p, c_void_p]),
    "mas_multihot_tiles_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mas_multihot_tiles_dev": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "mas_multihot_loss_fwd_tiles_dev": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                                c_int, c_float, c_int, c_void_p, c_void_p, c_void_p]),
    "mas_multihot_loss_bwd_tiles_dev": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                                c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p]),
    "mas_stage1_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "mas_stage1_loss_fwd_dev": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_float, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "mas_stage1_loss_bwd_dev": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "mas_multihot_loss_finish_dev": (c_int, [c_void_p, c_void_p, c_void_p]),
    "mas_multihot_loss_coef_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "mas_candidate_argmax_dev": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                         c_void_p, c_void_p]),
    "mas_proto_labeller_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "mas_proto_labeller_dev": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                                       c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mas_proto_labeller_src_dev": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                           c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mas_proto_labeller_batch_dev": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                             c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                             c_void_p, c_int, c_void_p]),
    "mas_multihot_labels_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mas_multihot_labels_dev": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                        c_void_p, c_void_p, c_size_t, c_void_p]),
    "mas_dominant_labels_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mas_dominant_labels_dev": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                        c_size_t, c_void_p]),
    "mas_miou_counts_dev": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int64, c_int, c_void_p, c_void_p]),
    "mas_multihot_loss_bwd_dev": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                          c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p]),
}

_lib = None


class MulActSegError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MulActSegError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -m mulactseg_b200.build`); there is no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().mas_last_error()
        raise MulActSegError(f"{what} failed (code {code}): {msg.decode() if msg else ''}")


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)
