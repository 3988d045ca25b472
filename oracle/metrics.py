"""CPU oracle: mIoU counters and dominant label assignment.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Restates in numpy
  * ``MeanIoU._after_step`` / ``_after_step_within_predregion`` / ``_after_epoch`` / ``_after_epoch_ipr``
    (``utils/miou.py:23-95``), and
  * the dominant label assignment of ``RegionCityscapesDominantAll.__getitem__`` (``dataloader/region_dataset.py:216-233``).
Pinned by ``tests/golden/metrics.npz``, produced by the unmodified reference classes (``oracle/gen_golden.py:gen_metrics``).
"""
from __future__ import annotations

import numpy as np


def miou_counts(outputs: np.ndarray, targets: np.ndarray, num_classes: int, ignore_label: int, by_output: bool = False):
    """-> (3, num_classes) int64: seen, correct, positive (utils/miou.py:23-54)."""
    outputs, targets = np.asarray(outputs).reshape(-1), np.asarray(targets).reshape(-1)
    keep = (outputs != ignore_label) if by_output else (targets != ignore_label)
    o, t = outputs[keep], targets[keep]
    counts = np.zeros((3, num_classes), dtype=np.int64)
    for i in range(num_classes):
        counts[0, i] = np.sum(t == i)
        counts[1, i] = np.sum((t == i) & (o == t))
        counts[2, i] = np.sum(o == i)
    return counts


def ious(counts: np.ndarray, ignore_label_list=None):
    seen, correct, positive = counts.astype(np.float64)
    out = []
    for i in range(counts.shape[1]):
        if ignore_label_list is not None and i in ignore_label_list:
            continue
        out.append(1 if seen[i] == 0 else correct[i] / (seen[i] + positive[i] - correct[i]))
    return [v * 100 for v in out]


def ious_precisions_recalls(counts: np.ndarray):
    seen, correct, positive = counts.astype(np.float64)
    i_, p_, r_ = [], [], []
    for i in range(counts.shape[1]):
        if seen[i] == 0:
            i_.append(1); p_.append(1); r_.append(1)
        else:
            with np.errstate(divide="ignore", invalid="ignore"):
                i_.append(correct[i] / (seen[i] + positive[i] - correct[i]))
                p_.append(correct[i] / positive[i])
                r_.append(correct[i] / seen[i])
    return [v * 100 for v in i_], [v * 100 for v in p_], [v * 100 for v in r_]


def dominant_target(target: np.ndarray, superpixel: np.ndarray, preserving_labels) -> np.ndarray:
    """Every listed superpixel takes its most frequent non-ignore label (np.unique order: the smallest wins a tie);
    ignore pixels stay 255, unlisted superpixels stay as they are."""
    h, w = target.shape
    t = target.reshape(-1).copy()
    s = superpixel.reshape(-1)
    ignore = t == 255
    for p in preserving_labels:
        mask = (s == p) & ~ignore
        u, c = np.unique(t[mask], return_counts=True)
        if c.size != 0:
            t[mask] = u[c.argmax()]
    t[ignore] = 255
    return t.reshape(h, w)
