"""Drop-in ``MeanIoU`` (reference ``utils/miou.py``): same constructor, hooks and return values, with the per-class
counts accumulated on the GPU by one kernel per step (``mas_miou_counts_dev``) instead of 3 x num_classes masked
reductions with a ``.item()`` sync each (``utils/miou.py:28-33``).  Nothing syncs before ``_after_epoch*`` (or a read
of ``total_seen`` / ``total_correct`` / ``total_positive``, which stay available as float numpy arrays like the
reference's).  numpy inputs (the reference's second branch, :34-38) are copied to the device first.
"""
from __future__ import annotations

from typing import Any, Dict

import numpy as np
import torch

from . import ops


class MeanIoU():
    def __init__(self, num_classes: int, ignore_label: int, output_tensor: str = 'outputs', target_tensor: str = 'targets',
                 name: str = 'iou', device=None) -> None:
        self.num_classes = num_classes
        self.ignore_label = ignore_label
        self.name = name
        self.output_tensor = output_tensor
        self.target_tensor = target_tensor
        self.device = device
        self._counts = None

    # ------------------------------------------------------------------ reference hooks
    def _before_epoch(self) -> None:
        self._counts = None

    def _accumulate(self, output_dict: Dict[str, Any], by_output: bool) -> None:
        outputs = output_dict[self.output_tensor]
        targets = output_dict[self.target_tensor]
        if isinstance(outputs, np.ndarray) or isinstance(targets, np.ndarray):
            dev = torch.device(self.device if self.device is not None else "cuda")
            outputs = torch.as_tensor(np.ascontiguousarray(outputs)).to(dev)
            targets = torch.as_tensor(np.ascontiguousarray(targets)).to(dev)
        if not outputs.is_cuda:
            raise RuntimeError("mulactseg_b200.miou needs CUDA tensors (there is no CPU path)")
        if self._counts is None or self._counts.device != outputs.device:
            fresh = torch.zeros(3 * self.num_classes, dtype=torch.int64, device=outputs.device)
            if self._counts is not None:
                fresh += self._counts.to(outputs.device)
            self._counts = fresh
        ops.miou_counts(outputs, targets, self.num_classes, self.ignore_label, by_output, self._counts)

    def _after_step(self, output_dict: Dict[str, Any]) -> None:
        self._accumulate(output_dict, False)

    def _after_step_within_predregion(self, output_dict: Dict[str, Any]) -> None:
        self._accumulate(output_dict, True)

    # ------------------------------------------------------------------ the reference's public counters
    def _host(self) -> np.ndarray:
        if self._counts is None:
            return np.zeros((3, self.num_classes))
        return self._counts.cpu().numpy().reshape(3, self.num_classes).astype(np.float64)

    @property
    def total_seen(self) -> np.ndarray:
        return self._host()[0]

    @property
    def total_correct(self) -> np.ndarray:
        return self._host()[1]

    @property
    def total_positive(self) -> np.ndarray:
        return self._host()[2]

    def _after_epoch(self, ignore_label_list=None):
        seen, correct, positive = self._host()
        ious = []
        for i in range(self.num_classes):
            if ignore_label_list is not None and i in ignore_label_list:
                continue
            if seen[i] == 0:
                ious.append(1)
            else:
                ious.append(correct[i] / (seen[i] + positive[i] - correct[i]))
        return [num * 100 for num in ious]

    def _after_epoch_ipr(self):
        seen, correct, positive = self._host()
        ious, precisions, recalls = [], [], []
        for i in range(self.num_classes):
            if seen[i] == 0:
                ious.append(1)
                precisions.append(1)
                recalls.append(1)
            else:
                with np.errstate(divide="ignore", invalid="ignore"):      # a class never predicted: x / 0.0 like the reference
                    ious.append(correct[i] / (seen[i] + positive[i] - correct[i]))
                    precisions.append(correct[i] / positive[i])
                    recalls.append(correct[i] / seen[i])
        return ([num * 100 for num in ious], [num * 100 for num in precisions], [num * 100 for num in recalls])
