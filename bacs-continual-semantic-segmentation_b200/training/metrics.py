"""IoU metric -- drop-in for the reference's ``training.metrics.IoU`` (training/metrics.py:20-103),
which subclasses torchmetrics 0.6.0's IoU/ConfusionMatrix.  The confusion matrix is an
integer histogram kernel (csrc/misc.cu: block-private K x K counters in shared memory);
the per-class metrics keep the reference's naming, including its swapped fp/fn (Q7)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
from torch import Tensor

from .. import ops


@dataclass
class IouMetric:
    iou_per_class: Tensor
    miou: Tensor
    accuracy: Tensor
    precision: Tensor
    recall: Tensor
    specificity: Tensor


class IoU(torch.nn.Module):
    """``iou(preds, target)`` accumulates; ``compute()`` returns an :class:`IouMetric`.

    Unlike torchmetrics' ``forward`` this does not also compute a batch-local value on
    every call (the reference's caller discards it, training/model.py:270-272)."""

    def __init__(self, num_classes: int = 11, ignore_indx: int = 255, sync_on_compute: bool = True):
        super().__init__()
        self.num_classes = num_classes
        self.ignore_index = ignore_indx
        self.sync_on_compute = sync_on_compute
        self.register_buffer("confmat", torch.zeros(num_classes, num_classes, dtype=torch.int64), persistent=False)
        self.register_buffer("out_of_range", torch.zeros(1, dtype=torch.int64), persistent=False)

    def update(self, preds: Tensor, target: Tensor) -> None:
        """metrics.py:38-50: flatten, cast to int32, keep 0 <= target < K, histogram of
        target*K + pred.  preds may be float (truncated like .int()) or int64."""
        if preds.dtype not in (torch.int64, torch.float32):
            preds = preds.long() if not preds.is_floating_point() else preds.float()
        ops.confmat_accumulate(preds.reshape(-1), target.reshape(-1).long(), self.num_classes, self.confmat,
                               self.out_of_range)

    def forward(self, preds: Tensor, target: Tensor) -> None:
        self.update(preds, target)

    def reset(self) -> None:
        self.confmat.zero_()
        self.out_of_range.zero_()

    def _synced_confmat(self) -> Tensor:
        if self.sync_on_compute and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            from ..distributed import allreduce_state
            cm = self.confmat.clone()
            allreduce_state(None, None, cm)
            return cm
        return self.confmat

    def compute(self) -> IouMetric:
        """metrics.py:52-88."""
        met = ops.confmat_metrics(self._synced_confmat())
        return IouMetric(iou_per_class=met[0], miou=met[5, 0], accuracy=met[1], precision=met[2], recall=met[3],
                         specificity=met[4])

    @staticmethod
    def get_mean_per_classes(metric_result: Tensor, classes: list):
        """metrics.py:90-102"""
        vals = metric_result.detach().cpu()
        return sum(float(vals[label]) for label in classes) / len(classes)


class PerStepResult:
    """Aggregates results per continual task (metrics.py:105-156); host-side bookkeeping."""

    def __init__(self, continual):
        self._per_step_result = {"mIoU": [], "IoU-Old": [], "IoU-Old-nobg": [], "IoU-New": []}
        self.metrics = self._per_step_result.keys()
        self.continual = continual
        self.task_id = 0

    def update(self, final_result):
        for metric in self.metrics:
            self._per_step_result[metric].append([])
        for dataset_id in range(len(final_result)):
            for metric in self.metrics:
                key = ("test.{}/Task {}/{}".format(dataset_id, self.task_id, metric) if self.continual
                       else "test.{}/{}".format(dataset_id, metric))
                if key in final_result[dataset_id]:
                    self._per_step_result[metric][-1].append(final_result[dataset_id][key])
        self.task_id += 1

    def get_metrics(self):
        if not self.continual:
            return ["mIoU"]
        return list(self.metrics) + ["Avg-IoU"]

    def get_avg_iou(self):
        return np.array(self._per_step_result["mIoU"]).mean(axis=0)

    def get_n_datasets(self):
        return len(self._per_step_result["mIoU"][-1])

    def compute(self):
        results = {metric: self._per_step_result[metric][-1] for metric in self.metrics}
        results["Avg-IoU"] = self.get_avg_iou()
        return results
