"""Continual-learning label remap on the GPU.

Mirrors the reference's ``TransformLabel`` (training/utils.py:225-261) and the label
transformation built per task by ``CityScapeScenario._get_label_transformation``
(dataset/cityscape_dataset.py:77-108; continuum's VOC/ADE scenarios build the same maps):
ignore-255 is kept, labels of hidden tasks collapse onto the masking value (background
shift), and the remap is applied with the reference's *sequential in-place* semantics, whose
aliasing under shuffled class orders (Q13) is reproduced exactly by composing the effective
LUT from the labels present in each image (csrc/labels.cu)."""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

IGNORE = 255
_LO, _N = -1, 257          # value domain [-1, 255]: Cityscapes raw ids start at -1


def build_inverted_order(class_order: Sequence[int], task_labels: Iterable[int], train: bool,
                         test_background: bool = True) -> Tuple[Dict[int, int], int]:
    """cityscape_dataset.py:96-106 -> (inverted_order, masking_value).  ``task_labels`` are the
    labels made visible: the current task's (overlap / disjoint training), every task's so far
    (sequential mode, and testing)."""
    class_order = list(class_order)
    inverted = {int(lab): class_order.index(lab) + 1 for lab in task_labels}
    inverted[IGNORE] = IGNORE
    masking = 0
    if not train:
        if test_background:
            inverted[0] = 0
        else:
            masking = IGNORE
    return inverted, masking


def _dense_map(mapping: Dict[int, int], masking: int) -> np.ndarray:
    dense = np.full(_N, masking, dtype=np.int32)
    for key, val in mapping.items():
        if not (_LO <= key < _LO + _N and _LO <= val < _LO + _N):
            raise ValueError("label remap: %d -> %d outside the supported domain [-1, 255]" % (key, val))
        dense[key - _LO] = val
    return dense


class TransformLabel:
    """Callable with the reference's constructor: ``TransformLabel(input_dict, masking_value,
    inverted_order=None, inverted_masking=None)``.  Accepts a CUDA int64 label tensor of shape
    [H,W] (one target, as the reference's dataset transform sees it) or [N,H,W] (N targets
    remapped independently in one launch) and returns the remapped tensor."""

    def __init__(self, input_dict, masking_value, inverted_order=None, inverted_masking=None):
        self.input_dict, self.masking_value = dict(input_dict), int(masking_value)
        self.inverted_order = None if inverted_order is None else dict(inverted_order)
        self.inverted_masking = None if inverted_masking is None else int(inverted_masking)
        self._map1 = _dense_map(self.input_dict, self.masking_value)
        self._map2 = None if self.inverted_order is None else _dense_map(self.inverted_order, self.inverted_masking)
        self._dev = {}

    def _device_maps(self, device):
        key = str(device)
        if key not in self._dev:
            m1 = torch.from_numpy(self._map1).to(device)
            m2 = None if self._map2 is None else torch.from_numpy(self._map2).to(device)
            self._dev[key] = (m1, m2)
        return self._dev[key]

    def __call__(self, lbl: torch.Tensor) -> torch.Tensor:
        squeeze = lbl.dim() == 2
        x = lbl.unsqueeze(0) if squeeze else lbl
        x = x.long().contiguous()
        m1, m2 = self._device_maps(x.device)
        out = ops.label_remap(x, m1, self.masking_value, m2,
                              0 if self.inverted_masking is None else self.inverted_masking, lo=_LO)
        return out[0] if squeeze else out

    def __repr__(self):
        return self.__class__.__name__
