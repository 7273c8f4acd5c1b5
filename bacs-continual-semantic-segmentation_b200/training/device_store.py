"""HBM-resident replay store.

The reference keeps the replay buffer in numpy memmaps (``training/buffer.py:25-30``: ``examples`` [N,3,H,W] fp32,
``logits`` [N,K,h,w] fp32, ``labels``, ``seen``) and feeds it to the step through two extra dataloaders with worker
processes (``dataloaders/base_datamodule.py:453-469``, ``dataset/base_segmentation_dataset.py:73-97``): every step
fancy-indexes the memmaps on the host and copies the minibatch to the GPU (75 MB of images at Br=24, 512x512).
``DeviceReplayStore`` mirrors the fields of a :class:`Buffer` in device memory once (300 VOC samples: 0.95 GB of the
180 GB) and serves minibatches with one gather kernel per field (``bacs_gather_rows``); the sampling decisions stay in
the ``Buffer`` (same ``np.random`` stream, same indices), and the on-disk ``.dat`` files stay the source of truth
for resume.  ``get_data`` returns exactly what ``Buffer.get_data(..., device=...)`` returns."""
from __future__ import annotations

from typing import Dict, Iterator, Optional

import numpy as np
import torch

from .. import ops


class DeviceReplayStore:
    def __init__(self, buffer, device, fields=None):
        self.buffer = buffer
        self.device = torch.device(device)
        self.fields = tuple(fields) if fields is not None else None
        self.maps: Dict[str, torch.Tensor] = {}
        self.n_uploaded_bytes = 0
        self.refresh()

    # ---- mirror --------------------------------------------------------------------------------------
    def refresh(self, indices=None) -> None:
        """(Re-)upload the buffer's fields: everything, or only the slots in ``indices`` (after ``add_data``)."""
        if self.buffer.dataset_map is None:
            return
        for name, dmap in self.buffer.dataset_map.items():
            if self.fields is not None and name not in self.fields:
                continue
            have = self.maps.get(name)
            host = dmap[:]
            if indices is None or have is None or tuple(have.shape) != tuple(host.shape):
                t = torch.from_numpy(np.ascontiguousarray(host))
                self.maps[name] = t.to(self.device)
                self.n_uploaded_bytes += t.numel() * t.element_size()
            else:
                idx = np.asarray(indices, dtype=np.int64)
                rows = torch.from_numpy(np.ascontiguousarray(host[idx]))
                have[torch.from_numpy(idx).to(self.device)] = rows.to(self.device)
                self.n_uploaded_bytes += rows.numel() * rows.element_size()

    def __len__(self) -> int:
        return int(self.buffer.buffer_size)

    # ---- Buffer.get_data on the device (training/buffer.py:346-389) ------------------------------------
    def gather(self, choice) -> Dict[str, torch.Tensor]:
        idx = torch.as_tensor(np.asarray(choice, dtype=np.int64)).to(self.device, non_blocking=True)
        return {name: ops.gather_rows(t, idx) for name, t in self.maps.items()}

    def get_data(self, size: int, return_indexes=False, same_task=False, task_num=None, mixup=False):
        if mixup and self.buffer.co_occurance_map is not None:
            raise NotImplementedError("DeviceReplayStore.get_data(mixup=True): co-occurrence mix-up is not on the BACS path")
        choice, task_id = self.buffer._sample_indices(size, same_task=same_task, task_num=task_num)
        ret = self.gather(choice)
        ret["n_classes"] = self.buffer._logits_n_classes[choice]
        if self.buffer.transformations is not None:
            ret["examples"] = self.buffer.transformations(ret["examples"])
        ret["task_id"] = task_id
        return (ret, choice) if return_indexes else ret

    # ---- the "bufferlogits" loader (base_datamodule.py:453-469) without workers or host copies ---------
    def logits_batches(self, batch_size: int, length: Optional[int] = None, shuffle: bool = True, transforms=None,
                       generator: Optional[torch.Generator] = None) -> Iterator[list]:
        """One epoch of ``[examples, logits, n_classes]`` batches over the first ``length`` slots, shuffled like a
        ``DataLoader(shuffle=True, drop_last=False)``; ``transforms`` is applied per sample as the reference's
        ``BaseMemMapDataset.__getitem__`` does."""
        n = int(length) if length is not None else len(self)
        order = torch.randperm(n, generator=generator) if shuffle else torch.arange(n)
        n_classes = torch.from_numpy(np.asarray(self.buffer._logits_n_classes))
        for start in range(0, n, batch_size):
            choice = order[start:start + batch_size]
            idx = choice.to(self.device, non_blocking=True)
            examples = ops.gather_rows(self.maps["examples"], idx)
            if transforms is not None:
                examples = torch.stack([transforms(e) for e in examples])
            yield [examples, ops.gather_rows(self.maps["logits"], idx), n_classes[choice]]
