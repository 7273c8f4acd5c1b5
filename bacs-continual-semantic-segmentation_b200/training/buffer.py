"""Replay buffer -- host-side store behind BACS's dark-experience replay
(reference: training/buffer.py).  It is "hot-path adjacent" (SURVEY 2, row 13): it feeds the
DER logit-MSE / dark++ kernels through the data loaders and is filled once per task, so it
stays host/disk code -- numpy memmaps with the reference's on-disk layout
``<cwd>/mem_maps/<buffer_name>/<field>_<n>.dat`` so that a run can be resumed by either
implementation.

Kept behaviours of the reference:
  * reservoir insertion with score-weighted eviction once the buffer is full;
  * the labels recorded for an inserted sample are the unique labels of the WHOLE batch;
  * ``DatasetMap.extend`` never writes slot 0 (``if index > 0``);
  * scores = min-max normalised 0.3 * importance / scale + 0.7 * balance."""
from __future__ import annotations

import os
from shutil import copyfile
from typing import Dict, Optional

import numpy as np
import torch


def _original_cwd() -> str:
    try:                                                   # the reference asks hydra for the launch dir
        from hydra.utils import get_original_cwd
        return get_original_cwd()
    except Exception:                                      # noqa: BLE001 -- hydra absent or not initialised
        return os.environ.get("BACS_BUFFER_ROOT", os.getcwd())


class DatasetMap:
    """One memmapped field of the buffer: ``[size, *data_size]`` of ``data_type``."""

    def __init__(self, size: int, data_size: tuple, data_type: str, path: str, name: str) -> None:
        self.name, self.size, self.data_size, self.path, self.data_type = name, size, tuple(data_size), path, data_type
        os.makedirs(self.path, exist_ok=True)
        self.file_path, self.increment = self._get_full_path(name)
        self.data_map = np.memmap(self.file_path, dtype=data_type, mode="w+", shape=(self.size, *self.data_size))
        self.length = 0

    def _get_full_path(self, name, increment=0):
        path = os.path.join(self.path, "{}_{}.dat".format(name, increment))
        while os.path.exists(path):                        # one file per process / re-creation
            increment += 1
            path = os.path.join(self.path, "{}_{}.dat".format(name, increment))
        return path, increment

    def __getitem__(self, index):
        return self.data_map[index, ...]

    def __len__(self) -> int:
        return self.length

    def add(self, item, index):
        self.data_map[index] = item
        self.length += 1

    def extend(self, items, indices):
        for index, item in zip(indices, items):
            if index > 0:                                  # reference quirk: slot 0 is never written
                self.add(item, index)

    def change_data_size(self, new_data_size):
        """Grows the leading data dimension (more classes in the stored logits), zero padded."""
        new_data_size = tuple(new_data_size)
        self.data_map.flush()
        tmp_path, self.increment = self._get_full_path(self.name, increment=self.increment)
        copyfile(self.file_path, tmp_path)
        old_shape = (self.size, *self.data_size)
        self.data_map = np.memmap(self.file_path, dtype=self.data_type, mode="w+", shape=(self.size, *new_data_size))
        old = np.memmap(tmp_path, dtype=self.data_type, mode="r", shape=old_shape)
        self.data_map[:, :self.data_size[0], ...] = old
        self.data_size = new_data_size
        del old
        try:
            os.remove(tmp_path)
        except OSError:
            pass


class Buffer:
    def __init__(self, buffer_size, buffer_name, same_task=False, task_num=-1, transformations=None) -> None:
        self.buffer_name = buffer_name
        self.buffer_size = buffer_size
        self._logits_n_classes = np.zeros(buffer_size, dtype="uint8")
        self._task_id_list = np.zeros(buffer_size, dtype="uint8")
        self.dataset_map: Optional[Dict[str, DatasetMap]] = None
        self.same_task = same_task
        self.task_num = task_num
        self._num_seen_examples = 0
        self.transformations = transformations
        self.importance_score = np.full(buffer_size, -np.inf, dtype="float")
        self.balance_score = np.full(buffer_size, -np.inf, dtype="float")
        self.scores = np.full(buffer_size, -np.inf, dtype="float")
        self._existing_indices = np.full(buffer_size, False)
        self.labels = {}                 # class -> number of stored samples containing it
        self._examples_labels = {}       # slot -> labels recorded at insertion
        self.img_paths, self.target_paths, self.target_trsf = {}, {}, {}
        self.co_occurance_map = None

    # ---- scores -------------------------------------------------------------------------------
    def get_importance(self):
        valid = self.importance_score != -np.inf
        return 10 if not valid.any() else np.median(-1 * self.importance_score[valid])

    def merge_scores(self, co_occurance_map=None):
        self.co_occurance_map = co_occurance_map
        for slot, labs in self._examples_labels.items():
            self.balance_score[slot] = min([self.labels[lab] for lab in labs if lab != 0])
        scale = np.mean(abs(self.importance_score)) * np.mean(abs(self.balance_score))
        pre = 0.3 * (self.importance_score / scale) + 0.7 * self.balance_score
        if pre.max() - pre.min() != 0:
            pre = (pre - np.min(pre)) / (np.max(pre) - np.min(pre))
        self.scores = pre / np.sum(pre)

    def functionalReservoir(self, N, m):
        if N < m:
            return N
        if np.random.randint(0, N) < m:
            self.merge_scores()
            return np.random.choice(range(m), p=self.scores, size=1)[0]
        return -1

    # ---- bookkeeping ----------------------------------------------------------------------------
    def update_task(self, task_num, new_class_size):
        self.task_num = task_num
        has_logits = self.dataset_map is not None and "logits" in self.dataset_map
        if has_logits and new_class_size > self._logits_n_classes.max() and self.num_seen_examples > 0:
            shape = self.dataset_map["logits"].data_size
            self.dataset_map["logits"].change_data_size([new_class_size, shape[1], shape[2]])

    @property
    def num_seen_examples(self):
        return self._num_seen_examples

    def _init_map(self, dict_data):
        root = os.path.join(_original_cwd(), "mem_maps", self.buffer_name)
        self.dataset_map = {
            key: DatasetMap(self.buffer_size, val.shape[1:], str(val.dtype).split(".")[-1], root, key)
            for key, val in dict_data.items()}

    def add_data(self, dict_data):
        """Reservoir insertion of a batch: ``examples`` (required), ``labels``, optional
        ``logits`` / ``seen`` / ``loss`` (importance) / ``img_paths`` + ``target_paths`` + ``target_trsf``."""
        dict_data = dict(dict_data)
        has_paths = "img_paths" in dict_data
        if has_paths:
            img_paths, target_paths = dict_data.pop("img_paths"), dict_data.pop("target_paths")
            target_trsf = dict_data.pop("target_trsf")
        loss_scores = dict_data.pop("loss").cpu().numpy() if "loss" in dict_data else None
        if self.dataset_map is None:
            self._init_map(dict_data)
        n_logit_classes = dict_data["logits"].shape[1] if "logits" in dict_data else None
        assert "examples" in dict_data
        batch_labels = dict_data["labels"].cpu().unique().numpy()       # whole batch (Q12)
        indices = []
        for index in range(dict_data["examples"].shape[0]):
            slot = self.functionalReservoir(self.num_seen_examples, self.buffer_size)
            if slot >= 0:
                if self.num_seen_examples >= self.buffer_size:
                    for former in self._examples_labels[slot]:
                        if former != 0 and former in self.labels:
                            self.labels[former] -= 1
                self._examples_labels[slot] = batch_labels
                for lab in batch_labels:
                    if lab != 0:
                        self.labels[lab] = self.labels.get(lab, 0) + 1
                if n_logit_classes is not None:
                    self._logits_n_classes[slot] = n_logit_classes
                if has_paths:
                    self.img_paths[slot] = img_paths[index]
                    self.target_paths[slot] = target_paths[index]
                    self.target_trsf[slot] = target_trsf[index]
                self._task_id_list[slot] = self.task_num
                self._existing_indices[slot] = True
                indices.append(slot)
                self.importance_score[slot] = loss_scores[index] if loss_scores is not None else -np.inf
            self._num_seen_examples += 1
        for key, val in dict_data.items():
            assert key in self.dataset_map
            self.dataset_map[key].extend(val.cpu().numpy(), indices)

    def get_available_tasks(self):
        return np.unique(self._task_id_list)

    # ---- sampling -------------------------------------------------------------------------------
    def _sample_indices(self, sample_size, same_task=False, task_num=None):
        sample_size = min(sample_size, self.num_seen_examples)
        existing = np.where(self._existing_indices)[0]
        pool = existing if self.num_seen_examples < self.buffer_size else self.buffer_size
        if not same_task:
            return np.random.choice(pool, size=sample_size, replace=False), -1
        task_id = np.random.choice(np.unique(self._task_id_list), size=1)[0] if task_num is None else task_num
        own = np.where((self._task_id_list == task_id) & self._existing_indices)[0]
        if own.size >= sample_size:
            return np.random.choice(own, size=sample_size, replace=False), task_id
        extra = np.random.choice(pool, size=sample_size - own.size, replace=False)
        return np.concatenate([extra, own], axis=0), task_id

    def get_data(self, size: int, return_indexes=False, same_task=False, task_num=None, mixup=False, device=None):
        if mixup and self.co_occurance_map is not None:
            raise NotImplementedError("Buffer.get_data(mixup=True): co-occurrence mix-up is not on the BACS path")
        choice, task_id = self._sample_indices(size, same_task=same_task, task_num=task_num)
        ret = {name: torch.tensor(np.asarray(dmap[choice]), device=device) for name, dmap in self.dataset_map.items()}
        ret["n_classes"] = self._logits_n_classes[choice]
        if self.transformations is not None:
            ret["examples"] = self.transformations(ret["examples"])
        ret["task_id"] = task_id
        return (ret, choice) if return_indexes else ret

    def is_empty(self) -> bool:
        return self.num_seen_examples == 0
