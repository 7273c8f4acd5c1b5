"""ExperienceReplay -- buffer lifecycle, sampling and the ER class-weighted CE
(reference: loss/experience_replay.py).  BACSLoss inherits the buffer plumbing; the CE
itself runs in the fused pixel kernel through ``compute_base_loss``."""
from __future__ import annotations

import numpy as np
import torch

from .. import _cabi, ops
from .base_loss import BaseLoss


class ExperienceReplay(BaseLoss):
    def __init__(self, name="Experience Replay", ignore_index=255, alpha: float = 1.0, buffer_size: int = 50,
                 replay_minibatch_size: int = 32, bg_weighted_ce: bool = False, same_task: bool = True):
        super().__init__(name, ignore_index=ignore_index)
        self.buffer_size = buffer_size
        self.replay_minibatch_size = replay_minibatch_size
        self.alpha = alpha
        self._use_er_loss = False
        self.buffer = None
        self.bg_weighted_ce = bg_weighted_ce
        self.same_task = same_task
        self._iter_indx = 0
        self.co_occurence_map = None
        self.grads = {}

    # ---- buffer lifecycle (experience_replay.py:34-97) ---------------------------------------
    def _init_buffer(self, task_num=0):
        from ..training.buffer import Buffer
        if task_num == 0:
            if self.same_task:
                self.buffer = [Buffer(self.buffer_size, "task_{}".format(0))]
            else:
                self.buffer = Buffer(self.buffer_size, "all_tasks")
        elif self.same_task:
            self.buffer.append(Buffer(self.buffer_size, "task_{}".format(task_num), task_num=task_num))
        self._get_current_buffer().update_task(task_num=task_num, new_class_size=self.nb_current_classes)

    def get_available_tasks(self):
        if self.buffer is None:
            return None
        if self.same_task:
            return range(len(self.buffer))
        return self.buffer.get_available_tasks()

    def _get_current_buffer(self):
        return self.buffer[-1] if self.same_task else self.buffer

    def update_buffer_scores(self):
        if self.same_task:
            for buffer in self.buffer:
                buffer.merge_scores(self.co_occurence_map)
        else:
            self._get_current_buffer().merge_scores(self.co_occurence_map)

    def _get_random_buffer(self):
        if not self.same_task:
            return self._get_current_buffer()
        n_buffers = len(self.buffer[:-1])
        if n_buffers > 1:
            scores = np.array([self.buffer[i].get_importance() for i in range(n_buffers)], dtype=np.float64)
            scores = scores / np.max(scores)
            e = np.exp(scores - scores.max())
            task_id = np.random.choice(range(n_buffers), p=e / e.sum(), size=1)[0]
        else:
            task_id = 0
        return self.buffer[task_id]

    # ---- events ------------------------------------------------------------------------------
    def on_train_batch_start(self, **kwargs):
        BaseLoss.on_train_batch_start(self, **kwargs)
        self._iter_indx = kwargs.get("batch_idx")

    def on_train_start(self, task_num, **kwargs):
        self._iter_indx = 0
        self._init_buffer(task_num=task_num)
        if task_num > 0:
            self._use_er_loss = True

    def _score_batch(self, model, images, labels, weights):
        """Per-image importance -(w_y * nll).mean over H*W (experience_replay.py:137-143,
        bacs_loss.py:183-189) from the fused kernel's SCORE mode."""
        fused = bool(getattr(self, "fused_logit_upsample", False))
        # fused: score straight from the head's low-res logits (the up-sample runs inside the kernel)
        logits = model(images, return_sem_logits=True) if fused else model(images)
        out = ops.pixel_loss(logits.detach(), labels, _cabi.PIX_SCORE, want_grad=False, want_preds=False,
                             class_w=weights, want_score=True, ignore_index=self.ignore_index, lowres=fused)
        return logits, out["score"].float()

    def on_train_end(self, **kwargs):
        """experience_replay.py:111-151: fill the ER buffer from the train loader."""
        super().on_train_end(**kwargs)
        if not kwargs.get("pre_last_tasks"):
            return
        model = kwargs.get("model", None)
        train_dataloader = kwargs.get("train_dataloader", None)
        if self.buffer is None:
            self._init_buffer()
        if model is not None and train_dataloader is not None:
            accelerator = kwargs.get("accelerator")
            model = model.to(accelerator.root_device)
            train_dataloader = accelerator.process_dataloader(train_dataloader)
            classes_weights = torch.ones(self.nb_current_classes, device=accelerator.root_device)
            classes_weights[0] = 0
            with torch.no_grad():
                for index, batch in enumerate(train_dataloader):
                    batch = accelerator.to_device(batch)
                    images, labels = batch[0], batch[1].long()
                    _, losses = self._score_batch(model, images, labels, classes_weights)
                    self._add_to_buffer(images, labels, losses)
                    if (index * images.shape[0]) >= self.buffer_size:
                        break
            self.update_buffer_scores()

    # ---- loss ----------------------------------------------------------------------------------
    def compute_loss(self, batch, model, train=True):
        """experience_replay.py:153-186."""
        if isinstance(batch, dict):
            img, mask = batch["main"][0], batch["main"][1]
        else:
            img, mask = batch[0], batch[1]
        loss, preds_mask = self.compute_base_loss(
            img, mask, model, train=train, use_weighted_ce=self.bg_weighted_ce and self._use_er_loss and train)
        preds_output = self._argmax(preds_mask)
        if train and self._use_er_loss:
            loss = loss + self.alpha * self._replay_er_loss(model, self._get_random_buffer())
        return loss, preds_output

    def _add_to_buffer(self, examples, labels, losses):
        with torch.no_grad():
            self._get_current_buffer().add_data({"examples": examples.detach().cpu(), "labels": labels.cpu(),
                                                 "loss": losses.detach().cpu()})

    def _update_prototype(self, feats, labels):
        if self._prototypes is not None and self.same_task:
            self._prototypes.update_feats_prototypes(feats, labels)

    def _sample_buffer(self, buffer=None, same_task=False, task_num=None, mixup=False, on_cpu=False):
        """experience_replay.py:207-242."""
        if self.same_task and task_num is not None and task_num < len(self.buffer):
            buffer = self.buffer[task_num]
        elif buffer is None:
            buffer = self._get_random_buffer()
        if buffer.is_empty():
            return None
        device = self.accelerator.root_device if self.accelerator is not None else self.device
        memory_dict = buffer.get_data(self.replay_minibatch_size, same_task=same_task, task_num=task_num,
                                      mixup=mixup, device=None if on_cpu else device)
        return (memory_dict, memory_dict["examples"], memory_dict.get("logits"), memory_dict["labels"],
                memory_dict["n_classes"], buffer.task_num if not same_task else memory_dict["task_id"])

    def _replay_er_loss(self, model, buffer):
        """experience_replay.py:244-272 (the reference multiplies by alpha here and again in
        compute_loss; kept)."""
        memory_data = self._sample_buffer(buffer)
        if memory_data is None or not self._use_er_loss:
            return 0
        _, memory_inputs, _, memory_labels, _, task_num = memory_data
        classes_weights = torch.zeros(self.nb_current_classes, device=self.device)
        if task_num > -1:
            classes_weights[1:self.get_n_old_classes(task_num + 1)] = 1
        else:
            classes_weights[1:self.old_classes] = 1
        loss, _ = self.compute_base_loss(memory_inputs, memory_labels, model, weights=classes_weights,
                                         task_num=task_num, train=True, use_weighted_ce=False,
                                         _loss_scale=self.alpha)
        return loss
