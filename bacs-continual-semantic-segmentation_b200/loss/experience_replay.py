"""ExperienceReplay -- replay-buffer bookkeeping, sampling and the ER class-weighted CE.

Interface mirror of the reference's ``loss/experience_replay.py`` (constructor keywords, hooks, attribute and method
names are what ``BACSLoss`` and the learners call); the arithmetic is not here: every loss
term is one launch of the fused pixel kernel through ``compute_base_loss`` / ``ops.pixel_loss`` (SCORE mode for the
per-image importance of the end-of-task pass).

``self.buffer`` keeps the reference's shape because the training loop reads it: a list with one ``Buffer`` per task
when ``same_task`` (the ER baseline), a single shared ``Buffer`` otherwise (BACS)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _cabi, ops
from .base_loss import BaseLoss


class ExperienceReplay(BaseLoss):
    def __init__(self, name="Experience Replay", ignore_index=255, alpha: float = 1.0, buffer_size: int = 50,
                 replay_minibatch_size: int = 32, bg_weighted_ce: bool = False, same_task: bool = True):
        super().__init__(name, ignore_index=ignore_index)
        self.alpha = alpha
        self.buffer_size, self.replay_minibatch_size = buffer_size, replay_minibatch_size
        self.bg_weighted_ce, self.same_task = bg_weighted_ce, same_task
        self.buffer = None
        self.co_occurence_map = None
        self.grads = {}
        self._use_er_loss = False
        self._iter_indx = 0

    # ------------------------------------------------------------------ buffers
    def _buffers(self):
        """every live buffer, oldest first (empty before the first task starts)"""
        if self.buffer is None:
            return []
        return list(self.buffer) if self.same_task else [self.buffer]

    def _get_current_buffer(self):
        return self._buffers()[-1]

    def _init_buffer(self, task_num=0):
        """reference experience_replay.py:34-52: per-task buffers are appended task by task, the shared buffer is
        created once at task 0; either way the current buffer learns the task's class count (its stored logits are
        zero-padded to it, training/buffer.py:63-93)."""
        from ..training.buffer import Buffer
        if self.same_task:
            fresh = Buffer(self.buffer_size, "task_%d" % task_num, **({"task_num": task_num} if task_num else {}))
            self.buffer = [fresh] if task_num == 0 else self.buffer + [fresh]
        elif task_num == 0:
            self.buffer = Buffer(self.buffer_size, "all_tasks")
        self._get_current_buffer().update_task(task_num=task_num, new_class_size=self.nb_current_classes)

    def get_available_tasks(self):
        if self.buffer is None:
            return None
        return range(len(self.buffer)) if self.same_task else self.buffer.get_available_tasks()

    def update_buffer_scores(self):
        """sampling scores = importance + class balance (training/buffer.py:150-161) of the buffers in use"""
        for buf in (self._buffers() if self.same_task else self._buffers()[-1:]):
            buf.merge_scores(self.co_occurence_map)

    def _get_random_buffer(self):
        """Which buffer a replay minibatch comes from (experience_replay.py:73-97): the shared buffer, or one of the
        EARLIER tasks' buffers drawn with probability softmax(importance / max importance)."""
        if not self.same_task:
            return self.buffer
        earlier = self.buffer[:-1]
        if len(earlier) <= 1:
            return self.buffer[0]
        score = np.asarray([b.get_importance() for b in earlier], dtype=np.float64)
        score = score / score.max()
        p = np.exp(score - score.max())
        return earlier[np.random.choice(range(len(earlier)), p=p / p.sum(), size=1)[0]]

    def _add_to_buffer(self, examples, labels, losses):
        with torch.no_grad():
            self._get_current_buffer().add_data({"examples": examples.detach().cpu(), "labels": labels.cpu(),
                                                 "loss": losses.detach().cpu()})

    def _update_prototype(self, feats, labels):
        if self._prototypes is not None and self.same_task:
            self._prototypes.update_feats_prototypes(feats, labels)

    # ------------------------------------------------------------------ hooks
    def on_train_batch_start(self, **kwargs):
        BaseLoss.on_train_batch_start(self, **kwargs)
        self._iter_indx = kwargs.get("batch_idx")

    def on_train_start(self, task_num, **kwargs):
        self._iter_indx = 0
        self._init_buffer(task_num=task_num)
        self._use_er_loss = self._use_er_loss or task_num > 0

    def _score_batch(self, model, images, labels, weights):
        """Per-image importance -(w_y * nll).mean over H*W (experience_replay.py:137-143, bacs_loss.py:183-189):
        SCORE mode of the fused kernel -- from the head's low-res logits when the up-sample is fused."""
        fused = bool(getattr(self, "fused_logit_upsample", False))
        logits = model(images, return_sem_logits=True) if fused else model(images)
        out = ops.pixel_loss(logits.detach(), labels, _cabi.PIX_SCORE, want_grad=False, want_preds=False,
                             class_w=weights, want_score=True, ignore_index=self.ignore_index, lowres=fused)
        return logits, out["score"].float()

    def on_train_end(self, **kwargs):
        """End of a task (experience_replay.py:111-151): score the task's training images with the background-free
        CE and offer them to the buffer until ``buffer_size`` images have been seen."""
        super().on_train_end(**kwargs)
        model, loader = kwargs.get("model"), kwargs.get("train_dataloader")
        if not kwargs.get("pre_last_tasks"):
            return
        if self.buffer is None:
            self._init_buffer()
        if model is None or loader is None:
            return
        accelerator = kwargs.get("accelerator")
        device = accelerator.root_device
        model = model.to(device)
        class_w = torch.ones(self.nb_current_classes, device=device)
        class_w[0] = 0
        with torch.no_grad():
            for step, batch in enumerate(accelerator.process_dataloader(loader)):
                images, labels = accelerator.to_device(batch)[:2]
                labels = labels.long()
                self._add_to_buffer(images, labels, self._score_batch(model, images, labels, class_w)[1])
                if step * images.shape[0] >= self.buffer_size:
                    break
        self.update_buffer_scores()

    # ------------------------------------------------------------------ loss
    def compute_loss(self, batch, model, train=True):
        """experience_replay.py:153-186: CE on the live batch (+ alpha * replay CE once a task has been finished)."""
        img, mask = (batch["main"] if isinstance(batch, dict) else batch)[:2]
        replaying = bool(train and self._use_er_loss)
        loss, logits = self.compute_base_loss(img, mask, model, train=train,
                                              use_weighted_ce=self.bg_weighted_ce and replaying)
        preds = self._argmax(logits)
        if replaying:
            loss = loss + self.alpha * self._replay_er_loss(model, self._get_random_buffer())
        return loss, preds

    def _sample_buffer(self, buffer=None, same_task=False, task_num=None, mixup=False, on_cpu=False):
        """One replay minibatch (experience_replay.py:207-242) as the reference's 6-tuple
        ``(dict, examples, logits, labels, n_classes, task)``; ``logits`` is None for buffers that store none (the
        ER buffer: the reference indexes ``["logits"]`` there and raises KeyError)."""
        if self.same_task and task_num is not None and task_num < len(self.buffer):
            buffer = self.buffer[task_num]
        buffer = self._get_random_buffer() if buffer is None else buffer
        if buffer.is_empty():
            return None
        device = None if on_cpu else (self.accelerator.root_device if self.accelerator is not None else self.device)
        got = buffer.get_data(self.replay_minibatch_size, same_task=same_task, task_num=task_num, mixup=mixup,
                              device=device)
        task = got["task_id"] if same_task else buffer.task_num
        return got, got["examples"], got.get("logits"), got["labels"], got["n_classes"], task

    def _replay_er_loss(self, model, buffer):
        """experience_replay.py:244-272: CE over the OLD foreground classes of a replayed minibatch.  The reference
        scales by alpha here and once more in compute_loss; both are kept (the first one as the kernel's loss scale)."""
        sample = self._sample_buffer(buffer) if self._use_er_loss else None
        if sample is None:
            return 0
        inputs, labels, task = sample[1], sample[3], sample[5]
        n_old = self.get_n_old_classes(task + 1) if task > -1 else self.old_classes
        class_w = torch.zeros(self.nb_current_classes, device=self.device)
        class_w[1:n_old] = 1
        return self.compute_base_loss(inputs, labels, model, weights=class_w, task_num=task, train=True,
                                      use_weighted_ce=False, _loss_scale=self.alpha)[0]
