"""Fake network objects satisfying the contract the loss relies on:
fixed tensors stand in for the DeepLabV3 forward so only the loss path runs."""
import torch
import torch.nn as nn


class FakeNet(nn.Module):
    """model(img, return_penultimate=True, return_attentions=True) -> (logits, pen, [att]);
    model(x, return_sem_logits=True) -> low-res logits.  Outputs are leaf tensors keyed
    by the id of the image tensor so main and replay batches get different outputs."""

    def __init__(self, seen_fg_network=None):
        super().__init__()
        self.seen_fg_network = seen_fg_network
        self.table = {}
        self.sem_table = {}

    def register(self, img, logits, pen, atts):
        self.table[id(img)] = (logits, pen, atts)

    def register_sem(self, img, sem):
        self.sem_table[id(img)] = sem

    def forward(self, x, return_attentions=False, return_penultimate=False,
                return_sem_logits=False, only_attentions=False):
        if return_sem_logits:
            return self.sem_table[id(x)] if id(x) in self.sem_table else next(iter(self.sem_table.values()))
        logits, pen, atts = self.table[id(x)] if id(x) in self.table else next(iter(self.table.values()))
        if return_penultimate and return_attentions:
            return logits, pen, atts
        if return_penultimate:
            return logits, pen
        if return_attentions:
            return logits, atts
        return logits

    def get_penultimate_layer_dim(self):
        return self.seen_fg_network.inter_channels


class EndOfTaskNet(nn.Module):
    """The model contract of the end-of-task loops (loss/bacs_loss.py:133-203, loss/prototypes.py:92-125):
    ``model(images)`` -> full-res logits, ``enable_caching_sem_logits`` / ``pop_sem_logits`` -> the low-res logits of
    the last forward, ``get_penultimate_output(images)``, ``clone()``, ``seen_fg_network``.  Batches are told apart
    by ``images[0, 0, 0, 0]`` (their index), so loaders may copy / move the tensors."""

    def __init__(self, seen_fg_network, logits, sems, pens):
        super().__init__()
        self.seen_fg_network = seen_fg_network
        self._logits, self._sems, self._pens = list(logits), list(sems), list(pens)
        self._last = None
        self._caching = False

    @staticmethod
    def _index(images):
        return int(round(float(images[0, 0, 0, 0])))

    def _pick(self, seq, images):
        return seq[self._index(images)].to(images.device)

    def forward(self, x, **kwargs):
        self._last = self._index(x)
        if kwargs.get("return_sem_logits"):
            return self._pick(self._sems, x)
        return self._pick(self._logits, x)

    def enable_caching_sem_logits(self):
        self._caching = True

    def pop_sem_logits(self):
        assert self._caching and self._last is not None
        self._caching = False
        return self._sems[self._last].to(self._logits[self._last].device if False else self._dev())

    def _dev(self):
        return getattr(self, "_device", "cpu")

    def to(self, device, *args, **kwargs):
        self._device = device
        return super().to(device, *args, **kwargs)

    def get_penultimate_output(self, images):
        return self._pick(self._pens, images)

    def clone(self):
        twin = EndOfTaskNet(self.seen_fg_network, self._logits, self._sems, self._pens)
        twin._device = self._dev()
        return twin


class EndOfTaskLoader(list):
    """A 'dataloader' for the end-of-task loops: a list of (images, labels) with the attributes the loops touch."""

    def __init__(self, batches, paths, target_paths):
        super().__init__(batches)
        import types
        self.shuffle = True
        self.dataset = types.SimpleNamespace(_x=paths, _y=target_paths, target_trsf=None)


class EndOfTaskAccelerator:
    def __init__(self, device):
        self.root_device = torch.device(device)

    def process_dataloader(self, loader):
        return loader

    def to_device(self, batch):
        return [t.to(self.root_device) for t in batch]


def end_of_task_case(synth, n_batches=3, name="tiny", seed=31):
    """Seeded inputs of the end-of-task parity case (shared by the golden generator and the GPU test)."""
    import numpy as np
    cfg = synth.CONFIGS[name]
    inps = [synth.make_step_inputs(cfg, seed=seed + i) for i in range(n_batches)]
    g = torch.Generator().manual_seed(seed)
    images, sems = [], []
    for i in range(n_batches):
        im = torch.rand(cfg.B, 3, cfg.H, cfg.W, generator=g)
        im[:, 0, 0, 0] = float(i)
        images.append(im)
        sems.append(torch.randn(cfg.B, cfg.K, cfg.h, cfg.w, generator=g))
    paths = np.array(["img_%d.jpg" % i for i in range(n_batches * cfg.B)])
    tpaths = np.array(["lbl_%d.png" % i for i in range(n_batches * cfg.B)])
    return cfg, inps, images, sems, paths, tpaths
