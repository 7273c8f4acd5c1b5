"""CPU oracle for the BACS per-pixel continual-learning loss path.

TEST INFRASTRUCTURE ONLY.  This file is a from-scratch CPU restatement (torch-CPU /
numpy, fp32 unless noted) of the reference's algorithm for the hot path named in
BASELINE.json.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the checker /
the timed CPU baseline -- never as the product.  The product path lives in
``bacs_b200`` (CUDA, sm_100a) and refuses to run without its compiled library.

Pinning status:
  * every function below is compared against the *imported* reference code in
    ``tests/test_oracle_vs_reference.py`` (runs only where /root/reference exists) and
    against the committed fixtures in ``tests/golden/*.npz`` (generated from the
    reference by ``tests/golden/make_golden.py``);
  * ``confusion_matrix``/``iou_metrics`` are additionally pinned by the reference's
    only known-answer vector (training/metrics.py:159-183 -> C=[[2,5],[5,4]]);
  * third-party arithmetic that is NOT under /root/reference and not installed here --
    ``segmentation_models_pytorch.losses.FocalLoss`` (binary mode), torchmetrics 0.6.0
    ``ConfusionMatrix``/``IoU`` and continuum 1.2.1's label transformation -- is restated
    from the published algorithm: **parity unpinned** for those three beyond the
    test_iou vector and the reference's own call sites.

Each function cites the reference file:line it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

IGNORE = 255


# --------------------------------------------------------------------------------------
# Row 0 -- continual-learning label remap
#   dataset/cityscape_dataset.py:77-108 (map construction), training/utils.py:225-261
#   (TransformLabel: sequential in-place application over sorted unique labels)
# --------------------------------------------------------------------------------------
def build_inverted_order(class_order: Sequence[int], task_labels: Iterable[int],
                         train: bool, test_background: bool = True) -> Tuple[Dict[int, int], int]:
    """Returns (inverted_order, masking_value) as cityscape_dataset.py:96-106 builds them.

    ``task_labels`` are the labels made visible (current task for overlap/disjoint
    training, all tasks so far for sequential mode and for testing)."""
    inv = {int(lab): class_order.index(lab) + 1 for lab in task_labels}
    inv[IGNORE] = IGNORE
    masking = 0
    if not train:
        if test_background:
            inv[0] = 0
        else:
            masking = IGNORE
    return inv, masking


def sequential_remap(lbl: np.ndarray, mapping: Dict[int, int], masking: int) -> np.ndarray:
    """training/utils.py:240-248.  Visits the labels present *at entry* in ascending
    order and rewrites pixels currently equal to each one -- so a pixel already
    rewritten to a larger, still-unvisited present label is rewritten again (Q13)."""
    out = np.array(lbl, copy=True)
    for cur in np.unique(out).tolist():
        cur = int(cur)
        out[out == cur] = mapping.get(cur, masking)
    return out


def transform_label(lbl: np.ndarray, input_dict: Dict[int, int], masking: int,
                    inverted_order: Optional[Dict[int, int]] = None,
                    inverted_masking: Optional[int] = None) -> np.ndarray:
    """training/utils.py:250-258: id->trainId pass then the CL pass."""
    out = sequential_remap(lbl, input_dict, masking)
    if inverted_order is not None:
        out = sequential_remap(out, inverted_order, inverted_masking)
    return out


def effective_sequential_lut(present: np.ndarray, mapping: Dict[int, int], masking: int,
                             lo: int, n: int) -> np.ndarray:
    """Closed form of ``sequential_remap`` as a LUT over values [lo, lo+n): follow
    v -> f(v) while the image is a *larger* value that was present at entry."""
    lut = np.zeros(n, dtype=np.int64)
    for v in range(lo, lo + n):
        cur = v
        while True:
            nxt = mapping.get(cur, masking)
            if nxt > cur and lo <= nxt < lo + n and present[nxt - lo]:
                cur = nxt
                continue
            lut[v - lo] = nxt
            break
    return lut


# --------------------------------------------------------------------------------------
# Row 3 -- nearest label down-sample + class -> task
# --------------------------------------------------------------------------------------
def nearest_src_index(out_size: int, in_size: int) -> np.ndarray:
    """Index rule of F.interpolate(mode='nearest') as used at loss/prototypes.py:181-186
    (SURVEY A1): float32 scale, floor, clamp."""
    if out_size == in_size:
        return np.arange(out_size, dtype=np.int64)
    if out_size == 2 * in_size:
        return np.arange(out_size, dtype=np.int64) >> 1
    scale = np.float32(in_size) / np.float32(out_size)
    idx = np.floor(np.arange(out_size, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, in_size - 1)


def downsample_labels(target: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """[B,H,W] int64 -> [B,h,w] int64 (loss/prototypes.py:181-186)."""
    ri = torch.from_numpy(nearest_src_index(h, target.shape[1])).to(target.device)
    ci = torch.from_numpy(nearest_src_index(w, target.shape[2])).to(target.device)
    return target[:, ri][:, :, ci].contiguous()


def class_to_task(labels, initial_classes: int, increment: int) -> np.ndarray:
    """loss/base_loss.py:98-107: rint (half-to-even) of max((c+1-initial)/increment, 0)."""
    labels = np.asarray(labels, dtype=np.int64)
    if increment <= 0:
        return np.zeros(labels.shape, dtype=np.int64)
    t = (labels + 1 - initial_classes) / increment
    t[t < 0] = 0
    return np.rint(t).astype(np.int64)


def class_task_lut(initial_classes: int, increment: int, ignore_index: int = IGNORE,
                   include_bg: bool = False) -> np.ndarray:
    """256-entry LUT class -> task, -1 for background / ignore (prototypes.py:191-204)."""
    lut = class_to_task(np.arange(256), initial_classes, increment)
    lut[ignore_index] = -1
    if not include_bg:
        lut[0] = -1
    return lut


# --------------------------------------------------------------------------------------
# Row 4/5 -- prototype accumulate / running-mean update / ready predicate
#   loss/prototypes.py:127-163, 31-40
# --------------------------------------------------------------------------------------
def proto_accumulate(features: torch.Tensor, target: torch.Tensor, initial_classes: int,
                     increment: int, n_tasks: int, ignore_index: int = IGNORE,
                     mode: str = "exact") -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-task feature sums S[g,:] (fp32) and masked-pixel counts N[g] (int64).

    mode='exact'   reproduces the reference's masked-index ``.view(D,-1)`` row split for
                   B>1: element (b,c,k) sits at flat position
                   D*sum_{b'<b} n_b' + c*n_b + k and lands in row pos // N_g.
    mode='channel' is the decomposable per-channel sum (identical when B == 1)."""
    B, D, h, w = features.shape
    feats = features.detach().float()
    ld = downsample_labels(target, h, w)
    lut = torch.from_numpy(class_task_lut(initial_classes, increment, ignore_index)).to(ld.device)
    valid = (ld >= 0) & (ld < 256)
    task = torch.full_like(ld, -1)
    task[valid] = lut[ld[valid]]
    sums = torch.zeros(n_tasks, D, dtype=torch.float32, device=feats.device)   # (the full-size GPU tests run this file on device tensors)
    counts = torch.zeros(n_tasks, dtype=torch.int64, device=feats.device)
    for g in range(n_tasks):
        m = task == g                                   # [B,h,w]
        n_g = int(m.sum())
        counts[g] = n_g
        if n_g == 0:
            continue
        if mode == "channel":
            sums[g] = (feats * m.unsqueeze(1)).double().sum(dim=(0, 2, 3)).float()
        else:
            flat = torch.cat([feats[b][:, m[b]].reshape(-1) for b in range(B)])
            sums[g] = flat.view(D, n_g).double().sum(-1).float()
    return sums, counts


def proto_update(proto: torch.Tensor, count: torch.Tensor, sums: torch.Tensor,
                 n: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """loss/prototypes.py:158-163 for every task row with n>0.  ``count`` keeps its dtype
    (int64 at task 0, float32 afterwards -- Q3); arithmetic is fp32 like torch's."""
    proto = proto.clone()
    count = count.clone()
    for g in range(proto.shape[0]):
        if int(n[g]) == 0:
            continue
        new = (sums[g] + count[g] * proto[g]) / (count[g] + n[g])
        count[g] += n[g]
        proto[g] = new
    return proto, count


def prototypes_ready(count: Optional[torch.Tensor]) -> bool:
    """loss/prototypes.py:31-40."""
    return count is not None and int(count.count_nonzero()) == count.shape[0]


# --------------------------------------------------------------------------------------
# Row 6 -- seen / unseen detector heads (networks/bg_detector.py:17-40,100-165)
# --------------------------------------------------------------------------------------
def seen_logits_lowres(pen: torch.Tensor, protos: torch.Tensor, weight: torch.Tensor,
                       bias: torch.Tensor) -> torch.Tensor:
    """z[b,t,i,j] = bias_t + sum_c w[t,c] * |sigmoid(pen[b,c,i,j]) - sigmoid(proto[t,c])|."""
    sx = torch.sigmoid(pen.float())                               # [B,D,h,w]
    sp = torch.sigmoid(protos.float())                            # [T,D]
    dist = (sx.unsqueeze(1) - sp[None, :, :, None, None]).abs()   # [B,T,D,h,w]
    return (dist * weight.float()[None, :, :, None, None]).sum(2) + bias.float()[None, :, None, None]


def _lerp_table(out_size: int, in_size: int, align_corners: bool):
    """Source indices/weights of torch's bilinear kernels (fp32 arithmetic)."""
    dst = torch.arange(out_size, dtype=torch.float32)
    if align_corners:
        scale = np.float32(in_size - 1) / np.float32(out_size - 1) if out_size > 1 else np.float32(0)
        src = dst * float(scale)
    else:
        scale = np.float32(in_size) / np.float32(out_size)
        src = (dst + 0.5) * float(scale) - 0.5
        src = torch.clamp(src, min=0.0)
    i0 = src.floor().long().clamp(max=in_size - 1)
    i1 = torch.where(i0 < in_size - 1, i0 + 1, i0)
    w1 = src - i0.float()
    return i0, i1, w1


def bilinear_upsample(x: torch.Tensor, out_hw: Tuple[int, int], align_corners: bool) -> torch.Tensor:
    """Explicit separable bilinear (checked against F.interpolate in the tests)."""
    H, W = out_hw
    y0, y1, wy = _lerp_table(H, x.shape[-2], align_corners)
    x0, x1, wx = _lerp_table(W, x.shape[-1], align_corners)
    if x.device.type != "cpu":   # the full-size GPU tests evaluate the oracle on device tensors
        y0, y1, wy, x0, x1, wx = (t.to(x.device) for t in (y0, y1, wy, x0, x1, wx))
    wy = wy.view(-1, 1)
    rows = x[..., y0, :] * (1 - wy) + x[..., y1, :] * wy
    return rows[..., x0] * (1 - wx) + rows[..., x1] * wx


def upsample_sem_logits(sem_logits: torch.Tensor, out_hw: Tuple[int, int]) -> torch.Tensor:
    """The network's last op (networks/deeplab_v3.py:154-160): F.interpolate(sem_logits, size=input_shape,
    mode="bilinear", align_corners=False) in fp32.  Oracle of the fused low-res path: every loss of
    this file applied to the result, gradients through it by autograd."""
    return bilinear_upsample(sem_logits.float(), out_hw, False)


def seen_probs(pen, protos, weight, bias, scale: int = 16) -> torch.Tensor:
    """get_seen_probs (bg_detector.py:141-165): sigmoid of the x16 align_corners=True
    up-sampled low-res logits, all T heads -> [B,T,H,W]."""
    z = seen_logits_lowres(pen, protos, weight, bias)
    return torch.sigmoid(bilinear_upsample(z, (z.shape[-2] * scale, z.shape[-1] * scale), True))


def seen_max(pen, protos, weight, bias, scale: int = 16) -> torch.Tensor:
    return seen_probs(pen, protos, weight, bias, scale).max(1)[0]


# --------------------------------------------------------------------------------------
# Row 7 -- seen-detector focal loss (base_loss.py:255-272 + smp FocalLoss binary)
# --------------------------------------------------------------------------------------
def focal_seen_loss(z_full: torch.Tensor, mask: torch.Tensor, gamma: float = 2.0,
                    alpha: Optional[float] = None, ignore_index: int = IGNORE) -> torch.Tensor:
    """z_full [B,1,H,W] logits of one head; target fg->1, bg->0, ignore dropped; returns 0
    when the batch has no background pixel (base_loss.py:260-262)."""
    if not bool((mask == 0).any()):
        return z_full.sum() * 0.0
    z = z_full.reshape(-1).float()
    y = mask.reshape(-1)
    keep = y != ignore_index
    z = z[keep]
    t = (y[keep] != 0).float()
    bce = torch.nn.functional.softplus(z) - t * z
    pt = torch.exp(-bce)
    loss = (1.0 - pt).pow(gamma) * bce
    if alpha is not None:
        loss = loss * (alpha * t + (1 - alpha) * (1 - t))
    return loss.mean()


# --------------------------------------------------------------------------------------
# Row 8 -- background-weighted unbiased CE (training/loss_utils.py:542-585), closed form
# --------------------------------------------------------------------------------------
def weighted_ce(logits: torch.Tensor, target: torch.Tensor, seen_max_prob: torch.Tensor,
                old_cl: int, gamma: float = 2.0, threshold: float = 0.5, ukd: bool = True,
                ignore_index: int = IGNORE) -> torch.Tensor:
    x = logits.float()
    lse = torch.logsumexp(x, dim=1)
    s = seen_max_prob.detach().float().clone()
    s[s > threshold] = 1.0
    is_ign = target == ignore_index
    is_bg = target == 0
    is_fg = ~is_bg & ~is_ign
    mod = (1.0 - s).pow(gamma)
    lse_fg = torch.logsumexp(x[:, 1:], dim=1)
    l1 = torch.where(is_bg, mod * (lse - x[:, 0]), torch.where(is_fg, lse - lse_fg, torch.zeros_like(lse)))
    is_old = (target < old_cl)                                       # includes bg; ignore (255) is never < old_cl
    y = target.clamp(max=x.shape[1] - 1)
    x_y = x.gather(1, y.unsqueeze(1)).squeeze(1)
    if ukd:
        old_term = lse - torch.logsumexp(x[:, :old_cl], dim=1)
    else:
        old_term = torch.zeros_like(lse)
    l2 = torch.where(is_old, old_term, torch.where(is_ign, torch.zeros_like(lse), lse - x_y))
    return (l1 + l2).mean()                                           # mean over ALL pixels (Q6)


# --------------------------------------------------------------------------------------
# Row 9 -- plain / class-weighted CE, per-image scoring
# --------------------------------------------------------------------------------------
def cross_entropy(logits, target, weight: Optional[torch.Tensor] = None,
                  ignore_index: int = IGNORE) -> torch.Tensor:
    """base_loss.py:237-240: sum_{y!=I} w_y (lse - x_y) / sum_{y!=I} w_y."""
    x = logits.float()
    K = x.shape[1]
    lse = torch.logsumexp(x, dim=1)
    keep = target != ignore_index
    y = torch.where(keep, target, torch.zeros_like(target))
    nll = lse - x.gather(1, y.unsqueeze(1)).squeeze(1)
    wy = torch.ones(K, device=x.device) if weight is None else weight.float().to(x.device)
    wpix = wy[y] * keep
    return (wpix * nll).sum() / wpix.sum()


def cross_entropy_per_image_score(logits, target, weight, ignore_index: int = IGNORE) -> torch.Tensor:
    """bacs_loss.py:183-189: -(w_y * nll) with reduction none, mean over H*W per image."""
    x = logits.float()
    lse = torch.logsumexp(x, dim=1)
    keep = target != ignore_index
    y = torch.where(keep, target, torch.zeros_like(target))
    nll = lse - x.gather(1, y.unsqueeze(1)).squeeze(1)
    per = weight.float().to(x.device)[y] * keep * nll
    return -per.view(x.shape[0], -1).mean(1)


# --------------------------------------------------------------------------------------
# Row 10 -- teacher distillation on the last attention map (bacs_loss.py:258-294)
# --------------------------------------------------------------------------------------
def teacher_distill(old_att: torch.Tensor, new_att: torch.Tensor, mask: torch.Tensor,
                    seen_max_prob: Optional[torch.Tensor], lkd: float = 0.25,
                    lkd_threshold: float = 0.5) -> torch.Tensor:
    m = mask == 0
    if seen_max_prob is not None:
        m = m & (seen_max_prob > lkd_threshold)
    H, W = mask.shape[-2:]
    uo = bilinear_upsample(old_att.float(), (H, W), False)
    un = bilinear_upsample(new_att.float(), (H, W), False)
    e = (uo * uo - un * un) * m.unsqueeze(1)
    return lkd * torch.linalg.vector_norm(e, 2.0, dim=-1).mean()   # sub-gradient 0 where a row is all-zero


# --------------------------------------------------------------------------------------
# Row 11 -- dark-experience-replay logit MSE with transplant (bacs_loss.py:387-431)
# --------------------------------------------------------------------------------------
def der_transplant_cut(n_classes: np.ndarray, K: int) -> np.ndarray:
    """cut[j] = first channel of replay sample j overwritten by the live logits.
    Follows the reference literally: for i, n in enumerate(unique(n_classes)) the *sample*
    touched is inverse[i] (Q5), not the samples whose class count is n."""
    n_classes = np.asarray(n_classes).astype(np.int64)
    uniq, inv = np.unique(n_classes, return_inverse=True)
    cut = np.full(n_classes.shape[0], K, dtype=np.int64)
    for i, n in enumerate(uniq.tolist()):
        j = int(inv[i])
        if n < K:
            cut[j] = min(cut[j], n)
    return cut


def der_mse(sem_logits: torch.Tensor, memory_logits: torch.Tensor, n_classes,
            ignore_rep_bg: bool = True, truncate: bool = True) -> torch.Tensor:
    """memory logits go through preprocess_batch's .long() (Q4) when ``truncate``."""
    s = sem_logits.float()
    m = memory_logits
    if truncate:
        m = m.long()
    m = m.float().clone()
    K = s.shape[1]
    cut = der_transplant_cut(np.asarray(n_classes), K)
    sd = s.detach()
    for j, c in enumerate(cut.tolist()):
        if c < K:
            m[j, c:] = sd[j, c:]
    if ignore_rep_bg:
        m[:, 0] = sd[:, 0]
    return ((m - s) ** 2).mean()


# --------------------------------------------------------------------------------------
# Row 13 -- MiB unbiased KD / unbiased CE (training/loss_utils.py:447-520)
# --------------------------------------------------------------------------------------
def unbiased_ce(logits, target, old_cl: int, ignore_index: int = IGNORE) -> torch.Tensor:
    x = logits.float()
    lse = torch.logsumexp(x, dim=1)
    keep = target != ignore_index
    y = torch.where(keep, target, torch.zeros_like(target))
    x_y = x.gather(1, y.unsqueeze(1)).squeeze(1)
    old_term = lse - torch.logsumexp(x[:, :old_cl], dim=1)
    per = torch.where(target < old_cl, old_term, lse - x_y) * keep
    return per.sum() / keep.sum()


def unbiased_kd(logits, old_logits, alpha: float = 1.0, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    x = logits.float()
    t = old_logits.float() * alpha
    Ko, K = t.shape[1], x.shape[1]
    lse = torch.logsumexp(x, dim=1)
    bkg = torch.logsumexp(torch.cat([x[:, :1], x[:, Ko:]], dim=1), dim=1) - lse
    q = torch.softmax(t, dim=1)
    per = (q[:, 0] * bkg + (q[:, 1:] * (x[:, 1:Ko] - lse.unsqueeze(1))).sum(1)) / Ko
    if mask is not None:
        per = per * mask.float()
    return -per.mean()


# --------------------------------------------------------------------------------------
# Rows 12/14 -- arg-max and confusion matrix / metrics
#   bacs_loss.py:255; training/metrics.py:38-88 over torchmetrics 0.6.0 ConfusionMatrix/IoU
# --------------------------------------------------------------------------------------
def class_sums(features: torch.Tensor, labels_down: torch.Tensor, K: int):
    """Per-class feature sums / pixel counts of the per-class prototype family (SDR, loss/sdr.py:120-159 sums the
    features of the pixels of each class present in the down-sampled labels): features [B,D,h,w], labels_down [B,h,w]
    -> (sums fp64 [K,D], counts int64 [K]); ids outside [0,K) are skipped."""
    B, D, h, w = features.shape
    f = features.double().permute(0, 2, 3, 1).reshape(-1, D)
    l = labels_down.reshape(-1).long()
    keep = (l >= 0) & (l < K)
    sums = torch.zeros(K, D, dtype=torch.float64, device=features.device)
    sums.index_add_(0, l[keep], f[keep])
    counts = torch.bincount(l[keep], minlength=K)[:K]
    return sums, counts


def class_distance(features: torch.Tensor, class_protos: torch.Tensor):
    """Optional per-class prototype family (no counterpart in the reference's arithmetic): squared
    Euclidean distance of every pixel feature [B,D,h,w] to every class prototype [Kc,D], evaluated directly as
    sum_d (f_d - c_d)^2 in fp64 (no |f|^2 + |c|^2 - 2 f.c expansion), and the nearest class (ties -> lowest)."""
    f = features.double().permute(0, 2, 3, 1)                       # [B,h,w,D]
    c = class_protos.double()                                       # [Kc,D]
    d2 = ((f.unsqueeze(3) - c) ** 2).sum(-1).permute(0, 3, 1, 2)    # [B,Kc,h,w]
    return d2, d2.argmin(1)


def argmax_first(logits: torch.Tensor) -> torch.Tensor:
    return logits.float().argmax(dim=1)


def confusion_matrix(preds, target, num_classes: int) -> np.ndarray:
    """rows = target, cols = prediction; only 0 <= target < K counted (metrics.py:45-50)."""
    t = np.asarray(target).reshape(-1).astype(np.int32).astype(np.int64)
    p = np.asarray(preds).reshape(-1).astype(np.int32).astype(np.int64)
    keep = (t >= 0) & (t < num_classes)
    idx = t[keep] * num_classes + p[keep]
    return np.bincount(idx, minlength=num_classes * num_classes).reshape(num_classes, num_classes)


def iou_metrics(confmat: np.ndarray) -> Dict[str, np.ndarray]:
    """metrics.py:52-88 (names swapped as in the reference: 'fn' = col sum - tp, 'fp' =
    row sum - tp) + torchmetrics' IoU-from-confmat with absent_score 0, reduction none."""
    c = torch.from_numpy(np.asarray(confmat, dtype=np.int64))
    tp = c.diagonal()
    fn = c.sum(0) - tp
    fp = c.sum(1) - tp
    total = c.sum()
    tn = total - (tp + fn + fp)

    def nz(v):
        v = v.clone()
        v[torch.isnan(v)] = 0
        return v

    acc = nz((tp + tn) / (tp + fp + fn + tn))
    prec = nz(tp / (tp + fp))
    rec = nz(tp / (tp + fn))
    spec = nz(tn / (tn + fp))
    union = c.sum(0) + c.sum(1) - tp
    iou = tp.float() / union.float()
    iou[union == 0] = 0.0
    return {"iou_per_class": iou.numpy(), "miou": iou.mean().numpy(), "accuracy": acc.numpy(),
            "precision": prec.numpy(), "recall": rec.numpy(), "specificity": spec.numpy()}


# --------------------------------------------------------------------------------------
# Whole step (rows 3-12) given network outputs -- used by tests, smoke() and the CPU
# baseline leg of bench.py.  Mirrors compute_loss at bacs_loss.py:212-256 with
# bg_weighted_ce=True, lkd>0 and (optionally) a replay batch.
# --------------------------------------------------------------------------------------
def bacs_step(logits, pen, old_att, new_att, mask, protos, counts, head_w, head_b, *,
              initial_classes: int, increment: int, old_cl: int, task_num: int,
              first_task: bool = False, epoch: int = 0, max_epochs: int = 30, gamma: float = 2.0,
              threshold: float = 0.5, ukd: bool = True, focal_gamma: float = 2.0,
              focal_alpha: Optional[float] = None, lkd: float = 0.25, lkd_threshold: float = 0.5,
              proto_mode: str = "exact", replay: Optional[dict] = None, alpha: float = 0.8,
              beta: float = 0.2, ignore_rep_bg: bool = True, nb_current_classes: Optional[int] = None):
    """Returns dict(loss, preds, protos, counts, seen_max).  Differentiable w.r.t. logits,
    new_att, head_w/head_b (and pen on the first task) through torch autograd."""
    T = protos.shape[0]
    sums, n = proto_accumulate(pen, mask, initial_classes, increment, T, mode=proto_mode)
    protos, counts = proto_update(protos, counts, sums, n)
    ready = prototypes_ready(counts)
    with torch.no_grad():
        smax = seen_max(pen.detach(), protos, head_w.detach(), head_b.detach())
    loss = weighted_ce(logits, mask, smax, old_cl, gamma, threshold, ukd)
    if ready:
        wt = max(0.0, 1.0 - math.exp(epoch - max_epochs))
        pen_in = pen if first_task else pen.detach()
        z = seen_logits_lowres(pen_in, protos[task_num:task_num + 1], head_w[task_num:task_num + 1],
                               head_b[task_num:task_num + 1])
        zf = bilinear_upsample(z, (z.shape[-2] * 16, z.shape[-1] * 16), True)
        loss = loss + wt * focal_seen_loss(zf, mask, focal_gamma, focal_alpha)
    if lkd > 0:
        loss = loss + teacher_distill(old_att, new_att, mask, smax, lkd, lkd_threshold)
    if replay is not None:
        K = logits.shape[1] if nb_current_classes is None else nb_current_classes
        cw = torch.zeros(K, device=logits.device)
        cw[(1 if ignore_rep_bg else 0):old_cl] = 1
        if beta != 0:
            rs, rn = proto_accumulate(replay["pen"], replay["mask"], initial_classes, increment, T,
                                      mode=proto_mode)
            protos, counts = proto_update(protos, counts, rs, rn)
            loss = loss + beta * cross_entropy(replay["logits"], replay["mask"], cw)
        if alpha != 0:
            loss = loss + alpha * der_mse(replay["sem_logits"], replay["memory_logits"],
                                          replay["n_classes"], ignore_rep_bg)
    preds = argmax_first(logits)
    return {"loss": loss, "preds": preds, "protos": protos, "counts": counts, "seen_max": smax}
