"""torch.autograd.Function wrappers of the fused forward+backward kernels.

Every Function computes its gradients in the forward launch (one read of the big tensors)
and keeps them for backward.  backward() multiplies the stored gradients by the upstream
gradient with a device-predicated in-place kernel that exits immediately when the upstream
gradient is 1 (the usual case without a GradScaler), so the step never synchronises with
the host and the [B,K,H,W] tensors are read once and written once.

Losses come back as fp32 tensors of shape [] attached to autograd, like the reference's."""
from __future__ import annotations

from typing import Optional

import torch

from . import _cabi, ops


_TWICE = ("bacs_b200: backward through this loss a second time: the gradient was computed in the forward launch and "
          "is scaled in place by the first backward; call the loss again instead of retain_graph=True")


def _take(ctx, name):
    """The gradient stored by forward, exactly once (torch raises for a second backward through a freed graph; a
    retained graph must not silently hand out None gradients)."""
    if getattr(ctx, "_consumed", False):
        raise RuntimeError(_TWICE)
    ctx._consumed = True
    value = getattr(ctx, name)
    setattr(ctx, name, None)
    return value


def _scaled(grad: Optional[torch.Tensor], g: torch.Tensor) -> Optional[torch.Tensor]:
    if grad is None:
        return None
    return ops.scale_inplace(grad, g.reshape(1))


# ---- side stream: the seen-head backward (16 us, small grid) runs next to the teacher-distill chain --------------
# PixelLossFunction.forward forks after the pixel kernel's reduction launch; whoever needs the head gradients
# (PixelLossFunction.backward) or ends the step's forward (BACSLoss.compute_loss) joins.  Everything the side
# kernel reads is kept alive until the join, so the caching allocator cannot hand it to a later main-stream op.
_SIDE = {}
_PENDING = {}
OVERLAP_HEAD_BACKWARD = True


def _side_stream(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    ent = _SIDE.get(key)
    if ent is None:
        side = torch.cuda.Stream(device=device)
        ent = (side, torch.cuda.Event(), torch.cuda.Event(), side.cuda_stream)
        _SIDE[key] = ent
    return key, ent


def join_side_stream():
    """Make the current stream wait for side-stream work forked by PixelLossFunction (no-op when none is pending)."""
    if not _PENDING:
        return
    for key in list(_PENDING):
        event, _keep = _PENDING.pop(key)
        torch.cuda.current_stream().wait_event(event)


class PixelLossFunction(torch.autograd.Function):
    """Fused per-pixel CE family (+ seen-detector focal term) + arg-max.

    forward(logits, features, head_weight, head_bias, labels, cfg) ->
        (loss [], preds int64 [B,H,W], distill_mask uint8 [B,H,W] or empty)
    Gradients: logits; head_weight / head_bias of the focal head; features only when
    cfg['features_grad'] (first task: stop_gradients is False, base_loss.py:265)."""

    @staticmethod
    def forward(ctx, logits, features, head_weight, head_bias, labels, cfg: dict):
        ctx.set_materialize_grads(False)     # no zero-filled "gradients" for the int64 preds / uint8 mask outputs
        mode = cfg["mode"]
        z = cfg.get("z")
        focal_head = cfg.get("focal_head", -1)
        want_grad = bool(cfg.get("want_grad", True)) and ctx.needs_input_grad[0]
        has_focal = z is not None and focal_head >= 0
        lowres = bool(cfg.get("lowres", False))      # logits are the network's low-res sem_logits (fused up-sample)
        B = logits.shape[0]
        H, W = (labels.shape[-2], labels.shape[-1]) if lowres else (logits.shape[2], logits.shape[3])
        scale = cfg.get("loss_scale", 1.0)
        over_wsum = mode != _cabi.PIX_WEIGHTED_CE
        main_coef = scale if over_wsum else scale / float(B * H * W)            # WEIGHTED_CE: mean over ALL pixels (Q6)
        # the reduction launch of the kernel also produces the focal normaliser and the loss scalar
        epilogue = {"ready": cfg.get("ready") if has_focal else None,
                    "focal_weight": cfg.get("focal_weight", 1.0) if has_focal else 0.0,
                    "loss_coef": main_coef, "over_wsum": over_wsum}
        out = ops.pixel_loss(
            logits, labels, mode, want_grad=want_grad, want_preds=True, z=z,
            want_distill_mask=bool(cfg.get("want_distill_mask", False)), focal_head=focal_head if has_focal else -1,
            class_w=cfg.get("class_w"), hist=cfg.get("hist"), old_cl=cfg.get("old_cl", 0), ukd=cfg.get("ukd", True),
            gamma=cfg.get("gamma", 2.0), threshold=cfg.get("threshold", 0.5),
            focal_gamma=cfg.get("focal_gamma", 2.0), focal_alpha=cfg.get("focal_alpha"),
            lkd_threshold=cfg.get("lkd_threshold", 0.5), ignore_index=cfg.get("ignore_index", 255),
            grad_scale=cfg.get("loss_scale", 1.0), seen_scale=cfg.get("seen_scale", 16),
            seen_max=cfg.get("seen_max"), epilogue=epilogue, lowres=lowres, gz=cfg.get("gz") if has_focal else None)
        acc = out["acc"]
        loss = out["loss"].reshape(())
        dweight = dbias = dfeat = None
        if has_focal and head_weight is not None and (ctx.needs_input_grad[2] or ctx.needs_input_grad[1]):
            want_df = bool(cfg.get("features_grad", False)) and ctx.needs_input_grad[1]
            proto_t = cfg["proto"][focal_head]
            hw_flat = head_weight.detach().reshape(-1).float().contiguous()
            # everything the side-stream kernel reads must exist BEFORE the fork event and stay alive until the join:
            # a channels_last / sliced feature map is made contiguous here, on the main stream
            features = features.contiguous()
            if OVERLAP_HEAD_BACKWARD and cfg.get("overlap", True):
                # launched on the side stream by handle (no current-stream switch: that costs ~30 us of CPU); the
                # outputs live in the current stream's pool and are only touched there after the join
                key, (side, fork, join, side_handle) = _side_stream(logits.device)
                join_side_stream()                                 # at most one fork in flight per device
                fork.record()
                side.wait_event(fork)
                dweight, dbias, dfeat = ops.seen_head_backward(features, proto_t, hw_flat, out["gz"],
                                                               out["focal_scale"], want_df, stream=side_handle)
                join.record(side)
                _PENDING[key] = (join, (features, proto_t, hw_flat, out["gz"], out["focal_scale"]))
            else:
                dweight, dbias, dfeat = ops.seen_head_backward(features, proto_t, hw_flat, out["gz"],
                                                               out["focal_scale"], want_df)
        ctx.grads = (out["dlogits"], dfeat, dweight, dbias)
        ctx.shapes = (None if head_weight is None else head_weight.shape, None if head_bias is None else head_bias.shape,
                      None if head_weight is None else head_weight.dtype)
        preds = out["preds"]
        dmask = out["distill_mask"] if out["distill_mask"] is not None else torch.empty(0, dtype=torch.uint8,
                                                                                        device=logits.device)
        ctx.mark_non_differentiable(preds, dmask)
        ctx.aux = {"acc": acc}
        return loss, preds, dmask

    @staticmethod
    def backward(ctx, g, _gp, _gm):
        dlogits, dfeat, dweight, dbias = _take(ctx, "grads")
        join_side_stream()
        if g is None:
            return None, None, None, None, None, None
        wshape, bshape, wdtype = ctx.shapes
        ops.scale_inplace_multi([dlogits, dfeat, dweight, dbias], g)   # one launch for all four gradients
        if dweight is not None:
            dweight = dweight.reshape(wshape).to(wdtype)
            dbias = dbias.reshape(bshape).to(wdtype)
        return dlogits, dfeat, dweight, dbias, None, None


class TeacherDistillFunction(torch.autograd.Function):
    """lkd * mean_{b,a,y} || m * (U(old)^2 - U(new)^2) ||_2 over x (bacs_loss.py:258-294)."""

    @staticmethod
    def forward(ctx, new_att, old_att, mask_u8, out_hw, lkd: float, addend=None):
        """``addend`` (optional fp32 scalar tensor, e.g. the loss terms computed before this one): the result is
        addend + distillation term, written by the kernel's own reduction launch (no separate add)."""
        B, A, h, w = new_att.shape
        H, W = out_hw
        coef = float(lkd) / float(B * A * H)
        _, dnew, loss = ops.teacher_distill(old_att.detach(), new_att.detach(), mask_u8, (H, W), coef,
                                            ctx.needs_input_grad[0], want_scaled=True,
                                            addend=None if addend is None else addend.detach())
        ctx.dnew = dnew
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        dnew = _take(ctx, "dnew")
        return _scaled(dnew, g), None, None, None, None, (g if ctx.needs_input_grad[5] else None)


class DerMseFunction(torch.autograd.Function):
    """alpha * MSE(transplanted memory logits, live low-res logits) (bacs_loss.py:387-431)."""

    @staticmethod
    def forward(ctx, sem_logits, memory_logits, cut, ignore_rep_bg: bool, truncate: bool, alpha: float):
        n = sem_logits.numel()
        coef = float(alpha) / float(n)
        total, dsem = ops.der_mse(sem_logits.detach(), memory_logits, cut, ignore_rep_bg, truncate, coef,
                                  ctx.needs_input_grad[0])
        ctx.dsem = dsem
        return ops.combine_scalars([(total, 0, coef)], sem_logits.device).reshape(())

    @staticmethod
    def backward(ctx, g):
        dsem = _take(ctx, "dsem")
        return _scaled(dsem, g), None, None, None, None, None


class UnbiasedKDFunction(torch.autograd.Function):
    """MiB unbiased KD, -mean(mask * per) (training/loss_utils.py:447-489); gradient to the new logits."""

    @staticmethod
    def forward(ctx, logits, old_logits, mask, alpha: float):
        B, _, H, W = logits.shape
        coef = 1.0 / float(B * H * W)
        total, dx = ops.unbiased_kd(logits.detach(), old_logits.detach(), mask, alpha, coef, ctx.needs_input_grad[0])
        ctx.dx = dx
        return ops.combine_scalars([(total, 0, -coef)], logits.device).reshape(())

    @staticmethod
    def backward(ctx, g):
        dx = _take(ctx, "dx")
        return _scaled(dx, g), None, None, None
