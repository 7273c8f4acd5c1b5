"""Cross-rank state of the BACS loss path: ONE packed fp64 all-reduce per step.

The reference never synchronises prototypes (each DDP rank keeps its own copy,
loss/prototypes.py:157-163) and lets torchmetrics all_gather the confusion matrix.  Here
each rank's per-task feature sums [T,D], pixel counts [T] and (at evaluation) the K x K
confusion matrix travel in one contiguous fp64 buffer -- integer counts stay exact below
2^53 -- through a single NCCL all-reduce over NVLink / NVSwitch, after which every rank
applies the identical running-mean update.  The payload is a few tens of KB, i.e.
latency-bound; there is no data-path collective anywhere else on this path."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_packed(sums: torch.Tensor, counts: torch.Tensor, group=None) -> None:
    """``sums`` [T,D] and ``counts`` [T] must be views of one contiguous fp64 buffer laid
    out [T*D | T] (ops.proto_accumulate allocates them that way): one collective moves both."""
    if world_size() == 1:
        return
    base = sums.untyped_storage().data_ptr()
    contiguous = (sums.dtype == torch.float64 and counts.dtype == torch.float64
                  and counts.data_ptr() == sums.data_ptr() + sums.numel() * 8
                  and sums.untyped_storage().data_ptr() == counts.untyped_storage().data_ptr() == base)
    if contiguous:
        packed = torch.empty(0, dtype=torch.float64, device=sums.device).set_(
            sums.untyped_storage(), sums.storage_offset(), (sums.numel() + counts.numel(),))
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    else:
        packed = torch.cat([sums.reshape(-1), counts.reshape(-1)])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        sums.copy_(packed[:sums.numel()].view_as(sums))
        counts.copy_(packed[sums.numel():])


def allreduce_state(sums: Optional[torch.Tensor], counts: Optional[torch.Tensor],
                    confmat: Optional[torch.Tensor], group=None) -> None:
    """Prototype sums / counts and an int64 confusion matrix in one fp64 all-reduce
    (bacs_pack_state / bacs_unpack_state do the int64 <-> fp64 conversion on the device)."""
    if world_size() == 1:
        return
    from . import ops
    device = (confmat if confmat is not None else sums).device
    packed = ops.pack_state(sums, counts, confmat, device)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    ops.unpack_state(packed, sums, counts, confmat)
