"""World-size-2 gloo test (CPU) of the N > 1 host logic: every rank contributes its
per-channel prototype sums / counts (and a confusion matrix) through ONE packed fp64
all-reduce; the merged update must equal the single-process result on the concatenated
batch.  The CUDA kernels are replaced by the oracle here -- only the
cross-rank protocol of bacs_b200.distributed is under test."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import bacs_oracle as O
    from bacs_b200 import synth
    from bacs_b200.distributed import allreduce_packed, world_size
    assert world_size() == world
    cfg = synth.CONFIGS["tiny"]
    inp = synth.make_step_inputs(cfg, seed=100 + rank)
    g = torch.Generator().manual_seed(7 + rank)
    mask = synth.make_labels(cfg, g, classes=list(range(1, cfg.K)))
    T, D = cfg.T, cfg.D
    s, n = O.proto_accumulate(inp.pen, mask, cfg.initial_classes, cfg.increment, T, mode="channel")
    packed = torch.zeros(T * D + T, dtype=torch.float64)
    sums, counts = packed[:T * D].view(T, D), packed[T * D:]
    sums.copy_(s.double())
    counts.copy_(n.double())
    allreduce_packed(sums, counts)                       # one collective for both
    # a confusion matrix travels the same way (exact below 2^53)
    cm = torch.full((4, 4), 2 ** 40 + rank, dtype=torch.int64)
    buf = cm.double()
    dist.all_reduce(buf)
    torch.save({"sums": sums.clone(), "counts": counts.clone(), "cm": buf.round().long(), "pen": inp.pen, "mask": mask},
               os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


def test_packed_allreduce_matches_single_process(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from oracle import bacs_oracle as O
    from bacs_b200 import synth
    cfg = synth.CONFIGS["tiny"]
    r = [torch.load(os.path.join(tmp_path, "rank%d.pt" % k)) for k in range(world)]
    # every rank ends with the same merged state
    assert torch.equal(r[0]["sums"], r[1]["sums"]) and torch.equal(r[0]["counts"], r[1]["counts"])
    # ... equal to the single-process per-channel sums over the concatenated batch
    pen = torch.cat([x["pen"] for x in r])
    mask = torch.cat([x["mask"] for x in r])
    s, n = O.proto_accumulate(pen, mask, cfg.initial_classes, cfg.increment, cfg.T, mode="channel")
    assert torch.equal(r[0]["counts"].long(), n)
    assert torch.allclose(r[0]["sums"].float(), s, rtol=1e-5, atol=1e-5 * float(s.abs().max()))
    # ... and to the reference's own (exact) mode evaluated image by image (B = 1 calls are per-channel)
    acc = torch.zeros_like(s)
    for b in range(pen.shape[0]):
        sb, _ = O.proto_accumulate(pen[b:b + 1], mask[b:b + 1], cfg.initial_classes, cfg.increment, cfg.T, mode="exact")
        acc += sb
    assert torch.allclose(acc, s, rtol=1e-5, atol=1e-5 * float(s.abs().max()))
    assert torch.equal(r[0]["cm"], torch.full((4, 4), 2 * 2 ** 40 + 1, dtype=torch.int64))
