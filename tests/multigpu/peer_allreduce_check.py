"""Multi-GPU check of the peer-memory all-reduce (csrc/peer.cu) against NCCL.  Run with
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu/peer_allreduce_check.py
Every rank prints 'peer all-reduce ok' (needs >= 2 GPUs; tests/test_gpu_multi.py launches it when they exist)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from bacs_b200 import ops
    from bacs_b200.distributed import PeerReducer
    T, D = 6, 512
    n = T * D + T
    red = PeerReducer.create(64 * 2048 + 64, dev)
    assert red is not None, "symmetric memory unavailable"
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    proto = torch.randn(T, D, device=dev, generator=g)
    proto_ref = proto.clone()
    count = torch.full((T,), 1000.0, device=dev)
    count_ref = count.clone()
    for step in range(5):
        packed = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
        packed[T * D:] = torch.randint(0, 50, (T,), device=dev, generator=g).double()
        if step == 2:
            packed[T * D + 1] = 0                                   # a task without pixels on this rank
        want = packed.clone()
        dist.all_reduce(want)
        # the sum in rank order, the same on every rank
        parts = [torch.empty_like(packed) for _ in range(world)]
        dist.all_gather(parts, packed)
        exact = parts[0].clone()
        for p in parts[1:]:
            exact += p
        ready = red.allreduce(packed, proto, count, T, D)
        torch.cuda.synchronize()
        assert int(red.error) == 0, "a peer timed out"
        assert torch.equal(packed, exact), (rank, step, float((packed - exact).abs().max()))
        assert torch.allclose(packed, want, rtol=1e-12, atol=1e-12)
        ready_ref = ops.proto_update(proto_ref, count_ref, exact[:T * D].view(T, D), exact[T * D:])
        assert torch.equal(proto, proto_ref) and torch.equal(count, count_ref) and int(ready) == int(ready_ref)
    # prototypes must be identical on all ranks
    allp = [torch.empty_like(proto) for _ in range(world)]
    dist.all_gather(allp, proto)
    # (they start different per rank here, so only the update arithmetic is compared above)
    # CUDA-graph replay of the plain all-reduce
    buf = torch.zeros(n, dtype=torch.float64, device=dev)
    src = torch.full((n,), float(rank + 1), dtype=torch.float64, device=dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        buf.copy_(src)
        red.allreduce(buf)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        buf.copy_(src)
        red.allreduce(buf)
    for _ in range(20):
        graph.replay()
    torch.cuda.synchronize()
    assert int(red.error) == 0
    assert float(buf[0]) == world * (world + 1) / 2 and torch.equal(buf, torch.full_like(buf, world * (world + 1) / 2))
    # latency of the one-launch all-reduce + update vs NCCL all-reduce + update kernel
    def timed(fn, iters=200):
        for _ in range(20):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1000
    pk = torch.ones(n, dtype=torch.float64, device=dev)
    t_peer = timed(lambda: red.allreduce(pk, proto, count, T, D))
    def nccl():
        dist.all_reduce(pk)
        ops.proto_update(proto_ref, count_ref, pk[:T * D].view(T, D), pk[T * D:])
    pk.fill_(1.0)
    t_nccl = timed(nccl)
    print("rank %d: peer all-reduce ok (world %d): peer+update %.1f us, nccl+update %.1f us" % (rank, world, t_peer, t_nccl),
          flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
