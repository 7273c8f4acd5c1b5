"""Multi-GPU check of the time-out path of the peer-memory all-reduce (csrc/peer.cu): rank 1 shows up later than the
time-out.  Rank 0 must leave its packed state, prototypes and counts untouched, report the exchange in the sticky error
word and PeerReducer.check() must raise; the late rank still completes its own exchange.  Run with
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu/peer_timeout_check.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    os.environ["BACS_PEER_TIMEOUT_MS"] = "1000"
    from bacs_b200.distributed import PeerReducer
    T, D = 6, 512
    red = PeerReducer.create(64 * 2048 + 64, dev)
    assert red is not None, "symmetric memory unavailable"
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    proto = torch.randn(T, D, device=dev, generator=g)
    count = torch.full((T,), 1000.0, device=dev)
    packed = torch.randn(T * D + T, dtype=torch.float64, device=dev, generator=g).abs()
    proto0, count0, packed0 = proto.clone(), count.clone(), packed.clone()
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 1:
        time.sleep(3.0)                                   # three times the time-out
    ready = red.allreduce(packed, proto, count, T, D)
    torch.cuda.synchronize()
    if rank == 0:
        assert int(red.error) != 0, "the missing peer went unnoticed"
        assert int(ready) == 0
        assert torch.equal(packed, packed0) and torch.equal(proto, proto0) and torch.equal(count, count0), \
            "a timed-out exchange must not touch the state"
        try:
            red.check()
        except RuntimeError as exc:
            assert "timed out" in str(exc)
        else:
            raise AssertionError("PeerReducer.check() did not raise")
        red.check()                                       # the error word is cleared by the raising check
    else:
        assert int(red.error) == 0                        # rank 0 had published its slot before giving up
        assert not torch.equal(proto, proto0)
    print("rank %d: peer time-out ok" % rank, flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
