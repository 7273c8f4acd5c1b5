"""Warm per-kernel times of the pre-pixel chain at the headline shape (label down-sample -> prototype sums ->
finalize/update -> seen logits), each entry point launched back to back, and the chain as a CUDA graph."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bacs_b200 import ops, synth


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def main():
    cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "voc15-1_b24"]
    inp = synth.make_step_inputs(cfg, seed=3, dtype=torch.bfloat16)
    dev = torch.device("cuda")
    pen, labels = inp.pen.to(dev), inp.mask.to(dev).long()
    B, D, h, w = pen.shape
    T = cfg.T
    lut = torch.full((256,), -1, dtype=torch.int32)
    for c in range(1, cfg.K):
        lut[c] = 0 if c <= cfg.initial_classes - 1 else 1 + (c - cfg.initial_classes) // cfg.increment
    lut = lut.to(dev)
    protos = inp.protos.to(dev).float().contiguous()
    counts = inp.counts.to(dev).float().contiguous()
    hw_, hb_ = inp.head_w.to(dev).float(), inp.head_b.to(dev).float()
    task, rank, n_bt, _ = ops.label_downsample_task(labels, h, w, lut, T)
    sums, cnts = ops.proto_accumulate(pen, task, rank, n_bt, T, 0)
    print("downsample_task   %.1f us" % timeit(lambda: ops.label_downsample_task(labels, h, w, lut, T)))
    print("proto_accumulate+finalize (exact) %.1f us" % timeit(lambda: ops.proto_accumulate(pen, task, rank, n_bt, T, 0)))
    print("proto_accumulate+finalize (chan)  %.1f us" % timeit(lambda: ops.proto_accumulate(pen, task, rank, n_bt, T, 1)))
    print("proto_update      %.1f us" % timeit(lambda: ops.proto_update(protos, counts, sums, cnts)))
    print("seen_logits       %.1f us" % timeit(lambda: ops.seen_logits(pen, protos, hw_.reshape(T, D), hb_.reshape(T))))

    def chain():
        t, r, n, _ = ops.label_downsample_task(labels, h, w, lut, T)
        s, c = ops.proto_accumulate(pen, t, r, n, T, 0)
        ops.proto_update(protos, counts, s, c)
        return ops.seen_logits(pen, protos, hw_.reshape(T, D), hb_.reshape(T))
    print("chain eager       %.1f us" % timeit(chain))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        chain()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            chain()
    print("chain graph       %.1f us" % timeit(g.replay))


main()
