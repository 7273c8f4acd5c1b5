import torch, sys
sys.path.insert(0, '/root/repo')
from bacs_b200 import synth, ops
cfg = synth.CONFIGS["small"]
inp = synth.make_step_inputs(cfg, seed=2)
loss_fn, net, batch, leaves = synth.build_bacs_step(cfg, inp)
def mk(parts):
    def step():
        for v in leaves.values(): v.grad = None
        loss = 0
        if "main" in parts:
            l, p = loss_fn.compute_loss([batch["main"][0], batch["main"][1]], net, train=True) if False else (None, None)
        if "pp" in parts:
            data = ({"examples": batch["buffer"][0]}, batch["buffer"][0], None, batch["buffer"][1], None, None)
            loss = loss + loss_fn._dark_pp(net, data, _scale=0.2)
        if "der" in parts:
            data = ({}, batch["bufferlogits"][0], batch["bufferlogits"][1], None, batch["bufferlogits"][2], None)
            loss = loss + loss_fn._dark_logits(net, data, _scale=0.8)
        if "full" in parts:
            loss, _ = loss_fn.compute_loss(batch, net, train=True)
        loss.backward()
    return step
for parts in (["pp"], ["der"], ["full"]):
    step = mk(parts)
    try:
        step()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s): step()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): step()
        g.replay(); torch.cuda.synchronize()
        print(parts, "capture ok")
    except Exception as e:
        print(parts, "FAILED", repr(e)[:200])
        torch.cuda.synchronize()
