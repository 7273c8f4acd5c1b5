import cProfile, pstats, sys, torch
sys.path.insert(0, '/root/repo')
from bacs_b200 import synth
cfg = synth.CONFIGS["voc15-1_b24"]
inp = synth.make_step_inputs(cfg, seed=0, dtype=torch.bfloat16)
loss_fn, net, batch, leaves = synth.build_bacs_step(cfg, inp, device="cuda")
def step():
    for v in leaves.values(): v.grad = None
    loss, preds = loss_fn.compute_loss(batch, net, train=True)
    loss.backward()
for _ in range(20): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(200): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("cpu issue time per step %.1f us" % ((t1 - t0) / 200 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
