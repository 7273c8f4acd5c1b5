"""Replay minibatch sampling: the reference's host memmap path (Buffer.get_data + H2D) vs the HBM-resident store."""
import os, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
os.environ["BACS_BUFFER_ROOT"] = tempfile.mkdtemp()
from bacs_b200.training.buffer import Buffer
from bacs_b200.training.device_store import DeviceReplayStore
N, K, H, W, Br = 300, 21, 512, 512, 24
rng = np.random.RandomState(0)
np.random.seed(0)
buf = Buffer(N, "all_tasks")
buf.update_task(task_num=0, new_class_size=K)
for i in range(N // 12):
    buf.add_data({"examples": torch.from_numpy(rng.rand(12, 3, H, W).astype(np.float32)),
                  "logits": torch.from_numpy(rng.randn(12, K, H // 16, W // 16).astype(np.float32)),
                  "labels": torch.from_numpy(rng.randint(0, K, size=(12, H, W)).astype(np.int64)),
                  "loss": torch.from_numpy(-rng.rand(12).astype(np.float32))})
buf.merge_scores()
t0 = time.perf_counter()
store = DeviceReplayStore(buf, "cuda", fields=("examples", "logits"))
torch.cuda.synchronize()
print("upload once: %.1f MB in %.0f ms" % (store.n_uploaded_bytes / 1e6, (time.perf_counter() - t0) * 1e3))
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3
host = timeit(lambda: buf.get_data(Br, device="cuda"))
dev = timeit(lambda: store.get_data(Br), 50)
print("replay minibatch Br=%d (examples %d MB + logits): host memmap + H2D %.2f ms | HBM-resident store %.3f ms (%.0fx)"
      % (Br, Br * 3 * H * W * 4 // 1000000, host, dev, host / dev))
