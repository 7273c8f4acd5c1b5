#!/usr/bin/env python
"""bench.py -- BACS loss step throughput (pixels/s, forward + backward) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]

One "step" = one pass of the whole BACS loss path over one batch of
synthetic network outputs: prototype update, seen-head logits, fused weighted-CE + focal +
arg-max + distill-mask kernel with d(loss)/d(logits), teacher-distill fwd+bwd, head
backward -- called through the reference-facing class interface
``BACSLoss.compute_loss(batch, network, train=True)`` followed by ``loss.backward()``.
The network itself is out of scope: a stand-in returns fixed tensors.

Workload at N = 1: BASELINE.json configs[1] -- VOC 15-1 overlap step, 512x512 crops, B=24,
K=21, T=6 prototypes, D=512, A=256, bf16 logits/features.  For N > 1 every rank runs the
same per-GPU workload (weak scaling) and the per-task prototype sums/counts travel in one
packed fp64 NCCL all-reduce per step.

Prints ONE JSON line (rank 0).  --impl reference times the CPU oracle port of the
reference's algorithm (oracle/bacs_oracle.py; the reference is pure Python and does not
travel to the GPU box) on the host cores, on a bounded sample of the same workload."""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "BACS loss step pixels/sec (fwd+bwd)"
UNIT = "pixels/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="voc15-1_b24")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32", "fp16"])
    ap.add_argument("--cpu-sample-batch", type=int, default=2)
    ap.add_argument("--no-graph", action="store_true", help="do not also time the CUDA-graph replay of the step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fused-lowres", action="store_true",
                    help="skip the extra timing of BACSLoss(fused_logit_upsample=True) (low-res logits in, N=1 only)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((clk, util))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "note": "NVML unavailable"}
        loaded = sorted(c for c, u in self.samples if u > 0) or sorted(c for c, _ in self.samples)
        return {"sm_mhz": loaded[len(loaded) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def oracle_step_time(cfg, batch, steps, warmup, dtype_name):
    """CPU baseline: the oracle port of the reference algorithm, fwd + bwd, all host threads."""
    import torch
    from oracle import bacs_oracle as O
    from bacs_b200 import synth
    import dataclasses
    small = dataclasses.replace(cfg, B=batch, Br=0)
    inp = synth.make_step_inputs(small, seed=0, dtype=torch.float32, with_replay=False)

    def step():
        leaf = lambda t: t.clone().requires_grad_(True)
        lg, na, hw, hb = leaf(inp.logits), leaf(inp.new_att), leaf(inp.head_w), leaf(inp.head_b)
        out = O.bacs_step(lg, inp.pen, inp.old_att, na, inp.mask, inp.protos, inp.counts, hw, hb,
                          initial_classes=small.initial_classes, increment=small.increment, old_cl=small.old_cl,
                          task_num=small.T - 1, epoch=3, max_epochs=30)
        out["loss"].backward()
        return float(out["loss"])

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    px = small.B * small.H * small.W
    return px / dt, dt, "oracle port, fp32, B=%d of the %dx%d K=%d D=%d A=%d step, %d timed steps" % (
        small.B, small.H, small.W, small.K, small.D, small.A, steps)


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bacs_b200 import synth
    # torchrun exports OMP_NUM_THREADS=1: the reference arm uses every host core this process may run on
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    torch.set_num_threads(max(1, ncores))
    cfg = synth.CONFIGS[args.config]
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    value, dt, sample = oracle_step_time(cfg, args.cpu_sample_batch, steps, warmup, args.dtype)
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(cfg), "per_step_sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(cfg):
    return ("%s: BACS step B=%d K=%d old_cl=%d T=%d %dx%d (features %dx%d) D=%d A=%d" %
            (cfg.name, cfg.B, cfg.K, cfg.old_cl, cfg.T, cfg.H, cfg.W, cfg.h, cfg.w, cfg.D, cfg.A))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from bacs_b200 import _cabi, ops, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the BACS loss path has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()
    dtype = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[args.dtype]
    cfg = synth.CONFIGS[args.config]
    inp = synth.make_step_inputs(cfg, seed=rank, dtype=dtype)
    # multi-GPU runs use the decomposable per-channel prototype sums
    loss_fn, net, batch, leaves = synth.build_bacs_step(cfg, inp, device=dev, exact_prototypes=(world == 1))
    protos0 = loss_fn._prototypes._prototypes_tensors.clone()
    counts0 = loss_fn._prototypes._count_features.clone()
    pixels = cfg.B * cfg.H * cfg.W

    def step():
        for v in leaves.values():
            v.grad = None
        loss, preds = loss_fn.compute_loss(batch, net, train=True)
        loss.backward()
        return loss, preds

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- the headline: K steps through the public class interface, inputs resident in HBM
    n0 = lib.bacs_launch_count()
    step()
    launches_per_step = int(lib.bacs_launch_count() - n0)
    ms_eager = timed(step, args.steps, max(args.warmup, 3))

    # ---- the same step captured once in a CUDA graph (the step never synchronises with the host)
    ms_graph = None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            ms_graph = timed(graph.replay, args.steps, max(args.warmup, 3))
        except Exception as exc:                                  # noqa: BLE001
            sys.stderr.write("bench.py: CUDA-graph replay unavailable: %r\n" % (exc,))
            ms_graph = None
    ms_step = ms_eager if ms_graph is None else min(ms_eager, ms_graph)
    value = world * pixels / (ms_step * 1e-3)

    # ---- dominant kernel (fused pixel kernel) timed alone, back to back, on its launch stream
    es = torch.finfo(dtype).bits // 8
    with torch.no_grad():
        P = loss_fn._prototypes
        w, b = loss_fn._stack_heads(net.seen_fg_network, cfg.T)
        z = ops.seen_logits(leaves["pen"].detach(), P._prototypes_tensors, w, b)
        lg, mk = leaves["logits"].detach(), batch["main"][1] if isinstance(batch, dict) else batch[1]

        variant = {}

        def pixel_only():
            variant["v"] = ops.pixel_loss(lg, mk, _cabi.PIX_WEIGHTED_CE, want_grad=True, z=z, want_distill_mask=True,
                                          focal_head=cfg.T - 1, old_cl=cfg.old_cl)["variant"]
        ms_pix = timed(pixel_only, args.steps, max(args.warmup, 3))
    kernel_name = {2: "pixel_wce_kernel", 1: "pixel_fast_kernel", 0: "pixel_loss_kernel (shared-memory tiles)",
                   4: "pixel_stream_stats_kernel + pixel_stream_grad_kernel"}.get(
        variant.get("v"), "pixel kernel")
    alg_bytes = pixels * (2 * cfg.K * es + 8 + 8 + 1) + z.numel() * 4
    peak, peak_src = peaks()
    achieved = alg_bytes / (ms_pix * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_pixel_wce_traffic.json")
    if os.path.exists(tpath) and args.config == "voc15-1_b24" and args.dtype == "bf16":
        with open(tpath) as f:
            traffic = json.load(f).get("traffic_bytes_per_launch")   # dram read + write of one launch (ncu --set full)
    roofline = {"bound": "hbm", "kernel": kernel_name + " (weighted CE + focal + argmax + distill mask + dlogits)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms_pix,
                "share_of_step": ms_pix / ms_step}

    # ---- end to end: host (pinned) inputs -> device -> step -> loss back on the host
    e2e = None
    if not args.no_e2e:
        host = {k: getattr(inp, k).pin_memory() for k in ("logits", "pen", "old_att", "new_att", "mask")}
        h2d = sum(t.numel() * t.element_size() for t in host.values())
        img = batch["main"][0] if isinstance(batch, dict) else batch[0]
        loss_host = torch.empty((), dtype=torch.float32).pin_memory()

        # Two device-side input sets: the copy stream uploads step i+1 while the compute stream runs step i
        # (every step's inputs cross PCIe inside the timed region; the loss of every step is read back).
        copy_stream = torch.cuda.Stream()
        dbuf = [{k: torch.empty_like(t, device=dev) for k, t in host.items()} for _ in range(2)]
        uploaded = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        state = {"i": 0, "primed": False}

        def upload(slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])            # the step that last used this slot is done
                for k, t in host.items():
                    dbuf[slot][k].copy_(t, non_blocking=True)
                uploaded[slot].record(copy_stream)

        def e2e_step():
            cur = torch.cuda.current_stream()
            slot = state["i"] & 1
            if not state["primed"]:
                upload(slot)
                state["primed"] = True
            upload(slot ^ 1)                                       # next step's inputs, overlapping this step
            cur.wait_event(uploaded[slot])
            d = dbuf[slot]
            lgt = d["logits"].detach().requires_grad_(True)
            nat = d["new_att"].detach().requires_grad_(True)
            net.register(img, lgt, d["pen"], [nat])
            loss_fn.prev_model.register(img, d["logits"], d["pen"], [d["old_att"]])
            bt = [img, d["mask"]] if not isinstance(batch, dict) else dict(batch, main=[img, d["mask"]])
            loss, _ = loss_fn.compute_loss(bt, net, train=True)
            loss.backward()
            loss_host.copy_(loss.detach(), non_blocking=True)
            consumed[slot].record(cur)
            state["i"] += 1
            if state["i"] >= 2:
                consumed[slot ^ 1].synchronize()                   # bound the queue: at most two steps in flight
        for ev in consumed:
            ev.record(torch.cuda.current_stream())
        ms_e2e = timed(e2e_step, max(3, min(args.steps, 20)), 3)
        torch.cuda.synchronize()
        e2e = {"value": world * pixels / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e}
        net.register(img, leaves["logits"], leaves["pen"], [leaves["new_att"]])

    # ---- extra (not the headline): the opt-in fused path -- the network hands over its low-res sem_logits and the
    # x16 bilinear up-sample, the loss and the adjoint of the up-sample run in one kernel
    fused = None
    if world == 1 and not args.no_fused_lowres:
        try:
            loss2, net2, batch2, leaves2 = synth.build_bacs_step(cfg, inp, device=dev, fused_logit_upsample=True)
            gen = torch.Generator().manual_seed(7)
            sems = []
            for key, nb in (("main", cfg.B), ("buffer", cfg.Br)):
                if isinstance(batch2, dict) and key in batch2 or key == "main":
                    im = batch2[key][0] if isinstance(batch2, dict) else batch2[0]
                    sem = torch.randn(nb, cfg.K, cfg.h, cfg.w, generator=gen).to(dtype).to(dev).requires_grad_(True)
                    net2.register_sem(im, sem)
                    sems.append(sem)

            def step2():
                for v in list(leaves2.values()) + sems:
                    v.grad = None
                loss, preds = loss2.compute_loss(batch2, net2, train=True)
                loss.backward()
            ms2 = timed(step2, args.steps, max(args.warmup, 3))
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step2()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph2):
                step2()
            ms2g = timed(graph2.replay, args.steps, max(args.warmup, 3))
            fused = {"what": "BACSLoss(fused_logit_upsample=True): sem_logits [B,K,H/16,W/16] in, up-sample + loss + "
                             "adjoint in one kernel; no [B,K,H,W] logits or gradient exist",
                     "ms_per_step_eager": ms2, "ms_per_step_graph": ms2g,
                     "value": pixels / (min(ms2, ms2g) * 1e-3), "unit": UNIT,
                     "logit_bytes_per_step": int(sems[0].numel() * es * 2)}
            graph2 = None
        except Exception as exc:                                  # noqa: BLE001
            sys.stderr.write("bench.py: fused low-res timing unavailable: %r\n" % (exc,))

    sampler.stop_flag = True
    sampler.join(timeout=2)
    clocks = sampler.summary()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
        except AttributeError:
            pass
        v, dt, sample = oracle_step_time(cfg, args.cpu_sample_batch, 3, 1, args.dtype)
        cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample,
               "ms_per_step": dt * 1e3}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": workload_name(cfg), "global_batch": world * cfg.B,
                           "parallelism": "dp%d" % world,
                           "l2_policy": "inputs larger than L2 (logits %.0f MB per step, L2 126 MB)"
                                        % (cfg.B * cfg.K * cfg.H * cfg.W * es / 1e6),
                           "timed_path": "cuda-graph replay of BACSLoss.compute_loss+backward" if ms_step == ms_graph
                                         else "BACSLoss.compute_loss+backward (eager)"},
                "ms_per_step_eager": ms_eager, "ms_per_step_graph": ms_graph,
                "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "fused_lowres": fused}
        print(json.dumps(line), flush=True)
    if world > 1:
        # A captured graph that holds NCCL kernels must be gone before the communicator is torn down; the
        # teardown itself is skipped (it can block on a communicator that was used under capture) and the
        # ranks leave together after a final barrier.
        graph = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
