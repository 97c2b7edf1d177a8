#!/usr/bin/env python
"""bench.py -- train slices/s of the CTA-GAN hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cyc|reg] [--impl ours|reference]

Workload at N=1 (default): BASELINE.json configs[1] = Cyc_Trainer full G/D step (2 generators + 2 PatchGAN discriminators,
LSGAN + cycle L1), batch 1 per GPU, 256x256, bf16 activations / fp32 master weights.  `--workload reg` runs configs[2]
(Reg_Trainer, batch 8 per GPU).  One "step" = one iteration body of the trainer (all forward/backward passes, the three
Adam updates).  N>1: torchrun, one rank per GPU, batch sharded by slice (weak scaling), NCCL gradient all-reduce.

Prints ONE JSON line (rank 0).  `value` = slices/s with inputs resident in HBM (CUDA events, max over ranks);
`e2e` = the same through trainer.step(host batch) incl. pinned H2D copy and D2H read of the loss each step;
`roofline` = the dominant kernel (3x3 256->256 res-block convolution) timed live with CUDA events;
`cpu_baseline` = the oracle's restated reference step (PyTorch fp32 on the host cores) on a bounded sample.
`--impl reference` times that CPU path as its own arm.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cta-gan_b200"))

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", choices=["cyc", "reg"], default="cyc")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--precision", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--no-graphs", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    return ap.parse_args()


def workload_config(args):
    batch = args.batch or (1 if args.workload == "cyc" else 8)
    base = {"noise_level": 1, "port": 8097, "save_root": "", "image_save": "", "Adv_lamda": 1, "Cyc_lamda": 10, "Corr_lamda": 20,
            "Smooth_lamda": 10, "epoch": 0, "n_epochs": 1, "batchSize": batch, "lr": 1e-4, "decay_epoch": 1, "size": args.size,
            "input_nc": 1, "output_nc": 1, "cuda": True, "n_cpu": 1, "precision": args.precision, "synthetic": True,
            "save_checkpoints": False, "log_every": 10 ** 9}
    base["name"] = "CycleGan" if args.workload == "cyc" else "RegGan"
    return base


WORKLOAD_NAME = {"cyc": "CycTrainer full G/D step (2 ResNet-9 G + 2 PatchGAN D, LSGAN + cycle L1), batch {b}/GPU, {s}x{s}",
                 "reg": "RegTrainer step (ResNet-9 G + Reg U-Net + warp + smoothness/L1 + PatchGAN D), batch {b}/GPU, {s}x{s}"}
# conv GFLOP per slice per step (minimal count), at 256^2; scales with pixels
GFLOP_PER_SLICE_256 = {"cyc": 1269.0, "reg": 488.0}


class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(index),
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except (ValueError, IndexError):
                continue

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU reference arm (the oracle's restated reference step == the reference's own PyTorch modules on the host cores)
# ----------------------------------------------------------------------------------------------------------------------


def cpu_reference_rate(args, budget_s, max_steps=None, warmup=1):
    from oracle import restate as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = workload_config(args)
    b, s = cfg["batchSize"], cfg["size"]
    random.seed(42); torch.manual_seed(42)
    if args.workload == "cyc":
        st = R.CycState(); step = lambda a, bb: R.cyc_step(st, a, bb)
    else:
        st = R.RegState(); step = lambda a, bb: R.reg_step(st, a, bb)
    batches = [R.synthetic_pair(b, s, seed=42 + i, phantom=True) for i in range(2)]
    times = []
    t_start = time.perf_counter()
    for i in range(warmup):
        step(*batches[i % 2])
    n = 0
    while True:
        t0 = time.perf_counter()
        step(*batches[n % 2])
        times.append(time.perf_counter() - t0)
        n += 1
        if max_steps is not None and n >= max_steps:
            break
        if time.perf_counter() - t_start + times[-1] > budget_s:
            break
    total = sum(times)
    return {"value": n * b / total, "unit": "slices/s", "cores": cores, "kind": "port",
            "sample": f"{n} step(s) of the same workload (batch {b}, {s}x{s}) after {warmup} warm-up, fp32 PyTorch CPU, "
                      f"{total / n:.2f} s/step"}, n, total


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload_config(args)
    res, n, total = cpu_reference_rate(args, budget_s=150.0, max_steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": "train slices/s", "value": res["value"], "unit": "slices/s", "n_gpus": args.gpus,
            "steps": n, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / n, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME[args.workload].format(b=cfg["batchSize"], s=cfg["size"]),
                       "note": "reference's CPU implementation of the path (restated iteration body on the reference's PyTorch ops), "
                               "rank 0 only, all host threads"},
            "cpu_baseline": res, "gpu_launches": 0,
            "e2e": {"value": res["value"], "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------


def time_dominant_kernel(precision, batch, size, reps=20, iters=10):
    """The res-block 3x3 256->256 convolution (89% of the generator's FLOPs) at this workload's shape, timed alone with CUDA
    events around CUDA-graph replays of `reps` back-to-back launches (so host launch overhead is not in the number).
    Returns (seconds per launch, flops per launch)."""
    from ctagan import engine as E
    T = torch.bfloat16 if precision == "bf16" else torch.float32
    h = size // 4
    x = torch.randn(batch, h + 2, h + 2, 256, device="cuda").to(T)
    w = torch.randn(256, 256, 3, 3, device="cuda") * 0.02
    prim = E.ConvPrim(w, None, 1, 0)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            prim.fprop(x, use_bias=False)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            prim.fprop(x, use_bias=False)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / (iters * reps)
    flops = 2.0 * batch * h * h * 256 * 256 * 9
    return sec, flops


def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ctagan
    from ctagan import ops
    from ctagan import trainers as TR
    from ctagan.graphs import GraphedTrainer
    cfg = workload_config(args)
    random.seed(42 + rank); torch.manual_seed(42)
    trainer = (TR.Cyc_Trainer if args.workload == "cyc" else TR.Reg_Trainer)(cfg)
    loader = TR.SyntheticSlices(cfg["batchSize"], cfg["size"], 10 ** 9, 42 + rank, trainer.data_keys, pool=8)
    host_batches = loader.batches
    dev_batches = [[b[k].cuda(non_blocking=True) for k in trainer.data_keys] for b in host_batches]
    runner = GraphedTrainer(trainer, enabled=not args.no_graphs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        runner.step_device(dev_batches[i % len(dev_batches)])
    barrier()
    launches0 = ops.launch_count()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        runner.step_device(dev_batches[i % len(dev_batches)])
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    launches = runner.launches_per_step() * args.steps if runner.enabled else ops.launch_count() - launches0
    clock_info = clocks.stop() if clocks else None

    # ---- end to end: pinned host batch -> H2D -> step -> D2H of the loss, every step --------------------------------------
    # The loss of step i is copied to pinned host memory asynchronously and consumed one step later (the way a training loop logs
    # without stalling the device); every step's value is read inside the timed region, the last one before the closing sync.
    host_loss = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    read_ev = [torch.cuda.Event() for _ in range(2)]
    for i in range(2):                                        # warm the e2e path (H2D staging, pinned buffers)
        runner.step_host(host_batches[i % len(host_batches)])
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    sink = 0.0
    for i in range(args.steps):
        losses = runner.step_host(host_batches[i % len(host_batches)])
        host_loss[i & 1].copy_(next(iter(losses.values())), non_blocking=True)
        read_ev[i & 1].record()
        if i > 0:
            read_ev[(i - 1) & 1].synchronize()
            sink += float(host_loss[(i - 1) & 1])
    read_ev[(args.steps - 1) & 1].synchronize()
    sink += float(host_loss[(args.steps - 1) & 1])
    t1.record()
    barrier()
    ms2 = torch.tensor([t0.elapsed_time(t1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms_total = float(ms2)

    if rank != 0:
        _finish(world)
        return
    b, s = cfg["batchSize"], cfg["size"]
    slices = args.steps * b * world
    value = slices / (ms_total * 1e-3)
    e2e_value = slices / (e2e_ms_total * 1e-3)
    h2d = sum(host_batches[0][k].numel() * 4 for k in trainer.data_keys)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak_tf = peaks.get("bf16_tflops", 1590.0)
    peak_src = "measured burst (MEASURED_PEAKS.json)" if "bf16_tflops" in peaks else "fallback 1.59 PFLOP/s"
    ksec, kflops = time_dominant_kernel(args.precision, b, s)
    achieved = kflops / ksec / 1e12
    # DRAM bytes per launch of this kernel from `ncu --set full` (profiles/ncu_full_conv_tc_valid_r1_n{1,8}.raw.csv: dram__bytes_read.sum
    # + dram__bytes_write.sum, L2-warm as in the step): operands and output stay in the 126 MB L2; what the kernel moves is L2->SM traffic
    # (l1tex__m_xbar2l1tex_read_bytes.sum = 78 MB at b=1, 469 MB at b=8).  Only the two profiled shapes have a figure.
    traffic = {(1, 256): 234e3, (8, 256): 116e3}.get((b, s)) if args.precision == "bf16" else None
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                "traffic_note": "DRAM bytes per launch (ncu --set full, L2-warm); L2->SM operand bytes per launch: 78e6 at b=1, 469e6 at b=8",
                "kernel": "res-block 3x3 256->256 conv fprop (tcgen05 implicit GEMM M=%d N=256 K=2304), timed alone over CUDA-graph replays, L2-warm" % (b * (s // 4) ** 2),
                "peak_source": peak_src,
                "step_conv_tflops": GFLOP_PER_SLICE_256[args.workload] * (s / 256.0) ** 2 * b * args.steps / (ms_total * 1e-3) / 1e3}
    line = {"metric": "train slices/s", "value": value, "unit": "slices/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME[args.workload].format(b=b, s=s), "global_batch": b * world,
                       "parallelism": f"dp{world}", "cuda_graphs": runner.enabled,
                       "l2": "per-step working set (fp32 master weights + grads + Adam moments, >0.4 GB) exceeds the 126 MB L2; "
                             "inputs rotate over 8 resident batches; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "slices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms_total / args.steps,
                    "note": "trainer step from a pinned host batch; the loss is read back every step through an async pinned copy consumed one step later"},
            "gpu_launches": int(launches), "clocks": clock_info, "roofline": roofline}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"], _, _ = cpu_reference_rate(args, budget_s=args.cpu_budget_s)
    elif world > 1:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    _finish(world)


def _finish(world):
    """Multi-rank runs end with os._exit: tearing down a NCCL communicator that captured CUDA graphs still reference can block."""
    sys.stdout.flush(); sys.stderr.flush()
    if world > 1:
        os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200 (sm_100a); the product path has no CPU fallback. Use --impl reference for the CPU arm.")
        run_ours(args)


if __name__ == "__main__":
    main()
