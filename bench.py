#!/usr/bin/env python
"""bench.py -- slices/s of the CTA-GAN hot path on B200 (BASELINE.json metric: train slices/s at 256^2 + conv tensor-pipe % of peak).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cyc|reg|hd|infer] [--impl ours|reference|cudnn]

Workloads (BASELINE.json `configs`):
  cyc    configs[1]  Cyc_Trainer full G/D step (2 generators + 2 PatchGAN discriminators, LSGAN + cycle L1), batch 1/GPU, 256x256  (default)
  reg    configs[2]  Reg_Trainer step (generator + Reg U-Net + warp + smoothness/L1 + PatchGAN D), batch 8/GPU, 256x256
  hd     configs[3]  Hd_Trainer_x2 step (Reg-GAN body + Discriminator_m/GANLoss + masked L1), batch 4/GPU, 512x512  (--hd-stage 1: Hd_Trainer_x1)
  infer  configs[4]  ResNet-9 generator forward (the reference's "pix2pix" generator), 512x512, batch --batch (default 16); --sweep: 1..256
One "step" = one iteration body of the trainer (all forward/backward passes and Adam updates), or one forward of one batch (infer).
N>1: torchrun, one rank per GPU, batch sharded by slice (weak scaling), NCCL gradient all-reduce (none for infer).

Prints ONE JSON line (rank 0):
  value         slices/s with inputs resident in HBM: CUDA events around K steps, max over ranks, median of `repeats` such regions (>= 1 s in all)
  e2e           the same through the trainer's public step from a pinned HOST batch: H2D inside the timed region, D2H of the loss (or of the
                generated slices for infer) every step
  roofline      the dominant kernel (3x3 256->256 res-block convolution + fused statistics, as the step runs it) timed INSIDE a dependent chain
                (conv -> norm_act_pad -> conv ...) with CUDA events; `isolated` = the same kernel timed alone
  roofline_hbm  the bandwidth kernels (norm_act_pad, norm backward, warp fwd/bwd) at the workload's shape against the measured HBM copy peak
  cpu_baseline  the oracle's restated reference step (PyTorch fp32 on the host cores), bounded sample
  library_bar   the reference's own op sequence on THIS GPU through PyTorch / cuDNN (eager fp32+TF32 and channels_last bf16 autocast, under a CUDA
                graph when it captures): the bar a hand-written path has to beat.  `--impl cudnn` prints it as its own arm.
`--impl reference` times the reference's CPU path (oracle port) as its own arm.
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

DEFAULT_BATCH = {"cyc": 1, "reg": 8, "hd": 4, "infer": 16}
DEFAULT_SIZE = {"cyc": 256, "reg": 256, "hd": 512, "infer": 512}
SWEEP = [1, 2, 4, 8, 16, 32, 64, 128, 256]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", choices=["cyc", "reg", "hd", "infer"], default="cyc")
    ap.add_argument("--impl", choices=["ours", "reference", "cudnn"], default="ours")
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--hd-stage", type=int, choices=[1, 2], default=2)
    ap.add_argument("--sweep", action="store_true", help="infer: also time batch 1..256 and add the table as `sweep`")
    ap.add_argument("--precision", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--no-graphs", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-bar", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    ap.add_argument("--min-seconds", type=float, default=1.0, help="total length of the timed regions (repeats of K steps)")
    args = ap.parse_args()
    args.size = args.size or DEFAULT_SIZE[args.workload]
    args.batch = args.batch or DEFAULT_BATCH[args.workload]
    return args


def workload_config(args):
    base = {"noise_level": 1, "port": 8097, "save_root": "", "image_save": "", "Adv_lamda": 1, "Cyc_lamda": 10, "Corr_lamda": 20,
            "Smooth_lamda": 10, "P2P_lamda": 100, "Adv_lamda1": 1, "Adv_lamda2": 0.1, "Corr_lamda1": 20, "Corr_lamda2": 2, "epoch": 0,
            "n_epochs": 1, "batchSize": args.batch, "lr": 1e-4, "lrd": 1e-4, "decay_epoch": 1, "size": args.size, "input_nc": 1, "output_nc": 1,
            "cuda": True, "n_cpu": 1, "precision": args.precision, "synthetic": True, "save_checkpoints": False, "log_every": 10 ** 9}
    base["name"] = {"cyc": "CycleGan", "reg": "RegGan", "hd": "HdGan", "infer": "P2p"}[args.workload]
    if os.environ.get("CTAGAN_FUSED_OPT") == "0":          # A/B switch: torch.optim.Adam + separate re-pack instead of the one-kernel optimiser
        base["fused_optimizer"] = False
    return base


def workload_name(args):
    b, s = args.batch, args.size
    return {"cyc": f"CycTrainer full G/D step (2 ResNet-9 G + 2 PatchGAN D, LSGAN + cycle L1), batch {b}/GPU, {s}x{s}",
            "reg": f"RegTrainer step (ResNet-9 G + Reg U-Net + warp + smoothness/L1 + PatchGAN D), batch {b}/GPU, {s}x{s}",
            "hd": f"HdTrainer stage-{args.hd_stage} step (ResNet-9 G + Reg U-Net + warp + smoothness/L1"
                  + (" + masked L1 + Discriminator_m/GANLoss" if args.hd_stage == 2 else " + PatchGAN D") + f"), batch {b}/GPU, {s}x{s}",
            "infer": f"ResNet-9 generator forward (p2pTrainer/CycTrainer test() hot loop), batch {b}/GPU, {s}x{s}"}[args.workload]


# conv GFLOP per slice per step (BASELINE.md section 3 / SURVEY.md 8d, minimal count) at 256^2; scales with the pixel count
GFLOP_PER_SLICE_256 = {"cyc": 1269.0, "reg": 488.0, "hd": 1957.0 / 4, "infer": 97.46}


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


def synthetic_batches(args, keys, seed, pool=8):
    """CT-like slices in [-1, 1] for the workload's input keys (pinned host tensors)."""
    from ctagan import trainers as TR
    return TR.SyntheticSlices(args.batch, args.size, 10 ** 9, seed, keys, pool=pool).batches


# ----------------------------------------------------------------------------------------------------------------------
# CPU reference arm (the oracle's restated reference step == the reference's own PyTorch ops on the host cores)
# ----------------------------------------------------------------------------------------------------------------------


def _oracle_step(args, device="cpu"):
    """(step(batch_index) -> None, slices per step) on the oracle's restated reference bodies."""
    from oracle import restate as R
    b, s = args.batch, args.size
    random.seed(42); torch.manual_seed(42)
    pairs = [R.synthetic_pair(b, s, seed=42 + i, phantom=True) for i in range(2)]
    if args.workload == "cyc":
        st = R.CycState(); return (lambda i: R.cyc_step(st, *pairs[i % 2])), b
    if args.workload == "reg" or (args.workload == "hd" and args.hd_stage == 1):
        st = R.RegState(); return (lambda i: R.reg_step(st, *pairs[i % 2])), b
    if args.workload == "hd":
        st = R.RegState(multiscale_d=True)
        return (lambda i: R.hd_x2_step(st, pairs[i % 2][0], (pairs[i % 2][1] * 1.7).clamp(-1, 1), pairs[i % 2][1])), b
    sd = R.init_generator(1, 1)

    def fwd(i):
        with torch.no_grad():
            R.generator_forward(sd, pairs[i % 2][0])
    return fwd, b


def cpu_reference_rate(args, budget_s, max_steps=None, warmup=1):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, b = _oracle_step(args)
    times = []
    t_start = time.perf_counter()
    for i in range(warmup):
        step(i)
    n = 0
    while True:
        t0 = time.perf_counter()
        step(n)
        times.append(time.perf_counter() - t0)
        n += 1
        if max_steps is not None and n >= max_steps:
            break
        if time.perf_counter() - t_start + times[-1] > budget_s:
            break
    total = sum(times)
    return {"value": n * b / total, "unit": "slices/s", "cores": cores, "kind": "port",
            "sample": f"{n} step(s) of the same workload (batch {b}, {args.size}x{args.size}) after {warmup} warm-up, fp32 PyTorch CPU, "
                      f"{total / n:.2f} s/step"}, n, total


def metric_name(args):
    return "infer slices/s" if args.workload == "infer" else "train slices/s"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = min(args.warmup, 1)
    res, n, total = cpu_reference_rate(args, budget_s=150.0, max_steps=args.steps, warmup=w)
    line = {"impl": "reference", "metric": metric_name(args), "value": res["value"], "unit": "slices/s", "n_gpus": args.gpus,
            "steps": n, "warmup": w, "ms_per_step": 1e3 * total / n, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": workload_name(args), "global_batch": args.batch, "parallelism": "cpu",
                       "note": "reference's CPU implementation of the path (restated iteration body on the reference's PyTorch ops), "
                               "rank 0 only, all host threads; steps bounded to ~150 s"},
            "cpu_baseline": res, "gpu_launches": 0,
            "e2e": {"value": res["value"], "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# library bar: the reference's op sequence on this GPU through PyTorch / cuDNN
# ----------------------------------------------------------------------------------------------------------------------


def _library_step_fn(args, autocast_bf16):
    """The reference iteration (oracle/restate.py forwards = the ATen/cuDNN ops the reference dispatches) on CUDA tensors, written
    without host syncs so that it can be captured in a CUDA graph.  The ReplayBuffer passes through (its first 50 pushes do)."""
    import contextlib
    from oracle import restate as R
    dev = "cuda"
    b, s = args.batch, args.size
    cl = torch.channels_last if autocast_bf16 else torch.contiguous_format
    ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if autocast_bf16 else contextlib.nullcontext

    def leaf(sd):
        return {k: (v.to(dev).to(memory_format=cl) if v.dim() == 4 else v.to(dev)).requires_grad_(True) for k, v in sd.items()}

    def adam(*sds):
        return torch.optim.Adam([p for sd in sds for p in sd.values()], lr=1e-4, betas=(0.5, 0.999), fused=True, capturable=True)

    random.seed(42); torch.manual_seed(42)
    a0, b0 = R.synthetic_pair(b, s, seed=42, phantom=True)
    rA, rB = a0.to(dev).to(memory_format=cl), b0.to(dev).to(memory_format=cl)
    mse = R.mse_vs_const
    if args.workload == "cyc":
        GA, DB, GB, DA = leaf(R.init_generator()), leaf(R.init_discriminator(1)), leaf(R.init_generator()), leaf(R.init_discriminator(1))
        oDB, oG, oDA = adam(DB), adam(GA, GB), adam(DA)

        def step():
            with ctx():
                oG.zero_grad(set_to_none=True)
                fB = R.generator_forward(GA, rA); fA = R.generator_forward(GB, rB)
                loss = (mse(R.discriminator_forward(DB, fB).float(), 1.0) + mse(R.discriminator_forward(DA, fA).float(), 1.0)
                        + 10 * R.l1_loss(R.generator_forward(GB, fB).float(), rA) + 10 * R.l1_loss(R.generator_forward(GA, fA).float(), rB))
            loss.backward(); oG.step()
            for D, o, real, fake in ((DA, oDA, rA, fA), (DB, oDB, rB, fB)):
                o.zero_grad(set_to_none=True)
                with ctx():
                    ld = mse(R.discriminator_forward(D, real).float(), 1.0) + mse(R.discriminator_forward(D, fake.detach()).float(), 0.0)
                ld.backward(); o.step()
            return loss
        return step
    if args.workload in ("reg", "hd"):
        multiscale = args.workload == "hd" and args.hd_stage == 2
        G, D, Rn = leaf(R.init_generator()), leaf(R.init_discriminator_m(1) if multiscale else R.init_discriminator(1)), leaf(R.init_reg(1, 1))
        oD, oR, oG = adam(D), adam(Rn), adam(G)
        rB1 = (rB * 1.7).clamp(-1, 1)

        def dloss(x, real):
            if multiscale:
                return R.gan_loss([[f.float() for f in sc] for sc in R.discriminator_m_forward(D, x)], real)
            return mse(R.discriminator_forward(D, x).float(), 1.0 if real else 0.0)

        def step():
            oR.zero_grad(set_to_none=True); oG.zero_grad(set_to_none=True)
            with ctx():
                fB = R.generator_forward(G, rA)
                tr = R.reg_forward(Rn, fB, rB).float()
                sr = R.warp(fB.float(), tr)
                loss = 10 * R.smoothing_loss(tr) + dloss(fB, True) + 20 * R.l1_loss(sr, rB)
                if multiscale:
                    loss = loss + 2 * R.masked_l1(sr, rB1, rB)
            loss.backward(); oR.step(); oG.step()
            oD.zero_grad(set_to_none=True)
            with ctx():
                with torch.no_grad():
                    fB = R.generator_forward(G, rA)
                ld = dloss(fB, False) + dloss(rB, True)
            ld.backward(); oD.step()
            return loss
        return step
    G = {k: (v.to(dev).to(memory_format=cl) if v.dim() == 4 else v.to(dev)) for k, v in R.init_generator().items()}

    def fwd():
        with torch.no_grad(), ctx():
            return R.generator_forward(G, rA).float().mean()
    return fwd


def _time_library(args, autocast_bf16, steps, warmup=3):
    step = _library_step_fn(args, autocast_bf16)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warmup):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graphed = False
    run = step
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        g.replay(); torch.cuda.synchronize()
        run, graphed = g.replay, True
    except Exception:                                    # noqa: BLE001 - capture is best effort; eager is the documented fallback
        torch.cuda.synchronize()
        step = _library_step_fn(args, autocast_bf16)
        for _ in range(2):
            step()
        run = step
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": args.batch / (ms * 1e-3), "ms_per_step": ms, "cuda_graph": graphed}


def library_bar(args, steps=10):
    """PyTorch / cuDNN on this GPU running the reference's op sequence (the oracle's functional restatement on CUDA tensors)."""
    torch.backends.cudnn.benchmark = True
    out = {"unit": "slices/s", "what": "the reference's ops through PyTorch " + torch.__version__ + " / cuDNN on this GPU, same workload, synthetic data"}
    for name, bf16 in (("eager_fp32_tf32", False), ("channels_last_bf16_autocast", True)):
        try:
            torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
            out[name] = _time_library(args, bf16, steps)
        except Exception as exc:                         # noqa: BLE001
            out[name] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        torch.cuda.synchronize(); torch.cuda.empty_cache()
    return out


def run_cudnn_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.cuda.set_device(0)
    bar = library_bar(args, steps=max(args.steps, 5))
    best = max((v for v in bar.values() if isinstance(v, dict) and "value" in v), key=lambda v: v["value"])
    line = {"impl": "cudnn", "metric": metric_name(args), "value": best["value"], "unit": "slices/s", "n_gpus": 1, "steps": max(args.steps, 5),
            "warmup": 3, "ms_per_step": best["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 autocast / fp32+tf32 (best of)", "data": "synthetic", "config": {"workload": workload_name(args)}, "library_bar": bar}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------


def graph_time(fn, reps=10, iters=10):
    """Seconds per call of fn(), timed with CUDA events around replays of a CUDA graph holding `reps` calls (no host launch overhead)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / (iters * reps)


def kernel_rooflines(args, peaks):
    """roofline (dominant conv, in-chain and isolated) and roofline_hbm (bandwidth kernels) at the workload's shape."""
    from ctagan import engine as E, lib as L, ops
    T = torch.bfloat16 if args.precision == "bf16" else torch.float32
    e = 2 if T == torch.bfloat16 else 4
    b, s = args.batch, args.size
    h = s // 4
    peak_tf = peaks.get("bf16_tflops", 1590.0)
    peak_sus = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_bw = peaks.get("hbm_gbs", 6650.0)
    src = "measured (MEASURED_PEAKS.json)" if "bf16_tflops" in peaks else "fallback (B200_PROFILING.md: 1.59 PFLOP/s, 6.65 TB/s)"
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(b, h + 2, h + 2, 256, device="cuda", generator=g).to(T)
    prims = [E.ConvPrim(torch.randn(256, 256, 3, 3, device="cuda", generator=g) * 0.02, None, 1, 0) for _ in range(2)]
    for p in prims:
        p.prepack(T)
    flops = 2.0 * b * h * h * 256 * 256 * 9

    def pool():
        return ops.ZeroPool(64 * b + 64, x.device) if T == torch.bfloat16 else None

    # -- the res-block chain exactly as the generator runs it: conv+stats -> norm/relu/pad -> conv+stats -> norm/+residual/pad --
    BLOCKS = 4

    def chain():
        X = x
        zp = pool()
        for _ in range(BLOCKS):
            ra, sa = prims[0].fprop_stats(X, zp)
            Tt = ops.norm_act_pad(ra, sa, L.ACT_RELU, 1)
            rb, sb = prims[1].fprop_stats(Tt, zp)
            X = ops.norm_act_pad(rb, sb, L.ACT_NONE, 1, res=X, res_pad=1)
        return X

    ra0, sa0 = prims[0].fprop_stats(x, pool())

    def chain_no_conv():
        X = x
        for _ in range(BLOCKS):
            Tt = ops.norm_act_pad(ra0, sa0, L.ACT_RELU, 1)
            X = ops.norm_act_pad(ra0, sa0, L.ACT_NONE, 1, res=X, res_pad=1)
        return X

    t_chain, t_rest = graph_time(chain, reps=3), graph_time(chain_no_conv, reps=3)
    t_conv_chain = max(t_chain - t_rest, 1e-9) / (2 * BLOCKS)
    t_conv_alone = graph_time(lambda: prims[0].fprop_stats(x, pool()), reps=20)
    t_conv_nostat = graph_time(lambda: prims[0].fprop(x, use_bias=False), reps=20)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = tr.get(f"conv_tc_b{b}_s{s}_{args.precision}", {}).get("dram_bytes")
    except (OSError, ValueError):
        pass
    achieved = flops / t_conv_chain / 1e12
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                "kernel": f"res-block 3x3 256->256 conv + fused InstanceNorm statistics (tcgen05 implicit GEMM M={b * h * h} N=256 K=2304), "
                          f"timed inside the generator's dependent chain (conv -> norm_act_pad -> conv -> norm_act_pad+residual, {BLOCKS} blocks, "
                          "CUDA-graph replays, chain minus the same chain without its convolutions), L2-warm",
                "us_per_launch": t_conv_chain * 1e6, "flops_per_launch": flops, "peak_source": src + ", burst",
                "frac_of_sustained": achieved / peak_sus,
                "isolated": {"with_statistics_us": t_conv_alone * 1e6, "tflops": flops / t_conv_alone / 1e12,
                             "without_statistics_us": t_conv_nostat * 1e6, "tflops_without_statistics": flops / t_conv_nostat / 1e12,
                             "note": "back-to-back replays of the one kernel overlap the tail of a launch with the head of the next"},
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from profiles/ncu_traffic.json (ncu --set full), null when "
                                "this shape has no committed capture"}

    # -- bandwidth kernels at the workload's shapes --
    hbm = []

    def add(name, bytes_, sec, note=""):
        hbm.append({"kernel": name, "bytes": bytes_, "us": sec * 1e6, "achieved": bytes_ / sec / 1e9, "peak": peak_bw, "unit": "GB/s",
                    "frac": bytes_ / sec / 1e9 / peak_bw, "note": note})

    Tn = b * h * h * 256
    resident = "tensor fits the 126 MB L2 (L2-resident, latency-bound at this size)" if Tn * e * 3 < 100e6 else "streams from HBM"
    add("norm_act_pad (InstanceNorm apply + ReLU + reflection pad, 256 ch)", 2 * Tn * e, graph_time(lambda: ops.norm_act_pad(ra0, sa0, L.ACT_RELU, 1)), resident)
    add("norm_act_pad + residual", 3 * Tn * e, graph_time(lambda: ops.norm_act_pad(ra0, sa0, L.ACT_NONE, 1, res=x, res_pad=1)), resident)
    gout = torch.randn(b, h + 2, h + 2, 256, device="cuda", generator=g).to(T)
    nb = int(L.load().ctagan_norm_act_pad_bwd_launches(1, h, h, 256, ops.dt(gout)))
    add(f"norm_act_pad backward ({'one cluster kernel' if nb == 1 else 'reduce + apply'}, zero-margined output)", (3 if nb == 1 else 5) * Tn * e,
        graph_time(lambda: ops.norm_act_pad_bwd(gout, ra0, sa0, L.ACT_RELU, 1, out_pad=2)), resident)
    if args.workload in ("reg", "hd") or True:
        P = b * s * s
        srcimg = torch.rand(b, 1, s, s, device="cuda", generator=g) * 2 - 1
        flow = torch.randn(b, 2, s, s, device="cuda", generator=g) * 1.5
        go = torch.randn(b, 1, s, s, device="cuda", generator=g) * 1e-4
        add("warp forward (Transformer_2D, fp32)", 4 * P * 4, graph_time(lambda: ops.warp_fwd(srcimg, flow)), "fp32 module boundary; 4*P*4 bytes")
        add("warp backward (gsrc + gflow, fixed-point scatter; 3 kernels + memset)", 7 * P * 4, graph_time(lambda: ops.warp_bwd(go, srcimg, flow)),
            "algorithmic 7*P*4 bytes; the deterministic scatter adds an 8-byte/pixel accumulator pass")
        add("L1 loss forward (single pass)", 2 * P * 4, graph_time(lambda: ops.l1_fwd(srcimg, go)), "")
    return roofline, hbm


def time_region(step, n_steps, barrier):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(n_steps):
        step(i)
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    sweep = None
    if args.workload == "infer":
        ctagan.set_precision(args.precision)
        net = ctagan.Generator(1, 1).cuda()
        net.prepack()
        host = synthetic_batches(args, ("A",), 42 + rank)
        dev = [b["A"].cuda() for b in host]
        out_host = torch.empty((args.batch, 1, args.size, args.size), dtype=torch.float32).pin_memory()
        static_in = torch.empty_like(dev[0])

        def make_runner(x_static):
            if args.no_graphs:
                def run():
                    with torch.no_grad():
                        return net(x_static)
                return run, None
            with torch.no_grad():
                side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(2):
                        net(x_static)
                torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
                n0 = ops.launch_count()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    y = net(x_static)
                nl = ops.launch_count() - n0

            def run():
                g.replay()
                return y
            return run, nl

        run, launches_per_step = make_runner(static_in)

        def step_device(i):
            static_in.copy_(dev[i % len(dev)], non_blocking=True)
            run()

        def step_host(i):
            static_in.copy_(host[i % len(host)]["A"], non_blocking=True)
            out_host.copy_(run(), non_blocking=True)
        h2d = d2h = args.batch * args.size * args.size * 4
        if launches_per_step is None:
            n0 = ops.launch_count(); step_device(0); launches_per_step = ops.launch_count() - n0
        data_keys = ("A",)
    else:
        cls = {"cyc": TR.Cyc_Trainer, "reg": TR.Reg_Trainer, "hd": TR.Hd_Trainer_x2 if args.hd_stage == 2 else TR.Hd_Trainer_x1}[args.workload]
        trainer = cls(cfg)
        data_keys = trainer.data_keys
        host = synthetic_batches(args, data_keys, 42 + rank)
        dev = [[b[k].cuda(non_blocking=True) for k in data_keys] for b in host]
        runner = GraphedTrainer(trainer, enabled=not args.no_graphs)
        host_loss = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        read_ev = [torch.cuda.Event() for _ in range(2)]
        sink = [0.0]

        def step_device(i):
            runner.step_device(dev[i % len(dev)])

        def step_host(i):
            # the loss of step i is copied to pinned host memory asynchronously and consumed one step later (the way a training loop logs
            # without stalling the device); every step's value is read inside the timed region
            losses = runner.step_host(host[i % len(host)])
            host_loss[i & 1].copy_(next(iter(losses.values())), non_blocking=True)
            read_ev[i & 1].record()
            if i > 0:
                read_ev[(i - 1) & 1].synchronize()
                sink[0] += float(host_loss[(i - 1) & 1])
        h2d, d2h = sum(host[0][k].numel() * 4 for k in data_keys), 4
        launches_per_step = None

    # ---- device-resident timing: repeats of K steps, >= min-seconds in all, median repeat reported ----------------------------
    W = max(args.warmup, 3)
    for i in range(W):
        step_device(i)
    barrier()
    if args.workload != "infer":
        launches_per_step = runner.launches_per_step() if runner.enabled else None
        if launches_per_step is None:
            n0 = ops.launch_count(); step_device(0); launches_per_step = ops.launch_count() - n0
    clocks = ClockSampler(local) if rank == 0 else None
    first = max_over_ranks(time_region(step_device, args.steps, barrier))
    repeats = [first]
    n_rep = int(min(max(3, -(-args.min_seconds * 1e3 // max(first, 1e-3))), 60))
    for _ in range(n_rep - 1):
        repeats.append(max_over_ranks(time_region(step_device, args.steps, barrier)))
    clock_info = clocks.stop() if clocks else None
    ms_total = sorted(repeats)[len(repeats) // 2]

    # ---- end to end: pinned host batch -> H2D -> step -> D2H, every step ----------------------------------------------------------
    for i in range(2):
        step_host(i)
    barrier()
    e2e_repeats = []
    for _ in range(max(3, min(n_rep, 10))):
        t = time_region(step_host, args.steps, lambda: (read_ev[(args.steps - 1) & 1].synchronize() if args.workload != "infer" else None, barrier())[-1])
        e2e_repeats.append(max_over_ranks(t))
    e2e_ms_total = sorted(e2e_repeats)[len(e2e_repeats) // 2]

    if args.workload == "infer" and args.sweep:
        sweep = []
        for bsz in SWEEP:
            try:
                xin = torch.rand(bsz, 1, args.size, args.size, device="cuda") * 2 - 1
                r, _ = make_runner(xin)
                for _ in range(2):
                    r()
                torch.cuda.synchronize()
                k = max(3, min(args.steps, int(2000 / bsz) + 1))
                ms = max_over_ranks(time_region(lambda i: r(), k, barrier))
                sweep.append({"batch": bsz, "slices_per_s": bsz * k * world / (ms * 1e-3), "ms_per_batch": ms / k})
                del r, xin
                torch.cuda.empty_cache()
            except RuntimeError as exc:                   # out of memory at the largest batches would show here
                sweep.append({"batch": bsz, "error": str(exc)[:120]})

    if rank != 0:
        _finish(world)
        return
    b, s = args.batch, args.size
    slices = args.steps * b * world
    value = slices / (ms_total * 1e-3)
    e2e_value = slices / (e2e_ms_total * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    roofline, roofline_hbm = kernel_rooflines(args, peaks)
    gflop = GFLOP_PER_SLICE_256[args.workload] * (s / 256.0) ** 2
    roofline["step_conv_tflops"] = gflop * slices / world / (ms_total * 1e-3) / 1e3
    roofline["step_frac_of_sustained"] = roofline["step_conv_tflops"] / peaks.get("bf16_tflops_sustained", 1400.0)
    line = {"metric": metric_name(args), "value": value, "unit": "slices/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": workload_name(args), "global_batch": b * world, "parallelism": f"dp{world}", "cuda_graphs": not args.no_graphs,
                       "repeats": len(repeats), "repeat_ms": [round(r, 3) for r in repeats[:12]], "timed_seconds": sum(repeats) * 1e-3,
                       "l2": "per-step working set (fp32 master weights + grads + Adam moments and the activations of the step, >0.4 GB) exceeds "
                             "the 126 MB L2; inputs rotate over 8 resident batches; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "slices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_total / args.steps,
                    "note": ("public step from a pinned host batch; the loss is read back every step through an async pinned copy consumed one step "
                             "later" if args.workload != "infer" else "pinned host batch -> H2D -> generator -> D2H of the generated slices, every step")},
            "gpu_launches": int(launches_per_step * args.steps), "clocks": clock_info, "roofline": roofline, "roofline_hbm": roofline_hbm}
    if sweep is not None:
        line["sweep"] = sweep
    if world == 1:
        if not args.no_cpu_baseline:
            line["cpu_baseline"], _, _ = cpu_reference_rate(args, budget_s=args.cpu_budget_s)
        if not args.no_library_bar:
            del dev
            torch.cuda.empty_cache()
            line["library_bar"] = library_bar(args)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    _finish(world)


def _finish(world):
    """Multi-rank runs end together (rank 0 still measures the kernel rooflines after the timed regions: the other ranks wait for it, a
    worker that exits early makes the launcher tear the job down) and with os._exit: tearing down a NCCL communicator that captured CUDA
    graphs still reference can block."""
    sys.stdout.flush(); sys.stderr.flush()
    if world > 1:
        import torch.distributed as dist
        try:
            dist.barrier()
        except Exception:                                # noqa: BLE001
            pass
        os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 (sm_100a); the product path has no CPU fallback. Use --impl reference for the CPU arm.")
    if args.impl == "cudnn":
        run_cudnn_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
