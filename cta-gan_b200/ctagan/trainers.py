"""Trainer entry points of the reference (trainer/{Cyc,Reg,Hd,p2p}Trainer.py) re-hosted on the sm_100a kernels.

Kept from the reference: class names, `__init__(config)` with the Yaml keys, `train()`, `test()`, `update_learning_rate()`
(bugs included: CycTrainer.py:117-126 forgets optimizer_D_A; HdTrainer.py:163-164 never decays D), Adam(betas=(0.5, 0.999)),
the exact order of forward/backward/step calls of every iteration body, checkpoint file names.
Replaced: DICOM datasets / Visdom (not in scope) by a synthetic CT-like slice stream and an
every-N-steps stdout logger without per-iteration host syncs; `.cuda()` hard-coding by LOCAL_RANK-aware devices; and a
data-parallel gradient all-reduce (NCCL) when launched under torchrun.

Each trainer exposes `step(batch)` = one iteration body, which is what bench.py times and tests/ compare with the oracle.
"""
from __future__ import annotations

import contextlib
import itertools
import os
import threading
from typing import Dict, Iterable, List

import torch
import torch.distributed as dist

from . import engine as E
from . import nn as N
from . import ops


# ----------------------------------------------------------------------------------------------------------------------
# plumbing
# ----------------------------------------------------------------------------------------------------------------------


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class SyntheticSlices:
    """CT-like slices in [-1, 1] (air = -1), as trainer/datasets.py:74-82 produces; pinned host tensors, per-rank seed."""

    def __init__(self, batch: int, size: int, n_batches: int, seed: int, keys=("A", "B"), pool: int = 4):
        g = torch.Generator().manual_seed(seed)
        yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing="ij")
        disc = ((yy - size / 2) ** 2 + (xx - size / 2) ** 2) <= (0.4 * size) ** 2
        self.batches = []
        for _ in range(pool):
            a = torch.where(disc, torch.rand(batch, 1, size, size, generator=g) * 0.6 - 0.3, torch.full((batch, 1, size, size), -1.0))
            b = (torch.roll(a, shifts=(2, -3), dims=(2, 3)) + 0.02 * torch.randn(batch, 1, size, size, generator=g)).clamp_(-1, 1)
            d = {}
            for k in keys:
                src = a if k.startswith("A") else b
                if k == "B1":     # windowed copy of B2 (centre 50 / width 400 HU window, trainer/datasets.py:45-56)
                    hu = (b * 0.5 + 0.5) * 4095 - 1024
                    src = (((hu - (50 - 200)) / 400).clamp(0, 1) - 0.5) / 0.5
                d[k] = src.clone().pin_memory() if torch.cuda.is_available() else src.clone()
            self.batches.append(d)
        self.n_batches = n_batches

    def __len__(self):
        return self.n_batches

    def __iter__(self):
        for i in range(self.n_batches):
            yield self.batches[i % len(self.batches)]


# ---- f4: checkpoints are written by a background thread ---------------------------------------------------------------
_PENDING_SAVES: List[threading.Thread] = []


def wait_checkpoints():
    """Block until every checkpoint handed to a writer thread is on disk (called before a checkpoint is read and at the end of train())."""
    while _PENDING_SAVES:
        _PENDING_SAVES.pop().join()


class GradSync:
    """Data-parallel gradient averaging (weak scaling over slices; every op on the path is per-sample, so N ranks x b slices ==
    1 rank x N*b slices up to reduction order).  With the fused optimiser the gradients already live in one flat bucket per optimiser
    group, which IS the NCCL buffer: no flatten copy.  The bucket is cut into chunks of ~`chunk_mb` at parameter boundaries, and a
    chunk goes to the all-reduce (on a communication stream) AS SOON AS the last weight-gradient launch of the step into it has been
    enqueued: every ConvPrim knows how many launches it receives per step (learnt from the previous step: the schedule is static) and
    calls `_final` after the last one; the chunk's all-reduce waits for those launches' events only, so it runs under the rest of
    the backward pass (gradients complete in reverse layer order, the head layers' last).  `__call__` (right before the optimiser
    step) reduces whatever is left -- everything, on the very first step -- and makes the training stream wait for the collectives.
    All ranks enqueue the collectives in the same order because they run the same host code."""

    def __init__(self, source, chunk_mb: float = None):
        self.opt = source if hasattr(source, "grad_flat") else None
        self.params = None if self.opt is not None else [p for p in source]
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if chunk_mb is None:
            chunk_mb = float(os.environ.get("CTAGAN_DDP_CHUNK_MB", "128"))
        self.chunk = int(chunk_mb * (1 << 20) / 4)
        self._comm = None
        # Measured on 2 B200s (Cyc step, batch 1 per GPU): one all-reduce per optimiser after the backward 5.42 ms/step; chunks of 8-24 MB
        # sent while the backward is still running 5.48-5.54 ms (the NCCL kernels take SMs from a latency-bound backward and every
        # collective pays its own launch latency); Reg step (batch 8): no difference (16.03-16.09 ms).  Early chunks are therefore opt-in.
        self.early = os.environ.get("CTAGAN_DDP_EARLY", "0") != "0"
        self.chunks = []              # [start, end, [prims]]
        self._chunk_of = {}           # id(prim) -> chunk index
        self._left, self._done = [], []
        self.n_early = 0              # chunks that went to the all-reduce before the end of their backward pass (all steps)
        if self.opt is not None and self.world > 1 and self.opt.grad_flat.is_cuda:
            self.opt.grad_sync = self
            self._build_chunks()

    def _build_chunks(self):
        opt = self.opt
        self.chunks, self._chunk_of = [], {}
        start, prims = 0, []
        n = len(opt.params)
        bias_ids = {id(q.b) for q in opt.prims if q.b is not None}
        for k, p in enumerate(opt.params):
            prim = opt._prim_of.get(id(p))
            if prim is not None:
                prims.append(prim)
            end = opt._offs[k + 1]
            # (a layer's bias gradient is written by the same launches as its weight gradient: never cut a chunk between the two)
            if (end - start >= self.chunk and not (k + 1 < n and id(opt.params[k + 1]) in bias_ids)) or k == n - 1:
                self.chunks.append([start, end, prims])
                start, prims = end, []
        for ci, (_, _, prims) in enumerate(self.chunks):
            for prim in prims:
                self._chunk_of[id(prim)] = ci
                prim.final_hook = self._final
        self._prims_key = tuple(id(q) for q in opt.prims)
        self.arm()

    def arm(self):
        """Start of a backward pass (FusedAdam.zero_grad)."""
        if self.opt is None or self.world == 1 or not self.chunks:
            return
        if tuple(id(q) for q in self.opt.prims) != self._prims_key:      # the optimiser re-bound its layers (a network rebuilt its plan)
            self._build_chunks()
            return
        # a chunk is reduced early only when every layer in it announces its last launch (expected_writes known from the previous step)
        self._left = [len(prims) if prims and all(q.expected_writes is not None for q in prims) else -1 for _, _, prims in self.chunks]
        self._done = [False] * len(self.chunks)

    def _comm_stream(self):
        if self._comm is None:
            self._comm = ops.named_stream("ddp.comm")
        return self._comm

    def _final(self, prim):
        """The last weight-gradient launch of this step into `prim`'s slice has just been enqueued (its event is prim.grad_event)."""
        ci = self._chunk_of.get(id(prim))
        if ci is None or not self.early or not self._left:
            return
        if prim.grad_writes > prim.expected_writes:
            # this step launches more weight gradients into the layer than the previous one did: the chunk must not go early any more
            if self._done[ci]:
                raise RuntimeError("GradSync: a gradient chunk was all-reduced early and then written again (the step's schedule changed); "
                                   "run one step with CTAGAN_DDP_EARLY=0 semantics first or keep the schedule static")
            self._left[ci] = -1
            return
        if self._left[ci] < 0 or self._done[ci]:
            return
        self._left[ci] -= 1
        if self._left[ci] > 0:
            return
        start, end, prims = self.chunks[ci]
        comm = self._comm_stream()
        for q in prims:
            comm.wait_event(q.grad_event)
        with torch.cuda.stream(comm):
            self._all_reduce(self.opt.grad_flat[start:end])
        self._done[ci] = True
        self.n_early += 1

    def _all_reduce(self, flat):
        if dist.get_backend() == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)              # the division happens inside the collective
        else:                                                        # gloo (CPU tests) has no AVG
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.div_(self.world)

    def __call__(self):
        if self.world == 1:
            return
        if self.opt is not None:
            flat = self.opt.grad_flat
            if not flat.is_cuda:
                self._all_reduce(flat)
                return
            comm = self._comm_stream()
            cur = torch.cuda.current_stream()
            comm.wait_stream(cur)
            with torch.cuda.stream(comm):
                if self.chunks:
                    for ci, (start, end, _) in enumerate(self.chunks):
                        if not (self._done and self._done[ci]):
                            self._all_reduce(flat[start:end])
                    self._done = [True] * len(self.chunks)
                else:
                    for o in range(0, flat.numel(), self.chunk):
                        self._all_reduce(flat[o:o + self.chunk])
            cur.wait_stream(comm)
            return
        owners = [p for p in self.params if p.grad is not None]
        grads = [p.grad for p in owners]
        flat = torch.cat([g.reshape(-1) for g in grads])
        self._all_reduce(flat)
        # the averaged gradients stay in the flat buffer: .grad becomes a view of it (no copy back)
        for p, c, g in zip(owners, flat.split([g.numel() for g in grads]), grads):
            p.grad = c.view_as(g)


class _TrainerBase:
    name = "base"
    data_keys = ("A", "B")

    def __init__(self, config: dict):
        if not torch.cuda.is_available():
            raise RuntimeError("the CTA-GAN B200 trainers need a CUDA device (sm_100a); there is no CPU fallback")
        self.config = dict(config)
        c = self.config
        c.setdefault("precision", "bf16")
        c.setdefault("log_every", 50)
        c.setdefault("synthetic", False)           # opt-in: a run on the reference's own Yaml must never silently train on synthetic discs
        c.setdefault("synthetic_batches", 100)
        c.setdefault("save_checkpoints", not c["synthetic"])      # synthetic runs do not overwrite trained models unless asked to
        self.rank, self.world, self.local_rank = _dist_env()
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.device)
        E.set_precision(c["precision"])
        self.step_count = 0
        self.last_losses: Dict[str, torch.Tensor] = {}

    # -- helpers -------------------------------------------------------------------------------------------------------
    def _adam(self, params, lr, nets):
        """Adam(betas=(0.5, 0.999)) of the reference.  Default: ctagan.optim.FusedAdam (one kernel per group: update + both packed bf16
        weight layouts, gradients in a flat bucket the weight-gradient kernels write into).  `fused_optimizer: false` (and the grouped
        Cyc schedule) fall back to torch.optim.Adam(fused, capturable).  Either way the learning rate lives in a device tensor, so
        update_learning_rate() also reaches the kernels captured in a CUDA graph."""
        c = self.config
        if c.get("fused_optimizer", True) and (c.get("cyc_schedule") or os.environ.get("CTAGAN_CYC_SCHEDULE", "streams")) != "grouped":
            from .optim import FusedAdam
            return FusedAdam(params, float(lr), nets)
        return torch.optim.Adam(params, lr=torch.tensor(float(lr), dtype=torch.float32, device=self.device), betas=(0.5, 0.999), fused=True,
                                capturable=True)

    @staticmethod
    def _fused(opt):
        return hasattr(opt, "grad_flat")

    @staticmethod
    def _set_lr(opt, lr):
        for g in opt.param_groups:
            if torch.is_tensor(g["lr"]):
                g["lr"].fill_(float(lr))
            else:
                g["lr"] = lr

    def _alloc_inputs(self):
        c = self.config
        shape = (c["batchSize"], c["input_nc"], c["size"], c["size"])
        return {k: torch.empty(shape, dtype=torch.float32, device=self.device) for k in self.data_keys}

    def _loader(self, list_key="train_list", seed_offset=0, train=True):
        """`synthetic: true` (explicit opt-in) -> the synthetic CT-like stream; otherwise the slice list named by the Yaml key, read and
        pre-processed by the GPU input pipeline (ctagan.data; trainer/datasets.py:36-119).  There is no silent fallback between them."""
        c = self.config
        if c["synthetic"]:
            return SyntheticSlices(c["batchSize"], c["size"], c["synthetic_batches"], 42 + self.rank + seed_offset, self.data_keys)
        path = c.get(list_key)
        if not (path and os.path.exists(path)):
            raise FileNotFoundError(f"{list_key}={path!r} does not exist (set `synthetic: true` for the synthetic CT-like stream)")
        from .data import SliceListLoader
        return SliceListLoader(path, c["batchSize"], c["size"], self.data_keys, device=self.device, rank=self.rank, world=self.world,
                               augment=train and self.name == "CycleGan", noise_level=c.get("noise_level", 0), seed=42 + self.rank + seed_offset)

    def sync_replicas(self):
        """Data parallelism only exchanges gradients, so every rank must START from the same weights: broadcast rank 0's parameters
        (after construction and after every checkpoint load), as DistributedDataParallel does."""
        if self.world <= 1 or not dist.is_initialized():
            return
        with torch.no_grad():
            for m in self.__dict__.values():
                if isinstance(m, torch.nn.Module):
                    for p in m.parameters():
                        dist.broadcast(p.data, src=0)
        E.invalidate_weight_cache()

    def load_checkpoint(self, net, fname, required=True):
        """load_state_dict(torch.load(save_root + fname)) (CycTrainer.py:239 etc.); a missing file is an error unless required=False."""
        wait_checkpoints()
        path = os.path.join(self.config.get("save_root") or "", fname)
        if not os.path.exists(path):
            if required:
                raise FileNotFoundError(f"checkpoint {path} not found")
            return False
        net.load_state_dict(torch.load(path, map_location=self.device, weights_only=True))
        return True

    def load_batch(self, batch):
        """H2D copy of one batch into the preallocated device inputs (CycTrainer.py:80-82,140-141)."""
        for k in self.data_keys:
            self.inputs[k].copy_(batch[k], non_blocking=True)
        return [self.inputs[k] for k in self.data_keys]

    def _log(self, epoch, i, n):
        """Every `log_every` steps the losses of the step are copied to a pinned host buffer (asynchronously, behind an event) and
        printed once the copy has landed -- at the latest at the next log point -- so the hot loop never waits for the device
        (the reference's Logger.log reads every loss on every iteration: trainer/utils.py:76,92)."""
        if self.rank != 0:
            return
        self._flush_log(block=False)
        if self.step_count % self.config["log_every"] != 0 or not self.last_losses:
            return
        self._flush_log(block=True)                     # (a previous read-back that is still pending: print it first, in order)
        keys = list(self.last_losses)
        vals = torch.stack([v.detach().float().reshape(()) for v in self.last_losses.values()])      # the graph overwrites its outputs: copy now
        ev = None
        if vals.is_cuda:
            host = torch.empty(vals.shape, dtype=torch.float32, pin_memory=True)
            host.copy_(vals, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        else:
            host = vals
        self._log_pending = (epoch, i, n, keys, host, ev)

    def _flush_log(self, block: bool):
        pend = getattr(self, "_log_pending", None)
        if pend is None:
            return
        epoch, i, n, keys, host, ev = pend
        if ev is not None:
            if not block and not ev.query():
                return
            ev.synchronize()
        msg = " | ".join(f"{k}: {float(v):.4f}" for k, v in zip(keys, host.tolist()))
        print(f"[{self.name}] epoch {epoch} batch {i + 1}/{n} -- {msg}", flush=True)
        self._log_pending = None

    def _save(self, epoch, nets: Dict[str, torch.nn.Module], wait: bool = False):
        """Rank 0 writes `state_dict`s under the reference's file names (RegTrainer.py:225-240, HdTrainer.py:275-280) WITHOUT stalling the
        training loop: the parameters are copied to pinned host memory on a copy stream (the training stream only waits for that copy,
        so the next optimiser step cannot overwrite weights that are still being read), and a writer thread pickles them and renames
        the file into place when it is complete.  The files are what `torch.save(net.state_dict())` writes, with CPU tensors."""
        c = self.config
        if self.rank != 0 or not c.get("save_checkpoints") or not c.get("save_root"):
            return
        os.makedirs(c["save_root"], exist_ok=True)
        on_gpu = any(p.is_cuda for net in nets.values() for p in net.parameters())
        copy_stream = ops.named_stream("ckpt.copy") if on_gpu else None
        if on_gpu:
            copy_stream.wait_stream(torch.cuda.current_stream())
        jobs = []
        with (torch.cuda.stream(copy_stream) if on_gpu else contextlib.nullcontext()):
            for fname, net in nets.items():
                state = net.state_dict()
                host = type(state)()
                for k, v in state.items():
                    if v.is_cuda:
                        h = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                        h.copy_(v.detach(), non_blocking=True)
                    else:
                        h = v.detach().clone()
                    host[k] = h
                if hasattr(state, "_metadata"):
                    host._metadata = state._metadata
                jobs.append((os.path.join(c["save_root"], fname.format(st=str(epoch))), host))
        ev = None
        if on_gpu:
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            torch.cuda.current_stream().wait_event(ev)

        def work():
            if ev is not None:
                ev.synchronize()
            for path, host in jobs:
                torch.save(host, path + ".tmp")
                os.replace(path + ".tmp", path)

        t = threading.Thread(target=work, name="ctagan-checkpoint-writer")
        t.start()
        _PENDING_SAVES.append(t)
        if wait:
            wait_checkpoints()

    def train(self):
        """The reference's epoch loop (CycTrainer.py:128-236 etc.).  Every iteration is a CUDA-graph replay (config `cuda_graphs`, default
        true; the learning rates live in device tensors, so update_learning_rate() reaches the captured Adam kernels)."""
        from .graphs import GraphedTrainer
        c = self.config
        runner = GraphedTrainer(self, enabled=bool(c.get("cuda_graphs", True)), warmup=1, replay_first=False)
        for epoch in range(c["epoch"] + 1, c["n_epochs"] + 1 + c["decay_epoch"]):
            if epoch > c["n_epochs"]:
                self.update_learning_rate()
            loader = self._loader()
            for i, batch in enumerate(loader):
                runner.step_host(batch)
                self._log(epoch, i, len(loader))
            self._save(epoch, self.checkpoint_nets())
        self._flush_log(block=True)
        wait_checkpoints()

    test_checkpoint = "aa.pth"          # CycTrainer.py:239 / RegTrainer.py:245: `save_root + 'aa.pth'`

    @torch.no_grad()
    def test(self, loader=None, checkpoint=None):
        """The reference's test() (CycTrainer.py:238-360): load the generator checkpoint, `fake_B = netG_A2B(real_A)` per batch, then the
        evaluation metrics -- computed ON THE GPU by fused reduction kernels (ctagan.evaluate: window, 0.3-threshold masks, MAE / PSNR /
        UQI / SSIM, int16 conversion), batched, with one host read at the end instead of a `.cpu().numpy()` round trip per slice.
        Batch-split across ranks without any collective.  checkpoint: file name under save_root (default: the reference's),
        False to evaluate the weights in memory."""
        from .evaluate import Evaluator
        if checkpoint is not False:
            self.load_checkpoint(self.netG_A2B, checkpoint or self.config.get("test_checkpoint") or self.test_checkpoint)
        loader = loader or self._loader("test_list", seed_offset=1000, train=False)
        ev = Evaluator(self.device)
        ka, kb = self.data_keys[0], self.data_keys[-1]
        for batch in loader:
            a = batch[ka].to(self.device, non_blocking=True)
            b = batch[kb].to(self.device, non_blocking=True)
            ev.add(self.netG_A2B(a), b)
        out = ev.result()
        if self.rank == 0:
            print(f"[{self.name}] test: {out}", flush=True)
        return out


# ----------------------------------------------------------------------------------------------------------------------
# CycleGAN  (trainer/CycTrainer.py)
# ----------------------------------------------------------------------------------------------------------------------


class Cyc_Trainer(_TrainerBase):
    name = "CycleGan"

    def __init__(self, config):
        super().__init__(config)
        c = self.config
        dev = self.device
        self.netG_A2B = N.Generator(c["input_nc"], c["output_nc"]).to(dev)            # CycTrainer.py:64-71 creation order
        self.netD_B = N.Discriminator(c["input_nc"]).to(dev)
        self.optimizer_D_B = self._adam(self.netD_B.parameters(), c["lr"], [self.netD_B])
        self.netG_B2A = N.Generator(c["input_nc"], c["output_nc"]).to(dev)
        self.netD_A = N.Discriminator(c["input_nc"]).to(dev)
        self.optimizer_G = self._adam(itertools.chain(self.netG_A2B.parameters(), self.netG_B2A.parameters()), c["lr"], [self.netG_A2B, self.netG_B2A])
        self.optimizer_D_A = self._adam(self.netD_A.parameters(), c["lr"], [self.netD_A])
        self.MSE_loss, self.L1_loss = N.MSELoss(), N.L1Loss()
        self.inputs = self._alloc_inputs()
        self.target_real, self.target_fake = 1.0, 0.0
        from .replay import ReplayBuffer
        self.fake_A_buffer, self.fake_B_buffer = ReplayBuffer(), ReplayBuffer()
        def sync(opt, params):
            return GradSync(opt if self._fused(opt) else params)
        self._sync_G = sync(self.optimizer_G, itertools.chain(self.netG_A2B.parameters(), self.netG_B2A.parameters()))
        self._sync_DA, self._sync_DB = sync(self.optimizer_D_A, self.netD_A.parameters()), sync(self.optimizer_D_B, self.netD_B.parameters())
        self.sync_replicas()

    def update_learning_rate(self):
        c = self.config
        lr = c["lr"] - c["lr"] / c["decay_epoch"]
        for opt in (self.optimizer_D_B, self.optimizer_G):          # optimizer_D_A is skipped, as in CycTrainer.py:117-126
            self._set_lr(opt, lr)
        c["lr"] = lr

    def checkpoint_nets(self):
        return {"{st}.pth": self.netG_A2B, "netD_B_{st}.pth": self.netD_B, "netG_B2A_{st}.pth": self.netG_B2A,
                "netD_A_{st}.pth": self.netD_A}

    # At batch 1 every kernel is latency-bound (66-132 CTAs for a few microseconds), so the two independent chains of the generator
    # phase (A: G_A2B(real_A) -> D_B -> G_B2A(fake_B);  B: G_B2A(real_B) -> D_A -> G_A2B(fake_A)) run on two streams; autograd
    # replays each backward node on its forward stream, so the backward chains overlap the same way.
    # phase_G / phase_DD / step_two_phase are the iteration in the reference's serial order (generator phase, host ReplayBuffer, both
    # discriminator phases): the cross-check of phase_all, the overlapped one-program schedule that step() and the CUDA graph use.
    def _side_streams(self):
        if not hasattr(self, "_streams"):
            prio = int(os.environ.get("CTAGAN_CHAIN_PRIO", "-1"))      # the input-gradient chains outrank the wgrad lanes: 5.93 -> 5.77 ms
            self._streams = tuple(ops.named_stream(f"cyc.{n}", prio if n != "repack" else 0) for n in ("chainA", "chainB", "repack"))
        return self._streams[:2]

    def phase_G(self, real_A, real_B):
        c = self.config
        cur = torch.cuda.current_stream()
        self.optimizer_G.zero_grad(set_to_none=True)
        for net in (self.netG_A2B, self.netG_B2A, self.netD_A, self.netD_B):
            net.prepack()                      # no-op in steady state: phase_DD re-packs every network right after its optimizer step
        # Each generator is used once per chain.  autograd would sum the two gradient contributions of a parameter on ONE stream and make
        # that stream wait for the other chain's producer, which serialises the chains; so the second use of each generator goes through
        # twin leaves aliasing the same storage, and the two gradient sets are added once after the join (one multi-tensor kernel).
        # (with the fused optimiser the weight-gradient kernels accumulate the second use in place: no twins needed)
        fused = self._fused(self.optimizer_G)
        twins_B2A = None if fused else [p.detach().requires_grad_() for p in self.netG_B2A.parameters()]
        twins_A2B = None if fused else [p.detach().requires_grad_() for p in self.netG_A2B.parameters()]
        sA, sB = self._side_streams()
        sA.wait_stream(cur); sB.wait_stream(cur)
        with torch.cuda.stream(sA):
            fake_B = self.netG_A2B(real_A)                                                 # CycTrainer.py:144-146
            loss_GAN_A2B = c["Adv_lamda"] * self.MSE_loss(self.netD_B(fake_B, freeze=True), self.target_real)
            recovered_A = self.netG_B2A(fake_B, params=twins_B2A)                          # :153-154
            loss_cycle_ABA = c["Cyc_lamda"] * self.L1_loss(recovered_A, real_A)
            loss_A = loss_GAN_A2B + loss_cycle_ABA
        with torch.cuda.stream(sB):
            fake_A = self.netG_B2A(real_B)                                                 # :148-150
            loss_GAN_B2A = c["Adv_lamda"] * self.MSE_loss(self.netD_A(fake_A, freeze=True), self.target_real)
            recovered_B = self.netG_A2B(fake_A, params=twins_A2B)                          # :156-157
            loss_cycle_BAB = c["Cyc_lamda"] * self.L1_loss(recovered_B, real_B)
            loss_B = loss_GAN_B2A + loss_cycle_BAB
        cur.wait_stream(sA); cur.wait_stream(sB)
        for t in (loss_A, loss_B, fake_A, fake_B):
            t.record_stream(cur)
        loss_Total = loss_A + loss_B                                                       # :160-162
        loss_Total.backward()
        for net, twins in (() if fused else ((self.netG_A2B, twins_A2B), (self.netG_B2A, twins_B2A))):
            own, extra = [], []
            for p_, t_ in zip(net.parameters(), twins):
                if t_.grad is None:
                    continue
                if p_.grad is None:
                    p_.grad = t_.grad
                else:
                    own.append(p_.grad); extra.append(t_.grad)
            if own:
                torch._foreach_add_(own, extra)
        self._sync_G()
        self.optimizer_G.step()
        return fake_A.detach(), fake_B.detach(), loss_Total.detach()

    def phase_D(self, netD, opt, sync, real, fake, repack=True, step_after=None, packed_after=None):
        """One discriminator update (:165-178 / :182-197).  real and fake go through the network as ONE batch: InstanceNorm is
        per sample, so this is the same arithmetic as the reference's two passes with half the kernel launches."""
        c = self.config
        opt.zero_grad(set_to_none=True)
        B = real.shape[0]
        pred = netD(torch.cat([real, fake], 0))
        loss_real = c["Adv_lamda"] * self.MSE_loss(pred[:B], self.target_real)
        loss_fake = c["Adv_lamda"] * self.MSE_loss(pred[B:], self.target_fake)
        loss_D = loss_real + loss_fake
        loss_D.backward()
        sync()
        if step_after is not None:             # an event after which the master weights (biases are read in place) may change
            torch.cuda.current_stream().wait_event(step_after)
        if packed_after is not None and self._fused(opt):      # the fused step rewrites the PACKED weights too: wait for their last reader
            torch.cuda.current_stream().wait_event(packed_after)
        opt.step()
        if repack and not self._fused(opt):
            netD.prepack(force=True)           # off the generator phase's critical path (it only reads the frozen discriminators)
        return loss_D.detach()

    def phase_DD(self, real_A, fake_A, real_B, fake_B):
        """Both discriminator updates; they touch disjoint networks, so they run concurrently on the two side streams."""
        cur = torch.cuda.current_stream()
        self.netD_A.prepack(); self.netD_B.prepack()
        sA, sB = self._side_streams()
        sC = self._streams[2]
        sA.wait_stream(cur); sB.wait_stream(cur); sC.wait_stream(cur)
        with torch.cuda.stream(sC):            # the generators were just updated (phase_G): re-pack them beside the discriminator phases
            if not self._fused(self.optimizer_G):
                self.netG_A2B.prepack(force=True); self.netG_B2A.prepack(force=True)
        with torch.cuda.stream(sA):
            loss_D_A = self.phase_D(self.netD_A, self.optimizer_D_A, self._sync_DA, real_A, fake_A)
        with torch.cuda.stream(sB):
            loss_D_B = self.phase_D(self.netD_B, self.optimizer_D_B, self._sync_DB, real_B, fake_B)
        cur.wait_stream(sA); cur.wait_stream(sB); cur.wait_stream(sC)
        loss_D_A.record_stream(cur); loss_D_B.record_stream(cur)
        return loss_D_A, loss_D_B

    # The whole iteration as ONE stream-ordered program.  The ReplayBuffer's random decisions do not depend on the data
    # (replay.py), so the host makes them up front and the device applies them with index tensors; the discriminator updates
    # then need nothing from the host and fork off as soon as their fake exists: they run beside the cycle forward and the
    # generator backward (which only READS the packed discriminator weights) instead of after them.
    def plan_replay(self, batch_size):
        """Host half of the two push_and_pop calls (:170, :189), in the reference's order.  Returns one int64 CPU tensor [4, B]."""
        sa, da = self.fake_A_buffer.plan(batch_size)
        sb, db = self.fake_B_buffer.plan(batch_size)
        return torch.tensor([sa, da, sb, db], dtype=torch.int64)

    def phase_all(self, real_A, real_B, sel):
        """sel: int64 device tensor [4, B] from plan_replay().  Returns (loss_G, loss_D_A, loss_D_B)."""
        c = self.config
        cur = torch.cuda.current_stream()
        self.optimizer_G.zero_grad(set_to_none=True)
        for net in (self.netG_A2B, self.netG_B2A, self.netD_A, self.netD_B):
            net.prepack()                      # no-op in steady state: every network is re-packed right after its optimizer step
        sA, sB = self._side_streams()
        if not hasattr(self, "_d_streams"):
            dprio = int(os.environ.get("CTAGAN_DUPD_PRIO", "0"))
            self._d_streams = (ops.named_stream("cyc.updateDA", dprio), ops.named_stream("cyc.updateDB", dprio))
        sDA, sDB = self._d_streams
        for s_ in (sA, sB, sDA, sDB):
            s_.wait_stream(cur)
        with torch.cuda.stream(sA):
            fake_B = self.netG_A2B(real_A)                                                 # CycTrainer.py:144-146
            ev_fake_B = torch.cuda.Event(); ev_fake_B.record(sA)
            loss_GAN_A2B = c["Adv_lamda"] * self.MSE_loss(self.netD_B(fake_B, freeze=True), self.target_real)
            ev_DB_read = torch.cuda.Event(); ev_DB_read.record(sA)
            recovered_A = self.netG_B2A(fake_B)                          # :153-154
            loss_A = loss_GAN_A2B + c["Cyc_lamda"] * self.L1_loss(recovered_A, real_A)
        with torch.cuda.stream(sB):
            fake_A = self.netG_B2A(real_B)                                                 # :148-150
            ev_fake_A = torch.cuda.Event(); ev_fake_A.record(sB)
            loss_GAN_B2A = c["Adv_lamda"] * self.MSE_loss(self.netD_A(fake_A, freeze=True), self.target_real)
            ev_DA_read = torch.cuda.Event(); ev_DA_read.record(sB)
            recovered_B = self.netG_A2B(fake_A)                          # :156-157
            loss_B = loss_GAN_B2A + c["Cyc_lamda"] * self.L1_loss(recovered_B, real_B)
        cur.wait_stream(sA); cur.wait_stream(sB)
        for t in (loss_A, loss_B):
            t.record_stream(cur)
        loss_Total = loss_A + loss_B                                                       # :160-162
        # Weight gradients are collected after the whole backward (engine.deferred_weight_grads): the two input-gradient chains never
        # wait for their lagging wgrad lanes at network boundaries, and the two uses of each generator are summed once at the end.
        with E.deferred_weight_grads() as dgrads:
            loss_Total.backward()
            cur.wait_stream(sA); cur.wait_stream(sB)          # the backward nodes ran on their forward streams
            ev_G_backward = torch.cuda.Event(); ev_G_backward.record(cur)
            dgrads.flush()
        # The discriminator updates are ISSUED here -- after the generator backward in host order, because the backward looks the
        # packed discriminator weights up when it runs and an optimizer step issued earlier would mark them stale -- but their
        # streams only wait for the fakes, so on the device (and as CUDA-graph branches) they run beside the generator backward.
        with torch.cuda.stream(sDA):                                                       # :165-178
            sDA.wait_event(ev_fake_A)
            pooled_A = self.fake_A_buffer.apply(fake_A, sel[0], sel[1])
            loss_D_A = self.phase_D(self.netD_A, self.optimizer_D_A, self._sync_DA, real_A, pooled_A, repack=False, step_after=ev_DA_read,
                                    packed_after=ev_G_backward)
            if not self._fused(self.optimizer_D_A):
                sDA.wait_event(ev_G_backward)  # the generator backward was the last reader of the packed discriminator weights
                self.netD_A.prepack(force=True)
        with torch.cuda.stream(sDB):                                                       # :182-197
            sDB.wait_event(ev_fake_B)
            pooled_B = self.fake_B_buffer.apply(fake_B, sel[2], sel[3])
            loss_D_B = self.phase_D(self.netD_B, self.optimizer_D_B, self._sync_DB, real_B, pooled_B, repack=False, step_after=ev_DB_read,
                                    packed_after=ev_G_backward)
            if not self._fused(self.optimizer_D_B):
                sDB.wait_event(ev_G_backward)
                self.netD_B.prepack(force=True)
        self._sync_G()
        self.optimizer_G.step()
        if not self._fused(self.optimizer_G):
            self.netG_A2B.prepack(force=True); self.netG_B2A.prepack(force=True)
        cur.wait_stream(sDA); cur.wait_stream(sDB)
        for t in (loss_D_A, loss_D_B, fake_A, fake_B):
            t.record_stream(cur)
        return loss_Total.detach(), loss_D_A, loss_D_B

    def phase_all_grouped(self, real_A, real_B, sel):
        """sel: int64 device tensor [4, B] from plan_replay().  Returns (loss_G, loss_D_A, loss_D_B).

        The two generators (and the two discriminators) always work on independent inputs at the same time, so they run GROUPED:
        one batch holding both inputs, one kernel launch per layer for both networks (nn.grouped_generators).  At batch 1 a
        single network's kernels are latency-bound and fill part of the chip.  MEASURED (B200, batch 1, 256x256): 6.37 ms per step
        against 5.90 ms for the two-stream schedule of phase_all -- a grouped conv takes as long as two concurrent single ones
        (18.8 us vs 2 x 8.9 us effective), and the 1-channel layers, which are not grouped, serialise.  Kept as an option
        (config["cyc_schedule"] = "grouped") and as the cross-check of ctagan_conv_gather_grouped / ctagan_conv_wgrad_grouped."""
        c = self.config
        B = real_A.shape[0]
        cur = torch.cuda.current_stream()
        self.optimizer_G.zero_grad(set_to_none=True)
        for net in (self.netG_A2B, self.netG_B2A, self.netD_A, self.netD_B):
            net.prepack()                      # no-op in steady state: every network is re-packed right after its optimizer step
        if not hasattr(self, "_d_stream"):
            self._d_stream = ops.named_stream("cyc.updateD")
        sD = self._d_stream
        sD.wait_stream(cur)
        fakes = N.grouped_generators((self.netG_A2B, self.netG_B2A), torch.cat([real_A, real_B]))      # CycTrainer.py:144-150
        fake_B, fake_A = fakes[:B], fakes[B:]
        ev_fakes = torch.cuda.Event(); ev_fakes.record(cur)
        pred = N.grouped_discriminators((self.netD_B, self.netD_A), fakes, freeze=True)
        ev_D_read = torch.cuda.Event(); ev_D_read.record(cur)
        loss_GAN_A2B = c["Adv_lamda"] * self.MSE_loss(pred[:B], self.target_real)
        loss_GAN_B2A = c["Adv_lamda"] * self.MSE_loss(pred[B:], self.target_real)
        recovered = N.grouped_generators((self.netG_B2A, self.netG_A2B), fakes)                        # :153-157
        loss_cycle_ABA = c["Cyc_lamda"] * self.L1_loss(recovered[:B], real_A)
        loss_cycle_BAB = c["Cyc_lamda"] * self.L1_loss(recovered[B:], real_B)
        loss_Total = loss_GAN_A2B + loss_GAN_B2A + loss_cycle_ABA + loss_cycle_BAB                     # :160-162
        loss_Total.backward()
        ev_G_backward = torch.cuda.Event(); ev_G_backward.record(cur)
        # The discriminator updates are ISSUED here -- after the generator backward in host order, because the backward looks the
        # packed discriminator weights up when it runs and an optimizer step issued earlier would mark them stale -- but their
        # stream only waits for the fakes, so on the device (and as a CUDA-graph branch) they run beside the generator backward.
        with torch.cuda.stream(sD):                                                        # :165-178, :182-197
            sD.wait_event(ev_fakes)
            pooled_A = self.fake_A_buffer.apply(fake_A, sel[0], sel[1])
            pooled_B = self.fake_B_buffer.apply(fake_B, sel[2], sel[3])
            self.optimizer_D_A.zero_grad(set_to_none=True); self.optimizer_D_B.zero_grad(set_to_none=True)
            pd = N.grouped_discriminators((self.netD_A, self.netD_B), torch.cat([real_A, pooled_A, real_B, pooled_B]))
            loss_D_A = c["Adv_lamda"] * self.MSE_loss(pd[:B], self.target_real) + c["Adv_lamda"] * self.MSE_loss(pd[B:2 * B], self.target_fake)
            loss_D_B = (c["Adv_lamda"] * self.MSE_loss(pd[2 * B:3 * B], self.target_real)
                        + c["Adv_lamda"] * self.MSE_loss(pd[3 * B:], self.target_fake))
            (loss_D_A + loss_D_B).backward()
            self._sync_DA(); self._sync_DB()
            sD.wait_event(ev_D_read)           # the master biases are read in place by the generator phase's discriminator forward
            self.optimizer_D_A.step(); self.optimizer_D_B.step()
            sD.wait_event(ev_G_backward)       # the generator backward was the last reader of the packed discriminator weights
            self.netD_A.prepack(force=True); self.netD_B.prepack(force=True)
            loss_D_A, loss_D_B = loss_D_A.detach(), loss_D_B.detach()
        self._sync_G()
        self.optimizer_G.step()
        self.netG_A2B.prepack(force=True); self.netG_B2A.prepack(force=True)
        cur.wait_stream(sD)
        for t in (loss_D_A, loss_D_B):
            t.record_stream(cur)
        fakes.record_stream(sD)
        return loss_Total.detach(), loss_D_A, loss_D_B

    def step(self, batch=None, tensors=None):
        real_A, real_B = tensors if tensors is not None else self.load_batch(batch)
        sel = self.plan_replay(real_A.shape[0]).to(real_A.device, non_blocking=True)
        loss_G, loss_D_A, loss_D_B = self.phase_fn()(real_A, real_B, sel)
        self.step_count += 1
        self.last_losses = {"loss_G": loss_G, "loss_D_A": loss_D_A, "loss_D_B": loss_D_B}
        return self.last_losses

    def phase_fn(self):
        sched = self.config.get("cyc_schedule") or os.environ.get("CTAGAN_CYC_SCHEDULE", "streams")
        return self.phase_all_grouped if sched == "grouped" else self.phase_all

    def step_two_phase(self, batch=None, tensors=None):
        """The same iteration in the reference's serial order (generator phase, then both discriminator phases); kept as the
        cross-check of the overlapped schedule."""
        real_A, real_B = tensors if tensors is not None else self.load_batch(batch)
        fake_A, fake_B, loss_G = self.phase_G(real_A, real_B)
        fake_A = self.fake_A_buffer.push_and_pop(fake_A)                                   # :170
        fake_B = self.fake_B_buffer.push_and_pop(fake_B)                                   # :189
        loss_D_A, loss_D_B = self.phase_DD(real_A, fake_A, real_B, fake_B)
        self.step_count += 1
        self.last_losses = {"loss_G": loss_G, "loss_D_A": loss_D_A, "loss_D_B": loss_D_B}
        return self.last_losses


# ----------------------------------------------------------------------------------------------------------------------
# Reg-GAN  (trainer/RegTrainer.py) and the two Hd stages (trainer/HdTrainer.py)
# ----------------------------------------------------------------------------------------------------------------------


class Reg_Trainer(_TrainerBase):
    name = "RegGan"
    lam_corr, lam_adv, lr_d = "Corr_lamda", "Adv_lamda", "lr"

    def _make_D(self):
        return N.Discriminator(self.config["input_nc"])

    def __init__(self, config):
        super().__init__(config)
        c = self.config
        dev = self.device
        self.netG_A2B = N.Generator(c["input_nc"], c["output_nc"]).to(dev)              # RegTrainer.py:94-101
        self.netD_B = self._make_D().to(dev)
        self.optimizer_D_B = self._adam(self.netD_B.parameters(), c[self.lr_d], [self.netD_B])
        self.R_A = N.Reg(c["size"], c["size"], c["input_nc"], c["input_nc"]).to(dev)
        self.spatial_transform = N.Transformer_2D().to(dev)
        self.optimizer_R_A = self._adam(self.R_A.parameters(), c["lr"], [self.R_A])
        self.optimizer_G = self._adam(self.netG_A2B.parameters(), c["lr"], [self.netG_A2B])
        self.MSE_loss, self.L1_loss = N.MSELoss(), N.L1Loss()
        self.criterionGAN = N.GANLoss()
        self.inputs = self._alloc_inputs()
        self.target_real, self.target_fake = 1.0, 0.0
        if self._fused(self.optimizer_G):
            sr, sg = GradSync(self.optimizer_R_A), GradSync(self.optimizer_G)
            self._sync_R, self._sync_G = sr, sg
            self._sync_GR = lambda: (sr(), sg())
            self._sync_D = GradSync(self.optimizer_D_B)
        else:
            self._sync_GR = GradSync(itertools.chain(self.R_A.parameters(), self.netG_A2B.parameters()))
            self._sync_D = GradSync(self.netD_B.parameters())
        self.sync_replicas()

    def update_learning_rate(self):
        c = self.config
        lr = c["lr"] - c["lr"] / c["decay_epoch"]
        for opt in (self.optimizer_D_B, self.optimizer_R_A, self.optimizer_G):           # RegTrainer.py:150-161
            self._set_lr(opt, lr)
        c["lr"] = lr

    def checkpoint_nets(self):
        return {"netG_A2B_{st}.pth": self.netG_A2B, "R_A_{st}.pth": self.R_A, "netD_B_{st}.pth": self.netD_B}

    def _side_stream(self):
        if not hasattr(self, "_side"):
            self._side = ops.named_stream("reg.adversarial")
        return self._side

    def _adv_G(self, fake_B):
        return self.MSE_loss(self.netD_B(fake_B, freeze=True), self.target_real)

    def _loss_D(self, fake_B, real_B):
        """fake and real slices go through D as ONE batch (InstanceNorm is per sample: same arithmetic, half the launches)."""
        c = self.config
        B = fake_B.shape[0]
        pred = self.netD_B(torch.cat([fake_B, real_B], 0))
        return c[self.lam_adv] * self.MSE_loss(pred[:B], self.target_fake) + c[self.lam_adv] * self.MSE_loss(pred[B:], self.target_real)

    def _extra_G_losses(self, SysRegist_A2B, tensors):
        return None

    def step(self, batch=None, tensors=None):
        c = self.config
        tensors = tensors if tensors is not None else self.load_batch(batch)
        real_A, real_B = tensors[0], tensors[-1]
        self.optimizer_R_A.zero_grad(set_to_none=True)                                     # RegTrainer.py:173-187
        self.optimizer_G.zero_grad(set_to_none=True)
        for net in (self.netG_A2B, self.R_A, self.netD_B):      # one re-pack launch per network (weights changed last step)
            net.prepack()
        fake_B = self.netG_A2B(real_A)
        # two independent consumers of fake_B: the registration branch (Reg -> warp -> L1, smoothness) and the adversarial branch
        cur = torch.cuda.current_stream()
        side = self._side_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            adv_loss = c[self.lam_adv] * self._adv_G(fake_B)
        Trans = self.R_A(fake_B, real_B)
        SysRegist_A2B = self.spatial_transform(fake_B, Trans)
        SR_loss = c[self.lam_corr] * self.L1_loss(SysRegist_A2B, real_B)
        SM_loss = c["Smooth_lamda"] * N.smooothing_loss(Trans)
        extra = self._extra_G_losses(SysRegist_A2B, tensors)
        cur.wait_stream(side)
        adv_loss.record_stream(cur)
        toal_loss = SM_loss + adv_loss + SR_loss
        if extra is not None:
            toal_loss = toal_loss + extra
        with E.deferred_weight_grads() as dgrads:     # the input-gradient chain does not wait for the wgrad lanes between networks
            toal_loss.backward()
            dgrads.flush()
        self._sync_GR()
        self.optimizer_R_A.step()
        self.optimizer_G.step()

        self.optimizer_D_B.zero_grad(set_to_none=True)                                     # :189-198
        self.netG_A2B.prepack()                                  # the generator was just updated
        with torch.no_grad():
            fake_B = self.netG_A2B(real_A)
        loss_D_B = self._loss_D(fake_B, real_B)
        loss_D_B.backward()
        self._sync_D()
        self.optimizer_D_B.step()
        self.step_count += 1
        self.last_losses = {"SR_loss": SR_loss.detach(), "adv_loss": adv_loss.detach(), "SM_loss": SM_loss.detach(),
                            "toal_loss": toal_loss.detach(), "loss_D_B": loss_D_B.detach()}
        return self.last_losses


class Hd_Trainer_x1(Reg_Trainer):
    """Stage 1 of CTA-GAN (trainer/HdTrainer.py:94-280): the Reg-GAN body on the A2/B2 inputs with the Hd loss weights."""
    name = "HdGan_x1"
    data_keys = ("A2", "B1", "B2")
    lam_corr, lam_adv, lr_d = "Corr_lamda1", "Adv_lamda1", "lrd"
    test_checkpoint = "netG_A2B_x_3.pth"          # HdTrainer.py:285

    def checkpoint_nets(self):
        """HdTrainer.py:275-280: the `_x_` names that stage 2 loads (`netG_A2B_x_45.pth`, `R_A_x_45.pth`, :697-699)."""
        return {"netG_A2B_x_{st}.pth": self.netG_A2B, "R_A_x_{st}.pth": self.R_A, "netD_B_x_{st}.pth": self.netD_B}

    def update_learning_rate(self):
        c = self.config
        lr = c["lr"] - c["lr"] / c["decay_epoch"]
        for opt in (self.optimizer_R_A, self.optimizer_G):
            self._set_lr(opt, lr)
        for g in self.optimizer_D_B.param_groups:      # HdTrainer.py:163-164 writes a no-op key: D never decays
            g["lrd"] = c["lrd"] - c["lrd"] / c["decay_epoch"]
        c["lr"] = lr


class Hd_Trainer_x2(Hd_Trainer_x1):
    """Stage 2 (trainer/HdTrainer.py:605-803): Discriminator_m + GANLoss and the additional masked L1 (:726-735)."""
    name = "HdGan_x2"

    def _make_D(self):
        return N.Discriminator_m(self.config["input_nc"])

    def _adv_G(self, fake_B):
        return self.criterionGAN(self.netD_B(fake_B, freeze=True), True)

    def _loss_D(self, fake_B, real_B):
        c = self.config
        return c[self.lam_adv] * (self.criterionGAN(self.netD_B(fake_B), False) + self.criterionGAN(self.netD_B(real_B), True)) / 2

    def _extra_G_losses(self, SysRegist_A2B, tensors):
        real_B1, real_B2 = tensors[1], tensors[2]
        return self.config["Corr_lamda2"] * N.masked_l1_loss(SysRegist_A2B, real_B1, real_B2)

    stage1_epoch = 45

    def load_stage1(self):
        """HdTrainer.py:697-699: stage 2 starts from the stage-1 generator and registration network.  Missing files are an error (the
        reference crashes in torch.load); `stage1_required: false` in the config starts from the current weights, loudly."""
        c = self.config
        ep = c.get("stage1_epoch", self.stage1_epoch)
        required = c.get("stage1_required", True)
        ok = [self.load_checkpoint(net, f"{prefix}_x_{ep}.pth", required=required)
              for prefix, net in (("netG_A2B", self.netG_A2B), ("R_A", self.R_A))]
        if not all(ok) and self.rank == 0:
            print(f"[{self.name}] WARNING: stage-1 checkpoints *_x_{ep}.pth not found under {c.get('save_root')!r}: "
                  "stage 2 starts from the weights in memory", flush=True)
        self.sync_replicas()
        return all(ok)

    def train(self):
        self.load_stage1()
        super().train()


Hd_Trainer_x = Hd_Trainer_x1          # train.py:42: "change the name Hd_Trainer_x1/Hd_Trainer_x2 to Hd_Trainer_x"


# ----------------------------------------------------------------------------------------------------------------------
# pix2pix  (trainer/p2pTrainer.py)
# ----------------------------------------------------------------------------------------------------------------------


class P2p_Trainer(_TrainerBase):
    name = "P2p"

    def __init__(self, config):
        super().__init__(config)
        c = self.config
        dev = self.device
        self.netG_A2B = N.Generator(c["input_nc"], c["output_nc"]).to(dev)               # p2pTrainer.py:60-63
        self.netD_B = N.Discriminator(c["input_nc"] * 2).to(dev)
        self.optimizer_D_B = self._adam(self.netD_B.parameters(), c["lr"], [self.netD_B])
        self.optimizer_G = self._adam(self.netG_A2B.parameters(), c["lr"], [self.netG_A2B])
        self.MSE_loss, self.L1_loss = N.MSELoss(), N.L1Loss()
        self.inputs = self._alloc_inputs()
        self.target_real, self.target_fake = 1.0, 0.0
        fz = self._fused(self.optimizer_G)
        self._sync_G = GradSync(self.optimizer_G if fz else self.netG_A2B.parameters())
        self._sync_D = GradSync(self.optimizer_D_B if fz else self.netD_B.parameters())
        self.sync_replicas()

    def update_learning_rate(self):
        c = self.config
        lr = c["lr"] - c["lr"] / c["decay_epoch"]
        for opt in (self.optimizer_D_B, self.optimizer_G):
            self._set_lr(opt, lr)
        c["lr"] = lr

    def checkpoint_nets(self):
        return {"netG_A2B_{st}.pth": self.netG_A2B, "netD_B_{st}.pth": self.netD_B}

    def step(self, batch=None, tensors=None):
        c = self.config
        real_A, real_B = tensors if tensors is not None else self.load_batch(batch)
        self.optimizer_G.zero_grad(set_to_none=True)                                       # p2pTrainer.py:127-137
        fake_B = self.netG_A2B(real_A)
        loss_L1 = self.L1_loss(fake_B, real_B) * c["P2P_lamda"]
        pred_fake = self.netD_B(torch.cat((real_A, fake_B), 1), freeze=True)
        loss_GAN_A2B = self.MSE_loss(pred_fake, self.target_real) * c["Adv_lamda"]
        toal_loss = loss_L1 + loss_GAN_A2B
        toal_loss.backward()
        self._sync_G()
        self.optimizer_G.step()

        self.optimizer_D_B.zero_grad(set_to_none=True)                                     # :139-148 (prediction is scaled)
        with torch.no_grad():
            fake_B = self.netG_A2B(real_A)
        pred_fake0 = self.netD_B(torch.cat((real_A, fake_B), 1)) * c["Adv_lamda"]
        pred_real = self.netD_B(torch.cat((real_A, real_B), 1)) * c["Adv_lamda"]
        loss_D_B = self.MSE_loss(pred_fake0, self.target_fake) + self.MSE_loss(pred_real, self.target_real)
        loss_D_B.backward()
        self._sync_D()
        self.optimizer_D_B.step()
        self.step_count += 1
        self.last_losses = {"loss_L1": loss_L1.detach(), "loss_GAN_A2B": loss_GAN_A2B.detach(), "toal_loss": toal_loss.detach(),
                            "loss_D_B": loss_D_B.detach()}
        return self.last_losses
