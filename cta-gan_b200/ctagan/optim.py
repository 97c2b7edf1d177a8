"""Optimiser of the training iterations: torch.optim.Adam(lr, betas=(0.5, 0.999)) of the reference trainers
(CycTrainer.py:67-73, RegTrainer.py:97-101, HdTrainer.py:101-105, p2pTrainer.py:62-63) as ONE kernel launch per parameter group.

* Gradients live in one flat fp32 bucket per optimiser.  Every parameter's `.grad` is a persistent view into it, and the
  weight-gradient kernels write there directly (ConvPrim.attach_grads): no per-step gradient allocation, no `zero_grad` pass (the
  first use of a layer in a step overwrites its slice, a second use accumulates), no flatten copy before the data-parallel
  all-reduce -- the bucket IS the all-reduce buffer.
* `step()` launches `ctagan_adam_pack_multi`: Adam with the arithmetic of PyTorch's fused kernel (bit-identical, see
  tests/test_gpu_optim.py) on weights, gradients and both moments read once, plus both packed bf16 layouts of every convolution
  weight written from the same tile.  The learning rate and the step counter live on the device, so the launch is captured in the
  iteration's CUDA graph and `update_learning_rate()` reaches it.

Parameters that never receive a gradient (the biases in front of a non-affine InstanceNorm: mathematically dead) keep a zero
gradient slice: Adam leaves them untouched, exactly like torch.optim.Adam skipping `grad is None`."""
from __future__ import annotations

import ctypes
import os
from typing import Iterable, List, Sequence

import torch

from . import engine as E
from . import lib as L
from . import ops

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}


def _prims_of(nets) -> dict:
    """{id(weight parameter): ConvPrim} of the given networks' plans."""
    out = {}
    for net in nets:
        plans = []
        if hasattr(net, "_scale_plan"):
            plans = [net._scale_plan(i) for i in range(net.num_D)]
        elif hasattr(net, "_get_plan"):
            plans = [net._get_plan()]
        for plan in plans:
            for prim in plan.prims():
                out[id(prim.w)] = prim
    return out


class FusedAdam:
    """Adam over `params` (the reference's betas; eps 1e-8; no weight decay).  nets: the networks owning the parameters -- their
    ConvPrims get their gradient slices attached and their packed weights refreshed by the optimiser kernel."""

    def __init__(self, params: Iterable[torch.Tensor], lr: float, nets: Sequence[torch.nn.Module], betas=(0.5, 0.999), eps=1e-8):
        self.params: List[torch.Tensor] = [p for p in params]
        assert self.params and all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in self.params)
        dev = self.params[0].device
        self.device = dev
        self.betas, self.eps = betas, eps
        self.lr = torch.tensor(float(lr), dtype=torch.float32, device=dev)
        self.param_groups = [{"params": self.params, "lr": self.lr, "betas": betas, "eps": eps}]      # torch.optim surface the trainers use
        self.step_count = torch.zeros((), dtype=torch.float32, device=dev)
        self.ticket = torch.zeros((2,), dtype=torch.int32, device=dev)      # [0]: the (late / only) launch, [1]: the early launch
        sizes = [p.numel() for p in self.params]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + ((n + 3) & ~3))                   # 16-byte aligned slices
        self.grad_flat = torch.zeros((offs[-1],), dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.grad_flat)
        self.exp_avg_sq = torch.zeros_like(self.grad_flat)
        # Weights of the layers the tensor-core engine serves keep their gradient in channels-last order ([O][KH][KW][I] storage behind a
        # [O][I][KH][KW] view: what torch calls torch.channels_last): the weight-gradient kernel writes whole rows, the optimiser
        # kernel reads the tile it needs either way.  Everything else (biases, the 1-2 channel layers) stays in PyTorch order.
        self.packed = [self._wants_packed(p) for p in self.params]
        self.views = [self.grad_flat[o:o + n].view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2) if pk
                      else self.grad_flat[o:o + n].view_as(p) for p, o, n, pk in zip(self.params, offs, sizes, self.packed)]
        self._offs, self._sizes = offs, sizes
        for p, v in zip(self.params, self.views):
            p.grad = v
        # attach the gradient slices to the ConvPrims (weight + its bias)
        self.nets = list(nets)
        self._early_plans = []            # GeneratorPlans of this optimiser (set by _bind_prims)
        self._early_seen = {}             # id(plan) -> body_done calls so far in this step
        self._bind_prims()
        self._table_key = None
        self._dtype = None
        self.grad_sync = None             # trainers.GradSync (data parallel): armed by zero_grad, fed by the layers' final-write hooks
        # Early launch (single GPU): the residual blocks and the tail of a generator hold 97 % of its parameters, and nothing reads
        # their packed weights or adds to their gradients once the backward pass of the generator's LAST use in the step has passed
        # the blocks.  From then on their optimiser kernel can run under the rest of the backward (the head layers, the lagging
        # weight-gradient lanes); step() only launches the remainder.  How many passes a generator sees per step is learnt from the
        # previous step (the schedule is static); data-parallel runs keep the single launch (the gradients must be averaged first).
        # MEASURED (B200, same box): Cyc 4.626 ms with the early launch against 4.591 ms without, Hd 26.26 against 26.02 ms -- the
        # optimiser kernel takes SMs and HBM from the tail of the backward it was meant to hide under.  Opt-in (CTAGAN_EARLY_ADAM=1).
        self.early_enabled = os.environ.get("CTAGAN_EARLY_ADAM", "0") != "0"
        self._early_expected = {}         # id(plan) -> body_done calls per step (from the previous step)
        self._early_events = []
        self._early_launched = False
        self._early_stream = None
        self._tables = {}

    def _bind_prims(self):
        prims = _prims_of(self.nets)
        self._plans = [self._plans_of(net) for net in self.nets]
        by_id = {id(p): v for p, v in zip(self.params, self.views)}
        self.prims = []
        for p in self.params:
            prim = prims.get(id(p))
            if prim is not None:
                prim.attach_grads(by_id[id(p)], by_id.get(id(prim.b)) if prim.b is not None else None)
                self.prims.append(prim)
        self._prim_of = {id(prim.w): prim for prim in self.prims}
        self._early_plans = []
        for plans in self._plans:
            for plan in plans:
                if isinstance(plan, E.GeneratorPlan):
                    plan.body_done_hook = self._body_done
                    self._early_plans.append(plan)
        self._early_seen = {id(pl): 0 for pl in self._early_plans}

    @staticmethod
    def _wants_packed(p) -> bool:
        return p.dim() == 4 and p.shape[0] % 32 == 0 and p.shape[1] % 32 == 0 and p.shape[2] * p.shape[3] > 1

    @staticmethod
    def _plans_of(net):
        if hasattr(net, "_scale_plan"):
            return [net._scale_plan(i) for i in range(net.num_D)]
        return [net._get_plan()] if hasattr(net, "_get_plan") else []

    # ---- torch.optim surface -----------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = True):
        """Start of a backward pass: nothing is cleared (the first weight-gradient kernel of each layer overwrites its slice);
        the per-step write counters are re-armed."""
        if any(p.grad is None or p.grad.data_ptr() != v.data_ptr() for p, v in zip(self.params, self.views)):
            for p, v in zip(self.params, self.views):                # someone replaced .grad (e.g. a foreign zero_grad): re-attach
                p.grad = v
        cur = [self._plans_of(net) for net in self.nets]
        if any(a is not b for pa, pb in zip(cur, self._plans) for a, b in zip(pa, pb)):
            self._bind_prims()            # a network rebuilt its plan (e.g. after .to()): attach the gradient slices to the new ConvPrims
            self._table_key = None
        for prim in self.prims:
            if prim.grad_writes > 0:
                prim.expected_writes = prim.grad_writes        # the schedule of a step is static: what the last step did, this one will
            prim.grad_writes, prim.grad_event = 0, None
        if self.grad_sync is not None:
            self.grad_sync.arm()
        for pl in self._early_plans:
            if self._early_seen.get(id(pl), 0) > 0:
                self._early_expected[id(pl)] = self._early_seen[id(pl)]
            self._early_seen[id(pl)] = 0
        self._early_events, self._early_launched = [], False

    # ---- early launch ----------------------------------------------------------------------------------------------------------
    def _body_done(self, plan):
        """generator_backward of `plan` has enqueued everything that touches its residual blocks / tail (on the current stream or on
        lanes whose events the ConvPrims hold)."""
        k = id(plan)
        self._early_seen[k] = self._early_seen.get(k, 0) + 1
        if (not self.early_enabled or self._early_launched or self.grad_sync is not None or self._table_key is None
                or "early" not in self._tables or self._early_expected.get(k) != self._early_seen[k]
                or self._table_key[0] != E.get_precision() or self._table_key[1][0] != self.params[0].data_ptr()):
            return
        ev = torch.cuda.Event()
        ev.record()                                   # covers the input-gradient kernels of this pass (they read the packed weights)
        self._early_events.append(ev)
        if any(self._early_expected.get(id(pl)) != self._early_seen.get(id(pl)) for pl in self._early_plans):
            return                                    # another generator of this optimiser still has a pass to go
        early = [q for pl in self._early_plans for q in pl.early_prims()]
        if any(q.grad_event is None or q.expected_writes is None or q.grad_writes != q.expected_writes for q in early):
            return                                    # (a layer without its final gradient: leave everything to step())
        if self._early_stream is None:
            self._early_stream = ops.named_stream("opt.early")
        st = self._early_stream
        for ev in self._early_events:
            st.wait_event(ev)
        for q in early:
            st.wait_event(q.grad_event)               # the last weight-gradient launch into every early layer (on its lane)
        with torch.cuda.stream(st):
            self._launch("early", 0)
        self._early_launched = True

    def _build_table(self, dtype):
        """Three item tables: every parameter ("all"), the early set (weights of the residual blocks / tails of the generators and their
        biases) and the rest ("late")."""
        early_w = {id(q.w) for pl in self._early_plans for q in pl.early_prims()}
        early_b = {id(q.b) for pl in self._early_plans for q in pl.early_prims() if q.b is not None}
        is_early = [id(p) in early_w or id(p) in early_b for p in self.params]
        self._tables = {"all": self._build_one(dtype, [True] * len(self.params))}
        if any(is_early) and not all(is_early):
            self._tables["early"] = self._build_one(dtype, is_early)
            self._tables["late"] = self._build_one(dtype, [not e for e in is_early])
        self._dtype = dtype

    def _build_one(self, dtype, select):
        lib = L.load()
        chosen = [(p, o, sz, pk) for (p, o, sz, pk), s_ in zip(zip(self.params, self._offs, self._sizes, self.packed), select) if s_]
        n = len(chosen)
        items = (L.AdamItem * n)()
        for k, (p, o, sz, pk) in enumerate(chosen):
            prim = self._prim_of.get(id(p))
            wp0 = wp1 = None
            if prim is not None:
                b0, b1 = prim.packed_buffers(dtype)
                wp0, wp1 = b0.data_ptr(), b1.data_ptr()
                O, I, KH, KW = p.shape
            else:
                O, I, KH, KW = p.numel(), 1, 1, 1
            if pk and prim is None:
                O, I, KH, KW = p.shape            # a packed gradient without a ConvPrim (unused layer): the optimiser still needs the geometry
            items[k] = L.AdamItem(p.data_ptr(), self.grad_flat[o:].data_ptr(), self.exp_avg[o:].data_ptr(), self.exp_avg_sq[o:].data_ptr(),
                                  wp0, wp1, O, I, KH, KW, 1 if pk else 0, 0)
        tiles = (ctypes.c_int * (n + 1))()
        L.check(lib.ctagan_adam_pack_tiles(ctypes.cast(items, ctypes.c_void_p), n, tiles))
        smem = int(lib.ctagan_adam_pack_smem_bytes(ctypes.cast(items, ctypes.c_void_p), n))
        raw = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8)
        return {"items": raw.to(self.device), "tiles": torch.tensor(list(tiles), dtype=torch.int32).to(self.device), "n": n,
                "total": int(tiles[n]), "smem": smem}

    def _launch(self, which: str, advance: int):
        """One optimiser kernel over table `which` on the current stream; advance: 1 = this launch moves the step counter on."""
        t = self._tables[which]
        ops._count(1)
        L.check(L.load().ctagan_adam_pack_multi(ctypes.c_void_p(t["items"].data_ptr()), ctypes.c_void_p(t["tiles"].data_ptr()), t["n"],
                                                t["total"], t["smem"], ctypes.c_void_p(self.lr.data_ptr()),
                                                ctypes.c_void_p(self.step_count.data_ptr()),
                                                ctypes.c_void_p(self.ticket[0 if advance else 1:].data_ptr()),
                                                float(self.betas[0]), float(self.betas[1]), float(self.eps), _DT[self._dtype], int(advance),
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def step(self):
        dtype = E.get_precision()
        key = (dtype, tuple(p.data_ptr() for p in self.params))
        if key != self._table_key:            # first step, or parameters were re-allocated / the precision changed: rebuild (outside captures)
            self._build_table(dtype)
            self._table_key = key
        ops.ensure_device()
        if self._early_launched and "late" in self._tables:
            torch.cuda.current_stream().wait_stream(self._early_stream)      # also orders the step counter: the early launch read it first
            self._launch("late", 1)
        else:
            self._launch("all", 1)
        for prim in self.prims:
            prim.mark_packed(dtype)
