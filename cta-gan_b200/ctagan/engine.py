"""Explicit forward/backward kernel schedules of the three networks on the hot path.

Each network pass is ONE autograd node whose forward enqueues the fused kernel sequence and whose backward enqueues the
hand-scheduled reverse sequence (dgrad / wgrad / InstanceNorm-backward with the reflection-pad fold), so no ATen kernel runs
between the module boundaries.  Activations are NHWC in `precision` (bf16 fast mode, fp32 validation mode); raw conv
outputs and (mean, rstd) are kept for backward, normalised activations are written once into the (reflection-)padded
buffer the next convolution reads.

Reference semantics followed (paths in the upstream tree): Model/CycleGan.py:6-103, Model/HdGan.py:148-256,
trainer/reg.py:31-132, trainer/layers.py:71-104,156-183,216-300.
"""
from __future__ import annotations

import os

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import lib as L
from . import ops

_PRECISION = {"value": torch.bfloat16}


def set_precision(p):
    """'bf16' (tensor-core fast mode) or 'fp32' (validation mode, <=1e-4 vs the fp32 reference)."""
    _PRECISION["value"] = {"bf16": torch.bfloat16, "fp32": torch.float32, torch.bfloat16: torch.bfloat16,
                           torch.float32: torch.float32}[p]


def get_precision() -> torch.dtype:
    return _PRECISION["value"]


_ENGINE = {"value": L.ENGINE_AUTO}
_CACHE_EPOCH = {"value": 0}


def invalidate_weight_cache():
    """Forget every packed (bf16 / re-laid-out) weight copy.  Needed (a) after out-of-band `.data` edits, which do not bump
    the tensor version the cache is keyed on, and (b) right before CUDA-graph capture, so that the pack kernels are part of
    the captured step and replays always see the current master weights."""
    _CACHE_EPOCH["value"] += 1


_PTR_EPOCH: Dict[int, int] = {}


def _after_optimizer_step(optimizer, args, kwargs):
    # torch's fused Adam updates parameters without bumping their version counters: a step invalidates the packed copies of exactly the
    # parameters this optimizer owns (keyed by storage address, so twin leaves aliasing a parameter are covered too)
    for group in optimizer.param_groups:
        for p in group["params"]:
            k = p.data_ptr()
            _PTR_EPOCH[k] = _PTR_EPOCH.get(k, 0) + 1


try:    # global hook: every torch.optim step (also user-written loops around the drop-in modules) refreshes the packed weights
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_hook
    _reg_hook(_after_optimizer_step)
except ImportError:     # pragma: no cover
    pass


def set_conv_engine(e):
    _ENGINE["value"] = {"auto": L.ENGINE_AUTO, "simt": L.ENGINE_SIMT, "tc": L.ENGINE_TC, "generic": L.ENGINE_GENERIC}[e]


class ConvPrim:
    """A convolution weight W[O][I][K][K] with stride s and zero padding p, offering the three GEMMs that
    Conv2d (fwd = fprop) and ConvTranspose2d (fwd = bprop) need.  Packed copies are cached per weight version."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int, pad: int):
        self.w, self.b = weight, bias
        self.O, self.I, self.K, _ = weight.shape
        self.s, self.p = stride, pad
        self._cache: Dict[Tuple[int, torch.dtype], Tuple[int, int, torch.Tensor]] = {}
        # persistent gradient buffers (slices of an optimiser's flat gradient bucket, ctagan.optim.FusedAdam): the weight-gradient
        # kernels write there directly; the first use of the layer in a step overwrites, a second use accumulates
        self.grad_w: Optional[torch.Tensor] = None
        self.grad_w_raw: Optional[torch.Tensor] = None
        self.grad_packed = False
        self.grad_b: Optional[torch.Tensor] = None
        self.grad_writes = 0
        self.grad_event = None
        # data-parallel overlap: how many weight-gradient launches this layer receives per step (learnt from the previous step) and who
        # wants to know that the last one has been enqueued (trainers.GradSync: the bucket chunk can go to the all-reduce)
        self.expected_writes: Optional[int] = None
        self.final_hook = None

    def attach_grads(self, grad_w, grad_b):
        """grad_w: the persistent gradient of the weight, logical shape [O][I][K][K]; either contiguous (PyTorch order) or a
        channels-last view of contiguous [O][K][K][I] storage (what the tensor-core kernel writes as whole rows)."""
        self.grad_w, self.grad_b, self.grad_writes, self.grad_event = grad_w, grad_b, 0, None
        self.grad_w_raw, self.grad_packed = grad_w, False
        if grad_w is not None and not grad_w.is_contiguous():
            raw = grad_w.permute(0, 2, 3, 1)
            assert raw.is_contiguous(), "gradient buffers must be contiguous or channels-last"
            self.grad_w_raw, self.grad_packed = raw, True

    def mark_packed(self, dtype: torch.dtype):
        """The optimiser kernel has just re-packed both copies of `dtype` from the updated master weights: every other cached copy is
        stale, these two are fresh."""
        ptr = self.w.data_ptr()
        _PTR_EPOCH[ptr] = _PTR_EPOCH.get(ptr, 0) + 1
        ver = self._version_key()
        for m in (0, 1):
            hit = self._cache.get((m, dtype))
            if hit is not None:
                self._cache[(m, dtype)] = (ver, hit[1])

    def packed_buffers(self, dtype: torch.dtype):
        """(mode-0 buffer, mode-1 buffer), allocated if need be (contents unspecified until packed)."""
        out = []
        for m in (0, 1):
            hit = self._cache.get((m, dtype))
            if hit is None or hit[1].dtype != dtype:
                shape = (self.O, self.K, self.K, self.I) if m == 0 else (self.I, self.K, self.K, self.O)
                self._cache[(m, dtype)] = (None, torch.empty(shape, dtype=dtype, device=self.w.device))
            out.append(self._cache[(m, dtype)][1])
        return out

    def packed(self, mode: int, dtype: torch.dtype) -> torch.Tensor:
        """The packed copy for `mode`, re-packed in place if stale (the buffer persists: CUDA graphs and grouped launches alias it)."""
        entries = self.stale_entries(dtype, modes=(mode,))
        if entries:
            ops.pack_weights_multi(entries, dtype)
        return self._cache[(mode, dtype)][1]

    def prepack(self, dtype: torch.dtype, modes=(0, 1)):
        """Materialise the packed copies on the CURRENT stream (call before forking work onto side streams)."""
        for m in modes:
            self.packed(m, dtype)

    def _version_key(self):
        ptr = self.w.data_ptr()
        return (self.w._version, ptr, _CACHE_EPOCH["value"], _PTR_EPOCH.get(ptr, 0))

    def stale_entries(self, dtype: torch.dtype, modes=(0, 1), force=False):
        """(w, destination buffer, mode) for every packed copy that is out of date; marks them fresh (the caller packs them all
        with ONE ctagan_pack_weights_multi launch).  Destination buffers persist, so CUDA graphs see stable pointers."""
        out = []
        ver = self._version_key()
        for m in modes:
            key = (m, dtype)
            hit = self._cache.get(key)
            if hit is not None and hit[0] == ver and not force:
                continue
            if hit is not None and hit[1].dtype == dtype:
                buf = hit[1]
            else:
                shape = (self.O, self.K, self.K, self.I) if m == 0 else (self.I, self.K, self.K, self.O)
                buf = torch.empty(shape, dtype=dtype, device=self.w.device)
            self._cache[key] = (ver, buf)
            out.append((self.w.detach(), buf, m))
        return out

    # I-channel input -> O-channel output, strided  (Conv2d forward / ConvTranspose2d input-gradient)
    def fprop(self, x, act=L.ACT_NONE, use_bias=True, pad=None):
        N, Hi, Wi, Ci = x.shape
        assert Ci == self.I, (Ci, self.I)
        p = self.p if pad is None else pad
        Ho = (Hi + 2 * p - self.K) // self.s + 1
        Wo = (Wi + 2 * p - self.K) // self.s + 1
        g = ops.make_geom(N, Hi, Wi, Ci, Ho, Wo, self.O, self.K, self.s, 1, p, act, ops.dt(x))
        bias = self.b.detach() if (use_bias and self.b is not None) else None
        return ops.conv_gather(x, self.packed(0, x.dtype), bias, g, _ENGINE["value"])

    def fprop_stats(self, x, pool):
        """fprop of a layer followed by InstanceNorm (bias dead): raw output + (mean, rstd), statistics fused when possible."""
        N, Hi, Wi, Ci = x.shape
        Ho = (Hi + 2 * self.p - self.K) // self.s + 1
        Wo = (Wi + 2 * self.p - self.K) // self.s + 1
        g = ops.make_geom(N, Hi, Wi, Ci, Ho, Wo, self.O, self.K, self.s, 1, self.p, L.ACT_NONE, ops.dt(x))
        return ops.conv_gather_stats(x, self.packed(0, x.dtype), None, g, pool, _ENGINE["value"])

    # O-channel input -> I-channel output of spatial size `out_hw`  (Conv2d input-gradient / ConvTranspose2d forward)
    def bprop(self, dy, out_hw, act=L.ACT_NONE, bias=None, pad=None):
        N, Ho, Wo, Co = dy.shape
        assert Co == self.O, (Co, self.O)
        p = self.p if pad is None else pad
        g = ops.make_geom(N, Ho, Wo, Co, out_hw[0], out_hw[1], self.I, self.K, 1, self.s, self.K - 1 - p, act, ops.dt(dy))
        return ops.conv_gather(dy, self.packed(1, dy.dtype), bias, g, _ENGINE["value"])

    def bprop_stats(self, dy, out_hw, pool):
        """ConvTranspose2d forward followed by InstanceNorm."""
        N, Ho, Wo, Co = dy.shape
        g = ops.make_geom(N, Ho, Wo, Co, out_hw[0], out_hw[1], self.I, self.K, 1, self.s, self.K - 1 - self.p, L.ACT_NONE, ops.dt(dy))
        return ops.conv_gather_stats(dy, self.packed(1, dy.dtype), None, g, pool, _ENGINE["value"])

    # dW[O][I][K][K] = sum gy(O-channel, strided side) x gx(I-channel, gathered side)
    def wgrad(self, gy, gx, want_bias=False, pad=None, gy_margin=0):
        N, Ho, Wo, Co = gy.shape
        _, Hi, Wi, Ci = gx.shape
        assert Co == self.O and Ci == self.I
        p = self.p if pad is None else pad
        g = ops.make_geom(N, Hi, Wi, Ci, Ho, Wo, Co, self.K, self.s, 1, p, L.ACT_NONE, ops.dt(gy), gy_margin)
        if _ACTIVE_LANE["value"] is not None:
            _ACTIVE_LANE["value"].keep += (gy, gx)
        if self.grad_w is None:
            return ops.conv_wgrad(gy, gx, g, want_bias, _ENGINE["value"])
        # gradient bucket: the first writer of the step overwrites and leaves an event, a later one (the other use of this network,
        # possibly on another stream) waits for it and accumulates
        acc = self.grad_writes > 0
        if acc and self.grad_event is not None and os.environ.get("CTAGAN_BUCKET_NOWAIT") != "1":
            torch.cuda.current_stream().wait_event(self.grad_event)
        _, db = ops.conv_wgrad(gy, gx, g, want_bias and self.grad_b is not None, _ENGINE["value"], out_w=self.grad_w_raw,
                               out_b=self.grad_b if want_bias else None, accumulate=acc, packed=self.grad_packed)
        out = (self.grad_w, db)
        self.grad_writes += 1
        self.grad_event = torch.cuda.Event()
        self.grad_event.record()
        if self.final_hook is not None and self.expected_writes is not None and self.grad_writes >= self.expected_writes:
            self.final_hook(self)          # (>: one launch more than in the previous step -- the hook's owner decides what that means)
        return out


_WGRAD_STREAMS: Dict[int, "torch.cuda.Stream"] = {}
_ACTIVE_LANE = {"value": None}
_WGRAD_OVERLAP = {"value": True}


def set_wgrad_overlap(flag: bool):
    _WGRAD_OVERLAP["value"] = bool(flag)


class _WgradLane:
    """Weight gradients are leaves of the backward dependency chain (nothing downstream of a layer's dgrad needs them), so they
    are enqueued on a side stream: the critical path of a backward pass is then norm-backward + dgrad per layer, and the
    latency-bound batch-1 wgrad GEMMs (split-K + reduce) fill the SMs the main chain leaves idle.

    The side stream reads tensors the main stream allocated (dy, the saved layer input).  The caching allocator would hand such a block
    to the next main-stream allocation as soon as the main chain drops it -- while the lagging wgrad may still be reading it -- so the
    lane keeps every wgrad operand alive until join() (record_stream is not an option under CUDA-graph capture)."""

    def __init__(self):
        self.main = torch.cuda.current_stream()
        self.sides = []
        self.used = set()
        self.turn = 0
        self.keep = []
        if _WGRAD_OVERLAP["value"]:
            key = self.main.cuda_stream
            sides = _WGRAD_STREAMS.get(key)
            if sides is None:
                # several side streams per chain, used round-robin: weight gradients of different layers are independent of each other,
                # and when a backward chain ends its leftover weight gradients then run side by side instead of one after the other
                n = max(1, int(os.environ.get("CTAGAN_WG_LANES", "4")))
                sides = [ops.named_stream(f"lane{k}:{key}") for k in range(n)]
                _WGRAD_STREAMS[key] = sides
            self.sides = sides

    def run(self, fn):
        if not self.sides:
            return fn()
        k = self.turn % len(self.sides)
        self.turn += 1
        side = self.sides[k]
        self.used.add(k)
        side.wait_stream(self.main)
        prev, _ACTIVE_LANE["value"] = _ACTIVE_LANE["value"], self
        try:
            with torch.cuda.stream(side):
                return fn()
        finally:
            _ACTIVE_LANE["value"] = prev

    def join(self):
        # (only when something was forked: waiting on a stream that never joined a CUDA-graph capture would invalidate the capture)
        d = _DEFERRED["value"]
        if d is not None:                 # the caller collects the weight gradients after the whole backward: nobody waits here
            for k in self.used:
                d.lanes[(self.main.cuda_stream, self.sides[k].cuda_stream)] = (self.main, self.sides[k])
            d.keep += self.keep
            self.keep, self.used = [], set()
            return
        for k in self.used:
            self.main.wait_stream(self.sides[k])
        self.keep, self.used = [], set()


_DEFERRED = {"value": None}


class deferred_weight_grads:
    """`with deferred_weight_grads() as d: loss.backward(); d.flush()` -- inside the block the network Functions hand their weight
    gradients to `d` instead of returning them to autograd, and their wgrad lanes are not joined at the end of each network's
    backward.  The input-gradient chain (the critical path of the backward pass) then never waits for the lagging weight-gradient
    lanes at network boundaries; flush() joins every lane once, on the current stream, and accumulates into `.grad` (`p.grad = g`
    for the first contribution, one multi-tensor add for the others, e.g. the second use of a generator in the cycle pass)."""

    def __init__(self):
        self.items = []        # (leaf tensor, gradient)
        self.lanes = {}
        self.keep = []
        self.flushed = False

    def __enter__(self):
        assert _DEFERRED["value"] is None, "deferred_weight_grads does not nest"
        if os.environ.get("CTAGAN_DEFER", "1") != "0":
            _DEFERRED["value"] = self
        return self

    def __exit__(self, *exc):
        _DEFERRED["value"] = None
        if exc[0] is None and not self.flushed:
            self.flush()
        return False

    def take(self, leaves, grads, needs):
        """Called by a Function's backward: keeps the gradients of the leaves that want one; returns what to hand to autograd.
        Gradients that the kernels have already written into the leaf's persistent gradient buffer are done."""
        for leaf, g, need in zip(leaves, grads, needs):
            if need and g is not None and not grad_in_place(leaf, g):
                self.items.append((leaf, g))
        return [None] * len(grads)

    def flush(self):
        cur = torch.cuda.current_stream()
        for main, side in self.lanes.values():
            cur.wait_stream(main)
            cur.wait_stream(side)
        own, extra = [], []
        for leaf, g in self.items:
            if leaf.grad is None:
                leaf.grad = g
            else:
                own.append(leaf.grad); extra.append(g)
        if own:
            torch._foreach_add_(own, extra)
        self.items, self.lanes, self.keep = [], {}, []
        self.flushed = True


def deferred():
    return _DEFERRED["value"]


def grad_in_place(leaf, g) -> bool:
    """True when `g` IS the leaf's persistent gradient buffer (the weight-gradient kernel wrote there): nothing to hand to autograd."""
    return leaf.grad is not None and g is not None and g.data_ptr() == leaf.grad.data_ptr()


def strip_in_place(leaves, grads):
    """Gradients for autograd: None for the ones already in place in their leaf's gradient buffer."""
    return [None if (g is None or grad_in_place(leaf, g)) else g for leaf, g in zip(leaves, grads)]


class _PairStore:
    """Packed-weight storage shared by ConvPrims of identical geometry: one buffer [G][...] per (mode, dtype) whose slices ARE the
    prims' own cache buffers, so a grouped launch reaches every member's weights through one tensor map."""

    def __init__(self, prims):
        self.prims = list(prims)
        self.bufs: Dict[Tuple[int, torch.dtype], torch.Tensor] = {}

    def buffer(self, mode: int, dtype: torch.dtype) -> torch.Tensor:
        key = (mode, dtype)
        buf = self.bufs.get(key)
        p0 = self.prims[0]
        if buf is None:
            shape = (p0.O, p0.K, p0.K, p0.I) if mode == 0 else (p0.I, p0.K, p0.K, p0.O)
            buf = torch.empty((len(self.prims), *shape), dtype=dtype, device=p0.w.device)
            self.bufs[key] = buf
        for k, prim in enumerate(self.prims):                     # (re-)install the slices; a foreign buffer means "stale"
            hit = prim._cache.get(key)
            if hit is None or hit[1].data_ptr() != buf[k].data_ptr():
                prim._cache[key] = (None, buf[k])
        return buf


_PAIR_STORES: Dict[Tuple[int, ...], _PairStore] = {}


class GroupedPrim:
    """G ConvPrims of identical geometry applied to G consecutive image groups of one batch: the ConvPrim interface, one launch per
    layer where the tcgen05 engine can (ctagan_conv_gather_grouped / ctagan_conv_wgrad_grouped), one launch per group otherwise
    (1-2 channel layers, live biases, fp32 validation mode).  Weight gradients come back stacked: dW[G][O][I][K][K]."""

    def __init__(self, prims: Sequence[ConvPrim]):
        p0 = prims[0]
        assert all((q.O, q.I, q.K, q.s, q.p) == (p0.O, p0.I, p0.K, p0.s, p0.p) for q in prims)
        self.prims = list(prims)
        self.G = len(prims)
        self.O, self.I, self.K, self.s, self.p, self.b = p0.O, p0.I, p0.K, p0.s, p0.p, p0.b
        order = sorted(range(self.G), key=lambda k: id(prims[k]))
        key = tuple(id(prims[k]) for k in order)
        store = _PAIR_STORES.get(key)
        if store is None:
            store = _PAIR_STORES[key] = _PairStore([prims[k] for k in order])
        self.store = store
        self.slots = [0] * self.G
        for s_, k in enumerate(order):
            self.slots[k] = s_

    def _packed(self, mode, dtype):
        buf = self.store.buffer(mode, dtype)
        entries = []
        for prim in self.prims:
            entries += prim.stale_entries(dtype, modes=(mode,))
        if entries:
            ops.pack_weights_multi(entries, dtype)
        return buf

    def _split(self, t):
        B = t.shape[0] // self.G
        return [t[k * B:(k + 1) * B] for k in range(self.G)]

    def _ok(self, g):
        return _ENGINE["value"] == L.ENGINE_AUTO and ops.conv_gather_grouped_supported(g, self.slots)

    def fprop(self, x, act=L.ACT_NONE, use_bias=True, pad=None):
        N, Hi, Wi, Ci = x.shape
        p = self.p if pad is None else pad
        Ho, Wo = (Hi + 2 * p - self.K) // self.s + 1, (Wi + 2 * p - self.K) // self.s + 1
        g = ops.make_geom(N, Hi, Wi, Ci, Ho, Wo, self.O, self.K, self.s, 1, p, act, ops.dt(x))
        if not (use_bias and self.b is not None) and self._ok(g):
            return ops.conv_gather_grouped(x, self._packed(0, x.dtype), g, self.slots)
        return torch.cat([q.fprop(xk, act, use_bias, pad) for q, xk in zip(self.prims, self._split(x))])

    def fprop_stats(self, x, pool):
        N, Hi, Wi, Ci = x.shape
        Ho, Wo = (Hi + 2 * self.p - self.K) // self.s + 1, (Wi + 2 * self.p - self.K) // self.s + 1
        g = ops.make_geom(N, Hi, Wi, Ci, Ho, Wo, self.O, self.K, self.s, 1, self.p, L.ACT_NONE, ops.dt(x))
        if pool is not None and self._ok(g):
            return ops.conv_gather_grouped(x, self._packed(0, x.dtype), g, self.slots, pool)
        parts = [q.fprop_stats(xk, pool) for q, xk in zip(self.prims, self._split(x))]
        return torch.cat([a for a, _ in parts]), torch.cat([b for _, b in parts])

    def bprop(self, dy, out_hw, act=L.ACT_NONE, bias=None, pad=None):
        N, Ho, Wo, Co = dy.shape
        p = self.p if pad is None else pad
        g = ops.make_geom(N, Ho, Wo, Co, out_hw[0], out_hw[1], self.I, self.K, 1, self.s, self.K - 1 - p, act, ops.dt(dy))
        if bias is None and self._ok(g):
            return ops.conv_gather_grouped(dy, self._packed(1, dy.dtype), g, self.slots)
        return torch.cat([q.bprop(dk, out_hw, act, bias, pad) for q, dk in zip(self.prims, self._split(dy))])

    def bprop_stats(self, dy, out_hw, pool):
        N, Ho, Wo, Co = dy.shape
        g = ops.make_geom(N, Ho, Wo, Co, out_hw[0], out_hw[1], self.I, self.K, 1, self.s, self.K - 1 - self.p, L.ACT_NONE, ops.dt(dy))
        if pool is not None and self._ok(g):
            return ops.conv_gather_grouped(dy, self._packed(1, dy.dtype), g, self.slots, pool)
        parts = [q.bprop_stats(dk, out_hw, pool) for q, dk in zip(self.prims, self._split(dy))]
        return torch.cat([a for a, _ in parts]), torch.cat([b for _, b in parts])

    def wgrad(self, gy, gx, want_bias=False, pad=None, gy_margin=0):
        N, Ho, Wo, Co = gy.shape
        _, Hi, Wi, Ci = gx.shape
        p = self.p if pad is None else pad
        g = ops.make_geom(N, Hi, Wi, Ci, Ho, Wo, Co, self.K, self.s, 1, p, L.ACT_NONE, ops.dt(gy), gy_margin)
        ws = ops.conv_wgrad_grouped_workspace(g, self.G) if _ENGINE["value"] == L.ENGINE_AUTO else 0
        if ws:
            if _ACTIVE_LANE["value"] is not None:
                _ACTIVE_LANE["value"].keep += (gy, gx)
            return ops.conv_wgrad_grouped(gy, gx, g, self.G, want_bias, ws)
        parts = [q.wgrad(a, b, want_bias, pad, gy_margin) for q, a, b in zip(self.prims, self._split(gy), self._split(gx))]
        return torch.stack([a for a, _ in parts]), (torch.stack([b for _, b in parts]) if want_bias else None)


def split_group_grads(grads, G):
    """Stacked per-layer gradients [G, ...] of a grouped plan -> flat list: all of member 0's gradients, then member 1's, ..."""
    return [None if t is None else t[k] for k in range(G) for t in grads]


def prepack_prims(prims, dtype, force=False):
    """Re-pack every stale (force: every) weight copy of a network with one kernel launch, into the persistent pack buffers."""
    entries = []
    for prim in prims:
        entries += prim.stale_entries(dtype, force=force)
    if entries:
        ops.pack_weights_multi(entries, dtype)


# ======================================================================================================================
# ResNet-9 generator  (Model/CycleGan.py:23-71)
# ======================================================================================================================


class GeneratorPlan:
    """Holds the ConvPrims of one Generator (parameters are the module's own nn.Parameters, fp32 OIHW)."""

    def __init__(self, params: Sequence[torch.Tensor], n_blocks: int):
        it = iter(params)
        nxt = lambda: (next(it), next(it))
        w, b = nxt(); self.head1 = ConvPrim(w, b, 1, 0)            # 7x7 on the reflect-padded image
        w, b = nxt(); self.head4 = ConvPrim(w, b, 2, 1)
        w, b = nxt(); self.head7 = ConvPrim(w, b, 2, 1)
        self.blocks = []
        for _ in range(n_blocks):
            w1, b1 = nxt(); w2, b2 = nxt()
            self.blocks.append((ConvPrim(w1, b1, 1, 0), ConvPrim(w2, b2, 1, 0)))
        # ConvTranspose2d weight [Cin_T][Cout_T][3][3] == a stride-2 conv weight W[O=Cin_T][I=Cout_T]
        w, b = nxt(); self.tail0 = ConvPrim(w, b, 2, 1)
        w, b = nxt(); self.tail3 = ConvPrim(w, b, 2, 1)
        w, b = nxt(); self.tail7 = ConvPrim(w, b, 1, 0)
        self.n_blocks = n_blocks
        self.params = list(params)
        # called (with need_dw) by generator_backward once the input-gradient chain has passed every residual block: from here on only
        # the three head layers still read packed weights / receive weight gradients (ctagan.optim.FusedAdam: early optimiser launch)
        self.body_done_hook = None

    def early_prims(self):
        """The layers whose weights are neither read nor differentiated after `body_done_hook`: the residual blocks and the tail."""
        out = []
        for c1, c2 in self.blocks:
            out += [c1, c2]
        return out + [self.tail0, self.tail3, self.tail7]

    def prims(self):
        out = [self.head1, self.head4, self.head7]
        for c1, c2 in self.blocks:
            out += [c1, c2]
        return out + [self.tail0, self.tail3, self.tail7]


class GroupedGeneratorPlan:
    """G generators of the same architecture over G consecutive image groups (see GroupedPrim); parameter order = member 0, member 1.."""

    def __init__(self, plans: Sequence[GeneratorPlan]):
        self.plans = list(plans)
        self.G = len(plans)
        self.n_blocks = plans[0].n_blocks
        grp = lambda name: GroupedPrim([getattr(p, name) for p in plans])
        self.head1, self.head4, self.head7 = grp("head1"), grp("head4"), grp("head7")
        self.blocks = [(GroupedPrim([p.blocks[i][0] for p in plans]), GroupedPrim([p.blocks[i][1] for p in plans]))
                       for i in range(self.n_blocks)]
        self.tail0, self.tail3, self.tail7 = grp("tail0"), grp("tail3"), grp("tail7")
        self.params = [t for p in plans for t in p.params]


def generator_forward(plan: GeneratorPlan, x_nchw: torch.Tensor, save: bool):
    T = get_precision()
    N, Cin, H, W = x_nchw.shape
    x0 = ops.nchw_to_nhwc(x_nchw, T)
    pool = ops.ZeroPool(2 * N * (64 + 128 + 256 + 512 * plan.n_blocks + 128 + 64) + 64, x0.device) if T == torch.bfloat16 else None
    P0 = ops.norm_act_pad(x0, None, L.ACT_NONE, 3)
    r1, s1 = plan.head1.fprop_stats(P0, pool)
    A1 = ops.norm_act_pad(r1, s1, L.ACT_RELU, 0)
    r2, s2 = plan.head4.fprop_stats(A1, pool)
    A2 = ops.norm_act_pad(r2, s2, L.ACT_RELU, 0)
    r3, s3 = plan.head7.fprop_stats(A2, pool)
    nb = plan.n_blocks
    X = ops.norm_act_pad(r3, s3, L.ACT_RELU, 1 if nb > 0 else 0)
    blocks = []
    for k, (c1, c2) in enumerate(plan.blocks):
        ra, sa = c1.fprop_stats(X, pool)
        Tt = ops.norm_act_pad(ra, sa, L.ACT_RELU, 1)
        rb, sb = c2.fprop_stats(Tt, pool)
        Xn = ops.norm_act_pad(rb, sb, L.ACT_NONE, 1 if k < nb - 1 else 0, res=X, res_pad=1)
        if save:
            blocks.append((X, ra, sa, Tt, rb, sb))
        X = Xn
    h4, w4 = X.shape[1], X.shape[2]
    r4, s4 = plan.tail0.bprop_stats(X, (2 * h4, 2 * w4), pool)
    A4 = ops.norm_act_pad(r4, s4, L.ACT_RELU, 0)
    r5, s5 = plan.tail3.bprop_stats(A4, (4 * h4, 4 * w4), pool)
    P5 = ops.norm_act_pad(r5, s5, L.ACT_RELU, 3)
    y = plan.tail7.fprop(P5, act=L.ACT_TANH, use_bias=True)
    out = ops.nhwc_to_nchw(y)
    saved = (P0, r1, s1, A1, r2, s2, A2, r3, s3, blocks, X, r4, s4, A4, r5, s5, P5, y) if save else None
    return out, saved


def generator_backward(plan: GeneratorPlan, saved, dout: torch.Tensor, need_dx: bool, need_dw: bool = True):
    (P0, r1, s1, A1, r2, s2, A2, r3, s3, blocks, X9, r4, s4, A4, r5, s5, P5, y) = saved
    T = y.dtype
    grads: List[Optional[torch.Tensor]] = []

    lane = _WgradLane()

    def wg(prim, gy, gx, bias=False, margin=0):
        if not need_dw:
            return None, None
        return lane.run(lambda: prim.wgrad(gy, gx, want_bias=bias, pad=(margin if margin else None), gy_margin=margin))

    bpool = ops.ZeroPool(2 * y.shape[0] * (64 + 128 + 256 + 512 * plan.n_blocks + 128 + 64), y.device)
    gy = ops.nchw_to_nhwc(dout, T)
    dy7 = ops.act_bwd(gy, y, L.ACT_TANH)
    dW7, db7 = wg(plan.tail7, dy7, P5, True)
    dP5 = plan.tail7.bprop(dy7, (P5.shape[1], P5.shape[2]))
    dr5 = ops.norm_act_pad_bwd(dP5, r5, s5, L.ACT_RELU, 3, pool=bpool)
    dA4 = plan.tail3.fprop(dr5, use_bias=False)
    dWt3, _ = wg(plan.tail3, A4, dr5)
    dr4 = ops.norm_act_pad_bwd(dA4, r4, s4, L.ACT_RELU, 0, pool=bpool)
    G = plan.tail0.fprop(dr4, use_bias=False)
    dWt0, _ = wg(plan.tail0, X9, dr4)
    block_grads = []
    # The gradient of a block's output has two consumers: the block's own second InstanceNorm and the skip connection.  It is
    # G_k = fold(dXp_{k+1}) + G_{k+1}; instead of materialising it with a kernel of its own, the norm-backward launch of the next
    # consumer folds dXp and adds the skip term itself and (when a further skip needs it) writes G_k as a second output.
    dXp = None                                                            # not-yet-folded input gradient of the block processed last
    for (c1, c2), (Xk, ra, sa, Tt, rb, sb) in zip(reversed(plan.blocks), reversed(blocks)):
        m = c2.K - 1                                                      # zero margin: dgrad becomes a VALID conv
        if dXp is None:
            drb = ops.norm_act_pad_bwd(G, rb, sb, L.ACT_NONE, 0, out_pad=m, pool=bpool)
        else:
            drb, G = ops.norm_act_pad_bwd(dXp, rb, sb, L.ACT_NONE, 1, addend=G, out_pad=m, pool=bpool, want_g=True)
        dW2, _ = wg(c2, drb, Tt, margin=m)
        dT = c2.bprop(drb, (Tt.shape[1], Tt.shape[2]), pad=m)
        dra = ops.norm_act_pad_bwd(dT, ra, sa, L.ACT_RELU, 1, out_pad=m, pool=bpool)
        dW1, _ = wg(c1, dra, Xk, margin=m)
        dXp = c1.bprop(dra, (Xk.shape[1], Xk.shape[2]), pad=m)
        block_grads.append((dW1, dW2))
    block_grads.reverse()
    if dXp is None:
        dr3 = ops.norm_act_pad_bwd(G, r3, s3, L.ACT_RELU, 0, pool=bpool)
    else:
        dr3 = ops.norm_act_pad_bwd(dXp, r3, s3, L.ACT_RELU, 1, addend=G, pool=bpool)
    if need_dw and getattr(plan, "body_done_hook", None) is not None:
        plan.body_done_hook(plan)        # every kernel that reads a residual-block / tail weight in this pass is enqueued (this stream or a lane)
    dWh7, _ = wg(plan.head7, dr3, A2)
    dA2 = plan.head7.bprop(dr3, (A2.shape[1], A2.shape[2]))
    dr2 = ops.norm_act_pad_bwd(dA2, r2, s2, L.ACT_RELU, 0, pool=bpool)
    dWh4, _ = wg(plan.head4, dr2, A1)
    dA1 = plan.head4.bprop(dr2, (A1.shape[1], A1.shape[2]))
    dr1 = ops.norm_act_pad_bwd(dA1, r1, s1, L.ACT_RELU, 0, pool=bpool)
    dWh1, _ = wg(plan.head1, dr1, P0)
    dx = None
    if need_dx:
        dP0 = plan.head1.bprop(dr1, (P0.shape[1], P0.shape[2]))
        dx0 = ops.norm_act_pad_bwd(dP0, None, None, L.ACT_NONE, 3)
        dx = ops.nhwc_to_nchw(dx0)
    lane.join()
    if need_dw:
        # biases in front of a non-affine InstanceNorm are mathematically dead: zero gradient
        z = lambda prim: None          # (None == no gradient: Adam leaves the dead parameter untouched)
        grads = [dWh1, z(plan.head1), dWh4, z(plan.head4), dWh7, z(plan.head7)]
        for (c1, c2), (dW1, dW2) in zip(plan.blocks, block_grads):
            grads += [dW1, z(c1), dW2, z(c2)]
        grads += [dWt0, z(plan.tail0), dWt3, z(plan.tail3), dW7, db7]
    else:
        grads = [None] * len(plan.params)
    return dx, grads


# ======================================================================================================================
# PatchGAN discriminator  (Model/CycleGan.py:73-103, Model/HdGan.py:148-256)
# ======================================================================================================================


class DiscriminatorPlan:
    def __init__(self, params: Sequence[torch.Tensor]):
        it = iter(params)
        strides = [2, 2, 2, 1, 1]
        self.convs = []
        for s in strides:
            w, b = next(it), next(it)
            self.convs.append(ConvPrim(w, b, s, 1))
        self.params = list(params)

    def prims(self):
        return list(self.convs)


class GroupedDiscriminatorPlan:
    def __init__(self, plans: Sequence[DiscriminatorPlan]):
        self.plans = list(plans)
        self.G = len(plans)
        self.convs = [GroupedPrim([p.convs[i] for p in plans]) for i in range(len(plans[0].convs))]
        self.params = [t for p in plans for t in p.params]


def discriminator_forward(plan: DiscriminatorPlan, x_nchw: torch.Tensor, save: bool):
    """Returns the last feature map [N,1,h,w] fp32 and (optionally) the NHWC intermediates [a0..a3]."""
    T = get_precision()
    x0 = ops.nchw_to_nhwc(x_nchw, T)
    c = plan.convs
    a0 = c[0].fprop(x0, act=L.ACT_LRELU, use_bias=True)
    acts, raws, stats = [a0], [], []
    a = a0
    pool = ops.ZeroPool(2 * x0.shape[0] * (128 + 256 + 512) + 8, x0.device) if T == torch.bfloat16 else None
    for i in (1, 2, 3):
        r, s = c[i].fprop_stats(a, pool)
        a = ops.norm_act_pad(r, s, L.ACT_LRELU, 0)
        raws.append(r); stats.append(s); acts.append(a)
    y4 = c[4].fprop(a, use_bias=True)
    out = ops.nhwc_to_nchw(y4)
    saved = (x0, acts, raws, stats, y4.shape, y4.dtype) if save else None
    return out, acts, saved


def discriminator_backward(plan: DiscriminatorPlan, saved, dout: torch.Tensor, need_dx: bool, need_dw: bool):
    x0, acts, raws, stats, yshape, T = saved
    c = plan.convs
    dy = ops.nchw_to_nhwc(dout, T)
    gw: List[Optional[torch.Tensor]] = [None] * 10
    lane = _WgradLane()
    if need_dw:
        gw[8], gw[9] = lane.run(lambda: c[4].wgrad(dy, acts[3], want_bias=True))
    da = c[4].bprop(dy, (acts[3].shape[1], acts[3].shape[2]))
    for i in (3, 2, 1):
        dr = ops.norm_act_pad_bwd(da, raws[i - 1], stats[i - 1], L.ACT_LRELU, 0)
        if need_dw:
            gw[2 * i], _ = lane.run(lambda dr=dr, i=i: c[i].wgrad(dr, acts[i - 1]))
            gw[2 * i + 1] = None            # bias in front of InstanceNorm: mathematically dead
        da = c[i].bprop(dr, (acts[i - 1].shape[1], acts[i - 1].shape[2]))
    dy0 = ops.act_bwd(da, acts[0], L.ACT_LRELU)
    if need_dw:
        gw[0], gw[1] = lane.run(lambda: c[0].wgrad(dy0, x0, want_bias=True))
    dx = None
    if need_dx:
        dx0 = c[0].bprop(dy0, (x0.shape[1], x0.shape[2]))
        dx = ops.nhwc_to_nchw(dx0)
    lane.join()
    return dx, gw


# ======================================================================================================================
# Registration U-Net  (trainer/reg.py:31-99, trainer/layers.py)
# ======================================================================================================================

REG_NDF = [32, 64, 64, 64, 64, 64, 64]
REG_NUF = [64, 64, 64, 64, 64, 64, 32]


class _ResBlock:
    """x + IN(conv3(RP1(relu(IN(conv3(RP1(x)))))))   trainer/layers.py:257-300 (also Model/CycleGan.py:6-21)."""

    def __init__(self, w1, b1, w2, b2):
        self.c1, self.c2 = ConvPrim(w1, b1, 1, 0), ConvPrim(w2, b2, 1, 0)

    def forward(self, a, save):
        Pa = ops.norm_act_pad(a, None, L.ACT_NONE, 1)
        pool = ops.ZeroPool(4 * a.shape[0] * self.c1.O + 4, a.device) if a.dtype == torch.bfloat16 else None
        ra, sa = self.c1.fprop_stats(Pa, pool)
        Tt = ops.norm_act_pad(ra, sa, L.ACT_RELU, 1)
        rb, sb = self.c2.fprop_stats(Tt, pool)
        out = ops.norm_act_pad(rb, sb, L.ACT_NONE, 0, res=a, res_pad=0)
        return out, ((Pa, ra, sa, Tt, rb, sb) if save else None)

    def backward(self, saved, G, pre_act=None, pre_act_kind=L.ACT_NONE, lane=None):
        """G: grad w.r.t. the block output.  Returns (grad w.r.t. block input [through `pre_act_kind` of the producer when
        pre_act is given], [dW1, db1, dW2, db2])."""
        Pa, ra, sa, Tt, rb, sb = saved
        run = lane.run if lane is not None else (lambda f: f())
        m = self.c2.K - 1                                                 # zero margin: dgrad becomes a VALID conv (tcgen05 form)
        drb = ops.norm_act_pad_bwd(G, rb, sb, L.ACT_NONE, 0, out_pad=m)
        dW2, _ = run(lambda: self.c2.wgrad(drb, Tt, pad=m, gy_margin=m))
        dT = self.c2.bprop(drb, (Tt.shape[1], Tt.shape[2]), pad=m)
        dra = ops.norm_act_pad_bwd(dT, ra, sa, L.ACT_RELU, 1, out_pad=m)
        dW1, _ = run(lambda: self.c1.wgrad(dra, Pa, pad=m, gy_margin=m))
        dPa = self.c1.bprop(dra, (Pa.shape[1], Pa.shape[2]), pad=m)
        Ga = ops.norm_act_pad_bwd(dPa, pre_act, None, pre_act_kind, 1, addend=G)
        return Ga, [dW1, None, dW2, None]      # biases in front of InstanceNorm are dead


class RegPlan:
    """Parameter order == state_dict order of Reg (80 tensors; oracle/restate.py:init_reg documents it)."""

    def __init__(self, params: Sequence[torch.Tensor]):
        it = iter(params)
        nxt = lambda: (next(it), next(it))
        self.down = []
        for _ in REG_NDF:
            w, b = nxt()
            w1, b1 = nxt(); w2, b2 = nxt()
            self.down.append((ConvPrim(w, b, 1, 1), _ResBlock(w1, b1, w2, b2)))
        w, b = nxt(); self.c1 = ConvPrim(w, b, 1, 0)
        self.t = []
        for _ in range(3):
            w1, b1 = nxt(); w2, b2 = nxt()
            self.t.append(_ResBlock(w1, b1, w2, b2))
        w, b = nxt(); self.c2 = ConvPrim(w, b, 1, 0)
        self.up = []
        for _ in REG_NUF:
            w, b = nxt(); self.up.append(ConvPrim(w, b, 1, 1))
        w1, b1 = nxt(); w2, b2 = nxt()
        self.refine0 = _ResBlock(w1, b1, w2, b2)
        w, b = nxt(); self.refine1 = ConvPrim(w, b, 1, 0)
        w, b = nxt(); self.output = ConvPrim(w, b, 1, 1)
        self.params = list(params)

    def prims(self):
        out = []
        for conv, rb in self.down:
            out += [conv, rb.c1, rb.c2]
        out.append(self.c1)
        for rb in self.t:
            out += [rb.c1, rb.c2]
        out.append(self.c2)
        out += list(self.up) + [self.refine0.c1, self.refine0.c2, self.refine1, self.output]
        return out


def reg_forward(plan: RegPlan, img_a: torch.Tensor, img_b: torch.Tensor, save: bool):
    T = get_precision()
    if img_a.shape[1] == 1 and img_b.shape[1] == 1:
        x_in = ops.interleave2(img_a, img_b, T)
    else:                                            # generic channel counts: boundary concat is plain plumbing
        x_in = ops.nchw_to_nhwc(torch.cat([img_a, img_b], 1), T)
    x = x_in
    downs, skips = [], []
    for conv, rb in plan.down:
        a = conv.fprop(x, act=L.ACT_LRELU)
        sk, rsaved = rb.forward(a, save)
        pooled = ops.maxpool2_fwd(sk)
        if save:
            downs.append((x, a, rsaved, sk))
        skips.append(sk)
        x = pooled
    m_in = x
    m1 = plan.c1.fprop(m_in, act=L.ACT_LRELU)
    tsaved, ts_in = [], []
    t = m1
    for rb in plan.t:
        ts_in.append(t)
        t, rs = rb.forward(t, save)
        tsaved.append(rs)
    m2 = plan.c2.fprop(t, act=L.ACT_LRELU)
    x = m2
    ups = []
    for conv, sk in zip(plan.up, reversed(skips)):
        u = ops.upsample2x_cat_fwd(x, sk)
        xo = conv.fprop(u, act=L.ACT_LRELU)
        if save:
            ups.append((u, xo, x.shape[3]))
        x = xo
    rf, rfsaved = plan.refine0.forward(x, save)
    r1 = plan.refine1.fprop(rf, act=L.ACT_LRELU)
    fl = plan.output.fprop(r1, act=L.ACT_NONE)
    out = ops.nhwc_to_nchw(fl)
    saved = (downs, m_in, m1, ts_in, tsaved, t, m2, ups, x, rfsaved, rf, r1, fl.dtype) if save else None
    return out, saved


def reg_backward(plan: RegPlan, saved, dflow: torch.Tensor, need_da: bool, need_db: bool, in_channels=(1, 1)):
    downs, m_in, m1, ts_in, tsaved, t_out, m2, ups, x_last, rfsaved, rf, r1, T = saved
    lane = _WgradLane()
    dfl = ops.nchw_to_nhwc(dflow, T)
    dWo, dbo = lane.run(lambda: plan.output.wgrad(dfl, r1, want_bias=True))
    dr1 = plan.output.bprop(dfl, (r1.shape[1], r1.shape[2]))
    dy = ops.act_bwd(dr1, r1, L.ACT_LRELU)
    dWr1, dbr1 = lane.run(lambda dy=dy: plan.refine1.wgrad(dy, rf, want_bias=True))
    Grf = plan.refine1.bprop(dy, (rf.shape[1], rf.shape[2]))
    # refine.0 res-block: its input is the LeakyReLU output of up_1 -> fold that activation's backward in
    Gx, g_refine0 = plan.refine0.backward(rfsaved, Grf, pre_act=x_last, pre_act_kind=L.ACT_LRELU, lane=lane)
    up_grads = []
    skip_grads = []
    dyo = Gx  # already multiplied by lrelu'(x_last)
    for conv, (u, xo, C1) in zip(reversed(plan.up), reversed(ups)):
        dW, db = lane.run(lambda dyo=dyo, conv=conv, u=u: conv.wgrad(dyo, u, want_bias=True))
        du = conv.bprop(dyo, (u.shape[1], u.shape[2]))
        gx, gskip = ops.upsample2x_cat_bwd(du, C1)
        up_grads.append((dW, db))
        skip_grads.append(gskip)
        dyo = gx  # grad w.r.t. the previous up conv's post-activation output (or m2): activation backward applied below
        # previous producer is a conv+lrelu: find its output tensor
        prev_out = None
        idx = len(up_grads)
        if idx < len(ups):
            prev_out = ups[len(ups) - 1 - idx][1]
            dyo = ops.act_bwd(dyo, prev_out, L.ACT_LRELU)
    up_grads.reverse()            # now in plan.up order
    # dyo is grad w.r.t. m2 (post-lrelu)
    dy = ops.act_bwd(dyo, m2, L.ACT_LRELU)
    dWc2, dbc2 = lane.run(lambda dy=dy: plan.c2.wgrad(dy, t_out, want_bias=True))
    Gt = plan.c2.bprop(dy, (t_out.shape[1], t_out.shape[2]))
    t_grads = []
    for i in (2, 1, 0):
        if i == 0:
            Gt, gr = plan.t[i].backward(tsaved[i], Gt, pre_act=m1, pre_act_kind=L.ACT_LRELU, lane=lane)
        else:
            Gt, gr = plan.t[i].backward(tsaved[i], Gt, lane=lane)
        t_grads.append(gr)
    t_grads.reverse()
    dWc1, dbc1 = lane.run(lambda Gt=Gt: plan.c1.wgrad(Gt, m_in, want_bias=True))
    Gp = plan.c1.bprop(Gt, (m_in.shape[1], m_in.shape[2]))      # grad w.r.t. pooled output of down_7
    down_grads = []
    # skip_grads was filled from up_1 (skip of down_1) to up_7 (skip of down_7): already in down-block order
    for n in range(len(plan.down) - 1, -1, -1):
        conv, rb = plan.down[n]
        x_in, a, rsaved, sk = downs[n]
        Gsk = ops.maxpool2_bwd(Gp, sk, addend=skip_grads[n])
        Ga, gr = rb.backward(rsaved, Gsk, pre_act=a, pre_act_kind=L.ACT_LRELU, lane=lane)
        dW, db = lane.run(lambda Ga=Ga, conv=conv, x_in=x_in: conv.wgrad(Ga, x_in, want_bias=True))
        down_grads.append((dW, db, gr))
        if n > 0:
            Gp = conv.bprop(Ga, (x_in.shape[1], x_in.shape[2]))
        elif need_da or need_db:
            Gp = conv.bprop(Ga, (x_in.shape[1], x_in.shape[2]))
    down_grads.reverse()
    da = db_ = None
    if need_da or need_db:
        if tuple(in_channels) == (1, 1):
            da, db_ = ops.deinterleave2(Gp, need_da, need_db)
        else:
            gin = ops.nhwc_to_nchw(Gp)
            Ca = in_channels[0]
            da = gin[:, :Ca].contiguous() if need_da else None
            db_ = gin[:, Ca:].contiguous() if need_db else None
    lane.join()
    grads: List[torch.Tensor] = []
    for dW, db, gr in down_grads:
        grads += [dW, db] + gr
    grads += [dWc1, dbc1]
    for gr in t_grads:
        grads += gr
    grads += [dWc2, dbc2]
    for dW, db in up_grads:
        grads += [dW, db]
    grads += g_refine0 + [dWr1, dbr1, dWo, dbo]
    return da, db_, grads
