"""Tensor-level wrappers over the C ABI.  torch is used only for device memory and streams: every function here
enqueues hand-written sm_100a kernels from libctagan.so on torch's current stream."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("ctagan ops need CUDA tensors (sm_100a); there is no CPU fallback")


_checked_devices = set()
_LAUNCHES = {"n": 0}


def launch_count() -> int:
    """Number of libctagan kernels enqueued so far by this process (memsets are not counted)."""
    return _LAUNCHES["n"]


def _count(n=1):
    _LAUNCHES["n"] += n


def ensure_device():
    dev = torch.cuda.current_device()
    if dev not in _checked_devices:
        L.check(L.load().ctagan_check_device(dev))
        _checked_devices.add(dev)


_NAMED_STREAMS = {}


def named_stream(name: str, priority: int = 0) -> "torch.cuda.Stream":
    """The process-wide dedicated CUDA stream for a role (e.g. "cyc.chainA", "lane:<main stream>"): created once through
    ctagan_stream_create, never taken from PyTorch's round-robin stream pool, so distinct roles are distinct CUDA streams."""
    key = (torch.cuda.current_device(), name)
    st = _NAMED_STREAMS.get(key)
    if st is None:
        ensure_device()
        h = ctypes.c_void_p()
        L.check(L.load().ctagan_stream_create(int(priority), ctypes.byref(h)))
        st = torch.cuda.ExternalStream(h.value)
        _NAMED_STREAMS[key] = st
    return st


def dt(t: torch.Tensor) -> int:
    return _DT[t.dtype]


def make_geom(N, Hi, Wi, Ci, Ho, Wo, Co, K, stride, dil, pad, act, dtype, gy_margin=0) -> L.ConvGeom:
    return L.ConvGeom(N, Hi, Wi, Ci, Ho, Wo, Co, K, K, stride, dil, pad, pad, act, dtype, gy_margin)


def conv_gather(x, wp, bias, g: L.ConvGeom, engine=L.ENGINE_AUTO):
    """x[N,Hi,Wi,Ci] -> y[N,Ho,Wo,Co] (see ctagan_conv_gather)."""
    _require_cuda(x, wp)
    ensure_device()
    y = torch.empty((g.N, g.Ho, g.Wo, g.Co), dtype=x.dtype, device=x.device)
    lib = L.load()
    ws_bytes = int(lib.ctagan_conv_gather_workspace_bytes(ctypes.byref(g), engine))
    if ws_bytes:            # 1-2 output channels on the tensor cores: per-tap planes Z^T in caller-owned scratch, then a gather
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
        _count(2)
        L.check(lib.ctagan_conv_gather_ws(ctypes.byref(g), _p(x), _p(wp), _p(bias), _p(y), _p(ws), ws_bytes, engine, _stream()))
        return y
    _count(1)
    L.check(lib.ctagan_conv_gather(ctypes.byref(g), _p(x), _p(wp), _p(bias), _p(y), engine, _stream()))
    return y


class ZeroPool:
    """One zero-filled buffer per network pass, handed out in slices: the arrival tickets of all deterministic split reductions of the
    pass (fused conv statistics, two-kernel norm backward) are cleared by a single memset instead of one per layer.  The kernels
    re-arm their tickets, so a slice is zero again when its kernel has finished."""

    def __init__(self, n_doubles: int, device):
        self.buf = torch.zeros((max(n_doubles, 2),), dtype=torch.float64, device=device)
        self.off = 0

    def take(self, n: int):
        """n zeroed doubles, 16-byte aligned (the kernels read accumulator pairs as double2)."""
        step = (n + 1) & ~1
        if self.off + step > self.buf.numel():
            return torch.zeros((step,), dtype=torch.float64, device=self.buf.device)[:n]
        out = self.buf[self.off:self.off + n]
        self.off += step
        return out


def _stat_buffers(g: L.ConvGeom, pool: "ZeroPool", scratch_bytes: int, device):
    """(tickets, scratch, stats) of a fused-statistics convolution: N * ceil(Co/32) zeroed uint32 tickets from the pool, the per-tile
    partial-sum slots (any content), the (mean, rstd) output."""
    n_tickets = g.N * ((g.Co + 31) // 32)
    tickets = pool.take((n_tickets + 1) // 2)
    scratch = torch.empty((scratch_bytes,), dtype=torch.uint8, device=device)
    stats = torch.empty((g.N, g.Co, 2), dtype=torch.float32, device=device)
    return tickets, scratch, stats


def conv_gather_stats(x, wp, bias, g: L.ConvGeom, pool: "ZeroPool", engine=L.ENGINE_AUTO):
    """Convolution + InstanceNorm statistics.  Returns (y, stats[N][Co][2] = (mean, rstd)).  On the tcgen05 engine the sums come out
    of the conv epilogue (no extra pass over y; per-tile partial sums added in tile order by the last CTA: deterministic); otherwise
    conv followed by the shifted-sum statistics kernels."""
    lib = L.load()
    if pool is not None and lib.ctagan_conv_gather_engine(ctypes.byref(g), engine) == 2:
        ensure_device()
        y = torch.empty((g.N, g.Ho, g.Wo, g.Co), dtype=x.dtype, device=x.device)
        nbytes = int(lib.ctagan_conv_gather_stats_scratch_bytes(ctypes.byref(g), engine))
        tickets, scratch, stats = _stat_buffers(g, pool, nbytes, x.device)
        _count(1)
        L.check(lib.ctagan_conv_gather_stats(ctypes.byref(g), _p(x), _p(wp), _p(bias), _p(y), _p(tickets), _p(scratch), nbytes, _p(stats),
                                             engine, _stream()))
        return y, stats
    y = conv_gather(x, wp, bias, g, engine)
    return y, instnorm_stats(y)


def _groups(slots) -> L.ConvGroups:
    gr = L.ConvGroups()
    gr.groups = len(slots)
    for k, s in enumerate(slots):
        gr.slot[k] = int(s)
    return gr


def conv_gather_grouped_supported(g: L.ConvGeom, slots) -> bool:
    return bool(L.load().ctagan_conv_gather_grouped_supported(ctypes.byref(g), ctypes.byref(_groups(slots))))


def conv_gather_grouped(x, wp, g: L.ConvGeom, slots, pool: "ZeroPool" = None):
    """Grouped convolution (see ctagan_conv_gather_grouped): image group k uses slot slots[k] of the packed buffer wp[slots][...].
    With a ZeroPool the InstanceNorm statistics come out of the epilogue: returns (y, stats), else y."""
    _require_cuda(x, wp)
    ensure_device()
    y = torch.empty((g.N, g.Ho, g.Wo, g.Co), dtype=x.dtype, device=x.device)
    tickets = scratch = stats = None
    nbytes = 0
    if pool is not None:
        nbytes = int(L.load().ctagan_conv_gather_stats_scratch_bytes(ctypes.byref(g), L.ENGINE_AUTO))
        tickets, scratch, stats = _stat_buffers(g, pool, nbytes, x.device)
    _count(1)
    L.check(L.load().ctagan_conv_gather_grouped(ctypes.byref(g), ctypes.byref(_groups(slots)), _p(x), _p(wp), None, _p(y), _p(tickets),
                                                _p(scratch), nbytes, _p(stats), _stream()))
    return (y, stats) if pool is not None else y


def conv_wgrad_grouped_workspace(g: L.ConvGeom, groups: int) -> int:
    return int(L.load().ctagan_conv_wgrad_grouped_workspace_bytes(ctypes.byref(g), groups))


def conv_wgrad_grouped(gy, gx, g: L.ConvGeom, groups: int, want_bias: bool, ws_bytes: int):
    """One weight gradient per image group: returns (dw[groups,Co,Ci,KH,KW], db[groups,Co] or None)."""
    _require_cuda(gy, gx)
    dw = torch.empty((groups, g.Co, g.Ci, g.KH, g.KW), dtype=torch.float32, device=gy.device)
    db = torch.empty((groups, g.Co), dtype=torch.float32, device=gy.device) if want_bias else None
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=gy.device)
    _count(2)
    L.check(L.load().ctagan_conv_wgrad_grouped(ctypes.byref(g), groups, _p(gy), _p(gx), _p(dw), _p(db), _p(ws), ws_bytes, _stream()))
    return dw, db


def conv_wgrad(gy, gx, g: L.ConvGeom, want_bias: bool, engine=L.ENGINE_AUTO, out_w=None, out_b=None, accumulate=False, packed=False):
    """Weight (and bias) gradient.  out_w / out_b: persistent gradient buffers to write into (e.g. slices of an optimiser's flat
    gradient bucket) instead of fresh tensors; accumulate: add to them instead of overwriting (the second use of a network in one
    backward pass); packed: out_w is the contiguous [Co][KH][KW][Ci] storage of a channels-last gradient."""
    _require_cuda(gy, gx)
    dw = out_w if out_w is not None else torch.empty((g.Co, g.Ci, g.KH, g.KW), dtype=torch.float32, device=gy.device)
    db = None
    if want_bias:
        db = out_b if out_b is not None else torch.empty((g.Co,), dtype=torch.float32, device=gy.device)
    assert dw.is_contiguous() and (db is None or db.is_contiguous())
    ws_bytes = L.load().ctagan_conv_wgrad_workspace_bytes(ctypes.byref(g), engine)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=gy.device) if ws_bytes else None
    _count(3 if ws_bytes and want_bias else (2 if ws_bytes else 1))
    L.check(L.load().ctagan_conv_wgrad(ctypes.byref(g), _p(gy), _p(gx), _p(dw), _p(db), _p(ws), ws_bytes, engine, int(bool(accumulate)) | (2 if packed else 0), _stream()))
    return dw, db


def pack_weights(w: torch.Tensor, mode: int, dtype: torch.dtype):
    _require_cuda(w)
    O, I, KH, KW = w.shape
    shape = (O, KH, KW, I) if mode == 0 else (I, KH, KW, O)
    wp = torch.empty(shape, dtype=dtype, device=w.device)
    _count(1)
    L.check(L.load().ctagan_pack_weights(_p(w), _p(wp), O, I, KH, KW, mode, _DT[dtype], _stream()))
    return wp


def instnorm_stats(x):
    N, H, W, C = x.shape
    stats = torch.empty((N, C, 2), dtype=torch.float32, device=x.device)
    n_acc = int(L.load().ctagan_instnorm_stats_scratch_doubles(N, H * W, C, dt(x)))
    acc = torch.empty((n_acc,), dtype=torch.float64, device=x.device)
    _count(2)
    L.check(L.load().ctagan_instnorm_stats(_p(x), _p(stats), _p(acc), N, H * W, C, dt(x), _stream()))
    return stats


def norm_act_pad(x, stats, act, pad, res=None, res_pad=0):
    N, H, W, C = x.shape
    out = torch.empty((N, H + 2 * pad, W + 2 * pad, C), dtype=x.dtype, device=x.device)
    _count(1)
    L.check(L.load().ctagan_norm_act_pad(_p(x), _p(stats), _p(res), res_pad, _p(out), N, H, W, C, pad, act, dt(x), _stream()))
    return out


def norm_act_pad_bwd(gout, x, stats, act, pad, addend=None, out_pad=0, pool=None, want_g=False):
    """Backward of norm_act_pad (see ctagan_norm_act_pad_bwd).  want_g: also return fold(gout) + addend (the skip-connection gradient)."""
    N, Hp, Wp, C = gout.shape
    H, W = Hp - 2 * pad, Wp - 2 * pad
    dx = torch.empty((N, H + 2 * out_pad, W + 2 * out_pad, C), dtype=gout.dtype, device=gout.device)
    g_out = torch.empty((N, H, W, C), dtype=gout.dtype, device=gout.device) if want_g else None
    lib = L.load()
    launches = lib.ctagan_norm_act_pad_bwd_launches(int(stats is not None), H, W, C, dt(gout))
    acc, scratch, zeroed = None, None, 0
    if stats is not None and launches == 2:
        if pool is not None:
            acc, zeroed = pool.take(N * C * 2 + N), 1
        else:
            acc = torch.empty((N * C * 2 + N,), dtype=torch.float64, device=gout.device)
        n_scr = int(lib.ctagan_norm_act_pad_bwd_scratch_doubles(1, N, H, W, C, dt(gout)))
        scratch = torch.empty((n_scr,), dtype=torch.float64, device=gout.device)
    _count(launches)
    L.check(lib.ctagan_norm_act_pad_bwd(_p(gout), _p(x), _p(stats), _p(addend), _p(dx), _p(g_out), _p(acc), zeroed, _p(scratch), N, H, W, C,
                                        pad, act, out_pad, dt(gout), _stream()))
    return (dx, g_out) if want_g else dx


def act_bwd(gy, y, act):
    dx = torch.empty_like(gy)
    _count(1)
    L.check(L.load().ctagan_act_bwd(_p(gy), _p(y), _p(dx), gy.numel(), act, dt(gy), _stream()))
    return dx


def maxpool2_fwd(x):
    N, H, W, C = x.shape
    y = torch.empty((N, H // 2, W // 2, C), dtype=x.dtype, device=x.device)
    _count(1)
    L.check(L.load().ctagan_maxpool2_fwd(_p(x), _p(y), N, H, W, C, dt(x), _stream()))
    return y


def maxpool2_bwd(gy, x, addend=None):
    N, H, W, C = x.shape
    gx = torch.empty_like(x)
    _count(1)
    L.check(L.load().ctagan_maxpool2_bwd(_p(gy), _p(x), _p(addend), _p(gx), N, H, W, C, dt(x), _stream()))
    return gx


def upsample2x_cat_fwd(x, skip):
    N, H, W, C1 = x.shape
    C2 = skip.shape[3]
    out = torch.empty((N, 2 * H, 2 * W, C1 + C2), dtype=x.dtype, device=x.device)
    _count(1)
    L.check(L.load().ctagan_upsample2x_cat_fwd(_p(x), _p(skip), _p(out), N, H, W, C1, C2, dt(x), _stream()))
    return out


def upsample2x_cat_bwd(gout, C1):
    N, Ho, Wo, C = gout.shape
    H, W, C2 = Ho // 2, Wo // 2, C - C1
    gx = torch.empty((N, H, W, C1), dtype=gout.dtype, device=gout.device)
    gskip = torch.empty((N, Ho, Wo, C2), dtype=gout.dtype, device=gout.device)
    _count(1)
    L.check(L.load().ctagan_upsample2x_cat_bwd(_p(gout), _p(gx), _p(gskip), N, H, W, C1, C2, dt(gout), _stream()))
    return gx, gskip


def nchw_to_nhwc(x: torch.Tensor, dtype: torch.dtype):
    """fp32 NCHW (boundary) -> NHWC `dtype` (internal)."""
    _require_cuda(x)
    ensure_device()
    x = x.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    N, C, H, W = x.shape
    out = torch.empty((N, H, W, C), dtype=dtype, device=x.device)
    _count(1)
    L.check(L.load().ctagan_nchw_to_nhwc(_p(x), _p(out), N, C, H * W, _DT[dtype], _stream()))
    return out


def nhwc_to_nchw(x: torch.Tensor):
    N, H, W, C = x.shape
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device)
    _count(1)
    L.check(L.load().ctagan_nhwc_to_nchw(_p(x), _p(out), N, C, H * W, dt(x), _stream()))
    return out


def plane_mean_fwd(x_nhwc):
    N, H, W, C = x_nhwc.shape
    out = torch.empty((N, C), dtype=torch.float32, device=x_nhwc.device)
    _count(1)
    L.check(L.load().ctagan_plane_mean_fwd(_p(x_nhwc), _p(out), N, H * W, C, dt(x_nhwc), _stream()))
    return out


def plane_mean_bwd(gout, shape, dtype):
    N, H, W, C = shape
    gx = torch.empty(shape, dtype=dtype, device=gout.device)
    _count(1)
    L.check(L.load().ctagan_plane_mean_bwd(_p(gout), _p(gx), N, H * W, C, _DT[dtype], _stream()))
    return gx


def warp_fwd(src, flow):
    B, C, H, W = src.shape
    out = torch.empty_like(src)
    _count(1)
    L.check(L.load().ctagan_warp_fwd(_p(src), _p(flow), _p(out), B, C, H, W, _stream()))
    return out


def warp_bwd(gout, src, flow, need_src=True, need_flow=True):
    B, C, H, W = src.shape
    gsrc = torch.empty_like(src) if need_src else None
    gflow = torch.empty_like(flow) if need_flow else None
    ws_bytes = int(L.load().ctagan_warp_bwd_workspace_bytes(B, C, H, W)) if need_src else 0
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=src.device) if ws_bytes else None
    _count(3 if need_src else 1)
    L.check(L.load().ctagan_warp_bwd(_p(gout), _p(src), _p(flow), _p(gsrc), _p(gflow), _p(ws), ws_bytes, B, C, H, W, _stream()))
    return gsrc, gflow


LOSS_ACC_DOUBLES = 1024        # CTAGAN_LOSS_ACC_DOUBLES


def _scalar_out(ref):
    return (torch.empty((), dtype=torch.float32, device=ref.device),
            torch.empty((LOSS_ACC_DOUBLES,), dtype=torch.float64, device=ref.device))


def l1_fwd(a, b):
    loss, acc = _scalar_out(a)
    _count(1)
    L.check(L.load().ctagan_l1_fwd(_p(a), _p(b), _p(loss), _p(acc), a.numel(), _stream()))
    return loss


def l1_bwd(a, b, gloss):
    ga = torch.empty_like(a)
    _count(1)
    L.check(L.load().ctagan_l1_bwd(_p(a), _p(b), _p(gloss), _p(ga), a.numel(), _stream()))
    return ga


def _mse_target(target):
    """(scalar, device pointer): a float constant, or a 1-element fp32 CUDA tensor that the kernel reads on the device."""
    if torch.is_tensor(target):
        return 0.0, _p(target)
    return float(target), None


def mse_const_fwd(p, target):
    loss, acc = _scalar_out(p)
    tv, tp = _mse_target(target)
    _count(1)
    L.check(L.load().ctagan_mse_const_fwd(_p(p), tv, tp, _p(loss), _p(acc), p.numel(), _stream()))
    return loss


def mse_const_bwd(p, target, gloss):
    gp = torch.empty_like(p)
    tv, tp = _mse_target(target)
    _count(1)
    L.check(L.load().ctagan_mse_const_bwd(_p(p), tv, tp, _p(gloss), _p(gp), p.numel(), _stream()))
    return gp


def smooth_fwd(flow):
    B, C, H, W = flow.shape
    loss, acc = _scalar_out(flow)
    _count(1)
    L.check(L.load().ctagan_smooth_fwd(_p(flow), _p(loss), _p(acc), B, C, H, W, _stream()))
    return loss


def smooth_bwd(flow, gloss):
    B, C, H, W = flow.shape
    g = torch.empty_like(flow)
    _count(1)
    L.check(L.load().ctagan_smooth_bwd(_p(flow), _p(gloss), _p(g), B, C, H, W, _stream()))
    return g


def masked_l1_fwd(warped, b1, b2):
    loss, acc = _scalar_out(warped)
    _count(1)
    L.check(L.load().ctagan_masked_l1_fwd(_p(warped), _p(b1), _p(b2), _p(loss), _p(acc), warped.numel(), _stream()))
    return loss


def masked_l1_bwd(warped, b1, b2, gloss):
    g = torch.empty_like(warped)
    _count(1)
    L.check(L.load().ctagan_masked_l1_bwd(_p(warped), _p(b1), _p(b2), _p(gloss), _p(g), warped.numel(), _stream()))
    return g


def interleave2(a, b, dtype):
    """cat([a, b], 1) for 1-channel fp32 NCHW images -> NHWC [N,H,W,2] of `dtype`."""
    _require_cuda(a, b)
    ensure_device()
    a = a.contiguous().float(); b = b.contiguous().float()
    N, _, H, W = a.shape
    out = torch.empty((N, H, W, 2), dtype=dtype, device=a.device)
    _count(1)
    L.check(L.load().ctagan_interleave2(_p(a), _p(b), _p(out), N * H * W, _DT[dtype], _stream()))
    return out


def deinterleave2(src, need_a=True, need_b=True):
    N, H, W, _ = src.shape
    a = torch.empty((N, 1, H, W), dtype=torch.float32, device=src.device) if need_a else None
    b = torch.empty((N, 1, H, W), dtype=torch.float32, device=src.device) if need_b else None
    _count(1)
    L.check(L.load().ctagan_deinterleave2(_p(src), _p(a), _p(b), N * H * W, dt(src), _stream()))
    return a, b


def pack_weights_multi(entries, dtype):
    """entries: list of (w fp32 OIHW, wp packed buffer, mode).  One kernel launch for all of them."""
    n = len(entries)
    arr = (L.PackItem * n)()
    for k, (w, wp, mode) in enumerate(entries):
        O, I, KH, KW = w.shape
        arr[k] = L.PackItem(w.data_ptr(), wp.data_ptr(), O, I, KH, KW, mode)
    ensure_device()
    # one launch per shared-memory class of filter size (<= 9 taps, <= 16 taps, larger), 48 table entries each
    per_class = {}
    for w, _, _ in entries:
        taps = w.shape[2] * w.shape[3]
        c = 0 if taps <= 9 else (1 if taps <= 16 else 2)
        per_class[c] = per_class.get(c, 0) + 1
    _count(sum((v + 47) // 48 for v in per_class.values()))
    L.check(L.load().ctagan_pack_weights_multi(arr, n, _DT[dtype], _stream()))
