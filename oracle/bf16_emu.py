"""bf16-emulating variant of the parity oracle -- TEST INFRASTRUCTURE ONLY (same rules as oracle/restate.py).

The benchmarked mode of the product computes in bf16 (bf16 activations and packed weights, fp32 accumulation, fp32 statistics,
losses and master weights).  Against the fp32 oracle such a run can only be checked loosely: operand rounding through 24
InstanceNorm-renormalised layers moves a random-init generator's output by ~3e-2 max-rel whatever the kernels do (SURVEY.md
App. C), so a kernel bug of that size would hide.  This module restates the SAME networks (same reference lines as
oracle/restate.py, cited there) on the CPU in fp32 arithmetic but with a rounding to bf16 at exactly the points where the sm_100a
kernels store bf16:

  * module inputs (`ctagan_nchw_to_nhwc`, `ctagan_interleave2`) and packed weights (`ctagan_pack_weights*`);
  * every convolution output: fp32 accumulator (+ bias, activation) -> bf16;  biases in front of a non-affine InstanceNorm are
    skipped, as in the kernels (they are mathematically dead);
  * InstanceNorm: (mean, rstd) in fp64 -> fp32 from the UNROUNDED fp32 accumulator when the layer runs on the tcgen05 engine
    (statistics fused into the conv epilogue) and from the stored bf16 output otherwise (`ctagan_instnorm_stats`); the normalised,
    activated (+ residual) value -> bf16 (`ctagan_norm_act_pad`);
  * bilinear upsampling output -> bf16; max-pool / concat / reflection pad are exact;
  * the same roundings on the gradients that flow backwards through those points (every backward kernel stores bf16), while
    weight gradients, losses and the warp stay fp32.

It is not bit-exact against the GPU (accumulation order inside a dot product differs), but the remaining difference is a few
bf16 ulps per tensor instead of the whole operand-rounding envelope: the bf16 tests compare against it with tolerances an order of
magnitude tighter than against the fp32 oracle.

Pinning: with rounding switched off (`emulate(enabled=False)`) every function here must reproduce oracle/restate.py -- and hence
the real reference modules (tests/golden, oracle/make_golden.py) -- to fp32 round-off (tests/test_oracle_golden.py).

Use:   with bf16_emu.emulate():  ref = restate.cyc_step(state, a, b)        # the restated iteration bodies, bf16-emulated
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn.functional as F

from . import restate as R

Tensor = torch.Tensor
_ON = {"v": True}


class _RoundSTE(torch.autograd.Function):
    """bf16 rounding of a stored activation; the gradient that flows back through the same point is stored in bf16 too."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def q(x: Tensor) -> Tensor:
    return _RoundSTE.apply(x) if _ON["v"] else x


def qw(w: Tensor) -> Tensor:
    """packed bf16 copy of an fp32 master weight; its gradient is the fp32 weight gradient (not rounded)."""
    return w + (w.to(torch.bfloat16).to(w.dtype) - w).detach() if _ON["v"] else w


def _on_tc(n, ci, co, ho, wo, strided) -> bool:
    """Does this conv (+ InstanceNorm statistics) run on the tcgen05 engine?  (conv_tc.cu: tc_gather_kind)"""
    if ci % 8 or ci < 32 or co % 32 or n * ho * wo < 32:
        return False
    return wo >= 2 if strided else True


def _on_fewin_tc(n, ci, co, k, ho, wo) -> bool:
    """Does this 1-2 input channel conv run on the patch-matrix tensor-core kernel (statistics from the fp32 accumulator)?
    (conv_tc.cu: plan_fewin_tc)"""
    return ci <= 2 and 32 < co <= 64 and co % 8 == 0 and 32 <= k * k * ci <= 64 and k <= 7 and n * ho * wo >= 4096 and wo >= 16


def _norm_act(c: Tensor, act, tc: bool, res: Tensor = None) -> Tensor:
    """c: fp32 conv accumulator.  Returns the stored bf16 activation act((r - mean) * rstd) (+ res)."""
    r = q(c)
    src = (c if tc else r).detach().double() if _ON["v"] else None
    if _ON["v"]:
        # statistics are a function of the conv output: keep them in the autograd graph (through fp32 `c` / stored `r`)
        base = c if tc else r
        mean = base.mean((2, 3), keepdim=True)
        var = (base * base).mean((2, 3), keepdim=True) - mean * mean
        # the VALUES are those of the kernels' fp64 reduction rounded to fp32
        m64 = src.mean((2, 3), keepdim=True)
        v64 = ((src * src).mean((2, 3), keepdim=True) - m64 * m64).clamp_min(0)
        mean = mean + (m64.float() - mean).detach()
        rstd = torch.rsqrt(var.clamp_min(0) + 1e-5)
        rstd = rstd + ((v64 + 1e-5).rsqrt().float() - rstd).detach()
        y = (r - mean) * rstd
    else:
        y = F.instance_norm(c, eps=1e-5)
    if act == "relu":
        y = F.relu(y)
    elif act == "lrelu":
        y = F.leaky_relu(y, 0.2)
    if res is not None:
        y = res + y
    return q(y)


def _bias(b):
    return None if _ON["v"] else b          # dead bias in front of InstanceNorm: skipped by the kernels


def generator_forward(sd, x: Tensor, n_blocks: int = 9) -> Tensor:
    """restate.generator_forward (Model/CycleGan.py:66-71) with the kernels' bf16 storage points."""
    w = lambda k: qw(sd[k + ".weight"])
    b = lambda k: sd[k + ".bias"]
    n = x.shape[0]
    x = q(x)
    c = F.conv2d(R._rpad(x, 3), w("model_head.1"), _bias(b("model_head.1")))
    # Cin = 1: patch-matrix tensor-core kernel with fused statistics on large maps, else CUDA-core kernel + instnorm_stats
    x = _norm_act(c, "relu", tc=_on_fewin_tc(n, x.shape[1], c.shape[1], 7, c.shape[2], c.shape[3]))
    for k in ("model_head.4", "model_head.7"):
        c = F.conv2d(x, w(k), _bias(b(k)), stride=2, padding=1)
        x = _norm_act(c, "relu", _on_tc(n, x.shape[1], c.shape[1], c.shape[2], c.shape[3], True))
    for i in range(n_blocks):
        k1, k2 = f"model_body.{i}.conv_block.1", f"model_body.{i}.conv_block.5"
        c = F.conv2d(R._rpad(x, 1), w(k1), _bias(b(k1)))
        tc = _on_tc(n, 256, 256, c.shape[2], c.shape[3], False)
        t = _norm_act(c, "relu", tc)
        c = F.conv2d(R._rpad(t, 1), w(k2), _bias(b(k2)))
        x = _norm_act(c, None, tc, res=x)
    for k in ("model_tail.0", "model_tail.3"):
        c = F.conv_transpose2d(x, w(k), _bias(b(k)), stride=2, padding=1, output_padding=1)
        x = _norm_act(c, "relu", c.shape[3] >= 32 and _on_tc(n, x.shape[1], c.shape[1], c.shape[2], c.shape[3], False))    # (output-phase launches need Wo >= 32)
    return q(torch.tanh(F.conv2d(R._rpad(x, 3), w("model_tail.7"), b("model_tail.7"))))


def discriminator_features(sd, x: Tensor, key_fmt: str = "model.{i}") -> List[Tensor]:
    """restate.discriminator_features (Model/CycleGan.py:78-94) with the kernels' bf16 storage points."""
    idx = [0, 2, 5, 8, 11]
    strides = [2, 2, 2, 1, 1]
    feats = []
    n = x.shape[0]
    x = q(x)
    for j, (i, s) in enumerate(zip(idx, strides)):
        name = key_fmt.format(i=i, j=j)
        wt, bs = qw(sd[name + ".weight"]), sd[name + ".bias"]
        if 1 <= j <= 3:
            c = F.conv2d(x, wt, _bias(bs), stride=s, padding=1)
            x = _norm_act(c, "lrelu", _on_tc(n, x.shape[1], c.shape[1], c.shape[2], c.shape[3], True))
        elif j == 0:
            x = q(F.leaky_relu(F.conv2d(x, wt, bs, stride=s, padding=1), 0.2))
        else:
            x = q(F.conv2d(x, wt, bs, stride=s, padding=1))
        feats.append(x)
    return feats


def _reg_conv(sd, name, x, k, act=True):
    c = F.conv2d(x, qw(sd[name + ".conv2d.weight"]), sd[name + ".conv2d.bias"], stride=1, padding=(k - 1) // 2)
    return q(F.leaky_relu(c, 0.2) if act else c)


def _reg_resblocks(sd, prefix, x, n):
    nb = x.shape[0]
    for i in range(n):
        k1, k2 = f"{prefix}.model.{i}.conv_block.1", f"{prefix}.model.{i}.conv_block.5"
        c = F.conv2d(R._rpad(x, 1), qw(sd[k1 + ".weight"]), _bias(sd[k1 + ".bias"]))
        tc = _on_tc(nb, x.shape[1], c.shape[1], c.shape[2], c.shape[3], False)
        t = _norm_act(c, "relu", tc)
        c = F.conv2d(R._rpad(t, 1), qw(sd[k2 + ".weight"]), _bias(sd[k2 + ".bias"]))
        x = _norm_act(c, None, tc, res=x)
    return x


def reg_forward(sd, img_a: Tensor, img_b: Tensor) -> Tensor:
    """restate.reg_forward (trainer/reg.py:76-99) with the kernels' bf16 storage points."""
    p = "offset_map."
    x = q(torch.cat([img_a, img_b], 1))
    skips = {}
    nd = len(R.REG_NDF)
    for n in range(1, nd + 1):
        x = _reg_conv(sd, f"{p}down_{n}.conv_0", x, 3)
        x = _reg_resblocks(sd, f"{p}down_{n}.conv_0.resnet_block", x, 1)
        skips[n] = x
        x = F.max_pool2d(x, 2)
    x = _reg_conv(sd, p + "c1", x, 1)
    x = _reg_resblocks(sd, p + "t", x, 3)
    x = _reg_conv(sd, p + "c2", x, 1)
    for n in range(nd, 0, -1):
        s = skips[n]
        x = q(F.interpolate(x, (s.size(2), s.size(3)), mode="bilinear"))
        x = torch.cat([x, s], 1)
        x = _reg_conv(sd, f"{p}up_{n}", x, 3)
    x = _reg_resblocks(sd, p + "refine.0", x, 1)
    x = _reg_conv(sd, p + "refine.1", x, 1)
    return _reg_conv(sd, p + "output", x, 3, act=False)


class emulate:
    """Context manager: inside it the restated networks of oracle/restate.py (and therefore its iteration bodies cyc_step / reg_step /
    hd_x2_step / p2p_step, which look the forwards up at call time) run with the bf16 storage points of the kernels."""

    _NAMES = ("generator_forward", "discriminator_features", "reg_forward")

    def __init__(self, enabled: bool = True):
        self.enabled = enabled

    def __enter__(self):
        self._saved = {k: getattr(R, k) for k in self._NAMES}
        self._prev = _ON["v"]
        _ON["v"] = self.enabled
        R.generator_forward, R.discriminator_features, R.reg_forward = generator_forward, discriminator_features, reg_forward
        return self

    def __exit__(self, *exc):
        for k, v in self._saved.items():
            setattr(R, k, v)
        _ON["v"] = self._prev
        return False
