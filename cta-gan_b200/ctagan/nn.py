"""nn.Module surface of the hot path: same constructors, sub-module names, parameter creation order and state_dict layout
as the reference (Model/CycleGan.py, Model/HdGan.py, trainer/reg.py, trainer/transformer.py, trainer/utils.py), with every
forward/backward executed by the sm_100a kernels of libctagan.so through ctagan.engine."""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

from . import engine as E
from . import ops


class _Slot(nn.Module):
    """Parameter-free placeholder keeping nn.Sequential indices identical to the reference (pads, norms, activations:
    their arithmetic is fused into the neighbouring kernels)."""

    def __init__(self, what: str):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return f"fused:{self.what}"

    def forward(self, x):  # pragma: no cover - never called; the owning module runs the fused schedule
        raise RuntimeError("fused placeholder: call the owning module")


def _fresh_leaves(params, detach: bool):
    return [p.detach() if detach else p for p in params]


# ----------------------------------------------------------------------------------------------------------------------
# Generator
# ----------------------------------------------------------------------------------------------------------------------


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, x, *params):
        need = any(ctx.needs_input_grad)      # all False under torch.no_grad()
        out, saved = E.generator_forward(plan, x, save=need)
        ctx.plan, ctx.saved = plan, saved
        ctx.leaves = params if need else None
        return out

    @staticmethod
    def backward(ctx, dout):
        need_dx = ctx.needs_input_grad[1]
        need_dw = any(ctx.needs_input_grad[2:])
        dx, grads = E.generator_backward(ctx.plan, ctx.saved, dout.contiguous(), need_dx, need_dw)
        ctx.saved = None
        if isinstance(ctx.plan, E.GroupedGeneratorPlan):
            grads = E.split_group_grads(grads, ctx.plan.G) if need_dw else [None] * len(ctx.plan.params)
        if need_dw and E.deferred() is not None:
            grads = E.deferred().take(ctx.leaves, grads, ctx.needs_input_grad[2:])
        elif need_dw:
            grads = E.strip_in_place(ctx.leaves, grads)
        ctx.leaves = None
        return (None, dx, *grads)


class ResidualBlock(nn.Module):
    """Model/CycleGan.py:6-21.  conv_block indices: 0 pad, 1 conv, 2 IN, 3 ReLU, 4 pad, 5 conv, 6 IN."""

    def __init__(self, in_features):
        super().__init__()
        self.conv_block = nn.Sequential(_Slot("reflect1"), nn.Conv2d(in_features, in_features, 3), _Slot("instnorm"),
                                        _Slot("relu"), _Slot("reflect1"), nn.Conv2d(in_features, in_features, 3),
                                        _Slot("instnorm"))

    def forward(self, x):
        """x + IN(conv3(RP1(relu(IN(conv3(RP1(x)))))))  (Model/CycleGan.py:20-21).  Inside a Generator the block runs as part of the fused
        schedule; called on its own it runs the same kernels through the single-layer functions below."""
        c1, c2 = self.conv_block[1], self.conv_block[5]
        return _to_nchw(_res_block(_to_nhwc(x), c1.weight, c1.bias, c2.weight, c2.bias))


class Generator(nn.Module):
    """ResNet generator, Model/CycleGan.py:23-71 (== Model/HdGan.py:65-113)."""

    def __init__(self, input_nc, output_nc, n_residual_blocks=9):
        super().__init__()
        self.model_head = nn.Sequential(
            _Slot("reflect3"), nn.Conv2d(input_nc, 64, 7), _Slot("instnorm"), _Slot("relu"),
            nn.Conv2d(64, 128, 3, stride=2, padding=1), _Slot("instnorm"), _Slot("relu"),
            nn.Conv2d(128, 256, 3, stride=2, padding=1), _Slot("instnorm"), _Slot("relu"))
        self.model_body = nn.Sequential(*[ResidualBlock(256) for _ in range(n_residual_blocks)])
        self.model_tail = nn.Sequential(
            nn.ConvTranspose2d(256, 128, 3, stride=2, padding=1, output_padding=1), _Slot("instnorm"), _Slot("relu"),
            nn.ConvTranspose2d(128, 64, 3, stride=2, padding=1, output_padding=1), _Slot("instnorm"), _Slot("relu"),
            _Slot("reflect3"), nn.Conv2d(64, output_nc, 7), _Slot("tanh"))
        self.n_residual_blocks = n_residual_blocks
        self._plan = None

    def _get_plan(self):
        params = list(self.parameters())
        if self._plan is None or any(a is not b for a, b in zip(self._plan.params, params)):
            self._plan = E.GeneratorPlan(params, self.n_residual_blocks)
        return self._plan

    def _apply(self, fn, *a, **k):
        self._plan = None
        return super()._apply(fn, *a, **k)

    def prepack(self, force=False):
        """Pack the current weights (bf16 / re-laid-out copies) on the current stream; lets several passes of this network run
        concurrently on side streams afterwards.  `force` re-packs even copies the cache believes fresh (used where a trainer
        knows the optimizer has just stepped, so that the pack kernels are captured in that position of a CUDA graph)."""
        E.prepack_prims(self._get_plan().prims(), E.get_precision(), force)

    def forward(self, x, params=None):
        """`params` (extension): alternative leaf tensors aliasing this module's parameters (e.g. `p.detach().requires_grad_()`), so that
        two passes of the network inside one autograd graph accumulate their gradients separately (see Cyc_Trainer.phase_G)."""
        if x.shape[2] % 4 or x.shape[3] % 4:
            raise ValueError("Generator needs H and W to be multiples of 4")
        # the kernels index with 32 bits: the largest activation (64 channels at full resolution, reflection-padded by 3) must stay below
        # 2^31 elements.  Every op of the network is per sample (InstanceNorm statistics are per (n, c)), so a larger batch -- the
        # 128- and 256-slice points of the inference sweep at 512^2 -- is run in chunks with identical results.
        per_sample = 64 * (x.shape[2] + 6) * (x.shape[3] + 6)
        if x.shape[0] * per_sample >= 2 ** 31:
            chunk = max(1, (2 ** 31 - 1) // per_sample)
            return torch.cat([self.forward(xc, params) for xc in x.split(chunk)])
        plan = self._get_plan()
        return _GeneratorFn.apply(plan, x, *(plan.params if params is None else params))


# ----------------------------------------------------------------------------------------------------------------------
# Discriminators
# ----------------------------------------------------------------------------------------------------------------------


class _DiscriminatorFn(torch.autograd.Function):
    """x -> last feature map [N,1,h,w]; intermediates are returned detached through `sink` (feature taps)."""

    @staticmethod
    def forward(ctx, plan, sink, x, *params):
        need = any(ctx.needs_input_grad)
        out, acts, saved = E.discriminator_forward(plan, x, save=need)
        if sink is not None:
            sink.extend(acts)
        ctx.plan, ctx.saved = plan, saved
        ctx.leaves = params if need else None
        return out

    @staticmethod
    def backward(ctx, dout):
        need_dx = ctx.needs_input_grad[2]
        need_dw = any(ctx.needs_input_grad[3:])
        dx, gw = E.discriminator_backward(ctx.plan, ctx.saved, dout.contiguous(), need_dx, need_dw)
        ctx.saved = None
        if isinstance(ctx.plan, E.GroupedDiscriminatorPlan):
            gw = E.split_group_grads(gw, ctx.plan.G)
        if need_dw and E.deferred() is not None:
            gw = E.deferred().take(ctx.leaves, gw, ctx.needs_input_grad[3:])
        elif need_dw:
            gw = E.strip_in_place(ctx.leaves, gw)
        ctx.leaves = None
        return (None, None, dx, *gw)


class _PlaneMeanFn(torch.autograd.Function):
    """F.avg_pool2d(x, x.size()[2:]).view(N, -1) for a 1-channel map (Model/CycleGan.py:103)."""

    @staticmethod
    def forward(ctx, x):
        N, C, H, W = x.shape
        ctx.shape = (N, H, W, C)
        xx = x.contiguous()
        if C != 1:
            xx = ops.nchw_to_nhwc(xx, torch.float32)
        return ops.plane_mean_fwd(xx.view(N, H, W, C))

    @staticmethod
    def backward(ctx, g):
        N, H, W, C = ctx.shape
        gx = ops.plane_mean_bwd(g.contiguous(), ctx.shape, torch.float32)
        return gx.view(N, C, H, W) if C == 1 else ops.nhwc_to_nchw(gx)


def plane_mean(x):
    return _PlaneMeanFn.apply(x)


class _DiscBase(nn.Module):
    _plan = None

    def _conv_params(self):
        return list(self.parameters())

    def _get_plan(self, detach=False):
        params = self._conv_params()
        if self._plan is None or any(a is not b for a, b in zip(self._plan.params, params)):
            self._plan = E.DiscriminatorPlan(params)
        return self._plan

    def _apply(self, fn, *a, **k):
        self._plan = None
        return super()._apply(fn, *a, **k)

    def prepack(self, force=False):
        E.prepack_prims(self._get_plan().prims(), E.get_precision(), force)

    def _run(self, x, sink=None, freeze=False):
        plan = self._get_plan()
        params = [p.detach() for p in plan.params] if freeze else plan.params
        return _DiscriminatorFn.apply(plan, sink, x, *params)


class Discriminator(_DiscBase):
    """PatchGAN discriminator, Model/CycleGan.py:73-103 (== Model/HdGan.py:115-145).

    `freeze=True` (extension) treats the weights as constants for this call: the generator phase of the trainers uses it,
    because the reference computes and then discards these weight gradients."""

    def __init__(self, input_nc):
        super().__init__()
        self.model = nn.Sequential(
            nn.Conv2d(input_nc, 64, 4, stride=2, padding=1), _Slot("lrelu"),
            nn.Conv2d(64, 128, 4, stride=2, padding=1), _Slot("instnorm"), _Slot("lrelu"),
            nn.Conv2d(128, 256, 4, stride=2, padding=1), _Slot("instnorm"), _Slot("lrelu"),
            nn.Conv2d(256, 512, 4, padding=1), _Slot("instnorm"), _Slot("lrelu"),
            nn.Conv2d(512, 1, 4, padding=1))

    def forward(self, x, freeze=False):
        return plane_mean(self._run(x, freeze=freeze))


_GROUPED_PLANS = {}


def _grouped_plan(nets, kind):
    plans = [n._get_plan() for n in nets]
    key = (kind, *[id(p) for p in plans])
    hit = _GROUPED_PLANS.get(key)
    if hit is None or any(a is not b for a, b in zip(hit.plans, plans)):
        hit = (E.GroupedGeneratorPlan if kind == "G" else E.GroupedDiscriminatorPlan)(plans)
        _GROUPED_PLANS[key] = hit
    return hit


def grouped_generators(nets, x):
    """[nets[0](x[:B]); nets[1](x[B:2B]); ...] for Generators of the same architecture, one kernel launch per layer for all of them
    (extension: the two generators of a CycleGAN iteration work on independent inputs at the same time, CycTrainer.py:144-157)."""
    if x.shape[0] % len(nets) or x.shape[2] % 4 or x.shape[3] % 4:
        raise ValueError("grouped_generators: batch must split evenly over the networks, H and W must be multiples of 4")
    plan = _grouped_plan(nets, "G")
    return _GeneratorFn.apply(plan, x, *plan.params)


def grouped_discriminators(nets, x, freeze=False):
    """The same for `Discriminator`s: returns the pooled predictions [N, 1] of nets[k] on the k-th image group."""
    plan = _grouped_plan(nets, "D")
    params = [p.detach() for p in plan.params] if freeze else plan.params
    return plane_mean(_DiscriminatorFn.apply(plan, None, x, *params))


class NLayerDiscriminator(_DiscBase):
    """Model/HdGan.py:148-205 for the configuration the trainers use (ndf=64, n_layers=3, InstanceNorm, no sigmoid)."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=None, use_sigmoid=False, getIntermFeat=False):
        super().__init__()
        if ndf != 64 or n_layers != 3 or use_sigmoid:
            raise NotImplementedError("only ndf=64, n_layers=3, use_sigmoid=False (the reference's live configuration)")
        self.getIntermFeat, self.n_layers = getIntermFeat, n_layers
        chans = [(input_nc, 64, 2), (64, 128, 2), (128, 256, 2), (256, 512, 1), (512, 1, 1)]
        seqs = []
        for j, (ci, co, s) in enumerate(chans):
            mods = [nn.Conv2d(ci, co, 4, stride=s, padding=1)]
            if 1 <= j <= 3:
                mods.append(_Slot("instnorm"))
            if j <= 3:
                mods.append(_Slot("lrelu"))
            seqs.append(mods)
        if getIntermFeat:
            for j, mods in enumerate(seqs):
                setattr(self, "model" + str(j), nn.Sequential(*mods))
        else:
            self.model = nn.Sequential(*[m for mods in seqs for m in mods])

    def forward(self, x):
        sink: List[torch.Tensor] = []
        last = self._run(x, sink=sink)
        if self.getIntermFeat:
            return [a.permute(0, 3, 1, 2).float() for a in sink] + [last]
        return last


class Discriminator_m(_DiscBase):
    """Model/HdGan.py:207-256.  num_D=1 (the default every trainer uses) runs natively; the centre-crop multi-scale
    branch (:251) is reproduced with one native discriminator pass per scale."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=None, use_sigmoid=False, num_D=1, getIntermFeat=True):
        super().__init__()
        self.num_D, self.n_layers, self.getIntermFeat = num_D, n_layers, getIntermFeat
        self._scales = []
        for i in range(num_D):
            netD = NLayerDiscriminator(input_nc, ndf, n_layers, norm_layer, use_sigmoid, getIntermFeat)
            if getIntermFeat:
                for j in range(n_layers + 2):
                    setattr(self, "scale" + str(i) + "_layer" + str(j), getattr(netD, "model" + str(j)))
            else:
                setattr(self, "layer" + str(i), netD.model)
        self._plans = {}

    def _apply(self, fn, *a, **k):
        self._plans = {}
        return super()._apply(fn, *a, **k)

    def _scale_params(self, i):
        if self.getIntermFeat:
            mods = [getattr(self, "scale" + str(i) + "_layer" + str(j)) for j in range(self.n_layers + 2)]
        else:
            mods = [getattr(self, "layer" + str(i))]
        return [p for m in mods for p in m.parameters()]

    def _scale_plan(self, i):
        params = self._scale_params(i)
        plan = self._plans.get(i)
        if plan is None or any(a is not b for a, b in zip(plan.params, params)):
            plan = E.DiscriminatorPlan(params)
            self._plans[i] = plan
        return plan

    def prepack(self, force=False):
        for i in range(self.num_D):
            E.prepack_prims(self._scale_plan(i).prims(), E.get_precision(), force)

    def forward(self, x, freeze=False):
        result = []
        cur = x
        for i in range(self.num_D):
            s = cur.size(2)
            plan = self._scale_plan(self.num_D - 1 - i)
            sink: List[torch.Tensor] = []
            params = [p.detach() for p in plan.params] if freeze else plan.params
            last = _DiscriminatorFn.apply(plan, sink, cur, *params)
            if self.getIntermFeat:
                result.append([a.permute(0, 3, 1, 2) for a in sink] + [last])
            else:
                result.append([last])
            if i != self.num_D - 1:                # tf.center_crop(input, int(s/2)), Model/HdGan.py:251 (torchvision's rounding rule)
                c = int(s / 2)
                top, left = int(round((cur.size(2) - c) / 2.0)), int(round((cur.size(3) - c) / 2.0))
                cur = cur[:, :, top:top + c, left:left + c].contiguous()
        return result


class GANLoss(nn.Module):
    """Model/HdGan.py:258-293 (LSGAN): avg-pool the last map, MSE against a constant, scale weights [1.8, 0.2]."""

    def __init__(self, use_lsgan=True, target_real_label=1.0, target_fake_label=0.0, tensor=None):
        super().__init__()
        if not use_lsgan:
            raise NotImplementedError("only the LSGAN form is on the hot path")
        self.real, self.fake = float(target_real_label), float(target_fake_label)

    def __call__(self, input, target_is_real):
        tgt = self.real if target_is_real else self.fake
        if isinstance(input[0], list):
            w = [1.8, 0.2]
            loss = 0
            for i, scale in enumerate(input):
                loss = loss + mse_const(plane_mean(scale[-1]), tgt) * w[i]
            return loss
        return mse_const(plane_mean(input[-1]), tgt)


# ----------------------------------------------------------------------------------------------------------------------
# Registration network + warp
# ----------------------------------------------------------------------------------------------------------------------


class _RegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, a, b, *params):
        need = any(ctx.needs_input_grad)
        out, saved = E.reg_forward(plan, a, b, save=need)
        ctx.plan, ctx.saved = plan, saved
        ctx.in_channels = (a.shape[1], b.shape[1])
        ctx.leaves = params if need else None
        return out

    @staticmethod
    def backward(ctx, dflow):
        da, db, grads = E.reg_backward(ctx.plan, ctx.saved, dflow.contiguous(), ctx.needs_input_grad[1], ctx.needs_input_grad[2],
                                       ctx.in_channels)
        ctx.saved = None
        if E.deferred() is not None and any(ctx.needs_input_grad[3:]):
            grads = E.deferred().take(ctx.leaves, grads, ctx.needs_input_grad[3:])
        elif any(ctx.needs_input_grad[3:]):
            grads = E.strip_in_place(ctx.leaves, grads)
        ctx.leaves = None
        return (None, da, db, *grads)


# ---- single-layer autograd functions on internal NHWC tensors (the standalone forwards of trainer/layers.py's modules; the networks
# themselves run the fused schedules of ctagan.engine) ----


class _ToNHWCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.nchw_to_nhwc(x, E.get_precision())

    @staticmethod
    def backward(ctx, g):
        return ops.nhwc_to_nchw(g.contiguous())


class _ToNCHWFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.dtype = x.dtype
        return ops.nhwc_to_nchw(x)

    @staticmethod
    def backward(ctx, g):
        return ops.nchw_to_nhwc(g.contiguous(), ctx.dtype)


def _to_nhwc(x):
    ops.ensure_device()
    return _ToNHWCFn.apply(x)


def _to_nchw(x):
    return _ToNCHWFn.apply(x)


_ACT_CODE = {None: E.L.ACT_NONE, "relu": E.L.ACT_RELU, "leaky_relu": E.L.ACT_LRELU, "tanh": E.L.ACT_TANH}


class _ConvActFn(torch.autograd.Function):
    """act(conv(x) + b) [or act(IN(conv(x))) with norm=True] for one Conv2d(stride, zero padding) on NHWC tensors."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad, act, norm):
        prim = E.ConvPrim(w.detach(), None if b is None else b.detach(), stride, pad)
        if norm:
            if act == E.L.ACT_TANH:
                raise NotImplementedError("InstanceNorm followed by tanh")
            pool = ops.ZeroPool(x.shape[0] * w.shape[0] + 8, x.device) if x.dtype == torch.bfloat16 else None
            r, st = prim.fprop_stats(x, pool)                 # the bias in front of a non-affine InstanceNorm is dead
            y = ops.norm_act_pad(r, st, act, 0)
            ctx.saved = (x, r, st)
        else:
            y = prim.fprop(x, act=act, use_bias=b is not None)
            ctx.saved = (x, y, None)
        ctx.prim, ctx.act, ctx.norm, ctx.has_bias = prim, act, norm, b is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, r, st = ctx.saved
        prim = ctx.prim
        g = g.contiguous()
        if ctx.norm:
            dy = ops.norm_act_pad_bwd(g, r, st, ctx.act, 0)
        else:
            dy = ops.act_bwd(g, r, ctx.act) if ctx.act != E.L.ACT_NONE else g
        dw = db = dx = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dw, db = prim.wgrad(dy, x, want_bias=ctx.has_bias and not ctx.norm)
        if ctx.needs_input_grad[0]:
            dx = prim.bprop(dy, (x.shape[1], x.shape[2]))
        return dx, dw, (db if ctx.has_bias and not ctx.norm else None), None, None, None, None


class _ResBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        rb = E._ResBlock(w1.detach(), b1, w2.detach(), b2)
        out, saved = rb.forward(x, save=True)
        ctx.rb, ctx.saved = rb, saved
        return out

    @staticmethod
    def backward(ctx, g):
        gx, (dw1, _, dw2, _) = ctx.rb.backward(ctx.saved, g.contiguous())
        ctx.saved = None
        return gx, dw1, None, dw2, None


def _res_block(x, w1, b1, w2, b2):
    if x.shape[1] < 2 or x.shape[2] < 2:
        raise ValueError("ReflectionPad2d(1) needs maps of at least 2x2")
    return _ResBlockFn.apply(x, w1, b1, w2, b2)


class _MaxPool2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.x = x
        return ops.maxpool2_fwd(x)

    @staticmethod
    def backward(ctx, g):
        return ops.maxpool2_bwd(g.contiguous(), ctx.x)


def get_init_function(activation, init_function, **kwargs):
    """trainer/layers.py:23-53 (same torch.nn.init calls, so the same random draws)."""
    from functools import partial
    a = 0.0
    if activation == "leaky_relu":
        a = 0.2 if "negative_slope" not in kwargs else kwargs["negative_slope"]
    gain = 0.02 if "gain" not in kwargs else kwargs["gain"]
    if isinstance(init_function, str):
        if init_function == "kaiming":
            activation = "relu" if activation is None else activation
            return partial(nn.init.kaiming_normal_, a=a, nonlinearity=activation, mode="fan_in")
        if init_function == "dirac":
            return nn.init.dirac_
        if init_function == "xavier":
            activation = "relu" if activation is None else activation
            return partial(nn.init.xavier_normal_, gain=nn.init.calculate_gain(nonlinearity=activation, param=a))
        if init_function == "normal":
            return partial(nn.init.normal_, mean=0.0, std=gain)
        if init_function == "orthogonal":
            return partial(nn.init.orthogonal_, gain=gain)
        if init_function == "zeros":
            return partial(nn.init.normal_, mean=0.0, std=1e-5)
        raise ValueError(f"unknown init function {init_function!r}")
    if init_function is None:
        if activation in ("relu", "leaky_relu"):
            return partial(nn.init.kaiming_normal_, a=a, nonlinearity=activation)
        if activation in ("tanh", "sigmoid"):
            return partial(nn.init.xavier_normal_, gain=nn.init.calculate_gain(nonlinearity=activation, param=a))
        raise ValueError("init_func=None needs an activation")
    return init_function


class Conv(nn.Module):
    """trainer/layers.py:71-104: Conv2d -> InstanceNorm (optional) -> activation -> ResnetTransformer (optional); same constructor,
    sub-module names (`conv2d`, `resnet_block`) and initialisation draws as the reference."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, bias=True, activation="relu", init_func="kaiming",
                 use_norm=False, use_resnet=False, **kwargs):
        super().__init__()
        if activation not in _ACT_CODE:
            raise NotImplementedError(f"activation {activation!r} is not on the hot path (relu / leaky_relu / tanh / None)")
        if kwargs.get("negative_slope", 0.2) != 0.2:
            raise NotImplementedError("LeakyReLU slopes other than 0.2")
        self.conv2d = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=bias)
        self.resnet_block = ResnetTransformer(out_channels, 1, init_func) if use_resnet else None
        self.use_norm, self.activation_name = bool(use_norm), activation
        get_init_function(activation, init_func)(self.conv2d.weight)
        if self.conv2d.bias is not None:
            self.conv2d.bias.data.zero_()

    def _forward_nhwc(self, x):
        c = self.conv2d
        x = _ConvActFn.apply(x, c.weight, c.bias, c.stride[0], c.padding[0], _ACT_CODE[self.activation_name], self.use_norm)
        if self.resnet_block is not None:
            x = self.resnet_block._forward_nhwc(x)
        return x

    def forward(self, x):
        return _to_nchw(self._forward_nhwc(_to_nhwc(x)))


class ResnetBlock(nn.Module):
    """trainer/layers.py:243-300 (reflect padding, InstanceNorm, no dropout: the only configuration the reference builds).
    conv_block indices 1 and 5 are the convolutions."""

    def __init__(self, dim, padding_type="reflect", norm_layer=None, use_dropout=False, use_bias=True):
        super().__init__()
        if padding_type != "reflect" or use_dropout:
            raise NotImplementedError("ResnetBlock: only padding_type='reflect' without dropout (trainer/layers.py:221-222)")
        self.conv_block = nn.Sequential(_Slot("reflect1"), nn.Conv2d(dim, dim, 3, bias=use_bias), _Slot("instnorm"), _Slot("relu"),
                                        _Slot("reflect1"), nn.Conv2d(dim, dim, 3, bias=use_bias), _Slot("instnorm"))

    def _forward_nhwc(self, x):
        c1, c2 = self.conv_block[1], self.conv_block[5]
        return _res_block(x, c1.weight, c1.bias, c2.weight, c2.bias)

    def forward(self, x):
        return _to_nchw(self._forward_nhwc(_to_nhwc(x)))


class ResnetTransformer(nn.Module):
    """trainer/layers.py:216-240 (all convs are created first, then re-drawn with init_func('relu') in traversal order)."""

    def __init__(self, dim, n_blocks, init_func="kaiming"):
        super().__init__()
        self.model = nn.Sequential(*[ResnetBlock(dim, "reflect", None, False, True) for _ in range(n_blocks)])
        init_ = get_init_function("relu", init_func)
        for m in self.model.modules():
            if type(m) == nn.Conv2d:
                init_(m.weight)
                if m.bias is not None:
                    m.bias.data.zero_()

    def _forward_nhwc(self, x):
        for blk in self.model:
            x = blk._forward_nhwc(x)
        return x

    def forward(self, x):
        return _to_nchw(self._forward_nhwc(_to_nhwc(x)))


class DownBlock(nn.Module):
    """trainer/layers.py:156-183: conv_0 (+ conv_1 with refine) -> MaxPool2d(pool_size=2); returns (pooled, skip) with skip=True."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, bias=False, activation="relu", init_func="kaiming",
                 use_norm=False, use_resnet=False, skip=True, refine=False, pool=True, pool_size=2, **kwargs):
        super().__init__()
        if pool and pool_size != 2:
            raise NotImplementedError("DownBlock: MaxPool2d(2) only")
        kwargs.pop("callback", None)
        self.conv_0 = Conv(in_channels, out_channels, kernel_size, stride, padding, bias=bias, activation=activation, init_func=init_func,
                           use_norm=use_norm, use_resnet=use_resnet, **kwargs)
        self.conv_1 = None
        if refine:
            self.conv_1 = Conv(out_channels, out_channels, kernel_size, stride, padding, bias=bias, activation=activation,
                               init_func=init_func, use_norm=use_norm, use_resnet=use_resnet, **kwargs)
        self.skip, self.pool = skip, bool(pool)

    def forward(self, x):
        x = skip = self.conv_0._forward_nhwc(_to_nhwc(x))
        if self.conv_1 is not None:
            x = skip = self.conv_1._forward_nhwc(x)
        if self.pool:
            if x.shape[1] % 2 or x.shape[2] % 2:
                raise ValueError("DownBlock: MaxPool2d(2) needs even H and W here")
            x = _MaxPool2Fn.apply(x)
        return (_to_nchw(x), _to_nchw(skip)) if self.skip else _to_nchw(x)


class ResUnet(nn.Module):
    """trainer/reg.py:31-75 for cfg 'A' (the only one defined)."""

    def __init__(self, nc_a, nc_b, cfg="A", init_func="kaiming", init_to_identity=True):
        super().__init__()
        in_nf = nc_a + nc_b
        skip = {}
        for n, out_nf in enumerate(E.REG_NDF, start=1):
            setattr(self, f"down_{n}", DownBlock(in_nf, out_nf, 3, 1, 1, bias=True, activation="leaky_relu", init_func=init_func,
                                                 use_norm=False, use_resnet=True, skip=True, refine=False, pool=True))
            skip[n] = out_nf
            in_nf = out_nf
        self.c1 = Conv(in_nf, 2 * in_nf, 1, 1, 0, bias=True, activation="leaky_relu", init_func=init_func, use_norm=False, use_resnet=False)
        self.t = ResnetTransformer(2 * in_nf, 3, init_func)
        self.c2 = Conv(2 * in_nf, in_nf, 1, 1, 0, bias=True, activation="leaky_relu", init_func=init_func, use_norm=False, use_resnet=False)
        n = len(E.REG_NDF)
        for out_nf in E.REG_NUF:
            setattr(self, f"up_{n}", Conv(in_nf + skip[n], out_nf, 3, 1, 1, bias=True, activation="leaky_relu", init_func=init_func,
                                          use_norm=False, use_resnet=False))
            in_nf = out_nf
            n -= 1
        self.refine = nn.Sequential(ResnetTransformer(in_nf, 1, init_func),
                                    Conv(in_nf, in_nf, 1, 1, 0, bias=True, activation="leaky_relu", init_func=init_func, use_norm=False,
                                         use_resnet=False))
        self.output = Conv(in_nf, 2, 3, 1, 1, bias=True, activation=None, init_func="zeros" if init_to_identity else init_func,
                           use_norm=False, use_resnet=False)


class Reg(nn.Module):
    """trainer/reg.py:101-132.  forward(img_a, img_b) -> (B, 2, H, W) flow in pixels (ch0 rows, ch1 cols)."""

    def __init__(self, height, width, in_channels_a, in_channels_b):
        super().__init__()
        if height % 128 or width % 128 or height < 256 or width < 256:
            raise ValueError("Reg needs H, W >= 256 and multiples of 128 (7 poolings; trainer/reg.py:15)")
        self.oh, self.ow = height, width
        self.in_channels_a, self.in_channels_b = in_channels_a, in_channels_b
        self.offset_map = ResUnet(in_channels_a, in_channels_b)
        self._plan = None

    def _apply(self, fn, *a, **k):
        self._plan = None
        return super()._apply(fn, *a, **k)

    def _get_plan(self):
        params = list(self.parameters())
        if self._plan is None or any(a is not b for a, b in zip(self._plan.params, params)):
            self._plan = E.RegPlan(params)
        return self._plan

    def prepack(self, force=False):
        E.prepack_prims(self._get_plan().prims(), E.get_precision(), force)

    def forward(self, img_a, img_b, apply_on=None):
        plan = self._get_plan()
        return _RegFn.apply(plan, img_a, img_b, *plan.params)


class _WarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, flow):
        src, flow = src.contiguous().float(), flow.contiguous().float()
        ctx.save_for_backward(src, flow)
        return ops.warp_fwd(src, flow)

    @staticmethod
    def backward(ctx, g):
        src, flow = ctx.saved_tensors
        gs, gf = ops.warp_bwd(g.contiguous(), src, flow, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return gs, gf


class Transformer_2D(nn.Module):
    """trainer/transformer.py:7-31: one fused bilinear warp kernel (no materialised grid, no per-call H2D copy)."""

    def forward(self, src, flow):
        if src.shape[0] != flow.shape[0] or src.shape[2:] != flow.shape[2:] or flow.shape[1] != 2:
            raise ValueError("Transformer_2D: src (B,C,H,W) and flow (B,2,H,W) must agree")
        ops.ensure_device()
        return _WarpFn.apply(src, flow)


# ----------------------------------------------------------------------------------------------------------------------
# Losses
# ----------------------------------------------------------------------------------------------------------------------


class _L1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous().float(), b.contiguous().float()
        ctx.save_for_backward(a, b)
        return ops.l1_fwd(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ga = ops.l1_bwd(a, b, g.contiguous()) if ctx.needs_input_grad[0] else None
        gb = ops.l1_bwd(b, a, g.contiguous()) if ctx.needs_input_grad[1] else None
        return ga, gb


class _MseConstFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, target):
        p = p.contiguous().float()
        ctx.save_for_backward(p)
        ctx.target = target
        return ops.mse_const_fwd(p, target)

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        return ops.mse_const_bwd(p, ctx.target, g.contiguous()), None


class _SmoothFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow):
        flow = flow.contiguous().float()
        ctx.save_for_backward(flow)
        return ops.smooth_fwd(flow)

    @staticmethod
    def backward(ctx, g):
        (flow,) = ctx.saved_tensors
        return ops.smooth_bwd(flow, g.contiguous())


class _MaskedL1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, warped, b1, b2):
        warped, b1, b2 = warped.contiguous().float(), b1.contiguous().float(), b2.contiguous().float()
        ctx.save_for_backward(warped, b1, b2)
        return ops.masked_l1_fwd(warped, b1, b2)

    @staticmethod
    def backward(ctx, g):
        warped, b1, b2 = ctx.saved_tensors
        return ops.masked_l1_bwd(warped, b1, b2, g.contiguous()), None, None


def l1_loss(a, b):
    ops.ensure_device()
    return _L1Fn.apply(a, b)


def mse_const(p, target):
    """mean((p - target)^2) against a constant: a Python number, or a 1-element tensor whose VALUE is read on the device at kernel time
    (never cached by object identity; CPU tensors are read on the host)."""
    ops.ensure_device()
    if torch.is_tensor(target):
        if target.numel() != 1:
            raise NotImplementedError("MSELoss here takes a broadcast constant target (CycTrainer.py:83-84)")
        if not target.is_cuda:
            target = float(target)
        else:
            target = target.detach().reshape(1).float()
    else:
        target = float(target)
    return _MseConstFn.apply(p, target)


def smooothing_loss(y_pred):
    """trainer/utils.py:165-173 (the reference's spelling is kept: it is the public name)."""
    ops.ensure_device()
    return _SmoothFn.apply(y_pred)


def masked_l1_loss(warped, real_b1, real_b2):
    """The masked L1 of trainer/HdTrainer.py:726-735 as one fused pass (no in-place edits of the inputs)."""
    ops.ensure_device()
    return _MaskedL1Fn.apply(warped, real_b1, real_b2)


class L1Loss(nn.Module):
    def forward(self, a, b):
        return l1_loss(a, b)


class MSELoss(nn.Module):
    """torch.nn.MSELoss for the only way the trainers use it: prediction vs a constant (1,1) target tensor (or a number)."""

    def forward(self, pred, target):
        return mse_const(pred, target)
