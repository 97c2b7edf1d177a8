"""CPU restatement of the CTA-GAN hot path -- TEST INFRASTRUCTURE ONLY.

This file is the parity oracle.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
path (``cta-gan_b200/``, ``Model/``, ``trainer/``) never does and hard-fails when the
CUDA library is missing.

Everything here is a *functional* restatement (plain ``torch.nn.functional`` calls on a
``state_dict``-shaped mapping, CPU, fp32 or fp64) of the reference's nn.Modules.  Each
function cites the reference file:line it follows (paths relative to the upstream tree).
The arithmetic itself lives in a third-party dependency of the reference (PyTorch; the
reference pins no version -- the oracle runs on the torch in this image, 2.11.0).

Pinning: the reference ships no tests, golden vectors or fixtures ("parity unpinned" by the
reference's own tests).  ``oracle/make_golden.py`` therefore imports the real reference
modules from ``/root/reference`` in the build container, checks every function below against
them bit-for-bit (same weights, same inputs) and freezes the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` re-checks this restatement against those frozen reference
outputs everywhere (no reference tree needed).
"""
from __future__ import annotations

import math
import random
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]

# --------------------------------------------------------------------------------------
# parameter construction (creation order + default initialisers == the reference's)
# --------------------------------------------------------------------------------------


def _torch_conv_init(cout: int, cin: int, k: int, transposed: bool = False) -> Tuple[Tensor, Tensor]:
    """Default nn.Conv2d / nn.ConvTranspose2d initialisation, drawn in torch's own order
    (weight first, then bias) so that ``torch.manual_seed`` reproduces the reference's
    random-init weights (train.py:22-28,49)."""
    if transposed:
        m = torch.nn.ConvTranspose2d(cin, cout, k, stride=2, padding=1, output_padding=1)
    else:
        m = torch.nn.Conv2d(cin, cout, k)
    return m.weight.detach().clone(), m.bias.detach().clone()


def init_generator(input_nc: int = 1, output_nc: int = 1, n_blocks: int = 9) -> State:
    """Parameters of Generator in creation order.  Model/CycleGan.py:24-64 (dup HdGan.py:65-113)."""
    sd: State = OrderedDict()

    def put(name, wb):
        sd[name + ".weight"], sd[name + ".bias"] = wb

    put("model_head.1", _torch_conv_init(64, input_nc, 7))          # :27-28
    put("model_head.4", _torch_conv_init(128, 64, 3))               # :36 (1st down)
    put("model_head.7", _torch_conv_init(256, 128, 3))              # :36 (2nd down)
    for b in range(n_blocks):                                        # :44-45, block ctor :10-16
        put(f"model_body.{b}.conv_block.1", _torch_conv_init(256, 256, 3))
        put(f"model_body.{b}.conv_block.5", _torch_conv_init(256, 256, 3))
    put("model_tail.0", _torch_conv_init(128, 256, 3, transposed=True))   # :51
    put("model_tail.3", _torch_conv_init(64, 128, 3, transposed=True))    # :51
    put("model_tail.7", _torch_conv_init(output_nc, 64, 7))               # :58-59
    return sd


def init_discriminator(input_nc: int = 1, key_fmt: str = "model.{i}") -> State:
    """Discriminator parameters.  Model/CycleGan.py:78-94; the Discriminator_m variant
    (Model/HdGan.py:156-180,216-220) has the same tensors under ``scale0_layer{j}.0``."""
    sd: State = OrderedDict()
    chans = [(input_nc, 64), (64, 128), (128, 256), (256, 512), (512, 1)]
    idx = [0, 2, 5, 8, 11]
    for j, ((ci, co), i) in enumerate(zip(chans, idx)):
        w, b = _torch_conv_init(co, ci, 4)
        name = key_fmt.format(i=i, j=j)
        sd[name + ".weight"], sd[name + ".bias"] = w, b
    return sd


def init_discriminator_m(input_nc: int = 1, num_D: int = 1) -> State:
    """Discriminator_m parameters, Model/HdGan.py:215-221: one NLayerDiscriminator per scale, created in scale order."""
    sd: State = OrderedDict()
    for i in range(num_D):
        sd.update(init_discriminator(input_nc, key_fmt="scale%d_layer{j}.0" % i))
    return sd


def _kaiming(cout, cin, k, act):
    """trainer/layers.py:23-33 -- kaiming_normal_(a, nonlinearity, fan_in); bias zero (:91-92,229-230)."""
    m = torch.nn.Conv2d(cin, cout, k)            # consumes RNG exactly like the reference ctor
    a = 0.2 if act == "leaky_relu" else 0.0
    nonlin = "relu" if act is None else act
    torch.nn.init.kaiming_normal_(m.weight, a=a, nonlinearity=nonlin, mode="fan_in")
    return m.weight.detach().clone(), torch.zeros(cout)


def _init_resnet_transformer(sd: State, prefix: str, dim: int, n: int):
    """trainer/layers.py:216-237: convs are created (default init) for all blocks first,
    then ``apply(init_weights)`` re-draws them in module traversal order."""
    mods = []
    for i in range(n):
        c1 = torch.nn.Conv2d(dim, dim, 3)
        c2 = torch.nn.Conv2d(dim, dim, 3)
        mods.append((i, c1, c2))
    for i, c1, c2 in mods:
        for j, c in ((1, c1), (5, c2)):
            torch.nn.init.kaiming_normal_(c.weight, a=0.0, nonlinearity="relu", mode="fan_in")
            sd[f"{prefix}.model.{i}.conv_block.{j}.weight"] = c.weight.detach().clone()
            sd[f"{prefix}.model.{i}.conv_block.{j}.bias"] = torch.zeros(dim)


REG_NDF = [32, 64, 64, 64, 64, 64, 64]   # trainer/reg.py:15
REG_NUF = [64, 64, 64, 64, 64, 64, 32]   # trainer/reg.py:18


def init_reg(in_a: int = 1, in_b: int = 1) -> State:
    """Reg / ResUnet parameters in creation order.  trainer/reg.py:32-75, layers.py:80-94."""
    sd: State = OrderedDict()
    p = "offset_map."
    in_nf = in_a + in_b
    skip = {}
    for n, out_nf in enumerate(REG_NDF, start=1):                     # reg.py:42-48
        w, b = None, None
        # Conv ctor order (layers.py:83-91): conv2d created, resnet_block created+initialised, then conv2d re-initialised
        conv = torch.nn.Conv2d(in_nf, out_nf, 3)
        tmp: State = OrderedDict()
        _init_resnet_transformer(tmp, f"{p}down_{n}.conv_0.resnet_block", out_nf, 1)
        torch.nn.init.kaiming_normal_(conv.weight, a=0.2, nonlinearity="leaky_relu", mode="fan_in")
        sd[f"{p}down_{n}.conv_0.conv2d.weight"] = conv.weight.detach().clone()
        sd[f"{p}down_{n}.conv_0.conv2d.bias"] = torch.zeros(out_nf)
        sd.update(tmp)
        skip[n] = out_nf
        in_nf = out_nf
    w, b = _kaiming(2 * in_nf, in_nf, 1, "leaky_relu")                # reg.py:51 c1
    sd[p + "c1.conv2d.weight"], sd[p + "c1.conv2d.bias"] = w, b
    _init_resnet_transformer(sd, p + "t", 2 * in_nf, 3)               # reg.py:53-54
    w, b = _kaiming(in_nf, 2 * in_nf, 1, "leaky_relu")                # reg.py:55 c2
    sd[p + "c2.conv2d.weight"], sd[p + "c2.conv2d.bias"] = w, b
    n = len(REG_NDF)
    for out_nf in REG_NUF:                                            # reg.py:59-64
        w, b = _kaiming(out_nf, in_nf + skip[n], 3, "leaky_relu")
        sd[f"{p}up_{n}.conv2d.weight"], sd[f"{p}up_{n}.conv2d.bias"] = w, b
        in_nf = out_nf
        n -= 1
    _init_resnet_transformer(sd, p + "refine.0", in_nf, 1)            # reg.py:66
    w, b = _kaiming(in_nf, in_nf, 1, "leaky_relu")                    # reg.py:67-69
    sd[p + "refine.1.conv2d.weight"], sd[p + "refine.1.conv2d.bias"] = w, b
    conv = torch.nn.Conv2d(in_nf, 2, 3)                               # reg.py:73-75 -> 'zeros' == normal(0, 1e-5), layers.py:44-45
    torch.nn.init.normal_(conv.weight, mean=0.0, std=1e-5)
    sd[p + "output.conv2d.weight"] = conv.weight.detach().clone()
    sd[p + "output.conv2d.bias"] = torch.zeros(2)
    return sd


# --------------------------------------------------------------------------------------
# forward passes
# --------------------------------------------------------------------------------------


def _inorm(x: Tensor) -> Tensor:
    """nn.InstanceNorm2d defaults: affine=False, no running stats, eps=1e-5, biased variance."""
    return F.instance_norm(x, eps=1e-5)


def _rpad(x: Tensor, p: int) -> Tensor:
    return F.pad(x, (p, p, p, p), mode="reflect")


def generator_forward(sd: State, x: Tensor, n_blocks: int = 9) -> Tensor:
    """Generator.forward, Model/CycleGan.py:66-71 (head :27-40, body :43-45 + :20-21, tail :48-60)."""
    w = lambda k: sd[k + ".weight"]
    b = lambda k: sd[k + ".bias"]
    x = F.relu(_inorm(F.conv2d(_rpad(x, 3), w("model_head.1"), b("model_head.1"))))
    x = F.relu(_inorm(F.conv2d(x, w("model_head.4"), b("model_head.4"), stride=2, padding=1)))
    x = F.relu(_inorm(F.conv2d(x, w("model_head.7"), b("model_head.7"), stride=2, padding=1)))
    for i in range(n_blocks):
        k1, k2 = f"model_body.{i}.conv_block.1", f"model_body.{i}.conv_block.5"
        t = F.relu(_inorm(F.conv2d(_rpad(x, 1), w(k1), b(k1))))
        x = x + _inorm(F.conv2d(_rpad(t, 1), w(k2), b(k2)))
    for k in ("model_tail.0", "model_tail.3"):
        x = F.relu(_inorm(F.conv_transpose2d(x, w(k), b(k), stride=2, padding=1, output_padding=1)))
    return torch.tanh(F.conv2d(_rpad(x, 3), w("model_tail.7"), b("model_tail.7")))


def discriminator_features(sd: State, x: Tensor, key_fmt: str = "model.{i}") -> List[Tensor]:
    """The 5-conv PatchGAN stack; returns every layer's (post-activation) output.
    Model/CycleGan.py:78-94 == Model/HdGan.py:156-180."""
    idx = [0, 2, 5, 8, 11]
    strides = [2, 2, 2, 1, 1]
    feats = []
    for j, (i, s) in enumerate(zip(idx, strides)):
        name = key_fmt.format(i=i, j=j)
        x = F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=s, padding=1)
        if 1 <= j <= 3:
            x = _inorm(x)
        if j <= 3:
            x = F.leaky_relu(x, 0.2)
        feats.append(x)
    return feats


def discriminator_forward(sd: State, x: Tensor) -> Tensor:
    """Discriminator.forward, Model/CycleGan.py:98-103: global average pool -> (B, 1)."""
    y = discriminator_features(sd, x)[-1]
    return F.avg_pool2d(y, y.shape[2:]).view(y.shape[0], -1)


def discriminator_m_forward(sd: State, x: Tensor, num_D: int = 1) -> List[List[Tensor]]:
    """Discriminator_m.forward with getIntermFeat=True, Model/HdGan.py:236-256: pass i runs scale (num_D-1-i) on the input centre-cropped
    i times to half its size (torchvision center_crop, :251)."""
    result = []
    for i in range(num_D):
        s = x.shape[2]
        result.append(discriminator_features(sd, x, key_fmt="scale%d_layer{j}.0" % (num_D - 1 - i)))
        if i != num_D - 1:
            c = int(s / 2)
            top, left = int(round((x.shape[2] - c) / 2.0)), int(round((x.shape[3] - c) / 2.0))      # torchvision.transforms.functional.center_crop
            x = x[:, :, top:top + c, left:left + c]
    return result


def gan_loss(pred, target_is_real: bool) -> Tensor:
    """GANLoss.__call__, Model/HdGan.py:269-293 (LSGAN; scale weights [1.8, 0.2])."""
    tgt = 1.0 if target_is_real else 0.0
    if isinstance(pred[0], list):
        wts = [1.8, 0.2]
        loss = 0
        for i, scale in enumerate(pred):
            y = scale[-1]
            p = F.avg_pool2d(y, y.shape[2:]).view(y.shape[0], -1)
            loss = loss + mse_vs_const(p, tgt) * wts[i]
        return loss
    y = pred[-1]
    p = F.avg_pool2d(y, y.shape[2:]).view(y.shape[0], -1)
    return mse_vs_const(p, tgt)


def mse_vs_const(pred: Tensor, target: float) -> Tensor:
    """torch.nn.MSELoss()(pred(B,1), const(1,1)) -- broadcast target, mean over B. CycTrainer.py:76,83-84,146."""
    return ((pred - target) ** 2).mean()


def l1_loss(a: Tensor, b: Tensor) -> Tensor:
    """torch.nn.L1Loss() (mean).  CycTrainer.py:77,154."""
    return (a - b).abs().mean()


def _reg_conv(sd, name, x, k, act=True):
    """trainer/layers.py:96-104 Conv.forward with use_norm=False: conv -> LeakyReLU(0.2)."""
    x = F.conv2d(x, sd[name + ".conv2d.weight"], sd[name + ".conv2d.bias"], stride=1, padding=(k - 1) // 2)
    return F.leaky_relu(x, 0.2) if act else x


def _reg_resblocks(sd, prefix, x, n):
    """trainer/layers.py:239-240,297-300 (block body :272-293)."""
    for i in range(n):
        k1, k2 = f"{prefix}.model.{i}.conv_block.1", f"{prefix}.model.{i}.conv_block.5"
        t = F.relu(_inorm(F.conv2d(_rpad(x, 1), sd[k1 + ".weight"], sd[k1 + ".bias"])))
        x = x + _inorm(F.conv2d(_rpad(t, 1), sd[k2 + ".weight"], sd[k2 + ".bias"]))
    return x


def reg_forward(sd: State, img_a: Tensor, img_b: Tensor) -> Tensor:
    """Reg.forward -> ResUnet.forward, trainer/reg.py:76-99,128-132.  Returns the (B,2,H,W) flow in pixels."""
    p = "offset_map."
    x = torch.cat([img_a, img_b], 1)                                  # :77
    skips = {}
    nd = len(REG_NDF)
    for n in range(1, nd + 1):                                        # :81-84, DownBlock layers.py:174-183
        x = _reg_conv(sd, f"{p}down_{n}.conv_0", x, 3)
        x = _reg_resblocks(sd, f"{p}down_{n}.conv_0.resnet_block", x, 1)
        skips[n] = x
        x = F.max_pool2d(x, 2)
    x = _reg_conv(sd, p + "c1", x, 1)                                 # :85-88
    x = _reg_resblocks(sd, p + "t", x, 3)
    x = _reg_conv(sd, p + "c2", x, 1)
    for n in range(nd, 0, -1):                                        # :90-96
        s = skips[n]
        x = F.interpolate(x, (s.size(2), s.size(3)), mode="bilinear")
        x = torch.cat([x, s], 1)
        x = _reg_conv(sd, f"{p}up_{n}", x, 3)
    x = _reg_resblocks(sd, p + "refine.0", x, 1)                      # :97
    x = _reg_conv(sd, p + "refine.1", x, 1)
    return _reg_conv(sd, p + "output", x, 3, act=False)               # :98


def warp(src: Tensor, flow: Tensor) -> Tensor:
    """Transformer_2D.forward, trainer/transformer.py:12-29, without the hard `.cuda()` at :21."""
    b, _, h, w = flow.shape
    grids = torch.meshgrid([torch.arange(0, h, device=flow.device), torch.arange(0, w, device=flow.device)], indexing="ij")
    grid = torch.stack(grids).to(flow.dtype).repeat(b, 1, 1, 1)
    new_locs = grid + flow
    shape = (h, w)
    parts = []
    for i in range(2):
        parts.append(2 * (new_locs[:, i, ...] / (shape[i] - 1) - 0.5))
    new_locs = torch.stack(parts, dim=1).permute(0, 2, 3, 1)[..., [1, 0]]
    return F.grid_sample(src, new_locs, align_corners=True, padding_mode="border")


def smoothing_loss(flow: Tensor) -> Tensor:
    """smooothing_loss, trainer/utils.py:165-173."""
    dy = (flow[:, :, 1:, :] - flow[:, :, :-1, :]).abs()
    dx = (flow[:, :, :, 1:] - flow[:, :, :, :-1]).abs()
    return (dx * dx).mean() + (dy * dy).mean()


def masked_l1(warped: Tensor, real_b1: Tensor, real_b2: Tensor) -> Tensor:
    """The masked L1 block, trainer/HdTrainer.py:726-735 (without its in-place aliasing side effects)."""
    bb = (real_b1 >= 0.3).to(real_b2.dtype)
    rb = real_b2 * bb
    rb = torch.where(rb == 0, torch.full_like(rb, -1.0), rb)
    sw = warped * bb
    sw = torch.where(sw == 0, torch.full_like(sw, -1.0), sw)
    return (sw - rb).abs().mean()


# --------------------------------------------------------------------------------------
# iteration bodies (restated): nets are dicts of leaf tensors with requires_grad=True
# --------------------------------------------------------------------------------------


class ReplayBuffer:
    """trainer/utils.py:120-140 -- 50-slot history buffer driven by Python's `random`."""

    def __init__(self, max_size: int = 50):
        self.max_size = max_size
        self.data: List[Tensor] = []

    def push_and_pop(self, data: Tensor) -> Tensor:
        out = []
        for element in data.detach():
            element = element.unsqueeze(0)
            if len(self.data) < self.max_size:
                self.data.append(element)
                out.append(element)
            elif random.uniform(0, 1) > 0.5:
                i = random.randint(0, self.max_size - 1)
                out.append(self.data[i].clone())
                self.data[i] = element
            else:
                out.append(element)
        return torch.cat(out)


def leafify(sd: State, dtype=torch.float32) -> State:
    return OrderedDict((k, v.detach().to(dtype).clone().requires_grad_(True)) for k, v in sd.items())


def make_adam(sds, lr):
    params = [p for sd in sds for p in sd.values()]
    return torch.optim.Adam(params, lr=lr, betas=(0.5, 0.999))


class CycState:
    """Networks + optimisers of Cyc_Trainer.__init__, trainer/CycTrainer.py:64-88."""

    def __init__(self, lr=1e-4, dtype=torch.float32, n_blocks=9):
        self.n_blocks = n_blocks
        self.G_A2B = leafify(init_generator(n_blocks=n_blocks), dtype)
        self.D_B = leafify(init_discriminator(1), dtype)
        self.G_B2A = leafify(init_generator(n_blocks=n_blocks), dtype)
        self.D_A = leafify(init_discriminator(1), dtype)
        self.opt_D_B = make_adam([self.D_B], lr)
        self.opt_G = make_adam([self.G_A2B, self.G_B2A], lr)
        self.opt_D_A = make_adam([self.D_A], lr)
        self.buf_A, self.buf_B = ReplayBuffer(), ReplayBuffer()


def cyc_step(st: CycState, real_A: Tensor, real_B: Tensor, adv=1.0, cyc=10.0) -> Dict[str, float]:
    """One iteration of Cyc_Trainer.train, trainer/CycTrainer.py:138-197."""
    nb = st.n_blocks
    st.opt_G.zero_grad()
    fake_B = generator_forward(st.G_A2B, real_A, nb)
    loss_GAN_A2B = adv * mse_vs_const(discriminator_forward(st.D_B, fake_B), 1.0)
    fake_A = generator_forward(st.G_B2A, real_B, nb)
    loss_GAN_B2A = adv * mse_vs_const(discriminator_forward(st.D_A, fake_A), 1.0)
    rec_A = generator_forward(st.G_B2A, fake_B, nb)
    loss_cyc_ABA = cyc * l1_loss(rec_A, real_A)
    rec_B = generator_forward(st.G_A2B, fake_A, nb)
    loss_cyc_BAB = cyc * l1_loss(rec_B, real_B)
    loss_G = loss_GAN_A2B + loss_GAN_B2A + loss_cyc_ABA + loss_cyc_BAB
    loss_G.backward()
    st.opt_G.step()

    st.opt_D_A.zero_grad()
    fa = st.buf_A.push_and_pop(fake_A)
    loss_D_A = adv * mse_vs_const(discriminator_forward(st.D_A, real_A), 1.0) + \
        adv * mse_vs_const(discriminator_forward(st.D_A, fa.detach()), 0.0)
    loss_D_A.backward()
    st.opt_D_A.step()

    st.opt_D_B.zero_grad()
    fb = st.buf_B.push_and_pop(fake_B)
    loss_D_B = adv * mse_vs_const(discriminator_forward(st.D_B, real_B), 1.0) + \
        adv * mse_vs_const(discriminator_forward(st.D_B, fb.detach()), 0.0)
    loss_D_B.backward()
    st.opt_D_B.step()
    return {"loss_G": float(loss_G.detach()), "loss_GAN_A2B": float(loss_GAN_A2B), "loss_GAN_B2A": float(loss_GAN_B2A),
            "loss_cycle_ABA": float(loss_cyc_ABA), "loss_cycle_BAB": float(loss_cyc_BAB),
            "loss_D_A": float(loss_D_A), "loss_D_B": float(loss_D_B)}


class RegState:
    """Networks + optimisers of Reg_Trainer.__init__ (RegTrainer.py:94-101) / Hd_Trainer_x1/x2 (HdTrainer.py:99-105,610-616)."""

    def __init__(self, lr=1e-4, lrd=None, dtype=torch.float32, multiscale_d=False, n_blocks=9):
        self.n_blocks = n_blocks
        self.multiscale_d = multiscale_d
        self.G = leafify(init_generator(n_blocks=n_blocks), dtype)
        self.D = leafify(init_discriminator_m(1) if multiscale_d else init_discriminator(1), dtype)
        self.opt_D = make_adam([self.D], lr if lrd is None else lrd)
        self.R = leafify(init_reg(1, 1), dtype)
        self.opt_R = make_adam([self.R], lr)
        self.opt_G = make_adam([self.G], lr)


def reg_step(st: RegState, real_A: Tensor, real_B: Tensor, corr=20.0, adv=1.0, smooth=10.0) -> Dict[str, float]:
    """One iteration of Reg_Trainer.train (RegTrainer.py:170-198) == Hd_Trainer_x1.train (HdTrainer.py:192-228)."""
    nb = st.n_blocks
    st.opt_R.zero_grad()
    st.opt_G.zero_grad()
    fake_B = generator_forward(st.G, real_A, nb)
    trans = reg_forward(st.R, fake_B, real_B)
    sysreg = warp(fake_B, trans)
    sr = corr * l1_loss(sysreg, real_B)
    advl = adv * mse_vs_const(discriminator_forward(st.D, fake_B), 1.0)
    sm = smooth * smoothing_loss(trans)
    total = sm + advl + sr
    total.backward()
    st.opt_R.step()
    st.opt_G.step()

    st.opt_D.zero_grad()
    with torch.no_grad():
        fake_B = generator_forward(st.G, real_A, nb)
    loss_D = adv * mse_vs_const(discriminator_forward(st.D, fake_B), 0.0) + \
        adv * mse_vs_const(discriminator_forward(st.D, real_B), 1.0)
    loss_D.backward()
    st.opt_D.step()
    return {"SR_loss": float(sr), "adv_loss": float(advl), "SM_loss": float(sm), "toal_loss": float(total),
            "loss_D_B": float(loss_D)}


def hd_x2_step(st: RegState, real_A2: Tensor, real_B1: Tensor, real_B2: Tensor,
               corr1=20.0, corr2=2.0, adv1=1.0, smooth=10.0) -> Dict[str, float]:
    """One iteration of Hd_Trainer_x2.train, trainer/HdTrainer.py:705-751 (Discriminator_m + GANLoss + masked L1)."""
    nb = st.n_blocks
    real_BB2 = real_B2.clone()
    st.opt_R.zero_grad()
    st.opt_G.zero_grad()
    fake_B = generator_forward(st.G, real_A2, nb)
    trans = reg_forward(st.R, fake_B, real_B2)
    sysreg = warp(fake_B, trans)
    sm = smooth * smoothing_loss(trans)
    sr = corr1 * l1_loss(sysreg, real_B2)
    advl = adv1 * gan_loss(discriminator_m_forward(st.D, fake_B), True)
    sr2 = corr2 * masked_l1(sysreg, real_B1, real_B2)
    total = sm + advl + sr + sr2
    total.backward()
    st.opt_R.step()
    st.opt_G.step()

    st.opt_D.zero_grad()
    with torch.no_grad():
        fake_B = generator_forward(st.G, real_A2, nb)
    loss_D = adv1 * (gan_loss(discriminator_m_forward(st.D, fake_B), False) +
                     gan_loss(discriminator_m_forward(st.D, real_BB2), True)) / 2
    loss_D.backward()
    st.opt_D.step()
    return {"SR_loss": float(sr), "SR_loss2": float(sr2), "adv_loss": float(advl), "SM_loss": float(sm),
            "toal_loss": float(total), "loss_D_B": float(loss_D)}


class P2pState:
    """P2p_Trainer.__init__, trainer/p2pTrainer.py:60-63."""

    def __init__(self, lr=1e-4, dtype=torch.float32, n_blocks=9):
        self.n_blocks = n_blocks
        self.G = leafify(init_generator(n_blocks=n_blocks), dtype)
        self.D = leafify(init_discriminator(2), dtype)
        self.opt_D = make_adam([self.D], lr)
        self.opt_G = make_adam([self.G], lr)


def p2p_step(st: P2pState, real_A: Tensor, real_B: Tensor, adv=1.0, p2p=100.0) -> Dict[str, float]:
    """One iteration of P2p_Trainer.train, trainer/p2pTrainer.py:122-148 (note: the D step scales the *prediction*)."""
    nb = st.n_blocks
    st.opt_G.zero_grad()
    fake_B = generator_forward(st.G, real_A, nb)
    l1 = l1_loss(fake_B, real_B) * p2p
    gan = mse_vs_const(discriminator_forward(st.D, torch.cat((real_A, fake_B), 1)), 1.0) * adv
    total = l1 + gan
    total.backward()
    st.opt_G.step()

    st.opt_D.zero_grad()
    with torch.no_grad():
        fake_B = generator_forward(st.G, real_A, nb)
    pf = discriminator_forward(st.D, torch.cat((real_A, fake_B), 1)) * adv
    pr = discriminator_forward(st.D, torch.cat((real_A, real_B), 1)) * adv
    loss_D = mse_vs_const(pf, 0.0) + mse_vs_const(pr, 1.0)
    loss_D.backward()
    st.opt_D.step()
    return {"loss_L1": float(l1), "loss_GAN_A2B": float(gan), "toal_loss": float(total), "loss_D_B": float(loss_D)}


# --------------------------------------------------------------------------------------
# synthetic CT-like slices
# --------------------------------------------------------------------------------------


def synthetic_pair(batch: int, size: int, seed: int = 42, phantom: bool = False) -> Tuple[Tensor, Tensor]:
    """Slices in [-1, 1] as trainer/datasets.py:74-82 produces (air = -1)."""
    g = torch.Generator().manual_seed(seed)
    if not phantom:
        a = torch.rand(batch, 1, size, size, generator=g) * 2 - 1
        b = torch.rand(batch, 1, size, size, generator=g) * 2 - 1
        return a, b
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing="ij")
    disc = ((yy - size / 2) ** 2 + (xx - size / 2) ** 2) <= (0.4 * size) ** 2
    a = torch.full((batch, 1, size, size), -1.0)
    vals = torch.rand(batch, 1, size, size, generator=g) * 0.6 - 0.3
    a = torch.where(disc, vals, a)
    b = torch.roll(a, shifts=(2, -3), dims=(2, 3)) + 0.02 * torch.randn(batch, 1, size, size, generator=g)
    return a, b.clamp_(-1, 1)
