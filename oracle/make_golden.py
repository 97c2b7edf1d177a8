"""Freeze golden vectors from the REAL reference modules and pin oracle/restate.py against them.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
Writes tests/golden/golden_v1.pt (a dict of small fp32 tensors + metadata).  TEST INFRASTRUCTURE ONLY.

For every module on the hot path the script
  1. seeds torch (42, as train.py:49), builds the reference nn.Module and the restated state dict with the same seed,
     and asserts that keys, shapes and values agree exactly (pins creation order + initialisers);
  2. runs reference module and restatement on the same synthetic slices and asserts bit-equality (CPU fp32);
  3. stores inputs seeds, outputs and gradient fingerprints as golden vectors.
Iteration bodies (the trainers cannot be instantiated) are driven here with the reference's own
nn.Modules + torch.optim.Adam following the cited trainer lines, and compared with restate.*_step.
"""
from __future__ import annotations

import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import restate as R  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

torch.set_num_threads(8)
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "golden_v1.pt")


def seed(s=42):
    random.seed(s)
    torch.manual_seed(s)


def check_state(ref_module, sd, what):
    rsd = ref_module.state_dict()
    assert list(rsd.keys()) == list(sd.keys()), (what, [k for k in rsd if k not in sd][:5], [k for k in sd if k not in rsd][:5])
    for k in rsd:
        assert rsd[k].shape == sd[k].shape, (what, k)
        assert torch.equal(rsd[k], sd[k]), (what, k, (rsd[k] - sd[k]).abs().max())
    return {k: (tuple(v.shape), float(v.double().sum()), float(v.double().abs().sum())) for k, v in rsd.items()}


def grad_fingerprint(params: dict):
    return {k: (float(p.grad.double().norm()), float(p.grad.double().sum())) for k, p in params.items() if p.grad is not None}


def main():
    ref = load_reference()
    G = {}
    S = 64

    # ---------------- Generator ----------------
    seed(); g_ref = ref.CycleGan.Generator(1, 1)
    seed(); g_sd = R.init_generator(1, 1)
    G["generator.state_fp"] = check_state(g_ref, g_sd, "Generator")
    seed(); g_hd = ref.HdGan.Generator(1, 1)
    check_state(g_hd, g_sd, "HdGan.Generator")
    a, b = R.synthetic_pair(1, S, seed=42)
    y_ref = g_ref(a)
    leaf = R.leafify(g_sd)
    y = R.generator_forward(leaf, a)
    assert torch.equal(y_ref, y), (y_ref - y).abs().max()
    (y_ref * b).sum().backward()
    (y * b).sum().backward()
    for (k, p), (k2, p2) in zip(g_ref.named_parameters(), leaf.items()):
        assert k == k2 and torch.equal(p.grad, p2.grad), k
    G["generator.out_64"] = y_ref.detach().clone()
    G["generator.grad_fp_64"] = grad_fingerprint(leaf)
    a2, _ = R.synthetic_pair(2, 128, seed=7, phantom=True)
    G["generator.out_128_phantom_b2"] = g_ref(a2).detach().clone()
    assert torch.equal(G["generator.out_128_phantom_b2"], R.generator_forward(g_sd, a2))

    # ---------------- Discriminator ----------------
    for nc in (1, 2):
        seed(); d_ref = ref.CycleGan.Discriminator(nc)
        seed(); d_sd = R.init_discriminator(nc)
        G[f"discriminator{nc}.state_fp"] = check_state(d_ref, d_sd, "Discriminator")
        x = torch.cat([a, b], 1)[:, :nc]
        x2 = torch.cat([x, -x], 0)
        p_ref = d_ref(x2)
        leaf = R.leafify(d_sd)
        p = R.discriminator_forward(leaf, x2)
        assert torch.equal(p_ref, p)
        l_ref = torch.nn.MSELoss()(p_ref, torch.ones(1, 1)); l_ref.backward()
        l = R.mse_vs_const(p, 1.0); l.backward()
        assert torch.equal(l_ref, l)
        for (k, q), (k2, q2) in zip(d_ref.named_parameters(), leaf.items()):
            assert k == k2 and torch.allclose(q.grad, q2.grad, rtol=0, atol=0), k
        G[f"discriminator{nc}.pred_64_b2"] = p_ref.detach().clone()
        G[f"discriminator{nc}.mse_real"] = l_ref.detach().clone()
        G[f"discriminator{nc}.grad_fp"] = grad_fingerprint(leaf)

    # ---------------- Discriminator_m + GANLoss ----------------
    seed(); dm_ref = ref.HdGan.Discriminator_m(1)
    seed(); dm_sd = R.init_discriminator_m(1)
    G["discriminator_m.state_fp"] = check_state(dm_ref, dm_sd, "Discriminator_m")
    feats_ref = dm_ref(a)
    feats = R.discriminator_m_forward(dm_sd, a)
    assert len(feats_ref) == 1 and len(feats_ref[0]) == 5
    for fr, f in zip(feats_ref[0], feats[0]):
        assert torch.equal(fr, f)
    gl = ref.HdGan.GANLoss()
    for flag in (True, False):
        assert torch.equal(gl(feats_ref, flag), R.gan_loss(feats, flag))
        G[f"discriminator_m.ganloss_{flag}"] = gl(feats_ref, flag).detach().clone()
    G["discriminator_m.last_64"] = feats_ref[0][-1].detach().clone()
    G["discriminator_m.feat_shapes"] = [tuple(f.shape) for f in feats_ref[0]]

    # ---------------- Reg (needs >= 256) ----------------
    seed(); r_ref = ref.reg.Reg(256, 256, 1, 1)
    seed(); r_sd = R.init_reg(1, 1)
    G["reg.state_fp"] = check_state(r_ref, r_sd, "Reg")
    ra, rb = R.synthetic_pair(1, 256, seed=3, phantom=True)
    fl_ref = r_ref(ra, rb)
    leaf = R.leafify(r_sd)
    fl = R.reg_forward(leaf, ra, rb)
    assert torch.equal(fl_ref, fl), (fl_ref - fl).abs().max()
    G["reg.flow_256"] = fl_ref.detach().clone()
    # make the flow non-trivial for the gradient fingerprint: scale output weights up
    seed(1); wbig = torch.randn_like(r_sd["offset_map.output.conv2d.weight"]) * 0.05
    r_ref.offset_map.output.conv2d.weight.data.copy_(wbig)
    leaf["offset_map.output.conv2d.weight"].data.copy_(wbig)
    fl_ref = r_ref(ra, rb); fl = R.reg_forward(leaf, ra, rb)
    assert torch.equal(fl_ref, fl)
    G["reg.flow_256_bigw"] = fl_ref.detach().clone()
    ref.utils.smooothing_loss(fl_ref).backward()
    R.smoothing_loss(fl).backward()
    for (k, q), (k2, q2) in zip(r_ref.named_parameters(), leaf.items()):
        assert k == k2 and torch.equal(q.grad, q2.grad), k
    G["reg.grad_fp_bigw"] = grad_fingerprint(leaf)
    G["reg.smooth_bigw"] = ref.utils.smooothing_loss(fl_ref).detach().clone()

    # ---------------- warp (Transformer_2D) + smoothness ----------------
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self        # transformer.py:21 hard-codes .cuda()
    try:
        tr = ref.transformer.Transformer_2D()
        for name, (Bn, Hh, Ww, mag) in {"small": (2, 32, 48, 3.0), "sq": (1, 64, 64, 1.5), "tiny_flow": (1, 64, 64, 1e-3),
                                       "oob": (1, 40, 40, 30.0)}.items():
            gsrc = torch.Generator().manual_seed(11)
            src = (torch.rand(Bn, 1, Hh, Ww, generator=gsrc) * 2 - 1).requires_grad_(True)
            flow = ((torch.rand(Bn, 2, Hh, Ww, generator=gsrc) * 2 - 1) * mag).requires_grad_(True)
            wt = torch.rand(Bn, 1, Hh, Ww, generator=gsrc)
            o_ref = tr(src, flow)
            (o_ref * wt).sum().backward()
            gs_ref, gf_ref = src.grad.clone(), flow.grad.clone()
            src.grad = None; flow.grad = None
            o = R.warp(src, flow)
            (o * wt).sum().backward()
            assert torch.equal(o_ref, o) and torch.equal(gs_ref, src.grad) and torch.equal(gf_ref, flow.grad), name
            G[f"warp.{name}"] = {"src": src.detach().clone(), "flow": flow.detach().clone(), "wt": wt,
                                 "out": o_ref.detach().clone(), "gsrc": gs_ref, "gflow": gf_ref}
            sl_ref = ref.utils.smooothing_loss(flow.detach())
            assert torch.equal(sl_ref, R.smoothing_loss(flow.detach()))
            G[f"smooth.{name}"] = sl_ref.clone()
    finally:
        torch.Tensor.cuda = orig_cuda

    # ---------------- ReplayBuffer ----------------
    random.seed(5); rb_ref = ref.utils.ReplayBuffer(max_size=4)
    picks_ref = [rb_ref.push_and_pop(torch.full((1, 1, 2, 2), float(i))).flatten()[0].item() for i in range(24)]
    random.seed(5); rb_my = R.ReplayBuffer(max_size=4)
    picks = [rb_my.push_and_pop(torch.full((1, 1, 2, 2), float(i))).flatten()[0].item() for i in range(24)]
    assert picks_ref == picks
    G["replay.picks_seed5_size4"] = picks_ref

    # ---------------- iteration bodies: reference modules driven per the trainer lines ----------------
    Sx = 64
    # Cyc (CycTrainer.py:138-197)
    seed()
    nets = [ref.CycleGan.Generator(1, 1), ref.CycleGan.Discriminator(1), ref.CycleGan.Generator(1, 1), ref.CycleGan.Discriminator(1)]
    gA2B, dB, gB2A, dA = nets
    import itertools
    oDB = torch.optim.Adam(dB.parameters(), lr=1e-4, betas=(0.5, 0.999))
    oG = torch.optim.Adam(itertools.chain(gA2B.parameters(), gB2A.parameters()), lr=1e-4, betas=(0.5, 0.999))
    oDA = torch.optim.Adam(dA.parameters(), lr=1e-4, betas=(0.5, 0.999))
    mse, l1 = torch.nn.MSELoss(), torch.nn.L1Loss()
    t1, t0 = torch.ones(1, 1), torch.zeros(1, 1)
    bufA, bufB = ref.utils.ReplayBuffer(), ref.utils.ReplayBuffer()
    seed(); st = R.CycState()
    cyc_losses = []
    for it in range(2):
        rA, rB = R.synthetic_pair(1, Sx, seed=100 + it, phantom=True)
        oG.zero_grad()
        fB = gA2B(rA); lg1 = 1 * mse(dB(fB), t1)
        fA = gB2A(rB); lg2 = 1 * mse(dA(fA), t1)
        lc1 = 10 * l1(gB2A(fB), rA); lc2 = 10 * l1(gA2B(fA), rB)
        lt = lg1 + lg2 + lc1 + lc2; lt.backward(); oG.step()
        oDA.zero_grad()
        lr_ = 1 * mse(dA(rA), t1); fA_ = bufA.push_and_pop(fA); lf_ = 1 * mse(dA(fA_.detach()), t0)
        lDA = lr_ + lf_; lDA.backward(); oDA.step()
        oDB.zero_grad()
        lr_ = 1 * mse(dB(rB), t1); fB_ = bufB.push_and_pop(fB); lf_ = 1 * mse(dB(fB_.detach()), t0)
        lDB = lr_ + lf_; lDB.backward(); oDB.step()
        mine = R.cyc_step(st, rA, rB)
        refl = {"loss_G": float(lt), "loss_D_A": float(lDA), "loss_D_B": float(lDB)}
        for k, v in refl.items():
            assert v == mine[k], (it, k, v, mine[k])
        cyc_losses.append(mine)
    G["cyc_step.losses_64"] = cyc_losses
    G["cyc_step.G_A2B_head_w_after2"] = st.G_A2B["model_head.1.weight"].detach().clone()
    assert torch.equal(gA2B.model_head[1].weight, st.G_A2B["model_head.1.weight"])

    # Reg (RegTrainer.py:170-198) at 256 (Reg's minimum), 3-block generator to keep CPU time low
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        seed()
        g = ref.CycleGan.Generator(1, 1, n_residual_blocks=3); d = ref.CycleGan.Discriminator(1)
        oD = torch.optim.Adam(d.parameters(), lr=1e-4, betas=(0.5, 0.999))
        r = ref.reg.Reg(256, 256, 1, 1); tr = ref.transformer.Transformer_2D()
        oR = torch.optim.Adam(r.parameters(), lr=1e-4, betas=(0.5, 0.999))
        oG = torch.optim.Adam(g.parameters(), lr=1e-4, betas=(0.5, 0.999))
        seed(); st = R.RegState(n_blocks=3)
        reg_losses = []
        for it in range(2):
            rA, rB = R.synthetic_pair(1, 256, seed=200 + it, phantom=True)
            oR.zero_grad(); oG.zero_grad()
            fB = g(rA); T = r(fB, rB); sr_ = tr(fB, T)
            SR = 20 * l1(sr_, rB); adv = 1 * mse(d(fB), t1); SM = 10 * ref.utils.smooothing_loss(T)
            tot = SM + adv + SR; tot.backward(); oR.step(); oG.step()
            oD.zero_grad()
            with torch.no_grad():
                fB = g(rA)
            lD = 1 * mse(d(fB), t0) + 1 * mse(d(rB), t1); lD.backward(); oD.step()
            mine = R.reg_step(st, rA, rB)
            for k, v in {"SR_loss": float(SR), "adv_loss": float(adv), "SM_loss": float(SM), "loss_D_B": float(lD)}.items():
                assert v == mine[k], (it, k, v, mine[k])
            reg_losses.append(mine)
        G["reg_step.losses_256_nb3"] = reg_losses
    finally:
        torch.Tensor.cuda = orig_cuda

    G["meta"] = {"torch": str(torch.__version__), "reference": ref.root, "note": "all tensors fp32 CPU, seed 42 init"}
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    torch.save(G, OUT)
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB;", len(G), "entries")


# ======================================================================================================================
# golden_v2: the iteration bodies the reference cannot run as a trainer are pinned by exec'ing THE REFERENCE'S OWN SOURCE LINES
# (read from the reference tree at run time, never copied into this repository) on the reference's own nn.Modules
# ======================================================================================================================
OUT2 = os.path.join(os.path.dirname(HERE), "tests", "golden", "golden_v2.pt")


def reference_lines(ref_root, rel_path, first_marker, last_marker, after=None, lo=None, hi=None):
    """Source lines [first_marker .. last_marker] (inclusive, dedented) of a reference file; the span is looked up by its first and
    last statement (starting the search after the line containing `after`) and must lie inside the cited range [lo, hi]."""
    import textwrap
    lines = open(os.path.join(ref_root, rel_path), encoding="utf-8").read().split("\n")
    start = 0
    if after is not None:
        start = next(i for i, l in enumerate(lines) if after in l)
    a = next(i for i in range(start, len(lines)) if first_marker in lines[i])
    b = next(i for i in range(a, len(lines)) if last_marker in lines[i])
    assert lo is None or (lo <= a + 1 and b + 1 <= hi), (rel_path, a + 1, b + 1, lo, hi)
    return textwrap.dedent("\n".join(lines[a:b + 1])), (a + 1, b + 1)


def _adam(params, lr=1e-4):
    return torch.optim.Adam(params, lr=lr, betas=(0.5, 0.999))


def weight_fp(sd):
    return {k: (float(v.detach().double().sum()), float(v.detach().double().abs().sum())) for k, v in sd.items()}


def main_v2():
    import copy
    import types
    ref = load_reference()
    G = {}
    mse, l1 = torch.nn.MSELoss(), torch.nn.L1Loss()
    t1, t0 = torch.ones(1, 1), torch.zeros(1, 1)
    cfg = {"Smooth_lamda": 10, "Corr_lamda1": 20, "Corr_lamda2": 2, "Adv_lamda1": 1, "Adv_lamda": 1, "P2P_lamda": 100}
    env = {"torch": torch, "copy": copy, "Variable": (lambda x, **k: x), "smooothing_loss": ref.utils.smooothing_loss, "D": 2}

    # ---------------- masked L1 (HdTrainer.py:726-735) on its own: value and gradient ----------------
    src, span = reference_lines(ref.root, "trainer/HdTrainer.py", "bb = real_B1", "SR_loss2 = self.config['Corr_lamda2']",
                                after="class Hd_Trainer_x2", lo=705, hi=751)
    G["masked_l1.ref_lines"] = span
    g = torch.Generator().manual_seed(21)
    for name, S in (("a", 48), ("b", 64)):
        b2 = (torch.rand(2, 1, S, S, generator=g) * 2 - 1)
        b1 = (b2 * 1.7 + 0.1 * torch.randn(2, 1, S, S, generator=g)).clamp(-1, 1)
        warped = (b2 + 0.2 * torch.randn(2, 1, S, S, generator=g)).clamp(-1, 1)
        warped[0, 0, :4] = 0.0                                   # exact zeros inside the mask: the `== 0 -> -1` fill of the warped image
        b1[0, 0, :4] = 0.9
        wv = warped.clone().requires_grad_(True)
        loc = {"self": types.SimpleNamespace(config=cfg, L1_loss=l1), "real_B1": b1.clone(), "real_B2": b2.clone(), "SysRegist_A2B": wv * 1.0}
        exec(src, dict(env), loc)
        loc["SR_loss2"].backward()
        wm = warped.clone().requires_grad_(True)
        mine = cfg["Corr_lamda2"] * R.masked_l1(wm, b1, b2)
        mine.backward()
        assert torch.equal(loc["SR_loss2"].detach(), mine.detach()) and torch.equal(wv.grad, wm.grad), name
        G[f"masked_l1.{name}"] = {"warped": warped, "b1": b1, "b2": b2, "loss": (loc["SR_loss2"].detach() / cfg["Corr_lamda2"]).clone(),
                                  "grad": (wv.grad / cfg["Corr_lamda2"]).clone()}

    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self        # transformer.py:21 hard-codes .cuda()
    try:
        def reg_like_self(multiscale):
            seed()
            s = types.SimpleNamespace(config=cfg, L1_loss=l1, MSE_loss=mse, target_real=t1, target_fake=t0)
            s.netG_A2B = ref.HdGan.Generator(1, 1, n_residual_blocks=3)          # HdTrainer.py:99-105 / :610-616 creation order
            s.netD_B = ref.HdGan.Discriminator_m(1) if multiscale else ref.HdGan.Discriminator(1)
            s.optimizer_D_B = _adam(s.netD_B.parameters())
            s.R_A = ref.reg.Reg(256, 256, 1, 1)
            s.spatial_transform = ref.transformer.Transformer_2D()
            s.optimizer_R_A = _adam(s.R_A.parameters())
            s.optimizer_G = _adam(s.netG_A2B.parameters())
            s.criterionGAN = ref.HdGan.GANLoss()
            for k in ("input_A2", "input_B", "input_B2"):
                setattr(s, k, torch.empty(1, 1, 256, 256))
            return s

        def hd_batch(it):
            rA, rB = R.synthetic_pair(1, 256, seed=300 + it, phantom=True)
            return {"A2": rA, "B1": (rB * 1.7).clamp(-1, 1), "B2": rB}

        # ---------------- Hd stage 2 (HdTrainer.py:705-751) ----------------
        src, span = reference_lines(ref.root, "trainer/HdTrainer.py", "real_A2 = Variable(self.input_A2.copy_(batch['A2']))",
                                    "self.optimizer_D_B.step()", after="class Hd_Trainer_x2", lo=705, hi=751)
        G["hd_x2_step.ref_lines"] = span
        s = reg_like_self(True)
        seed(); st = R.RegState(multiscale_d=True, n_blocks=3)
        losses = []
        for it in range(2):
            batch = hd_batch(it)
            loc = {"self": s, "batch": {k: v.clone() for k, v in batch.items()}}
            exec(src, dict(env), loc)
            mine = R.hd_x2_step(st, batch["A2"], batch["B1"], batch["B2"])
            for k in ("SR_loss", "SR_loss2", "adv_loss", "SM_loss", "toal_loss", "loss_D_B"):
                assert float(loc[k]) == mine[k], (it, k, float(loc[k]), mine[k])
            losses.append(mine)
        for (k, p_), (k2, p2) in zip(s.netG_A2B.named_parameters(), st.G.items()):
            assert k == k2 and torch.equal(p_, p2), k
        G["hd_x2_step.losses_256_nb3"] = losses
        G["hd_x2_step.G_fp_after2"] = weight_fp(st.G)
        G["hd_x2_step.D_fp_after2"] = weight_fp(st.D)

        # ---------------- Hd stage 1 (HdTrainer.py:192-228) ----------------
        src, span = reference_lines(ref.root, "trainer/HdTrainer.py", "real_A2 = Variable(self.input_A2.copy_(batch['A2']))",
                                    "self.optimizer_D_B.step()", after="class Hd_Trainer_x1", lo=192, hi=228)
        G["hd_x1_step.ref_lines"] = span
        s = reg_like_self(False)
        seed(); st = R.RegState(n_blocks=3)
        losses = []
        for it in range(2):
            batch = hd_batch(it)
            loc = {"self": s, "batch": {k: v.clone() for k, v in batch.items()}}
            exec(src, dict(env), loc)
            mine = R.reg_step(st, batch["A2"], batch["B2"], corr=cfg["Corr_lamda1"], adv=cfg["Adv_lamda1"], smooth=cfg["Smooth_lamda"])
            for k in ("SR_loss", "adv_loss", "SM_loss", "toal_loss", "loss_D_B"):
                assert float(loc[k]) == mine[k], (it, k, float(loc[k]), mine[k])
            losses.append(mine)
        G["hd_x1_step.losses_256_nb3"] = losses
        G["hd_x1_step.G_fp_after2"] = weight_fp(st.G)
    finally:
        torch.Tensor.cuda = orig_cuda

    # ---------------- pix2pix (p2pTrainer.py:122-148) ----------------
    src, span = reference_lines(ref.root, "trainer/p2pTrainer.py", "real_A = Variable(self.input_A.copy_(batch['A']))",
                                "self.optimizer_D_B.step()", lo=122, hi=148)
    G["p2p_step.ref_lines"] = span
    seed()
    s = types.SimpleNamespace(config=cfg, L1_loss=l1, MSE_loss=mse, target_real=t1, target_fake=t0)
    s.netG_A2B = ref.CycleGan.Generator(1, 1)                                  # p2pTrainer.py:60-63
    s.netD_B = ref.CycleGan.Discriminator(2)
    s.optimizer_D_B = _adam(s.netD_B.parameters())
    s.optimizer_G = _adam(s.netG_A2B.parameters())
    s.input_A, s.input_B = torch.empty(1, 1, 64, 64), torch.empty(1, 1, 64, 64)
    seed(); st = R.P2pState()
    losses = []
    for it in range(2):
        rA, rB = R.synthetic_pair(1, 64, seed=400 + it, phantom=True)
        loc = {"self": s, "batch": {"A": rA.clone(), "B": rB.clone()}}
        exec(src, dict(env), loc)
        mine = R.p2p_step(st, rA, rB)
        for k in ("loss_L1", "loss_GAN_A2B", "toal_loss", "loss_D_B"):
            assert float(loc[k]) == mine[k], (it, k, float(loc[k]), mine[k])
        losses.append(mine)
    G["p2p_step.losses_64"] = losses
    G["p2p_step.G_fp_after2"] = weight_fp(st.G)

    # ---------------- Discriminator_m(num_D=2) + GANLoss weights [1.8, 0.2] (HdGan.py:236-256,269-293) ----------------
    seed(); dm = ref.HdGan.Discriminator_m(1, num_D=2)
    seed(); dm_sd = R.init_discriminator_m(1, num_D=2)
    G["discriminator_m2.state_fp"] = check_state(dm, dm_sd, "Discriminator_m(num_D=2)")
    x, _ = R.synthetic_pair(2, 128, seed=9, phantom=True)
    feats_ref = dm(x)
    feats = R.discriminator_m_forward(dm_sd, x, num_D=2)
    assert len(feats_ref) == 2
    for sr_, sm_ in zip(feats_ref, feats):
        for fr, f in zip(sr_, sm_):
            assert torch.equal(fr, f)
    gl = ref.HdGan.GANLoss()
    for flag in (True, False):
        assert torch.equal(gl(feats_ref, flag), R.gan_loss(feats, flag))
        G[f"discriminator_m2.ganloss_{flag}"] = gl(feats_ref, flag).detach().clone()
    G["discriminator_m2.last"] = [sc[-1].detach().clone() for sc in feats_ref]
    G["discriminator_m2.feat_shapes"] = [[tuple(f.shape) for f in sc] for sc in feats_ref]

    # ---------------- standalone layer modules: trainer/layers.py Conv / DownBlock / ResnetBlock, CycleGan.ResidualBlock ----------------
    def layer_case(name, make, shape, tuple_out=False):
        seed(7); m = make()
        gx = torch.Generator().manual_seed(31)
        x = torch.randn(*shape, generator=gx).requires_grad_(True)
        y = m(x)
        ys = y if tuple_out else (y,)
        wt = [torch.randn(t.shape, generator=gx) for t in ys]
        sum((t * w_).sum() for t, w_ in zip(ys, wt)).backward()
        G[f"layer.{name}"] = {"state": {k: v.detach().clone() for k, v in m.state_dict().items()}, "x": x.detach().clone(),
                              "y": [t.detach().clone() for t in ys], "wt": wt, "gx": x.grad.clone(),
                              "gparams": {k: p_.grad.clone() for k, p_ in m.named_parameters() if p_.grad is not None}}

    Lr = ref.layers
    layer_case("conv_lrelu_resnet", lambda: Lr.Conv(8, 16, 3, 1, 1, activation="leaky_relu", init_func="kaiming", bias=True, use_resnet=True,
                                                    use_norm=False), (2, 8, 16, 16))
    layer_case("conv_norm_relu", lambda: Lr.Conv(8, 16, 3, 1, 1, activation="relu", init_func="kaiming", bias=True, use_resnet=False,
                                                 use_norm=True), (2, 8, 12, 12))
    layer_case("conv_1x1_none", lambda: Lr.Conv(16, 8, 1, 1, 0, activation=None, init_func="zeros", bias=True), (1, 16, 8, 8))
    layer_case("downblock", lambda: Lr.DownBlock(2, 16, 3, 1, 1, activation="leaky_relu", init_func="kaiming", bias=True, use_resnet=True,
                                                 use_norm=False), (2, 2, 16, 16), tuple_out=True)
    layer_case("resnet_block", lambda: Lr.ResnetBlock(16, "reflect", Lr.norm_layer, False, True), (2, 16, 10, 10))
    layer_case("resnet_transformer", lambda: Lr.ResnetTransformer(16, 2, "kaiming"), (1, 16, 8, 8))
    layer_case("residual_block", lambda: ref.CycleGan.ResidualBlock(16), (2, 16, 12, 12))

    G["meta"] = {"torch": str(torch.__version__), "reference": ref.root,
                 "note": "iteration bodies: exec of the reference's own source lines on the reference's own modules; fp32 CPU"}
    torch.save(G, OUT2)
    print("wrote", OUT2, os.path.getsize(OUT2) / 1e6, "MB;", len(G), "entries")


# ======================================================================================================================
# golden_v3: the data-side / evaluation-side arithmetic (oracle/restate_eval.py) pinned against the reference's own source lines, PIL and torch
# ======================================================================================================================
OUT3 = os.path.join(os.path.dirname(HERE), "tests", "golden", "golden_v3.pt")


def main_v3():
    import numpy as np
    from PIL import Image
    import torchvision.transforms.functional as TF
    from oracle import restate_eval as RE
    ref = load_reference()
    G = {}
    rng = np.random.default_rng(7)
    # to_windowdata (CycTrainer.py:34-57) and the metric methods (:362-398): exec the reference's lines
    src_w, span_w = reference_lines(ref.root, "trainer/CycTrainer.py", "def to_windowdata(image,WC,WW):", "return image", lo=34, hi=57)
    src_m, span_m = reference_lines(ref.root, "trainer/CycTrainer.py", "def PSNR(self, fake, real):", "return UQI", lo=362, hi=398)
    env = {"np": np}
    exec(src_w, env)
    exec(src_m, env)
    G["eval.ref_lines"] = {"to_windowdata": span_w, "metrics": span_m}
    src_b, span_b = reference_lines(ref.root, "trainer/CycTrainer.py", "b = to_windowdata(real_B, WC, WW)", "fake_B[fake_B==0]=-1", lo=286, hi=320)
    # keep only the image-building statements of that span (the metric calls in between need skimage / lpips)
    keep = [ln for ln in src_b.split("\n") if not any(t in ln for t in ("self.", "measure.", "loss_fn", "+=", "LPIPSw", "#"))]
    src_b = "\n".join(keep)
    cases = {}
    for name, (H, W) in {"a": (64, 80), "b": (96, 96)}.items():
        yy, xx = np.mgrid[0:H, 0:W]
        disc = ((yy - H / 2) ** 2 + (xx - W / 2) ** 2) <= (0.4 * min(H, W)) ** 2
        real = np.where(disc, rng.uniform(-0.55, -0.35, (H, W)), -1.0).astype(np.float32)       # soft tissue ~ 0..400 HU inside, air outside
        vessels = rng.uniform(0, 1, (H, W)) > 0.9
        real = np.where(disc & vessels, real + 0.12, real).astype(np.float32)
        fake = (real + rng.normal(0, 0.03, (H, W))).astype(np.float32).clip(-1, 1)
        WC, WW = 40.0, 400.0
        loc = {"real_B": real.copy(), "fake_B": fake.copy(), "WC": WC, "WW": WW}
        exec(src_b, dict(env), loc)
        c, b, fm, rm = RE.eval_images(fake.copy(), real.copy(), WC, WW)
        for mine, key in ((c, "c"), (b, "b"), (fm, "fake_B"), (rm, "real_B")):
            assert np.array_equal(mine, loc[key]), (name, key)
        assert np.array_equal(RE.to_windowdata(real.copy().astype(np.float64), WC, WW), env["to_windowdata"](real.copy().astype(np.float64), WC, WW))
        vals = {}
        for tag, (f_, r_) in {"w": (c, b), "raw": (fm, rm)}.items():
            for fn in ("MAE", "PSNR", "UQI"):
                v_ref = float(env[fn](None, f_, r_)); v = float(getattr(RE, fn)(f_, r_))
                assert v == v_ref, (name, tag, fn, v, v_ref)
                vals[f"{fn}_{tag}"] = v_ref
            vals[f"SSIM_{tag}"] = float(RE.SSIM(f_, r_))
        cases[name] = {"fake": torch.from_numpy(fake), "real": torch.from_numpy(real), "WC": WC, "WW": WW, "metrics": vals,
                       "int16": torch.from_numpy(RE.to_dicom_int16(fake))}
    # an all-air slice: the "no valid pixel" branches of MAE / PSNR
    air = np.full((40, 40), -1.0, np.float32)
    c, b, fm, rm = RE.eval_images(air.copy(), air.copy(), 40.0, 400.0)
    assert float(env["MAE"](None, fm, rm)) == float(RE.MAE(fm, rm)) and float(env["PSNR"](None, fm, rm)) == float(RE.PSNR(fm, rm))
    G["eval.cases"] = cases
    # read_dicom / read_ori_w arithmetic (datasets.py:74-82, 45-65) on synthetic stored pixel values
    src_d, span_d = reference_lines(ref.root, "trainer/datasets.py", "image2[image2<0]=0", "image2 = (image2 - 0.5)/0.5", after="def read_dicom", lo=74, hi=82)
    raw = rng.integers(-50, 4096, (32, 48)).astype(np.int16)
    loc = {"image2": raw.astype(np.int64)}
    exec(src_d, {"np": np}, loc)
    assert np.array_equal(loc["image2"], RE.read_dicom_norm(raw))
    src_o, span_o = reference_lines(ref.root, "trainer/datasets.py", "center =50", "image1 = (image1 - 0.5)/0.5", after="def read_ori_w", lo=36, hi=71)
    hu = rng.integers(-1100, 3000, (32, 48)).astype(np.int16)
    loc = {"data1": hu.astype(np.int64)}
    exec(src_o, {"np": np}, loc)
    assert np.array_equal(loc["image1"], RE.window_image(hu.astype(np.float64), 50, 400))
    G["data.ref_lines"] = {"read_dicom": span_d, "read_ori_w": span_o}
    G["data.raw"] = torch.from_numpy(raw); G["data.raw_norm"] = torch.from_numpy(RE.read_dicom_norm(raw).astype(np.float32))
    G["data.hu"] = torch.from_numpy(hu); G["data.hu_window"] = torch.from_numpy(RE.window_image(hu.astype(np.float64), 50, 400).astype(np.float32))
    # RandomAffine resampling against PIL, Resize against torch
    img = rng.standard_normal((97, 131)).astype(np.float32)
    aff = []
    for angle, tr, sc in ((1.0, (2, -1), 1.02), (-7.5, (5, 3), 0.9), (0.3, (0, 0), 1.0), (33.0, (-20, 11), 1.3)):
        m = TF._get_inverse_affine_matrix((131 * 0.5, 97 * 0.5), angle, list(tr), sc, [0.0, 0.0])
        pil = np.asarray(Image.fromarray(img, mode="F").transform((131, 97), Image.AFFINE, m, resample=Image.NEAREST, fillcolor=-1))
        assert np.array_equal(pil, RE.affine_nearest(img, m, -1.0))
        aff.append({"m": torch.tensor(m, dtype=torch.float64), "out": torch.from_numpy(pil.copy()), "angle": angle, "translate": tr, "scale": sc})
    G["data.affine_src"] = torch.from_numpy(img); G["data.affine"] = aff
    x = torch.from_numpy(img)[None, None]
    G["data.resize"] = {(h, w): torch.nn.functional.interpolate(x, size=[h, w])[0, 0].clone() for h, w in ((64, 64), (200, 257), (97, 131))}
    G["meta"] = {"torch": str(torch.__version__), "note": "eval/data arithmetic: reference source lines exec'd; affine vs PIL " + Image.__version__}
    torch.save(G, OUT3)
    print("wrote", OUT3, os.path.getsize(OUT3) / 1e6, "MB;", len(G), "entries")


if __name__ == "__main__":
    if "--v3-only" in sys.argv:
        main_v3()
    else:
        if "--v2-only" not in sys.argv:
            main()
        main_v2()
        main_v3()
