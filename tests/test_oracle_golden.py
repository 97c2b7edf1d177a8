"""CPU: the oracle restatement (oracle/restate.py) against golden vectors frozen from the REAL reference
modules by oracle/make_golden.py.  No reference tree needed."""
import random

import pytest

import torch

from oracle import restate as R

torch.set_num_threads(8)


def _seed(s=42):
    random.seed(s)
    torch.manual_seed(s)


def _fp(sd):
    return {k: (tuple(v.shape), float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items()}


def test_state_dict_layout_and_init(golden):
    for name, fn in (("generator", lambda: R.init_generator(1, 1)), ("discriminator1", lambda: R.init_discriminator(1)),
                     ("discriminator2", lambda: R.init_discriminator(2)), ("discriminator_m", lambda: R.init_discriminator_m(1)),
                     ("reg", lambda: R.init_reg(1, 1))):
        _seed()
        fp = _fp(fn())
        g = golden[name + ".state_fp"]
        assert list(fp) == list(g), name
        for k in fp:
            assert fp[k][0] == g[k][0] and abs(fp[k][1] - g[k][1]) <= 1e-9 * max(1, abs(g[k][2])), (name, k)
    assert len(golden["generator.state_fp"]) == 48 and len(golden["reg.state_fp"]) == 80


def test_generator(golden):
    _seed(); sd = R.init_generator(1, 1)
    a, b = R.synthetic_pair(1, 64, seed=42)
    leaf = R.leafify(sd)
    y = R.generator_forward(leaf, a)
    assert torch.allclose(y, golden["generator.out_64"], rtol=0, atol=2e-6)
    (y * b).sum().backward()
    for k, (nrm, s) in golden["generator.grad_fp_64"].items():
        got = float(leaf[k].grad.double().norm())
        assert abs(got - nrm) <= 2e-4 * max(nrm, 1e-6) + 1e-7, k
    a2, _ = R.synthetic_pair(2, 128, seed=7, phantom=True)
    assert torch.allclose(R.generator_forward(sd, a2), golden["generator.out_128_phantom_b2"], rtol=0, atol=2e-6)


def test_discriminators(golden):
    a, b = R.synthetic_pair(1, 64, seed=42)
    for nc in (1, 2):
        _seed(); sd = R.init_discriminator(nc)
        x = torch.cat([a, b], 1)[:, :nc]
        x2 = torch.cat([x, -x], 0)
        p = R.discriminator_forward(sd, x2)
        assert torch.allclose(p, golden[f"discriminator{nc}.pred_64_b2"], rtol=1e-5, atol=1e-7)
        assert torch.allclose(R.mse_vs_const(p, 1.0), golden[f"discriminator{nc}.mse_real"], rtol=1e-5)
    _seed(); sd = R.init_discriminator_m(1)
    feats = R.discriminator_m_forward(sd, a)
    assert [tuple(f.shape) for f in feats[0]] == golden["discriminator_m.feat_shapes"]
    assert torch.allclose(feats[0][-1], golden["discriminator_m.last_64"], rtol=1e-5, atol=1e-6)
    for flag in (True, False):
        assert torch.allclose(R.gan_loss(feats, flag), golden[f"discriminator_m.ganloss_{flag}"], rtol=1e-5)


def test_reg(golden):
    _seed(); sd = R.init_reg(1, 1)
    ra, rb = R.synthetic_pair(1, 256, seed=3, phantom=True)
    fl = R.reg_forward(sd, ra, rb)
    ref = golden["reg.flow_256"]
    assert (fl - ref).abs().max() <= 1e-5 * ref.abs().max() + 1e-9
    _seed(1); wbig = torch.randn_like(sd["offset_map.output.conv2d.weight"]) * 0.05
    leaf = R.leafify(sd)
    leaf["offset_map.output.conv2d.weight"].data.copy_(wbig)
    fl = R.reg_forward(leaf, ra, rb)
    ref = golden["reg.flow_256_bigw"]
    assert (fl - ref).abs().max() <= 1e-4 * ref.abs().max()
    sm = R.smoothing_loss(fl)
    assert torch.allclose(sm, golden["reg.smooth_bigw"], rtol=1e-4)


def test_warp_and_smooth(golden):
    for name in ("small", "sq", "tiny_flow", "oob"):
        g = golden[f"warp.{name}"]
        src = g["src"].clone().requires_grad_(True)
        flow = g["flow"].clone().requires_grad_(True)
        o = R.warp(src, flow)
        (o * g["wt"]).sum().backward()
        assert torch.allclose(o, g["out"], rtol=0, atol=1e-6), name
        assert torch.allclose(src.grad, g["gsrc"], rtol=0, atol=1e-6), name
        assert torch.allclose(flow.grad, g["gflow"], rtol=0, atol=1e-5), name
        assert torch.allclose(R.smoothing_loss(g["flow"]), golden[f"smooth.{name}"], rtol=1e-6), name


def test_replay_buffer(golden):
    random.seed(5); rb = R.ReplayBuffer(max_size=4)
    picks = [rb.push_and_pop(torch.full((1, 1, 2, 2), float(i))).flatten()[0].item() for i in range(24)]
    assert picks == golden["replay.picks_seed5_size4"]


def test_cyc_step_losses(golden):
    _seed(); st = R.CycState()
    for it, ref in enumerate(golden["cyc_step.losses_64"]):
        rA, rB = R.synthetic_pair(1, 64, seed=100 + it, phantom=True)
        mine = R.cyc_step(st, rA, rB)
        for k, v in ref.items():
            assert abs(mine[k] - v) <= 2e-4 * abs(v) + 1e-7, (it, k, mine[k], v)
    assert torch.allclose(st.G_A2B["model_head.1.weight"], golden["cyc_step.G_A2B_head_w_after2"], atol=1e-5)


# ---- golden_v2: vectors produced by exec'ing the reference's own source lines (oracle/make_golden.py:main_v2) ----


def _weight_fp(sd):
    return {k: (float(v.detach().double().sum()), float(v.detach().double().abs().sum())) for k, v in sd.items()}


def _fp_close(a, b, rtol=1e-5):
    assert a.keys() == b.keys()
    for k in a:
        for x, y in zip(a[k], b[k]):
            assert abs(x - y) <= rtol * max(abs(y), 1e-3), (k, x, y)


def test_masked_l1_pin(golden):
    """HdTrainer.py:726-735 (exec'd from the reference tree when the golden file was made): value and gradient."""
    assert 705 <= golden["masked_l1.ref_lines"][0] and golden["masked_l1.ref_lines"][1] <= 751
    for name in ("a", "b"):
        g = golden[f"masked_l1.{name}"]
        w = g["warped"].clone().requires_grad_(True)
        loss = R.masked_l1(w, g["b1"], g["b2"])
        loss.backward()
        assert torch.equal(loss.detach(), g["loss"]) and torch.equal(w.grad, g["grad"])


def test_p2p_step_pin(golden):
    """p2pTrainer.py:122-148, two iterations."""
    _seed(); st = R.P2pState()
    for it, ref in enumerate(golden["p2p_step.losses_64"]):
        a, b = R.synthetic_pair(1, 64, seed=400 + it, phantom=True)
        mine = R.p2p_step(st, a, b)
        for k, v in ref.items():
            assert abs(mine[k] - v) <= 1e-5 * abs(v) + 1e-8, (it, k, mine[k], v)
    _fp_close(_weight_fp(st.G), golden["p2p_step.G_fp_after2"])


@pytest.mark.timeout(600)
def test_hd_steps_pin(golden):
    """Hd stage 1 (HdTrainer.py:192-228) and stage 2 (:705-751: Discriminator_m + GANLoss + masked L1), first iteration each
    (256x256 with a 3-block generator, as frozen)."""
    for key, multiscale in (("hd_x1_step", False), ("hd_x2_step", True)):
        _seed(); st = R.RegState(multiscale_d=multiscale, n_blocks=3)
        a, b = R.synthetic_pair(1, 256, seed=300, phantom=True)
        b1 = (b * 1.7).clamp(-1, 1)
        mine = R.hd_x2_step(st, a, b1, b) if multiscale else R.reg_step(st, a, b, corr=20, adv=1, smooth=10)
        for k, v in golden[f"{key}.losses_256_nb3"][0].items():
            assert abs(mine[k] - v) <= 2e-5 * abs(v) + 1e-9, (key, k, mine[k], v)


def test_discriminator_m_two_scales_pin(golden):
    """Discriminator_m(num_D=2): scale order, centre crop (HdGan.py:236-256) and the GANLoss scale weights [1.8, 0.2] (:273)."""
    _seed(); sd = R.init_discriminator_m(1, num_D=2)
    for k, (shape, s, sa) in golden["discriminator_m2.state_fp"].items():
        assert tuple(sd[k].shape) == tuple(shape) and abs(float(sd[k].double().sum()) - s) <= 1e-6 * max(1, abs(s))
    x, _ = R.synthetic_pair(2, 128, seed=9, phantom=True)
    feats = R.discriminator_m_forward(sd, x, num_D=2)
    assert [[tuple(f.shape) for f in sc] for sc in feats] == [[tuple(t) for t in sc] for sc in golden["discriminator_m2.feat_shapes"]]
    for mine, ref in zip(feats, golden["discriminator_m2.last"]):
        assert torch.allclose(mine[-1], ref, rtol=1e-5, atol=1e-6)
    for flag in (True, False):
        assert torch.allclose(R.gan_loss(feats, flag), golden[f"discriminator_m2.ganloss_{flag}"], rtol=1e-5)


# ---- the bf16-emulating oracle (oracle/bf16_emu.py) ----


def test_bf16_emulator_disabled_equals_restatement():
    """With rounding switched off the emulator's networks ARE the pinned restatement (so its only additions are the roundings)."""
    from oracle import bf16_emu as B
    _seed(); g, d, r = R.init_generator(1, 1, 2), R.init_discriminator(2), R.init_reg(1, 1)
    a, b = R.synthetic_pair(1, 256, seed=5, phantom=True)
    ref = (R.generator_forward(g, a[:, :, :64, :64], 2), R.discriminator_forward(d, torch.cat([a, b], 1)), R.reg_forward(r, a, b))
    with B.emulate(enabled=False):
        got = (R.generator_forward(g, a[:, :, :64, :64], 2), R.discriminator_forward(d, torch.cat([a, b], 1)), R.reg_forward(r, a, b))
    assert R.generator_forward.__module__ == "oracle.restate"             # restored on exit
    for x, y in zip(got, ref):
        assert torch.equal(x, y)


def test_bf16_emulator_stays_in_the_bf16_envelope():
    """Rounding at the kernels' storage points moves a random-init generator by about the operand-rounding envelope measured for the
    reference's own bf16 (~3e-2 max-rel), and gradients flow through the rounding points."""
    from oracle import bf16_emu as B
    _seed(); sd = R.leafify(R.init_generator(1, 1, 3))
    a, b = R.synthetic_pair(1, 64, seed=42)
    ref = R.generator_forward(sd, a, 3).detach()
    with B.emulate():
        y = R.generator_forward(sd, a, 3)
        R.l1_loss(y, b).backward()
    err = float((y.detach() - ref).abs().max() / ref.abs().max())
    assert 1e-4 < err < 8e-2, err
    assert torch.equal(y.detach(), y.detach().bfloat16().float())         # the module output is a stored bf16 tensor
    gw = sd["model_body.1.conv_block.1.weight"].grad
    assert gw is not None and float(gw.abs().max()) > 0


# ---- golden_v3: data-side / evaluation-side arithmetic (oracle/restate_eval.py) ----


def test_eval_and_data_restatement_pins(golden):
    import numpy as np
    from oracle import restate_eval as RE
    for name, case in golden["eval.cases"].items():
        vals = RE.slice_metrics(case["fake"].numpy(), case["real"].numpy(), case["WC"], case["WW"])
        ref = case["metrics"]
        order = ("MAE_w", "PSNR_w", "SSIM_w", "UQI_w", "MAE_raw", "PSNR_raw", "SSIM_raw", "UQI_raw")
        for v, k in zip(vals, order):
            assert abs(float(v) - ref[k]) <= 1e-12 * max(1.0, abs(ref[k])), (name, k, v, ref[k])
        assert np.array_equal(RE.to_dicom_int16(case["fake"].numpy()), case["int16"].numpy())
    assert np.array_equal(RE.read_dicom_norm(golden["data.raw"].numpy()).astype(np.float32), golden["data.raw_norm"].numpy())
    assert np.array_equal(RE.window_image(golden["data.hu"].numpy().astype(np.float64), 50, 400).astype(np.float32), golden["data.hu_window"].numpy())
    for a in golden["data.affine"]:
        assert np.array_equal(RE.affine_nearest(golden["data.affine_src"].numpy(), a["m"].tolist(), -1.0), a["out"].numpy())
