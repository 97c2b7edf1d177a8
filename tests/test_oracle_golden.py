"""CPU: the oracle restatement (oracle/restate.py) against golden vectors frozen from the REAL reference
modules by oracle/make_golden.py.  No reference tree needed."""
import random

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
