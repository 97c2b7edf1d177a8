"""GPU parity of whole training iterations (the trainer entry points) against the oracle's restated iteration bodies
(oracle/restate.py: cyc_step / reg_step / hd_x2_step / p2p_step) and the golden losses frozen from the real reference."""
import random

import pytest
import torch

from util import grad_report, trainer_leaves

pytestmark = pytest.mark.gpu

# fp32 validation mode: EVERY parameter gradient of every network of the iteration is compared with the restated reference body
# (per tensor and globally).  That is what makes these tests decisive about the plumbing of a step -- loss weights, signs, which
# gradient reaches which network (warp -> fake_B and flow, flow -> Reg -> fake_B, D -> fake_B, ...): a wrong route or sign changes
# whole tensors by O(1), three orders of magnitude above the fp32 tolerance, whereas a loss VALUE may barely notice it.
G_TENSOR, G_GLOBAL = 5e-3, 2e-3
# In the Reg / Hd / P2p bodies the discriminator update comes AFTER the generator's Adam step and re-runs the updated generator
# (RegTrainer.py:189-198).  Adam's first step moves every weight by lr * sign(gradient): the sign of near-zero-gradient weights is
# round-off, so the two updated generators -- and with them the fake slices the discriminator sees -- differ at the 1e-3 level,
# and the discriminator gradients at ~1e-2 (measured 1.1-1.2e-2 global).  (The Cyc body takes its fakes from BEFORE the update:
# there the discriminator gradients agree to 1e-6.)
D_AFTER_G = (5e-2, 3e-2)


def _seed(s=42):
    random.seed(s)
    torch.manual_seed(s)


def _cfg(name, size, batch=1, **kw):
    c = {"name": name, "noise_level": 1, "port": 8097, "save_root": "", "image_save": "", "Adv_lamda": 1, "Cyc_lamda": 10,
         "Corr_lamda": 20, "Smooth_lamda": 10, "P2P_lamda": 100, "Adv_lamda1": 1, "Adv_lamda2": 0.1, "Corr_lamda1": 20,
         "Corr_lamda2": 2, "epoch": 0, "n_epochs": 1, "batchSize": batch, "lr": 1e-4, "lrd": 1e-4, "decay_epoch": 1, "size": size,
         "input_nc": 1, "output_nc": 1, "cuda": True, "n_cpu": 1, "precision": "fp32", "synthetic": True, "save_checkpoints": False}
    c.update(kw)
    return c


def _close(a, b, tol):
    return abs(a - b) <= tol * abs(b) + 1e-7


def test_cyc_step_matches_reference_losses(golden):
    from oracle import restate as R
    from trainer import Cyc_Trainer
    _seed(); tr = Cyc_Trainer(_cfg("CycleGan", 64))
    _seed(); st = R.CycState()
    for it, ref in enumerate(golden["cyc_step.losses_64"]):
        rA, rB = R.synthetic_pair(1, 64, seed=100 + it, phantom=True)
        out = tr.step({"A": rA, "B": rB})
        for k in ("loss_G", "loss_D_A", "loss_D_B"):
            assert _close(float(out[k]), ref[k], 2e-3 if it else 2e-4), (it, k, float(out[k]), ref[k])
        if it == 0:        # every parameter gradient of the four networks (both uses of each generator summed)
            R.cyc_step(st, rA, rB)
            for mod, leaf in ((tr.netG_A2B, st.G_A2B), (tr.netG_B2A, st.G_B2A), (tr.netD_A, st.D_A), (tr.netD_B, st.D_B)):
                grad_report(*trainer_leaves([(mod, leaf)]), tensor_tol=2e-2, glob_tol=5e-3)
    # the UPDATE itself, not closeness to the initial weights.  Two Adam steps move every weight by ~lr * sign(gradient) each (the
    # first steps of Adam are sign-like, and the second gradient already differs by a few per cent because the first update flipped
    # the sign of near-zero-gradient weights in all four networks), so the comparison is element-wise: the large majority of the
    # elements must agree to 30 % of the step size -- an optimizer that never stepped (update 0) agrees nowhere.
    w = tr.netG_A2B.model_head[1].weight.detach().cpu()
    _seed(); w0 = R.init_generator(1, 1)["model_head.1.weight"]
    upd, upd_ref = w - w0, golden["cyc_step.G_A2B_head_w_after2"] - w0
    assert float(upd_ref.abs().max()) > 1e-4 and float(upd.abs().max()) > 1e-4
    agree = float(((upd - upd_ref).abs() <= 0.3 * 2e-4).float().mean())
    assert agree >= 0.85, agree


def test_reg_step_matches_oracle():
    from oracle import restate as R
    from trainer import Reg_Trainer
    _seed(); tr = Reg_Trainer(_cfg("RegGan", 256))
    _seed(); st = R.RegState()
    rA, rB = R.synthetic_pair(1, 256, seed=200, phantom=True)
    out = tr.step({"A": rA, "B": rB})
    ref = R.reg_step(st, rA, rB)
    for k in ("SR_loss", "adv_loss", "loss_D_B", "toal_loss"):
        assert _close(float(out[k]), ref[k], 3e-4), (k, float(out[k]), ref[k])
    assert _close(float(out["SM_loss"]), ref["SM_loss"], 2e-2), (float(out["SM_loss"]), ref["SM_loss"])   # ~1e-7 magnitude term
    # gradients: generator (fed by warp, Reg and D), registration net (fed by warp and smoothness), discriminator
    for mod, leaf in ((tr.netG_A2B, st.G), (tr.R_A, st.R)):
        grad_report(*trainer_leaves([(mod, leaf)]), tensor_tol=2e-2, glob_tol=5e-3)
    grad_report(*trainer_leaves([(tr.netD_B, st.D)]), tensor_tol=D_AFTER_G[0], glob_tol=D_AFTER_G[1])


def test_hd_x2_and_p2p_steps_match_oracle():
    from oracle import restate as R
    from trainer import Hd_Trainer_x2, P2p_Trainer
    _seed(); tr = Hd_Trainer_x2(_cfg("HdGan", 256))
    _seed(); st = R.RegState(multiscale_d=True)
    rA, rB = R.synthetic_pair(1, 256, seed=300, phantom=True)
    rB1 = (rB * 1.7).clamp(-1, 1)
    out = tr.step({"A2": rA, "B1": rB1, "B2": rB})
    ref = R.hd_x2_step(st, rA, rB1, rB)
    for k in ("SR_loss", "adv_loss", "loss_D_B", "toal_loss"):
        assert _close(float(out[k]), ref[k], 3e-4), (k, float(out[k]), ref[k])
    for mod, leaf in ((tr.netG_A2B, st.G), (tr.R_A, st.R)):
        grad_report(*trainer_leaves([(mod, leaf)]), tensor_tol=2e-2, glob_tol=5e-3)
    grad_report(*trainer_leaves([(tr.netD_B, st.D)]), tensor_tol=D_AFTER_G[0], glob_tol=D_AFTER_G[1])

    _seed(); tp = P2p_Trainer(_cfg("P2p", 64))
    _seed(); sp = R.P2pState()
    rA, rB = R.synthetic_pair(1, 64, seed=400, phantom=True)
    out = tp.step({"A": rA, "B": rB})
    ref = R.p2p_step(sp, rA, rB)
    for k in ("loss_L1", "loss_GAN_A2B", "loss_D_B"):
        assert _close(float(out[k]), ref[k], 3e-4), (k, float(out[k]), ref[k])
    grad_report(*trainer_leaves([(tp.netG_A2B, sp.G)]), tensor_tol=G_TENSOR, glob_tol=G_GLOBAL)
    grad_report(*trainer_leaves([(tp.netD_B, sp.D)]), tensor_tol=D_AFTER_G[0], glob_tol=D_AFTER_G[1])


def test_hd_x1_step_matches_oracle():
    """Hd stage 1 (HdTrainer.py:192-228): the Reg-GAN body with the Hd loss weights and inputs (A2 -> B2), fp32, losses + gradients."""
    from oracle import restate as R
    from trainer import Hd_Trainer_x1
    _seed(); tr = Hd_Trainer_x1(_cfg("HdGan", 256, Corr_lamda1=15, Adv_lamda1=0.5))       # non-default weights: the x1 keys are the ones read
    _seed(); st = R.RegState()
    rA, rB = R.synthetic_pair(1, 256, seed=310, phantom=True)
    out = tr.step({"A2": rA, "B1": (rB * 1.7).clamp(-1, 1), "B2": rB})
    ref = R.reg_step(st, rA, rB, corr=15, adv=0.5, smooth=10)
    for k in ("SR_loss", "adv_loss", "loss_D_B", "toal_loss"):
        assert _close(float(out[k]), ref[k], 3e-4), (k, float(out[k]), ref[k])
    for mod, leaf in ((tr.netG_A2B, st.G), (tr.R_A, st.R)):
        grad_report(*trainer_leaves([(mod, leaf)]), tensor_tol=2e-2, glob_tol=5e-3)
    grad_report(*trainer_leaves([(tr.netD_B, st.D)]), tensor_tol=D_AFTER_G[0], glob_tol=D_AFTER_G[1])


def test_bf16_cyc_step_runs_and_tracks(golden):
    from oracle import restate as R
    from trainer import Cyc_Trainer
    _seed(); tr = Cyc_Trainer(_cfg("CycleGan", 64, precision="bf16"))
    ref = golden["cyc_step.losses_64"][0]
    rA, rB = R.synthetic_pair(1, 64, seed=100, phantom=True)
    out = tr.step({"A": rA, "B": rB})
    for k in ("loss_G", "loss_D_A", "loss_D_B"):
        assert _close(float(out[k]), ref[k], 5e-2), (k, float(out[k]), ref[k])
    import ctagan
    ctagan.set_precision("bf16")


def _run_cyc(mode, steps, size=64, precision="fp32"):
    """`steps` Cyc iterations with a 2-slot ReplayBuffer (so the swap path is exercised from the third push on).  Returns, per
    iteration, the losses and a copy of EVERY parameter of the four networks after the iteration."""
    from oracle import restate as R
    from trainer import Cyc_Trainer
    from ctagan.graphs import GraphedTrainer
    from ctagan.replay import ReplayBuffer
    import ctagan
    _seed(); tr = Cyc_Trainer(_cfg("CycleGan", size, precision=precision))
    tr.fake_A_buffer, tr.fake_B_buffer = ReplayBuffer(2), ReplayBuffer(2)
    batches = [R.synthetic_pair(1, size, seed=500 + i, phantom=True) for i in range(steps)]
    random.seed(7)
    nets = (tr.netG_A2B, tr.netG_B2A, tr.netD_A, tr.netD_B)

    def snap(losses):
        torch.cuda.synchronize()
        return ({k: float(v) for k, v in losses.items()}, [p.detach().clone() for n in nets for p in n.parameters()])

    out = []
    if mode == "graph":
        g = GraphedTrainer(tr, warmup=1, replay_first=False)       # first call = the eager warm-up iteration + the capture
        for a, b in batches:
            out.append(snap(g.step_device((a.cuda(), b.cuda()))))
    else:
        fn = tr.step if mode == "fused" else tr.step_two_phase
        for a, b in batches:
            out.append(snap(fn({"A": a, "B": b})))
    ctagan.set_precision("bf16")
    return out


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cyc_overlapped_schedule_equals_serial_order(precision):
    """The one-program schedule (device-side ReplayBuffer moves, discriminator updates beside the generator backward, weight-gradient
    lanes, deferred gradient collection) must compute what the reference's serial order computes -- including the buffer's random
    swaps -- eagerly and as a CUDA graph.  The library is bit-reproducible (tests/test_gpu_determinism.py) and the three schedules run
    the same kernels on the same operands, so the comparison is EXACT: after every iteration every parameter of the four networks
    (i.e. every gradient that went through Adam) and every loss must be bit-identical.  A missed dependency or a wrong swap
    cannot hide behind a tolerance."""
    serial = _run_cyc("two_phase", 4, precision=precision)
    fused = _run_cyc("fused", 4, precision=precision)
    graph = _run_cyc("graph", 4, precision=precision)
    for name, run in (("fused", fused), ("graph", graph)):
        for i in range(4):
            assert run[i][0] == serial[i][0], (name, i, run[i][0], serial[i][0])
            bad = [k for k, (p, q) in enumerate(zip(run[i][1], serial[i][1])) if not torch.equal(p, q)]
            assert not bad, (name, "iteration", i, "parameters", bad[:8],
                             float((run[i][1][bad[0]] - serial[i][1][bad[0]]).abs().max()))


def test_train_loop_runs_graphed_and_decays_lr():
    """trainer.train(): the reference's epoch loop on CUDA-graph replays; the decayed learning rate must reach the captured Adam."""
    from trainer import Cyc_Trainer
    _seed(); tr = Cyc_Trainer(_cfg("CycleGan", 64, precision="bf16", n_epochs=1, decay_epoch=2, synthetic_batches=3, log_every=1000))
    w0 = tr.netG_A2B.model_head[1].weight.detach().clone()
    tr.train()                                        # epochs 1 (lr 1e-4), 2 (5e-5), 3 (2.5e-5): 3 batches each, every batch exactly once
    assert tr.step_count == 9
    lr = tr.optimizer_G.param_groups[0]["lr"]
    assert torch.is_tensor(lr) and float(lr) == pytest.approx(2.5e-5), lr   # lr -= lr / decay_epoch, twice (CycTrainer.py:117-126)
    assert float(tr.optimizer_D_A.param_groups[0]["lr"]) == pytest.approx(1e-4)     # optimizer_D_A never decays (CycTrainer.py:117-126)
    w1 = tr.netG_A2B.model_head[1].weight.detach()
    assert torch.isfinite(w1).all() and float((w1 - w0).abs().max()) > 1e-5
    for v in tr.last_losses.values():
        assert torch.isfinite(v).all()
    import ctagan
    ctagan.set_precision("bf16")
