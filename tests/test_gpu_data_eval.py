"""GPU input pipeline (f1) and on-GPU evaluation (f2) against the numpy restatement of the reference's CPU code
(oracle/restate_eval.py, pinned against the reference's own source lines / PIL / torch in golden_v3)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_normalise_window_resize_affine(golden):
    from ctagan import data as D
    raw = golden["data.raw"].cuda()
    assert torch.equal(D.hu_to_unit(raw).cpu(), golden["data.raw_norm"])
    assert torch.equal(D.hu_window(golden["data.hu"].cuda(), 50, 400).cpu(), golden["data.hu_window"])
    src = golden["data.affine_src"].cuda()
    for (h, w), ref in golden["data.resize"].items():
        assert torch.equal(D.resize_nearest(src[None], (h, w))[0].cpu(), ref), (h, w)
    ms = torch.stack([a["m"] for a in golden["data.affine"]])
    out = D.affine_nearest(src[None].expand(len(ms), -1, -1).contiguous(), ms, -1.0)
    for k, a in enumerate(golden["data.affine"]):
        assert torch.equal(out[k].cpu(), a["out"]), (k, a["angle"])                    # bit-exact against PIL's own resampling


def test_random_affine_parameters_follow_torchvision():
    """The host-side draw is torchvision's own (get_params + _get_inverse_affine_matrix): same RNG stream, same matrices."""
    import torchvision.transforms as T
    import torchvision.transforms.functional as TF
    from ctagan import data as D
    level, H, W = 2, 96, 128
    torch.manual_seed(5)
    mine = D.random_affine_matrices(3, (H, W), level)
    torch.manual_seed(5)
    for k in range(3):
        angle, tr, sc, sh = T.RandomAffine.get_params([-level, level], [0.02 * level] * 2, [1 - 0.02 * level, 1 + 0.02 * level], None, [W, H])
        ref = TF._get_inverse_affine_matrix((W * 0.5, H * 0.5), angle, list(tr), sc, list(sh))
        assert torch.allclose(mine[k], torch.tensor(ref, dtype=torch.float64), rtol=0, atol=1e-12), k


def test_eval_metrics_and_int16(golden):
    from ctagan import evaluate as EV
    from oracle import restate_eval as RE
    for name, case in golden["eval.cases"].items():
        f, r = case["fake"], case["real"]
        m = EV.slice_metrics(torch.stack([f, r, f]).cuda(), torch.stack([r, r, f]).cuda(), case["WC"], case["WW"]).cpu()
        ref0 = case["metrics"]
        order = ("MAE_w", "PSNR_w", "SSIM_w", "UQI_w", "MAE_raw", "PSNR_raw", "SSIM_raw", "UQI_raw")
        for j, k in enumerate(order):
            # (the masks are pixel-identical; the reference SUMS its fp32 arrays in fp32, the kernels in fp64: agreement to fp32 round-off)
            assert abs(float(m[0, j]) - ref0[k]) <= 2e-6 * max(1.0, abs(ref0[k])), (name, k, float(m[0, j]), ref0[k])
        for row, (a, b) in ((1, (r, r)), (2, (f, f))):                                 # identical images: PSNR = 100 branch, SSIM = 1
            ref = RE.slice_metrics(a.numpy(), b.numpy(), case["WC"], case["WW"])
            for j in range(8):
                assert abs(float(m[row, j]) - float(ref[j])) <= 2e-6 * max(1.0, abs(float(ref[j]))), (name, row, j)
        assert torch.equal(EV.to_dicom_int16(f.cuda()).cpu(), case["int16"])
    air = torch.full((2, 1, 40, 40), -1.0).cuda()                                     # no valid pixel: the `+1e-10` branches
    m = EV.slice_metrics(air, air).cpu()
    ref = RE.slice_metrics(np.full((40, 40), -1.0, np.float32), np.full((40, 40), -1.0, np.float32))
    for j in range(8):
        assert abs(float(m[0, j]) - float(ref[j])) <= 2e-6 * max(1.0, abs(float(ref[j]))), j


def test_slice_list_loader_and_trainer_test(tmp_path):
    """A slice list of .npy files through the pipeline (reader thread, pinned staging, copy stream, kernels) equals the reference
    arithmetic per slice; and Cyc_Trainer.test() evaluates a saved checkpoint on it with the on-GPU metrics."""
    import random
    from ctagan import data as D
    from oracle import restate_eval as RE
    from test_gpu_steps import _cfg
    from trainer import Cyc_Trainer, Hd_Trainer_x1
    rng = np.random.default_rng(3)
    root = tmp_path / "ST0"
    (root / "SE0").mkdir(parents=True); (root / "SE1").mkdir(parents=True)
    paths = []
    for i in range(5):
        a = rng.integers(-20, 2500, (80, 80)).astype(np.int16)
        np.save(root / "SE0" / f"IM{i}.npy", a)
        np.save(root / "SE1" / f"IM{i}.npy", (a + rng.integers(0, 300, (80, 80))).astype(np.int16))
        paths.append(str(root / "SE0" / f"IM{i}.npy"))
    lst = tmp_path / "list.txt"
    lst.write_text("\n".join(paths) + "\n")
    ld = D.SliceListLoader(str(lst), 2, 64, ("A2", "B1", "B2"), torch.device("cuda"))
    assert len(ld) == 2
    seen = 0
    for bi, batch in enumerate(ld):
        for k in range(2):
            a = np.load(sorted(paths)[bi * 2 + k]); b = np.load(sorted(paths)[bi * 2 + k].replace("SE0", "SE1"))
            rs = lambda x: torch.nn.functional.interpolate(torch.from_numpy(x.astype(np.float32))[None, None], size=[64, 64])[0, 0]
            assert torch.equal(batch["A2"][k, 0].cpu(), rs(RE.read_dicom_norm(a)))
            assert torch.equal(batch["B2"][k, 0].cpu(), rs(RE.read_dicom_norm(b)))
            assert torch.equal(batch["B1"][k, 0].cpu(), rs(RE.window_image(b.astype(np.float64) - 1024, 50, 400)))
        seen += 1
    assert seen == 2
    # trainer.test(): explicit list, checkpoint required
    random.seed(1); torch.manual_seed(1)
    cfg = _cfg("CycleGan", 64, precision="bf16", synthetic=False, test_list=str(lst), save_root=str(tmp_path) + "/", train_list=str(lst))
    tr = Cyc_Trainer(cfg)
    with pytest.raises(FileNotFoundError):
        tr.test()                                                                      # no aa.pth: fail clearly, never evaluate random weights
    torch.save(tr.netG_A2B.state_dict(), tmp_path / "aa.pth")
    out = tr.test()
    assert out["slices"] == 5 and all(np.isfinite(out[k]) for k in ("MAE", "PSNR", "SSIM", "UQI", "MAEw", "PSNRw", "SSIMw", "UQIw"))
    # the Hd stage hand-off: stage 1 writes the _x_ names stage 2 loads
    tr1 = Hd_Trainer_x1(_cfg("HdGan", 256, precision="bf16", save_root=str(tmp_path) + "/", save_checkpoints=True))
    tr1._save(45, tr1.checkpoint_nets())                                               # pinned copies + a writer thread: the call does not wait
    from ctagan.trainers import wait_checkpoints
    wait_checkpoints()
    assert all(os.path.exists(tmp_path / n) for n in ("netG_A2B_x_45.pth", "R_A_x_45.pth", "netD_B_x_45.pth"))
    assert not any(n.endswith(".tmp") for n in os.listdir(tmp_path))
    from trainer import Hd_Trainer_x2
    random.seed(9); torch.manual_seed(9)
    tr2 = Hd_Trainer_x2(_cfg("HdGan", 256, precision="bf16", save_root=str(tmp_path) + "/"))
    assert not torch.equal(tr2.netG_A2B.model_head[1].weight, tr1.netG_A2B.model_head[1].weight)
    assert tr2.load_stage1()                                                           # HdTrainer.py:697-699
    assert torch.equal(tr2.netG_A2B.model_head[1].weight, tr1.netG_A2B.model_head[1].weight)
    assert torch.equal(tr2.R_A.offset_map.c1.conv2d.weight, tr1.R_A.offset_map.c1.conv2d.weight)
    tr3 = Hd_Trainer_x2(_cfg("HdGan", 256, precision="bf16", save_root=str(tmp_path / "nowhere") + "/"))
    with pytest.raises(FileNotFoundError):
        tr3.load_stage1()                                                              # stage 2 never starts from random init silently
    import ctagan
    ctagan.set_precision("bf16")
