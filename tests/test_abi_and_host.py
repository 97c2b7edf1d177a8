"""CPU-only checks: the C-ABI library loads and exports every symbol include/ctagan.h declares (no compute calls), the host-side
logic mirrors the reference (ReplayBuffer, state_dict layouts, config keys), and the product path refuses to run without CUDA."""
import os
import random
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ctagan import lib
    protos = lib.parse_header()
    declared = set(re.findall(r"\b(ctagan_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", open(lib.HEADER_PATH).read(), flags=re.S)))
    declared -= {"ctagan_conv_geom"}
    assert declared == set(protos), declared ^ set(protos)
    assert len(protos) >= 30
    handle = lib.load()                       # raises if the .so is missing or a symbol does not resolve
    for name in protos:
        assert hasattr(handle, name), name
    assert handle.ctagan_version() >= 100
    assert lib.ConvGeom._fields_[-1][0] == "gy_margin" and len(lib.ConvGeom._fields_) == 16


def test_product_path_has_no_cpu_fallback():
    import ctagan
    from ctagan import ops
    net = ctagan.Generator(1, 1, n_residual_blocks=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 1, 16, 16))
    with pytest.raises(RuntimeError):
        ops.pack_weights(torch.zeros(4, 4, 3, 3), 0, torch.float32)
    from ctagan.trainers import Cyc_Trainer
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            Cyc_Trainer({"batchSize": 1})


def test_product_never_imports_oracle():
    bad = []
    for base in ("cta-gan_b200", "Model", "trainer"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh")):
                    src = open(os.path.join(dp, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
    assert not re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(ROOT, "train.py")).read(), flags=re.M)


def test_state_dict_layouts_match_reference(golden):
    import Model.CycleGan as M
    import Model.HdGan as H
    from trainer.reg import Reg
    random.seed(42); torch.manual_seed(42)
    for name, net in (("generator", M.Generator(1, 1)),):
        fp = golden[name + ".state_fp"]
        sd = net.state_dict()
        assert list(sd) == list(fp)
        for k, v in sd.items():
            assert tuple(v.shape) == fp[k][0] and abs(float(v.double().sum()) - fp[k][1]) <= 1e-9 * max(1.0, fp[k][2]), k
    torch.manual_seed(42); d = M.Discriminator(2)
    assert list(d.state_dict()) == list(golden["discriminator2.state_fp"])
    torch.manual_seed(42); dm = H.Discriminator_m(1)
    fp = golden["discriminator_m.state_fp"]
    assert list(dm.state_dict()) == list(fp)
    for k, v in dm.state_dict().items():
        assert abs(float(v.double().sum()) - fp[k][1]) <= 1e-9 * max(1.0, fp[k][2]), k
    torch.manual_seed(42); r = Reg(256, 256, 1, 1)
    fp = golden["reg.state_fp"]
    assert list(r.state_dict()) == list(fp)
    for k, v in r.state_dict().items():
        assert abs(float(v.double().sum()) - fp[k][1]) <= 1e-9 * max(1.0, fp[k][2]), k


def test_header_is_plain_c(tmp_path):
    """include/ctagan.h is the drop-in boundary: it must compile as C99 (no C++ in the signatures) and its structs must have the
    layout the ctypes binding assumes."""
    import ctypes
    import shutil
    import subprocess
    from ctagan import lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text('#include "ctagan.h"\n#include <stdio.h>\nint main(void){ printf("%zu %zu %zu\\n", sizeof(ctagan_conv_geom), '
                   'sizeof(ctagan_conv_groups), sizeof(ctagan_pack_item)); return 0; }\n')
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert sizes == [ctypes.sizeof(lib.ConvGeom), ctypes.sizeof(lib.ConvGroups), ctypes.sizeof(lib.PackItem)]


def test_replay_buffer_matches_reference(golden):
    from trainer.utils import ReplayBuffer
    random.seed(5); rb = ReplayBuffer(max_size=4)
    picks = [rb.push_and_pop(torch.full((1, 1, 2, 2), float(i))).flatten()[0].item() for i in range(24)]
    assert picks == golden["replay.picks_seed5_size4"]


def test_replay_buffer_plan_apply_equals_oracle_for_batches():
    """The split ReplayBuffer (host decisions up front, index-tensor data movement) against the oracle's restatement of
    trainer/utils.py:120-140 for batches > 1, where a later element of a call can receive an earlier one back."""
    from oracle import restate as R
    from trainer.utils import ReplayBuffer
    g = torch.Generator().manual_seed(1)
    batches = [torch.randn(3, 1, 4, 4, generator=g) for _ in range(12)]
    random.seed(11); ref = R.ReplayBuffer(max_size=4)
    want = [ref.push_and_pop(b) for b in batches]
    random.seed(11); rb = ReplayBuffer(max_size=4)
    got = []
    for b in batches:
        src, dst = rb.plan(b.shape[0])                              # (what the graphed trainer does: decisions first, data later)
        got.append(rb.apply(b, torch.tensor(src), torch.tensor(dst)))
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    assert len(rb.data) == 4 and all(torch.equal(x, y) for x, y in zip(rb.data, ref.data))


def test_yaml_configs_keep_reference_keys():
    from trainer.utils import get_config
    ref_keys = {"CycleGan": ["name", "noise_level", "port", "save_root", "image_save", "Adv_lamda", "Cyc_lamda", "Corr_lamda", "Smooth_lamda",
                             "epoch", "n_epochs", "batchSize", "train_list", "val_list", "test_list", "lr", "decay_epoch", "size", "input_nc",
                             "output_nc", "cuda", "n_cpu"],
                "HdGan": ["Adv_lamda1", "Adv_lamda2", "Corr_lamda1", "Corr_lamda2", "Smooth_lamda", "lrd", "lr", "size", "batchSize"],
                "P2p": ["Adv_lamda", "P2P_lamda", "lr", "size"]}
    for name, keys in ref_keys.items():
        cfg = get_config(os.path.join(ROOT, "Yaml", name + ".yaml"))
        assert cfg["name"] == name
        for k in keys:
            assert k in cfg, (name, k)
    assert get_config(os.path.join(ROOT, "Yaml", "CycleGan.yaml"))["Cyc_lamda"] == 10


def test_trainer_package_exports():
    import trainer
    for n in ("Cyc_Trainer", "P2p_Trainer", "Reg_Trainer", "Hd_Trainer_x", "Hd_Trainer_x1", "Hd_Trainer_x2"):
        assert hasattr(trainer, n), n
    from trainer.transformer import Transformer_2D  # noqa: F401
    from trainer.utils import smooothing_loss  # noqa: F401


def test_packed_weight_cache_is_invalidated_per_optimizer_and_keeps_its_buffers():
    """Host bookkeeping of the packed (bf16 / re-laid-out) weight copies, no kernel involved: an optimizer step marks exactly its own
    parameters stale (torch's fused Adam does not bump tensor versions, so a global post-step hook does it by storage address), in-place
    edits do it through the version counter, `force` re-packs regardless, and the destination buffers persist (CUDA graphs and
    grouped launches alias them)."""
    from ctagan import engine as E
    g = torch.Generator().manual_seed(0)
    w1 = torch.nn.Parameter(torch.randn(8, 4, 3, 3, generator=g))
    w2 = torch.nn.Parameter(torch.randn(8, 4, 3, 3, generator=g))
    p1, p2 = E.ConvPrim(w1, None, 1, 1), E.ConvPrim(w2, None, 1, 1)
    first = p1.stale_entries(torch.float32) + p2.stale_entries(torch.float32)
    assert len(first) == 4                                                   # two layouts per weight
    assert p1.stale_entries(torch.float32) == [] and p2.stale_entries(torch.float32) == []
    bufs = {id(p): [p._cache[(m, torch.float32)][1].data_ptr() for m in (0, 1)] for p in (p1, p2)}
    opt1 = torch.optim.SGD([w1], lr=0.1)
    w1.grad = torch.ones_like(w1)
    opt1.step()
    assert len(p1.stale_entries(torch.float32)) == 2 and p2.stale_entries(torch.float32) == []
    with torch.no_grad():
        w2.add_(1.0)
    assert len(p2.stale_entries(torch.float32)) == 2
    assert len(p1.stale_entries(torch.float32, force=True)) == 2
    E.invalidate_weight_cache()
    assert len(p1.stale_entries(torch.float32)) == 2 and len(p2.stale_entries(torch.float32)) == 2
    assert bufs == {id(p): [p._cache[(m, torch.float32)][1].data_ptr() for m in (0, 1)] for p in (p1, p2)}

    # grouped launches: both orders of a pair share ONE store; slots follow the members, slices are the members' own cache buffers
    ga, gb = E.GroupedPrim([p1, p2]), E.GroupedPrim([p2, p1])
    assert ga.store is gb.store and ga.slots == list(reversed(gb.slots)) and sorted(ga.slots) == [0, 1]
    buf = ga.store.buffer(0, torch.float32)
    assert tuple(buf.shape) == (2, 8, 3, 3, 4)
    assert p1._cache[(0, torch.float32)][1].data_ptr() == buf[ga.slots[0]].data_ptr()
    assert p2._cache[(0, torch.float32)][1].data_ptr() == buf[ga.slots[1]].data_ptr()
    assert len(p1.stale_entries(torch.float32, modes=(0,))) == 1             # re-homed copies are stale until re-packed


def test_checkpoints_are_written_by_a_background_thread(tmp_path):
    """f4: _save hands the state_dicts to a writer thread (atomic rename into the reference's file names) and returns; load_checkpoint /
    wait_checkpoints join it.  Host logic only (CPU tensors take the same path minus the pinned copies)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cta-gan_b200"))
    import torch
    from ctagan import trainers as TR

    class Stub:
        rank = 0
        device = torch.device("cpu")
        config = {"save_checkpoints": True, "save_root": str(tmp_path) + "/"}
        _save = TR._TrainerBase._save
        load_checkpoint = TR._TrainerBase.load_checkpoint

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 3), torch.nn.Conv2d(4, 1, 3))
    st = Stub()
    st._save(7, {"netG_A2B_{st}.pth": net, "R_A_x_{st}.pth": net})
    TR.wait_checkpoints()
    assert sorted(os.listdir(tmp_path)) == ["R_A_x_7.pth", "netG_A2B_7.pth"]
    sd = torch.load(tmp_path / "netG_A2B_7.pth", weights_only=True)
    assert list(sd) == list(net.state_dict()) and all(torch.equal(sd[k], v) for k, v in net.state_dict().items())
    other = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 3), torch.nn.Conv2d(4, 1, 3))
    st._save(8, {"netG_A2B_{st}.pth": net})
    assert st.load_checkpoint(other, "netG_A2B_8.pth")            # joins the writer first
    assert all(torch.equal(a, b) for a, b in zip(other.state_dict().values(), net.state_dict().values()))
    st.rank = 1
    st._save(9, {"netG_A2B_{st}.pth": net})                       # only rank 0 writes
    TR.wait_checkpoints()
    assert "netG_A2B_9.pth" not in os.listdir(tmp_path)


def test_loss_log_is_deferred_not_synchronous(capsys):
    """f4: _log copies the losses of a log step and prints them when the copy has landed (immediately for CPU tensors, at the next call or
    at the end of train() for device tensors) -- in order, once each."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cta-gan_b200"))
    import torch
    from ctagan import trainers as TR

    class Stub:
        rank, name = 0, "Stub"
        config = {"log_every": 2}
        step_count = 0
        last_losses = {}
        _log = TR._TrainerBase._log
        _flush_log = TR._TrainerBase._flush_log

    st = Stub()
    for step in range(1, 7):
        st.step_count = step
        st.last_losses = {"loss_G": torch.tensor(float(step)), "loss_D": torch.tensor(0.5)}
        st._log(1, step - 1, 6)
    st._flush_log(block=True)
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("[Stub]")]
    assert [l.split("loss_G: ")[1].split(" ")[0] for l in lines] == ["2.0000", "4.0000", "6.0000"]
    assert all("loss_D: 0.5000" in l for l in lines)
