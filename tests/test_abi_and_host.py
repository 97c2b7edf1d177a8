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
