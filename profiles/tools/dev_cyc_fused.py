import sys, os, random
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import _ctagan_path
import torch
import test_gpu_steps as T
golden = torch.load("tests/golden/golden_v1.pt", weights_only=False)
if os.environ.get("PRE", "1") == "1":
    T.test_cyc_step_matches_reference_losses(golden)
    T.test_reg_step_matches_oracle()
    T.test_hd_x2_and_p2p_steps_match_oracle()
    T.test_bf16_cyc_step_runs_and_tracks(golden)

def run(mode, steps=3):
    from oracle import restate as R
    from trainer import Cyc_Trainer
    from ctagan.graphs import GraphedTrainer
    from ctagan.replay import ReplayBuffer
    T._seed(); tr = Cyc_Trainer(T._cfg("CycleGan", 64, precision="fp32"))
    tr.fake_A_buffer, tr.fake_B_buffer = ReplayBuffer(2), ReplayBuffer(2)
    batches = [R.synthetic_pair(1, 64, seed=500 + i, phantom=True) for i in range(steps)]
    random.seed(7)
    nets = {"GA": tr.netG_A2B, "GB": tr.netG_B2A, "DA": tr.netD_A, "DB": tr.netD_B}
    def sig():
        return {k: round(float(sum(p.double().abs().sum() for p in n.parameters())), 6) for k, n in nets.items()}
    g = GraphedTrainer(tr, warmup=1) if mode == "graph" else None
    for i in range(steps):
        a, b = batches[0] if i == 1 else batches[i]
        if mode == "graph":
            if i == 0: continue
            l = g.step_device((a.cuda(), b.cuda()))
        else:
            l = (tr.step if mode == "fused" else tr.step_two_phase)({"A": a, "B": b})
        torch.cuda.synchronize()
        print(mode, i, {k: round(float(v), 5) for k, v in l.items()}, sig(), "streams", [s.cuda_stream % 100000 for s in tr._streams] if hasattr(tr, "_streams") else None,
              [s.cuda_stream % 100000 for s in getattr(tr, "_d_streams", [])])
    from ctagan import engine as E
    print("lanes", {k % 100000: v.cuda_stream % 100000 for k, v in E._WGRAD_STREAMS.items()})
for m in ("two_phase", "fused", "graph"):
    run(m)
