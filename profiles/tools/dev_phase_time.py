import sys, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cta-gan_b200"))
import torch
import bench
from ctagan import trainers as TR
from ctagan.graphs import GraphedTrainer
import argparse
args = argparse.Namespace(workload="cyc", batch=None, size=256, precision="bf16")
cfg = bench.workload_config(args)
random.seed(42); torch.manual_seed(42)
tr = TR.Cyc_Trainer(cfg)
loader = TR.SyntheticSlices(1, 256, 4, 42, tr.data_keys, pool=2)
dev = [[b[k].cuda() for k in tr.data_keys] for b in loader.batches]
run = GraphedTrainer(tr, enabled=True)
for i in range(4):
    run.step_device(dev[i % 2])
torch.cuda.synchronize()
gG, gD = run._graphs
def t(g, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(f"G-phase graph {t(gG):.2f} ms, D-phases graph {t(gD):.2f} ms")
# single generator forward / forward+backward, eager-captured in a graph
import ctagan
x = dev[0][0].clone().requires_grad_(True)
net = tr.netG_A2B
def fwd_only():
    with torch.no_grad():
        return net(x)
def fwd_bwd():
    for p in net.parameters():
        p.grad = None
    y = net(x)
    ctagan.l1_loss(y, dev[0][1]).backward()
for name, fn in (("G fwd", fwd_only), ("G fwd+bwd", fwd_bwd)):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    print(f"{name}: {t(g):.3f} ms")

# --- concurrency probe: two independent generator passes on two streams inside one graph ---
netB = tr.netG_B2A
xa, xb = dev[0][0].clone(), dev[0][1].clone()
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
def two_fwd(parallel):
    cur = torch.cuda.current_stream()
    net.prepack(); netB.prepack()
    if parallel:
        sA.wait_stream(cur); sB.wait_stream(cur)
        with torch.cuda.stream(sA), torch.no_grad():
            ya = net(xa)
        with torch.cuda.stream(sB), torch.no_grad():
            yb = netB(xb)
        cur.wait_stream(sA); cur.wait_stream(sB)
    else:
        with torch.no_grad():
            ya = net(xa); yb = netB(xb)
    return ya, yb
def two_fwd_bwd(parallel):
    cur = torch.cuda.current_stream()
    for p in list(net.parameters()) + list(netB.parameters()):
        p.grad = None
    net.prepack(); netB.prepack()
    if parallel:
        sA.wait_stream(cur); sB.wait_stream(cur)
        with torch.cuda.stream(sA):
            la = ctagan.l1_loss(net(xa), xb)
        with torch.cuda.stream(sB):
            lb = ctagan.l1_loss(netB(xb), xa)
        cur.wait_stream(sA); cur.wait_stream(sB)
    else:
        la = ctagan.l1_loss(net(xa), xb); lb = ctagan.l1_loss(netB(xb), xa)
    (la + lb).backward()
for name, fn in (("2x G fwd serial", lambda: two_fwd(False)), ("2x G fwd parallel", lambda: two_fwd(True)),
                 ("2x G fwd+bwd serial", lambda: two_fwd_bwd(False)), ("2x G fwd+bwd parallel", lambda: two_fwd_bwd(True))):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    print(f"{name}: {t(g):.3f} ms")
