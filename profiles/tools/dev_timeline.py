"""Developer probe: CUPTI timeline (torch.profiler) of one graph-replayed Cyc iteration -> per-stream busy time, overlap, gaps."""
import sys, os, random, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cta-gan_b200"))
import torch
import bench
from ctagan import trainers as TR
from ctagan.graphs import GraphedTrainer
import argparse
wl = sys.argv[1] if len(sys.argv) > 1 else "cyc"
args = argparse.Namespace(workload=wl, batch=(1 if wl == "cyc" else 8), size=256, precision="bf16", hd_stage=2)
cfg = bench.workload_config(args)
if os.environ.get("CTAGAN_FUSED_OPT") == "0":
    cfg["fused_optimizer"] = False
random.seed(42); torch.manual_seed(42)
tr = (TR.Cyc_Trainer if wl == "cyc" else TR.Reg_Trainer)(cfg)
loader = TR.SyntheticSlices(cfg["batchSize"], 256, 4, 42, tr.data_keys, pool=2)
dev = [[b[k].cuda() for k in tr.data_keys] for b in loader.batches]
run = GraphedTrainer(tr, enabled=True)
for i in range(5):
    run.step_device(dev[i % 2])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run.step_device(dev[0])
    torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/trace.json")
ev = [e for e in json.load(open("/tmp/trace.json"))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ev)
print(f"events {len(ev)}  span {(t1 - t0) / 1e3:.3f} ms")
by_stream = collections.defaultdict(list)
for e in ev:
    by_stream[e["args"].get("stream")].append(e)
for sid, es in sorted(by_stream.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    busy = sum(e["dur"] for e in es)
    print(f"  stream {sid}: {len(es)} events, busy {busy / 1e3:.3f} ms, first {(es[0]['ts'] - t0) / 1e3:.3f} last {(es[-1]['ts'] + es[-1]['dur'] - t0) / 1e3:.3f}")
# concurrency profile: sweep line
pts = []
for e in ev:
    pts.append((e["ts"], 1)); pts.append((e["ts"] + e["dur"], -1))
pts.sort()
cur = 0; last = pts[0][0]; hist = collections.Counter()
for t, d in pts:
    hist[cur] += t - last; last = t; cur += d
tot = sum(hist.values())
print("concurrency (kernels in flight -> share of span):", {k: f"{100 * v / tot:.1f}%" for k, v in sorted(hist.items())})
# per-kernel-name busy
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    n = e["name"].replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:70]
    agg[n][0] += 1; agg[n][1] += e["dur"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"  {v[1] / 1e3:7.3f} ms  n={v[0]:4d} avg {v[1] / v[0]:6.1f} us  {k}")
# timeline in 0.5 ms buckets: busy per stream
nb = int((t1 - t0) / 500) + 1
print("bucket(0.5ms) -> busy us per stream")
sids = sorted(by_stream, key=lambda s: -sum(e["dur"] for e in by_stream[s]))[:5]
for b in range(nb):
    row = []
    for sid in sids:
        lo, hi = t0 + b * 500, t0 + (b + 1) * 500
        row.append(sum(max(0, min(e["ts"] + e["dur"], hi) - max(e["ts"], lo)) for e in by_stream[sid]))
    print(f"  {b * 0.5:4.1f} ms: " + "  ".join(f"{r:5.0f}" for r in row))

# per-stream kernel sequence summary: consecutive runs of the same kernel family
def fam(e):
    return e["name"].replace("(anonymous namespace)::", "").replace("void ", "").split("<")[0].split("(")[0][:28]
for sid in sids[:6]:
    es = by_stream[sid]
    print(f"--- stream {sid}")
    # print a coarse trace: every 12th event with time and name
    for e in es[::max(1, len(es) // 28)]:
        print(f"   {(e['ts'] - t0) / 1e3:7.3f} ms  {e['dur']:6.1f} us  {fam(e)}")

# full event list for offline inspection
os.makedirs("gpurun_out", exist_ok=True)
with open(f"gpurun_out/timeline_{wl}_events.csv", "w") as f:
    f.write("start_us,dur_us,stream,name\n")
    for e in ev:
        f.write(f"{e['ts'] - t0:.1f},{e['dur']:.1f},{e['args'].get('stream')},{fam(e)}\n")
