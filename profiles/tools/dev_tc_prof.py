import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import _ctagan_path  # noqa
import torch
from ctagan import engine as E
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
os.environ["CTAGAN_TC_PROF"] = str(buf.data_ptr())
x = torch.randn(1, 66, 66, 256, device="cuda").bfloat16()
w = torch.randn(256, 256, 3, 3, device="cuda") * 0.02
prim = E.ConvPrim(w, None, 1, 0)
for _ in range(3):
    prim.fprop(x, use_bias=False)
torch.cuda.synchronize()
b = buf.cpu().tolist()
t0 = b[0]
names = ["start", "setup done", "producer issued all", "mma: last full wait done", "mma: final commit issued", "epi: tmem_full seen", "epi: done", "kernel end"]
for i, n in enumerate(names):
    print(f"{n:28s} {b[i] - t0:8d} cyc")
print("producer issue stamps (it 0..7):", [v - t0 for v in b[16:24]])
print("mma before-wait stamps     :", [v - t0 for v in b[32:40]])
print("mma full-wait stamps (it 0..7):", [v - t0 for v in b[24:32]])
print("mma after-commit stamps    :", [v - t0 for v in b[40:48]])
