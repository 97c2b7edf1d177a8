"""Developer probe (not a pytest): tcgen05 conv engines vs the CUDA-core engine + graph-replay kernel timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import _ctagan_path  # noqa
import torch
from ctagan import engine as E, lib as L, ops

torch.manual_seed(0)


def graph_time(fn, reps=20, iters=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters * reps)


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max())


def run(N, H, W, Ci, Co, K):
    """H, W = spatial size of the (padded) conv input; output (H-K+1)."""
    x = torch.randn(N, H, W, Ci, device="cuda").bfloat16()
    w = torch.randn(Co, Ci, K, K, device="cuda") / (Ci * K * K) ** 0.5
    b = torch.randn(Co, device="cuda")
    prim = E.ConvPrim(w, b, 1, 0)
    Ho, Wo = H - K + 1, W - K + 1
    E.set_conv_engine("simt"); ref = prim.fprop(x, act=L.ACT_RELU, use_bias=True)
    E.set_conv_engine("tc"); out = prim.fprop(x, act=L.ACT_RELU, use_bias=True)
    e_f = rel(out, ref)
    us_f = graph_time(lambda: prim.fprop(x, use_bias=False))
    fl = 2.0 * N * Ho * Wo * Ci * Co * K * K
    # input gradient: dy zero-margined by K-1
    m = K - 1
    dy = torch.randn(N, Ho, Wo, Co, device="cuda").bfloat16()
    dyz = torch.zeros(N, Ho + 2 * m, Wo + 2 * m, Co, device="cuda", dtype=torch.bfloat16)
    dyz[:, m:m + Ho, m:m + Wo] = dy
    E.set_conv_engine("simt"); dref = prim.bprop(dy, (H, W))
    E.set_conv_engine("tc"); dx = prim.bprop(dyz, (H, W), pad=m)
    e_d = rel(dx, dref)
    us_d = graph_time(lambda: prim.bprop(dyz, (H, W), pad=m))
    # weight gradient
    E.set_conv_engine("simt"); wref, _ = prim.wgrad(dy, x)
    E.set_conv_engine("tc"); dw, db = prim.wgrad(dyz, x, want_bias=True, pad=m, gy_margin=m)
    e_w = rel(dw, wref)
    e_b = rel(db, dy.float().sum((0, 1, 2)))
    us_w = graph_time(lambda: prim.wgrad(dyz, x, pad=m, gy_margin=m))
    print(f"N={N} in {H}x{W} {Ci}->{Co} k{K}: fprop {e_f:.2e} {us_f:.1f}us {fl/us_f/1e6:.0f}TF | dgrad {e_d:.2e} {us_d:.1f}us {fl/us_d/1e6:.0f}TF"
          f" | wgrad {e_w:.2e} (db {e_b:.1e}) {us_w:.1f}us {fl/us_w/1e6:.0f}TF", flush=True)
    return max(e_f, e_d, e_w)


if __name__ == '__main__':
    errs = []
    errs.append(run(1, 66, 66, 256, 256, 3))
    errs.append(run(8, 66, 66, 256, 256, 3))
    errs.append(run(2, 35, 35, 256, 512, 4))
    errs.append(run(1, 130, 130, 128, 128, 3))
    errs.append(run(4, 34, 34, 128, 128, 3))
    print("worst", max(errs))
    assert max(errs) < 2e-2


    def run_strided(N, H, W, Ci, Co, K, s, p):
        """Conv2d(Ci->Co, K, stride s, zero pad p) on an unpadded H x W input: fprop (4-D boxes), dgrad (output phases), wgrad."""
        x = torch.randn(N, H, W, Ci, device="cuda").bfloat16()
        w = torch.randn(Co, Ci, K, K, device="cuda") / (Ci * K * K) ** 0.5
        prim = E.ConvPrim(w, None, s, p)
        Ho, Wo = (H + 2 * p - K) // s + 1, (W + 2 * p - K) // s + 1
        dy = torch.randn(N, Ho, Wo, Co, device="cuda").bfloat16()
        res = []
        for eng in ("simt", "auto"):
            E.set_conv_engine(eng)
            res.append((prim.fprop(x, use_bias=False), prim.bprop(dy, (H, W)), prim.wgrad(dy, x)[0]))
        e = [rel(a, b) for a, b in zip(res[1], res[0])]
        fl = 2.0 * N * Ho * Wo * Ci * Co * K * K
        us = [graph_time(lambda: prim.fprop(x, use_bias=False)), graph_time(lambda: prim.bprop(dy, (H, W))), graph_time(lambda: prim.wgrad(dy, x))]
        print(f"N={N} {H}x{W} {Ci}->{Co} k{K} s{s} p{p}: fprop {e[0]:.2e} {us[0]:.1f}us {fl/us[0]/1e6:.0f}TF | dgrad {e[1]:.2e} {us[1]:.1f}us "
              f"{fl/us[1]/1e6:.0f}TF | wgrad {e[2]:.2e} {us[2]:.1f}us {fl/us[2]/1e6:.0f}TF", flush=True)
        return max(e)


    errs = [run_strided(1, 256, 256, 64, 128, 3, 2, 1), run_strided(1, 128, 128, 128, 256, 3, 2, 1), run_strided(2, 128, 128, 64, 128, 4, 2, 1),
            run_strided(1, 64, 64, 128, 256, 4, 2, 1), run_strided(2, 32, 32, 256, 512, 4, 1, 1), run_strided(1, 64, 64, 64, 64, 3, 1, 1)]
    print("worst strided", max(errs))
    assert max(errs) < 2e-2
