import torch


def maxrel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def l2rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def grad_report(named_params, leaf, tensor_tol, glob_tol, significant=0.05):
    """Global L2-rel over all live parameters (dead pre-InstanceNorm biases excluded on both sides) and the worst per-tensor L2-rel
    among the tensors that carry at least `significant` of the largest tensor-gradient norm.  (Near-dead tensors -- e.g. the 2x2-pixel
    bottleneck layers of Reg, whose InstanceNorm runs over 4 values -- are pure rounding noise per tensor; they stay in the global
    number, weighted by what they contribute.)"""
    rows = []
    ref_max = max(float(p.grad.abs().max()) for p in leaf.values() if p.grad is not None)
    for k, p in named_params:
        ref = leaf[k].grad
        if ref is None or float(ref.abs().max()) <= 1e-6 * ref_max:
            continue
        if p.grad is None:        # biases in front of a non-affine InstanceNorm get no gradient here: the reference's is round-off noise
            assert float(ref.abs().max()) <= 1e-4 * ref_max, (k, float(ref.abs().max()), ref_max)
            continue
        d = (p.grad.detach().double().cpu() - ref.double())
        rows.append((k, float((d * d).sum()), float((ref.double() ** 2).sum())))
    glob = (sum(r[1] for r in rows) / sum(r[2] for r in rows)) ** 0.5
    big = max(r[2] for r in rows)
    worst = max(((k, (n / d_) ** 0.5) for k, n, d_ in rows if d_ >= significant ** 2 * big), key=lambda t: t[1])
    print(f"[grad report] {len(rows)} tensors, global L2-rel {glob:.3e}, worst significant tensor {worst}", flush=True)
    assert glob <= glob_tol, ("global", glob, worst)
    assert worst[1] <= tensor_tol, ("tensor", worst, glob)
    return glob, worst


def trainer_leaves(pairs):
    """[(module, oracle leaf dict)] -> (named params with unique names, merged leaf dict)"""
    named, leaves = [], {}
    for tag, (mod, leaf) in enumerate(pairs):
        for k, p in mod.named_parameters():
            named.append((f"{tag}:{k}", p))
        for k, v in leaf.items():
            leaves[f"{tag}:{k}"] = v
    return named, leaves
