"""f3: the one-kernel optimiser step (ctagan.optim.FusedAdam: Adam + both packed bf16 weight layouts, gradients in a flat bucket)."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adam_is_bit_identical_to_torch_adam_and_repacks():
    """Same gradients, several steps, a learning-rate change in between: parameters and both moments equal torch.optim.Adam(fused)
    BIT FOR BIT (the kernel repeats PyTorch's arithmetic operation for operation), and the packed copies it emits equal a fresh
    ctagan_pack_weights of the updated master weights."""
    import ctagan
    from ctagan import ops
    from ctagan.optim import FusedAdam
    import Model.CycleGan as M
    ctagan.set_precision("bf16")
    random.seed(3); torch.manual_seed(3)
    net = M.Discriminator(2).cuda()
    ref = [p.detach().clone().requires_grad_(True) for p in net.parameters()]
    opt = FusedAdam(net.parameters(), 1e-4, [net])
    topt = torch.optim.Adam(ref, lr=torch.tensor(1e-4, device="cuda"), betas=(0.5, 0.999), fused=True, capturable=True)
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(5):
        if step == 3:
            opt.param_groups[0]["lr"].fill_(5e-5)
            topt.param_groups[0]["lr"].fill_(5e-5)
        opt.zero_grad()
        for p, r in zip(net.parameters(), ref):
            grad = torch.randn(p.shape, device="cuda", generator=g) * (10.0 ** random.uniform(-6, 0))
            p.grad.copy_(grad)                      # the bucket slice
            r.grad = grad.clone()
        opt.step(); topt.step()
        for k, (p, r) in enumerate(zip(net.parameters(), ref)):
            assert torch.equal(p.detach(), r.detach()), (step, k, float((p.detach() - r.detach()).abs().max()))
    st = topt.state[ref[0]]
    n0 = ref[0].numel()
    assert torch.equal(opt.exp_avg[:n0].view_as(ref[0]), st["exp_avg"]) and torch.equal(opt.exp_avg_sq[:n0].view_as(ref[0]), st["exp_avg_sq"])
    assert float(opt.step_count) == 5.0
    for prim in net._get_plan().prims():
        for mode in (0, 1):
            cached = prim._cache[(mode, torch.bfloat16)]
            assert cached[0] == prim._version_key()                               # marked fresh: no separate re-pack launch follows
            assert torch.equal(cached[1], ops.pack_weights(prim.w.detach(), mode, torch.bfloat16))


def test_gradient_bucket_overwrites_then_accumulates():
    """Two uses of one generator inside one backward pass (the cycle pass): the first weight-gradient launch of a layer overwrites its
    slice of the bucket, the second accumulates -- equal to autograd's accumulation of two separate passes, without zero_grad."""
    import ctagan
    from ctagan.optim import FusedAdam
    import Model.CycleGan as M
    ctagan.set_precision("fp32")
    random.seed(4); torch.manual_seed(4)
    net = M.Generator(1, 1, n_residual_blocks=2).cuda()
    twin = M.Generator(1, 1, n_residual_blocks=2).cuda()
    twin.load_state_dict(net.state_dict())
    x1 = torch.rand(1, 1, 32, 32, device="cuda") * 2 - 1
    x2 = torch.rand(1, 1, 32, 32, device="cuda") * 2 - 1
    (ctagan.l1_loss(twin(x1), x2) + ctagan.l1_loss(twin(x2), x1)).backward()            # plain autograd accumulation
    opt = FusedAdam(net.parameters(), 1e-4, [net])
    for _ in range(2):                                                                  # twice: stale values from the first round must not leak
        opt.zero_grad()
        (ctagan.l1_loss(net(x1), x2) + ctagan.l1_loss(net(x2), x1)).backward()
        torch.cuda.synchronize()
        for (k, p), q in zip(net.named_parameters(), twin.parameters()):
            if q.grad is None:
                assert float(p.grad.abs().max()) == 0.0, k                            # dead bias: its slice stays zero
            else:
                assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-8), (k, float((p.grad - q.grad).abs().max()))
    ctagan.set_precision("bf16")
