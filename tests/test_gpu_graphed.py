"""The CUDA-graphed training step (stain2stain_b200/graphed.py) against the eager step of the same engine and against the
fp32 oracle + torch.optim.Adam: same inputs, same t, k steps.  Also: the capture leaves the model untouched, dropout masks
and Adam's bias corrections advance across replays, and eager code after replays sees the updated weights."""
import copy
import functools

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"

SMALL = dict(dim=[3, 64, 64], num_channels=64, num_res_blocks=1, attention_resolutions="16,8", dropout=0.0,
             use_scale_shift_norm=True, num_heads=4, num_head_channels=32, channel_mult=[1, 2, 2, 4])


def _lit(cfg, lr=5e-4, wd=1e-5):
    from oracle import unet as ounet
    from stain2stain_b200.flow_matching import ConditionalFlowMatcher
    from stain2stain_b200.lit import ConditionalFlowMatchingLitModule
    from stain2stain_b200.optim import FusedAdam
    from stain2stain_b200.unet import UNetModel
    torch.manual_seed(0)
    ref = ounet.UNetModel(**cfg)
    ounet.dezero_(ref, seed=1984)
    net = UNetModel(**cfg)
    net.load_state_dict(ref.state_dict(), strict=True)
    lit = ConditionalFlowMatchingLitModule(net=net.to(DEV), flow_matcher=ConditionalFlowMatcher(0.0),
                                           optimizer=functools.partial(FusedAdam, lr=lr, weight_decay=wd))
    return ref.to(DEV), lit


def _data(n, B=4, H=64, seed=3):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return [(torch.rand(B, 3, H, H, device=DEV, generator=g) * 2 - 1, torch.rand(B, 3, H, H, device=DEV, generator=g) * 2 - 1,
             torch.rand(B, device=DEV, generator=g)) for _ in range(n)]


def test_graphed_step_equals_eager_and_oracle_trajectory():
    from oracle import flow as oflow
    from stain2stain_b200.graphed import GraphedTrainStep
    ref, lit_e = _lit(SMALL)
    _, lit_g = _lit(SMALL)
    lit_e.train(), lit_g.train(), ref.train()
    opt_e = lit_e.configure_optimizers()["optimizer"]
    opt_g = lit_g.configure_optimizers()["optimizer"]
    opt_r = torch.optim.Adam(ref.parameters(), lr=5e-4, weight_decay=1e-5)
    w0 = [p.detach().clone() for p in lit_g.parameters()]
    gs = GraphedTrainStep(lit_g, opt_g, (4, 3, 64, 64), DEV)
    for p, w in zip(lit_g.parameters(), w0):
        assert torch.equal(p, w), "capture / warm-up changed the model"
    assert all(float(st["step"]) == 0.0 and float(st["exp_avg"].abs().max()) == 0.0 for st in opt_g.state.values())
    losses_e, losses_g, losses_r = [], [], []
    for x0, x1, t in _data(5):
        opt_e.zero_grad(set_to_none=True)
        le = lit_e.model_step((x0, x1), t=t)
        le.backward()
        opt_e.step()
        losses_e.append(float(le))
        losses_g.append(float(gs(x0, x1, t=t)))
        opt_r.zero_grad(set_to_none=True)
        lr_ = oflow.model_step(ref, oflow.ConditionalFlowMatcher(0.0), (x0, x1), t=t)
        lr_.backward()
        opt_r.step()
        losses_r.append(float(lr_))
    for a, b, c in zip(losses_g, losses_e, losses_r):
        assert abs(a - b) <= 2e-3 * abs(b), (losses_g, losses_e)      # same kernels; only split-K atomics reorder sums
        assert abs(a - c) <= 3e-2 * abs(c), (losses_g, losses_r)      # the oracle's trajectory
    num = den = 0.0
    for p, q, w in zip(lit_g.parameters(), lit_e.parameters(), w0):
        num += float((p - q).double().pow(2).sum())
        den += float((q - w).double().pow(2).sum())
    assert (num / den) ** 0.5 <= 5e-2, (num / den) ** 0.5  # accumulated movement agrees with the eager engine
    sd_g, sd_e = opt_g.state_dict(), opt_e.state_dict()
    assert all(float(sd_g["state"][k]["step"]) == float(sd_e["state"][k]["step"]) == 5.0 for k in sd_g["state"])
    # eager code after replays sees the updated weights (packed operands are re-packed, not one step old)
    x0, x1, t = _data(1, seed=9)[0]
    lit_g.eval(), lit_e.eval()
    with torch.no_grad():
        a, b = lit_g.net(t, x0), lit_e.net(t, x0)
    assert float((a - b).norm() / b.norm()) <= 2e-2


def test_graphed_step_advances_dropout_and_bias_correction():
    from stain2stain_b200.graphed import GraphedTrainStep
    cfg = dict(SMALL, dropout=0.1)
    _, lit = _lit(cfg, lr=0.0, wd=0.0)  # lr = 0: the weights stay put, so any change of the loss is the dropout mask
    lit.train()
    opt = lit.configure_optimizers()["optimizer"]
    gs = GraphedTrainStep(lit, opt, (4, 3, 64, 64), DEV)
    x0, x1, t = _data(1)[0]
    l1 = float(gs(x0, x1, t=t))
    l2 = float(gs(x0, x1, t=t))
    l3 = float(gs(x0, x1, t=t))
    assert len({l1, l2, l3}) == 3, (l1, l2, l3)          # a frozen seed would repeat the mask and the loss exactly
    assert max(l1, l2, l3) / min(l1, l2, l3) < 1.2       # ... and they are the same loss up to mask noise
    # the moments follow torch's recursion with a LIVE step count: after 3 steps exp_avg_sq's bias correction is 1 - b2^3
    st = next(iter(opt.state.values()))
    assert float(st["step"]) == 3.0
    # t drawn like torchcfm when not given: CPU default generator
    torch.manual_seed(123)
    want = torch.rand(4)
    torch.manual_seed(123)
    gs(x0, x1)
    assert torch.equal(gs.t.cpu(), want)


def test_graphed_adam_matches_torch_adam_over_many_steps():
    """Only the optimizer part: gradients are forced to known values through the flat buffer; 12 device-counted steps of
    s2s_adam_multi_step against torch.optim.Adam (bias corrections must come from the live device counter)."""
    from stain2stain_b200.optim import FusedAdam
    g = torch.Generator(device=DEV).manual_seed(5)
    pa = [torch.randn(s, device=DEV, generator=g).requires_grad_() for s in [(64, 32, 3, 3), (513,), (7, 9)]]
    pb = [p.detach().clone().requires_grad_() for p in pa]
    oa, ob = FusedAdam(pa, lr=1e-2, weight_decay=1e-4), torch.optim.Adam(pb, lr=1e-2, weight_decay=1e-4)
    step_dev = torch.zeros((), dtype=torch.int64, device=DEV)
    grads = {p: torch.zeros_like(p) for p in pa}
    graph = torch.cuda.CUDAGraph()
    oa.step_on_device_counter(step_dev.add_(1), grads)  # eager warm-up builds the pointer table
    with torch.no_grad():
        for p, q in zip(pa, pb):
            p.copy_(q)
    for st in oa.state.values():
        st["exp_avg"].zero_(), st["exp_avg_sq"].zero_()
    step_dev.zero_()
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        step_dev.add_(1)
        oa.step_on_device_counter(step_dev, grads)
    for it in range(12):
        for p, q in zip(pa, pb):
            gr = torch.randn(p.shape, device=DEV, generator=g)
            grads[p].copy_(gr)
            q.grad = gr.clone()
        graph.replay()
        ob.step()
    assert int(step_dev) == 12
    for p, q in zip(pa, pb):
        assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), float((p - q).abs().max())
