"""CPU tests of the oracle (the fp32 restatement of torchcfm's UNet / matcher and torchdyn's solver).

The reference holds no golden vector on this path ("parity unpinned"), so the oracle is pinned by
the known-answer checks SURVEY.md 8(c) lists: parameter counts (70 954 883 / 70 956 419 for configs A / B,
35 746 307 for torchcfm's published 35.75 M CIFAR-10 model), zero output at initialisation, the state_dict key
scheme the reference's checkpoints use, and closed-form properties of the matcher and the solvers.
"""
import math

import pytest
import torch

from oracle import flow as oflow
from oracle import unet as ounet

TINY = dict(dim=[3, 32, 32], num_channels=32, num_res_blocks=1, attention_resolutions="16,8", dropout=0.0,
            use_scale_shift_norm=True, num_heads=4, num_head_channels=16, channel_mult=[1, 2, 2, 4])


def _nparams(m):
    return sum(p.numel() for p in m.parameters())


def test_param_counts_pin_the_constructor_rule():
    with torch.device("meta"):
        assert _nparams(ounet.UNetModel(**ounet.CONFIG_A)) == 70_954_883
        assert _nparams(ounet.UNetModel(**ounet.CONFIG_B)) == 70_956_419
        cifar = ounet.UNetModel(dim=(3, 32, 32), num_channels=128, num_res_blocks=2, channel_mult=[1, 2, 2, 2],
                                attention_resolutions="16", num_heads=4, num_head_channels=64, dropout=0.1)
        assert _nparams(cifar) == 35_746_307  # torchcfm's published "35.75 M": only with ds = image_size // res


def test_config_a_structure_and_state_dict_keys():
    with torch.device("meta"):
        net = ounet.UNetModel(**ounet.CONFIG_A)
    keys = list(net.state_dict().keys())
    assert len(net.input_blocks) == 12 and len(net.output_blocks) == 12
    n_attn = sum(isinstance(m, ounet.AttentionBlock) for m in net.modules())
    assert n_attn == 1  # SURVEY finding 6: only the middle block attends with the shipped yaml
    assert net.middle_block[1].num_heads == 16
    for k in ("time_embed.0.weight", "input_blocks.0.0.weight", "input_blocks.1.0.in_layers.2.weight",
              "input_blocks.1.0.emb_layers.1.weight", "input_blocks.1.0.out_layers.3.bias",
              "input_blocks.3.0.op.weight", "input_blocks.4.0.skip_connection.weight", "middle_block.1.qkv.weight",
              "middle_block.1.proj_out.bias", "output_blocks.2.1.conv.weight", "out.0.weight", "out.2.bias"):
        assert k in keys, k
    sd = net.state_dict()
    assert tuple(sd["middle_block.1.qkv.weight"].shape) == (1536, 512, 1)
    assert tuple(sd["input_blocks.4.0.emb_layers.1.weight"].shape) == (512, 512)
    assert tuple(sd["output_blocks.0.0.in_layers.2.weight"].shape) == (512, 1024, 3, 3)
    assert not any("running" in k for k in keys)


def test_zero_output_at_init_and_dezero():
    torch.manual_seed(0)
    net = ounet.UNetModel(**TINY).eval()
    x = torch.randn(2, 3, 32, 32)
    t = torch.rand(2)
    with torch.no_grad():
        assert float(net(t, x).abs().max()) == 0.0  # SURVEY finding 7
        ounet.dezero_(net)
        v = net(t, x)
    assert float(v.abs().mean()) > 1e-3
    # t conventions of the call sites: [B], 0-dim (ODE solver), [B,1,1,1]
    with torch.no_grad():
        assert torch.allclose(net(t[0], x[:1]), net(t[:1], x[:1]))
        assert torch.allclose(net(t[:, None, None, None], x), v)


def test_class_conditional_requires_y():
    net = ounet.UNetModel(**dict(TINY, class_cond=True, num_classes=3)).eval()
    x, t = torch.randn(2, 3, 32, 32), torch.rand(2)
    with pytest.raises(AssertionError):
        net(t, x)
    assert net(t, x, y=torch.tensor([0, 2])).shape == x.shape
    with pytest.raises(AssertionError):
        ounet.UNetModel(**TINY)(t, x, y=torch.tensor([0, 2]))


def test_timestep_embedding_cos_first():
    e = ounet.timestep_embedding(torch.tensor([0.0, 0.5]), 8)
    assert torch.allclose(e[0], torch.tensor([1.0, 1, 1, 1, 0, 0, 0, 0]))
    f = torch.exp(-math.log(10000.0) * torch.arange(4) / 4)
    assert torch.allclose(e[1], torch.cat([torch.cos(0.5 * f), torch.sin(0.5 * f)]))


def test_matcher_known_answers():
    fm = oflow.ConditionalFlowMatcher(0.0)
    x0, x1 = torch.randn(3, 3, 8, 8), torch.randn(3, 3, 8, 8)
    t, xt, ut = fm.sample_location_and_conditional_flow(x0, x1, t=torch.tensor([0.0, 1.0, 0.25]))
    assert torch.equal(xt[0], x0[0]) and torch.equal(xt[1], x1[1])
    assert torch.allclose(xt[2], 0.25 * x1[2] + 0.75 * x0[2])
    assert torch.equal(ut, x1 - x0)
    t2, _, _ = fm.sample_location_and_conditional_flow(x0, x1)
    assert t2.shape == (3,) and float(t2.min()) >= 0 and float(t2.max()) < 1
    with pytest.raises(AssertionError):
        fm.sample_location_and_conditional_flow(x0, x1, t=torch.rand(2))


def test_solvers_known_answers():
    x = torch.randn(2, 3, 4, 4)
    c = torch.randn_like(x)
    for solver in ("euler", "midpoint", "rk4", "dopri5"):
        for n in (2, 5):
            traj = oflow.NeuralODE(lambda t, z: c, solver=solver, atol=1e-4, rtol=1e-4).trajectory(
                x, torch.linspace(0, 1, n))
            assert traj.shape == (n, *x.shape)
            assert torch.allclose(traj[-1], x + c, atol=1e-5), solver  # constant field: exact for every solver
    # dx/dt = -x : dopri5 at 1e-4 must hit exp(-1) to ~1e-4; Euler with 50 steps is first-order
    f = lambda t, z: -z  # noqa: E731
    d = oflow.NeuralODE(f, solver="dopri5", atol=1e-4, rtol=1e-4).trajectory(x, torch.linspace(0, 1, 2))[-1]
    assert float((d - x * math.exp(-1)).abs().max()) < 5e-4 * float(x.abs().max())
    e = oflow.NeuralODE(f, solver="euler").trajectory(x, torch.linspace(0, 1, 51))[-1]
    assert torch.allclose(e, x * (1 - 1 / 50) ** 50, atol=1e-5)


def test_model_step_and_generate_follow_the_reference_litmodule():
    torch.manual_seed(0)
    net = ounet.dezero_(ounet.UNetModel(**TINY)).eval()
    x0, x1 = torch.rand(2, 3, 32, 32) * 2 - 1, torch.rand(2, 3, 32, 32) * 2 - 1
    t = torch.tensor([0.3, 0.7])
    loss = oflow.model_step(net, oflow.ConditionalFlowMatcher(0.0), (x0, x1), t=t)
    tt = t[:, None, None, None]
    want = torch.mean((net(t, tt * x1 + (1 - tt) * x0) - (x1 - x0)) ** 2)
    assert torch.allclose(loss, want)
    out = oflow.generate(net, x0[0], num_steps=3, solver="euler")  # (C,H,W) input gets a batch dim
    assert out.shape == (1, 3, 32, 32)
    with torch.no_grad():
        x = x0[:1]
        for k in range(2):
            x = x + 0.5 * net(torch.tensor(0.5 * k), x)
    assert torch.allclose(out, x, atol=1e-5)
    assert oflow.psnr(out, out) == float("inf") and abs(oflow.psnr(out, out + 0.02) - 40.0) < 1e-3
