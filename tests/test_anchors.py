"""Independent anchors for the parts of the path whose reference source is absent (torchcfm UNet blocks, torchdyn dopri5).

oracle/unet.py, oracle/flow.py and the product (stain2stain_b200/unet.py, neural_ode.py) restate those packages; a shared
misreading would pass every oracle-vs-engine test.  The checks below therefore compare BOTH against third implementations
that share no code with either: `scipy.integrate.solve_ivp(RK45)` (SciPy's own Dormand-Prince pair), the published order
conditions / stability polynomial of DOPRI5, and block formulas written here from the guided-diffusion paper text with
stock `torch.nn.functional` ops only (nothing imported from oracle/ for the expected values).

CPU tests pin the oracle and the product's generic solver; the `gpu` tests pin the engine's blocks.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ dopri5
def _fields():
    lam = torch.tensor([-1.0, -0.3, 0.5])

    def lin(t, x):  # x' = lam * x + sin(3 t): closed form below
        return lam * x + torch.sin(3.0 * t)

    def lin_exact(t, x0):
        l = lam.double()
        # x(t) = e^{lt} x0 + int_0^t e^{l(t-s)} sin(3s) ds
        part = (3.0 * torch.exp(l * t) - l * math.sin(3.0 * t) - 3.0 * math.cos(3.0 * t)) / (l * l + 9.0)
        return torch.exp(l * t) * x0.double() + part

    def vdp(t, x):  # van der Pol, mu = 1 (nonlinear; no closed form: SciPy is the judge)
        return torch.stack([x[..., 1], (1.0 - x[..., 0] ** 2) * x[..., 1] - x[..., 0]], dim=-1)
    return lin, lin_exact, vdp


def _scipy_rk45(f, x0, t_span):
    from scipy.integrate import solve_ivp
    sol = solve_ivp(lambda t, y: f(torch.tensor(t), torch.tensor(y)).numpy(), (float(t_span[0]), float(t_span[-1])),
                    x0.double().numpy(), method="RK45", rtol=1e-4, atol=1e-4, t_eval=t_span.double().numpy())
    assert sol.success
    return torch.from_numpy(sol.y.T.copy())


def _solvers():
    from oracle import flow as oflow
    from stain2stain_b200 import neural_ode as node
    return {"oracle": oflow.odeint, "product": node.odeint}


@pytest.mark.parametrize("which", ["oracle", "product"])
def test_dopri5_matches_scipy_rk45_and_the_closed_form(which):
    odeint = _solvers()[which]
    lin, lin_exact, vdp = _fields()
    t_span = torch.linspace(0.0, 2.0, 9)
    # linear field: both SciPy and the analytic solution
    x0 = torch.tensor([1.0, -2.0, 0.5])
    _, sol = odeint(lin, x0, t_span, solver="dopri5", atol=1e-4, rtol=1e-4)
    want = torch.stack([lin_exact(float(t), x0) for t in t_span])
    assert sol.shape == (9, 3)
    assert float((sol.double() - want).abs().max()) < 2e-3          # the tolerance the controller was given
    assert float((sol.double() - _scipy_rk45(lin, x0, t_span)).abs().max()) < 2e-3
    # nonlinear field: SciPy at the same tolerances, and a tight-tolerance SciPy run as ground truth
    y0 = torch.tensor([2.0, 0.0])
    _, sol = odeint(vdp, y0, t_span, solver="dopri5", atol=1e-4, rtol=1e-4)
    from scipy.integrate import solve_ivp
    truth = solve_ivp(lambda t, y: vdp(torch.tensor(t), torch.tensor(y)).numpy(), (0.0, 2.0), y0.double().numpy(),
                      method="DOP853", rtol=1e-10, atol=1e-12, t_eval=t_span.double().numpy()).y.T
    assert float((sol.double() - torch.from_numpy(truth.copy())).abs().max()) < 3e-3
    assert float((sol.double() - _scipy_rk45(vdp, y0, t_span)).abs().max()) < 3e-3


@pytest.mark.parametrize("which", ["oracle", "product"])
def test_dopri5_tableau_order_conditions_and_stability_polynomial(which):
    """Published facts about the Dormand-Prince 5(4) pair, independent of any implementation: b^T A^k 1 = 1/(k+1)! for
    k = 0..4 (order 5), = 1/600 for k = 5 (its stability polynomial ends in z^6/600); the embedded weights are order 4;
    c_i = sum_j a_ij; FSAL (last row of A == b5)."""
    if which == "oracle":
        from oracle import flow as m
    else:
        from stain2stain_b200 import neural_ode as m
    n = 7
    A = np.zeros((n, n))
    for i, row in enumerate(m._DOPRI_A):
        A[i, :len(row)] = row
    b5, b4, c = np.array(m._DOPRI_B5), np.array(m._DOPRI_B4), np.array(m._DOPRI_C)
    one = np.ones(n)
    assert np.allclose(A @ one, c, atol=1e-15)
    assert np.allclose(A[6, :6], b5[:6], atol=0) and b5[6] == 0.0
    v = one.copy()
    for k in range(6):
        coef = b5 @ v
        assert abs(coef - (1.0 / math.factorial(k + 1) if k < 5 else 1.0 / 600.0)) < 1e-14, (k, coef)
        v = A @ v
    v = one.copy()
    for k in range(4):
        assert abs(b4 @ v - 1.0 / math.factorial(k + 1)) < 1e-14, k
        v = A @ v
    # quadrature conditions sum b_i c_i^q = 1/(q+1)
    for q in range(5):
        assert abs(b5 @ c ** q - 1.0 / (q + 1)) < 1e-14
    for q in range(4):
        assert abs(b4 @ c ** q - 1.0 / (q + 1)) < 1e-14


@pytest.mark.parametrize("which", ["oracle", "product"])
@pytest.mark.parametrize("solver,order", [("euler", 1), ("midpoint", 2), ("rk4", 4)])
def test_fixed_step_solvers_converge_at_their_order(which, solver, order):
    odeint = _solvers()[which]
    lin, lin_exact, _ = _fields()
    x0 = torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    errs = []
    for n in (8, 16):
        t_span = torch.linspace(0.0, 1.0, n + 1, dtype=torch.float64)
        f = lambda t, x: lin(t.double(), x.double())  # noqa: E731
        if which == "product":  # the product casts t_span to fp32; keep the state fp64 so the order is measurable
            _, sol = odeint(f, x0, t_span, solver=solver)
        else:
            _, sol = odeint(f, x0, t_span, solver=solver)
        errs.append(float((sol[-1].double() - lin_exact(1.0, x0)).abs().max()))
    rate = math.log2(errs[0] / errs[1])
    assert abs(rate - order) < 0.35, (solver, errs, rate)


# ------------------------------------------------------------------------------------------------ UNet blocks, written here
def _gn(x, groups, w, b, eps=1e-5):
    """GroupNorm from its definition (no torch.nn.GroupNorm / F.group_norm)."""
    B, C = x.shape[:2]
    xg = x.reshape(B, groups, -1).double()
    mean = xg.mean(dim=2, keepdim=True)
    var = ((xg - mean) ** 2).mean(dim=2, keepdim=True)  # biased, as normalisation layers use
    xn = ((xg - mean) / torch.sqrt(var + eps)).reshape(x.shape)
    shape = [1, C] + [1] * (x.dim() - 2)
    return (xn * w.double().reshape(shape) + b.double().reshape(shape)).to(x.dtype)


def _silu(x):
    return x / (1.0 + torch.exp(-x))


def _resblock_expected(x, emb, p, skip_w=None, skip_b=None):
    """guided-diffusion ResBlock with use_scale_shift_norm=True, eval mode (dropout off)."""
    h = F.conv2d(_silu(_gn(x, 32, p["gn1w"], p["gn1b"])), p["c1w"], p["c1b"], padding=1)
    e = F.linear(_silu(emb), p["ew"], p["eb"])
    co = h.shape[1]
    scale, shift = e[:, :co, None, None], e[:, co:, None, None]
    h = _gn(h, 32, p["gn2w"], p["gn2b"]) * (1.0 + scale) + shift
    h = F.conv2d(_silu(h), p["c2w"], p["c2b"], padding=1)
    sk = x if skip_w is None else F.conv2d(x, skip_w, skip_b)
    return sk + h


def _attention_expected(x, gnw, gnb, qkv_w, qkv_b, proj_w, proj_b, heads, legacy=True):
    """x + proj(softmax(q k^T / sqrt(d)) v) over H*W tokens; `legacy` = head-major channel interleave [h][q|k|v][d]."""
    B, C, H, W = x.shape
    T, d = H * W, C // heads
    a = _gn(x, 32, gnw, gnb).reshape(B, C, T)
    qkv = torch.einsum("oc,bct->bot", qkv_w.reshape(3 * C, C), a) + qkv_b[None, :, None]
    if legacy:
        q, k, v = qkv.reshape(B, heads, 3, d, T).unbind(2)
    else:
        q, k, v = qkv.reshape(B, 3, heads, d, T).unbind(1)
    # stock fused attention op as the third implementation: [B, heads, T, d]
    o = F.scaled_dot_product_attention(q.transpose(2, 3), k.transpose(2, 3), v.transpose(2, 3))
    # and the definition itself
    w = torch.softmax(torch.einsum("bhdt,bhds->bhts", q, k) / math.sqrt(d), dim=-1)
    o2 = torch.einsum("bhts,bhds->bhtd", w, v)
    assert torch.allclose(o, o2, atol=2e-5, rtol=1e-4)
    o = o.transpose(2, 3).reshape(B, C, T)
    out = torch.einsum("oc,bct->bot", proj_w.reshape(C, C), o) + proj_b[None, :, None]
    return x + out.reshape(B, C, H, W)


def _timestep_embedding_expected(t, dim):
    half = dim // 2
    i = torch.arange(half, dtype=torch.float64)
    ang = t.double()[:, None] * (10000.0 ** (-i / half))[None]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=1).float()


def _rb_params(cin, cout, g):
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    return dict(gn1w=1 + 0.1 * r(cin), gn1b=0.1 * r(cin), c1w=r(cout, cin, 3, 3) / math.sqrt(9 * cin), c1b=0.1 * r(cout),
                ew=r(2 * cout, 512) / math.sqrt(512), eb=0.1 * r(2 * cout), gn2w=1 + 0.1 * r(cout), gn2b=0.1 * r(cout),
                c2w=r(cout, cout, 3, 3) / math.sqrt(9 * cout), c2b=0.1 * r(cout))


def _load_resblock(blk, p, skip_w=None, skip_b=None):
    with torch.no_grad():
        blk.in_layers[0].weight.copy_(p["gn1w"]); blk.in_layers[0].bias.copy_(p["gn1b"])
        blk.in_layers[2].weight.copy_(p["c1w"]); blk.in_layers[2].bias.copy_(p["c1b"])
        blk.emb_layers[1].weight.copy_(p["ew"]); blk.emb_layers[1].bias.copy_(p["eb"])
        blk.out_layers[0].weight.copy_(p["gn2w"]); blk.out_layers[0].bias.copy_(p["gn2b"])
        blk.out_layers[3].weight.copy_(p["c2w"]); blk.out_layers[3].bias.copy_(p["c2b"])
        if skip_w is not None:
            blk.skip_connection.weight.copy_(skip_w); blk.skip_connection.bias.copy_(skip_b)


@pytest.mark.parametrize("cin,cout", [(64, 64), (64, 128)])
def test_oracle_resblock_matches_the_written_out_formula(cin, cout):
    from oracle import unet as ounet
    g = torch.Generator().manual_seed(5)
    p = _rb_params(cin, cout, g)
    sw = torch.randn(cout, cin, 1, 1, generator=g) / math.sqrt(cin) if cin != cout else None
    sb = 0.1 * torch.randn(cout, generator=g) if cin != cout else None
    blk = ounet.ResBlock(cin, 512, 0.1, out_channels=cout, use_scale_shift_norm=True).eval()
    _load_resblock(blk, p, sw, sb)
    x = torch.randn(2, cin, 16, 16, generator=g)
    emb = torch.randn(2, 512, generator=g)
    with torch.no_grad():
        got = blk(x, emb)
    want = _resblock_expected(x, emb, p, sw, sb)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-4), float((got - want).abs().max())


@pytest.mark.parametrize("legacy", [True, False])
def test_oracle_attention_matches_the_written_out_formula(legacy):
    from oracle import unet as ounet
    g = torch.Generator().manual_seed(6)
    C, heads = 128, 4
    blk = ounet.AttentionBlock(C, num_head_channels=32, use_new_attention_order=not legacy).eval()
    assert blk.num_heads == heads
    with torch.no_grad():
        for q in blk.parameters():
            q.copy_(torch.randn(q.shape, generator=g) * (0.1 if q.dim() == 1 else 1.0 / math.sqrt(C)))
        blk.norm.weight.add_(1.0)
    x = torch.randn(2, C, 8, 8, generator=g)
    with torch.no_grad():
        got = blk(x)
    want = _attention_expected(x, blk.norm.weight, blk.norm.bias, blk.qkv.weight, blk.qkv.bias, blk.proj_out.weight,
                               blk.proj_out.bias, heads, legacy)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-4), float((got - want).abs().max())


def test_oracle_and_product_timestep_embedding_match_the_formula():
    from oracle import unet as ounet
    from stain2stain_b200 import unet as punet
    t = torch.tensor([0.0, 0.25, 0.9, 1.0])
    want = _timestep_embedding_expected(t, 128)
    assert torch.allclose(ounet.timestep_embedding(t, 128), want, atol=1e-6)
    assert torch.allclose(punet.timestep_embedding(t, 128), want, atol=1e-6)
    assert float(want[1, 0]) == pytest.approx(math.cos(0.25)) and float(want[1, 64]) == pytest.approx(math.sin(0.25))


def test_oracle_resampling_layers_match_stock_ops():
    from oracle import unet as ounet
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 32, 8, 8, generator=g)
    up = ounet.Upsample(32, True).eval()
    down = ounet.Downsample(32, True).eval()
    with torch.no_grad():
        want_up = F.conv2d(x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3), up.conv.weight, up.conv.bias, padding=1)
        assert torch.allclose(up(x), want_up, atol=1e-6)
        assert torch.allclose(down(x), F.conv2d(x, down.op.weight, down.op.bias, stride=2, padding=1), atol=1e-6)


# ------------------------------------------------------------------------------------------------ the engine's blocks (GPU)
def _to_engine(x):
    from stain2stain_b200 import kernels as K
    return K.nchw_to_nhwc16(x.float().contiguous(), K.ACT)


def _from_engine(y):
    from stain2stain_b200 import kernels as K
    return K.nhwc16_to_nchw(y, K.ACT)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout,hw", [(64, 64, 32), (128, 256, 16)])
def test_engine_resblock_matches_the_written_out_formula(cin, cout, hw):
    from stain2stain_b200 import unet as punet
    g = torch.Generator().manual_seed(8)
    p = _rb_params(cin, cout, g)
    sw = torch.randn(cout, cin, 1, 1, generator=g) / math.sqrt(cin) if cin != cout else None
    sb = 0.1 * torch.randn(cout, generator=g) if cin != cout else None
    blk = punet.ResBlock(cin, 512, 0.1, out_channels=cout, use_scale_shift_norm=True).eval()
    _load_resblock(blk, p, sw, sb)
    blk = blk.cuda()
    x = torch.randn(2, cin, hw, hw, generator=g)
    emb = torch.randn(2, 512, generator=g)
    want = _resblock_expected(x, emb, p, sw, sb)
    with torch.no_grad():
        got = _from_engine(blk([_to_engine(x.cuda())], F.silu(emb.cuda()))).cpu()  # the engine takes SiLU(emb) (applied once)
    assert _rel(got, want) <= 5e-3, _rel(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("legacy", [True, False])
def test_engine_attention_matches_the_written_out_formula(legacy):
    from stain2stain_b200 import unet as punet
    g = torch.Generator().manual_seed(9)
    C, heads, hw = 128, 4, 16
    blk = punet.AttentionBlock(C, num_head_channels=32, use_new_attention_order=not legacy).eval()
    with torch.no_grad():
        for q in blk.parameters():
            q.copy_(torch.randn(q.shape, generator=g) * (0.1 if q.dim() == 1 else 1.0 / math.sqrt(C)))
        blk.norm.weight.add_(1.0)
    x = torch.randn(2, C, hw, hw, generator=g)
    want = _attention_expected(x, blk.norm.weight, blk.norm.bias, blk.qkv.weight, blk.qkv.bias, blk.proj_out.weight,
                               blk.proj_out.bias, heads, legacy)
    blk = blk.cuda()
    with torch.no_grad():
        got = _from_engine(blk([_to_engine(x.cuda())])).cpu()
    assert _rel(got, want) <= 5e-3, _rel(got, want)


@pytest.mark.gpu
def test_engine_dopri5_on_device_matches_scipy():
    from stain2stain_b200.neural_ode import NeuralODE
    lin, lin_exact, _ = _fields()
    lam = torch.tensor([-1.0, -0.3, 0.5], device="cuda")
    f = lambda t, x: lam * x + torch.sin(3.0 * t)  # noqa: E731
    x0 = torch.tensor([[1.0, -2.0, 0.5]], device="cuda")
    t_span = torch.linspace(0.0, 2.0, 5, device="cuda")
    sol = NeuralODE(f, solver="dopri5", atol=1e-4, rtol=1e-4).trajectory(x0, t_span)
    want = _scipy_rk45(lin, x0[0].cpu(), t_span.cpu())
    assert float((sol[:, 0].double().cpu() - want).abs().max()) < 2e-3
