"""Phase-decomposed Upsample conv (s2s_upconv_*) against the reference composition
`F.interpolate(x, scale_factor=2, mode="nearest")` -> `F.conv2d(3x3, padding=1)` in true fp32, forward, data gradient,
weight gradient, bias gradient and epilogue statistics; inputs pre-rounded to the 16-bit storage formats."""
import math

import pytest
import torch
import torch.nn.functional as F

from test_gpu_kernels import DEV, K, assert_close_act, assert_close_bf16, nchw, nhwc, rb, rel_l2

pytestmark = pytest.mark.gpu

SHAPES = [(2, 32, 32, 128, 128), (2, 16, 16, 256, 256), (1, 32, 24, 512, 512), (3, 64, 64, 256, 128), (1, 20, 12, 128, 256)]


def _ref(x, w, b):
    return F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, b, padding=1)


def _packed(k, w, cin, cout, dgrad):
    if not dgrad:
        wp = torch.zeros((cout, 16 * cin), dtype=k.T16, device=DEV)
        for ph in range(4):
            k.pack_conv_weight(w, wp, k_off=ph * 4 * cin, ci_begin=0, ci_count=cin, fmt=k.ACT, mode=1 + ph)
        return wp
    wd = torch.zeros((cin, 16 * cout), dtype=k.T16, device=DEV)
    for ph in range(4):
        k.pack_conv_weight(w, wd, k_off=ph * 4 * cout, ci_begin=0, ci_count=cin, transpose_flip=True, fmt=k.GRAD, mode=1 + ph)
    return wd


@pytest.mark.parametrize("B,H,W,Cin,Cout", SHAPES)
def test_upconv_forward_and_statistics(B, H, W, Cin, Cout):
    k = K()
    assert k.upconv_supported(Cin, Cout)
    g = torch.Generator(device=DEV).manual_seed(H * 7 + Cin)
    x = rb(torch.randn(B, Cin, H, W, device=DEV, generator=g))
    w = torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / math.sqrt(9 * Cin)
    b = torch.randn(Cout, device=DEV, generator=g) * 0.1
    # the tap sums are formed in fp32 BEFORE rounding: compare against the fp32 weights (the engine's single rounding of
    # each summed tap is no worse than the plain conv's rounding of each tap)
    ref = _ref(x, w, b)
    out, stats = k.upconv_fwd(nhwc(x), _packed(k, w, Cin, Cout, False), Cout, b, want_stats=True)
    assert tuple(out.shape) == (B, 2 * H, 2 * W, Cout)
    got = nchw(out)
    assert_close_act(got, ref, "upconv fwd") if k.ACT != k.FMT_F16 else None
    r = rel_l2(got, ref)
    assert r < 2e-3, f"upconv fwd rel-L2 {r}"
    # every phase / border pixel individually (a wrong tap map shows up at the image border first)
    err = (got - ref).abs()
    assert float(err.max()) <= 2 ** -6 * float(ref.abs().max()), float(err.max())
    if stats is not None:
        s = stats.sum(dim=1)  # [B, Cout, 2]
        gq = got.double()
        assert torch.allclose(s[..., 0].double(), gq.sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)
        assert torch.allclose(s[..., 1].double(), (gq * gq).sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("B,H,W,Cin,Cout", SHAPES)
def test_upconv_backward(B, H, W, Cin, Cout):
    k = K()
    g = torch.Generator(device=DEV).manual_seed(H * 11 + Cout)
    x = rb(torch.randn(B, Cin, H, W, device=DEV, generator=g), "grad").requires_grad_()
    w = (torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / math.sqrt(9 * Cin)).requires_grad_()
    b = torch.zeros(Cout, device=DEV, requires_grad=True)
    dy = rb(torch.randn(B, Cout, 2 * H, 2 * W, device=DEV, generator=g), "grad")
    _ref(x, w, b).backward(dy)
    dyh = nhwc(dy, "grad")
    dx = nchw(k.upconv_dgrad(dyh, _packed(k, w.detach(), Cin, Cout, True), Cin), "grad")
    assert_close_bf16(dx, x.grad, "upconv dgrad")
    dw = k.upconv_wgrad(dyh, nhwc(x.detach(), "grad"))
    assert tuple(dw.shape) == (Cout, Cin, 3, 3)
    r = rel_l2(dw, w.grad)
    assert r < 2e-3, f"upconv wgrad rel-L2 {r}"
    # per-tap check: each of the 9 taps must be right on its own (the phase -> tap fold is where an index slip would hide)
    for tap in range(9):
        rt = rel_l2(dw.flatten(2)[:, :, tap], w.grad.flatten(2)[:, :, tap])
        assert rt < 4e-3, (tap, rt)


def test_upsample_module_uses_the_phase_path_and_matches_the_materialised_one():
    from stain2stain_b200 import ops
    from stain2stain_b200.unet import Upsample
    k = K()
    g = torch.Generator(device=DEV).manual_seed(3)
    up = Upsample(128, True).to(DEV)
    x = rb(torch.randn(2, 128, 16, 16, device=DEV, generator=g))
    xh = nhwc(x).requires_grad_()
    n0 = k.LAUNCHES[0]
    y = up([xh])
    assert k.LAUNCHES[0] - n0 >= 4 and ops.stats_of(y) is not None  # four phase launches (+ the weight pack)
    dy = rb(torch.randn(2, 128, 32, 32, device=DEV, generator=g), "grad")
    y.backward(nhwc(dy, "grad"))
    xr = x.clone().requires_grad_()
    ref = _ref(xr, up.conv.weight, up.conv.bias)
    gw, gb, gx = torch.autograd.grad(ref, [up.conv.weight, up.conv.bias, xr], dy)
    assert rel_l2(nchw(y.detach()), ref) < 2e-3
    assert rel_l2(up.conv.weight.grad, gw) < 3e-3 and rel_l2(up.conv.bias.grad, gb) < 3e-3
    assert_close_bf16(nchw(xh.grad, "grad"), gx, "Upsample module dgrad")
    # optimizer step -> the phase-summed operands must be re-packed
    with torch.no_grad():
        up.conv.weight.mul_(2.0)
    y2 = up([xh.detach()])
    ref2 = _ref(x, up.conv.weight, up.conv.bias)
    assert rel_l2(nchw(y2), ref2) < 2e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 128, 128), (1, 32, 16, 256, 256), (2, 8, 24, 128, 256)])
def test_downconv_dgrad_matches_autograd_of_the_stride2_conv(B, H, W, Cin, Cout):
    """Phase-decomposed data gradient of Downsample (conv3x3, stride 2, pad 1): H, W are the OUTPUT dims."""
    k = K()
    g = torch.Generator(device=DEV).manual_seed(H * 13 + Cin)
    x = torch.zeros(B, Cin, 2 * H, 2 * W, device=DEV, requires_grad=True)
    w = rb(torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / math.sqrt(9 * Cin), "grad")
    dy = rb(torch.randn(B, Cout, H, W, device=DEV, generator=g), "grad")
    F.conv2d(x, w, None, stride=2, padding=1).backward(dy)
    wd = torch.zeros((Cin, 9 * Cout), dtype=k.T16, device=DEV)
    k.pack_conv_weight(w, wd, ci_begin=0, ci_count=Cin, transpose_flip=True, fmt=k.GRAD)
    dx = nchw(k.downconv_dgrad(nhwc(dy, "grad"), wd, Cin), "grad")
    assert_close_bf16(dx, x.grad, "downconv dgrad")
    # and through the autograd op the UNet's Downsample uses
    from stain2stain_b200 import ops
    from stain2stain_b200.ops import ConvPlan, Seg
    plan = ConvPlan((Seg(0, 0, 0, Cin, 9, 2),), Cout)
    xin = rb(torch.randn(B, Cin, 2 * H, 2 * W, device=DEV, generator=g))
    xh = nhwc(xin).requires_grad_()
    wp = w.clone().requires_grad_()
    bias = torch.zeros(Cout, device=DEV, requires_grad=True)
    n0 = k.LAUNCHES[0]
    y = ops.fused_conv(plan, [xh], [wp], [bias])
    y.backward(nhwc(dy, "grad"))
    xr = xin.clone().requires_grad_()
    ref = F.conv2d(xr, w, None, stride=2, padding=1)
    gx, = torch.autograd.grad(ref, [xr], dy)
    assert_close_bf16(nchw(xh.grad, "grad"), gx, "Downsample dgrad through fused_conv")
