"""GPU parity of each C-ABI kernel against the same op in plain PyTorch fp32 (inputs pre-rounded to bf16).

Tolerances: conv / GEMM outputs are bf16-rounded -> max |err| <= 2^-7 * max|ref| (+ accumulation-order noise);
fp32 outputs (wgrad, stats, loss) -> rel-L2 <= 2e-3.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def K():
    from stain2stain_b200 import kernels
    return kernels


def _fmt(kind):
    k = K()
    return k.ACT if kind == "act" else k.GRAD


def rb(x, kind="act"):  # round to the storage format and back
    k = K()
    return k.to_float(k.from_float(x, _fmt(kind)), _fmt(kind))


def nhwc(x, kind="act"):  # fp32 NCHW -> opaque 16-bit NHWC
    return K().nchw_to_nhwc16(x.contiguous(), _fmt(kind))


def nchw(x, kind="act"):  # opaque 16-bit NHWC -> fp32 NCHW
    return K().nhwc16_to_nchw(x, _fmt(kind))


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def assert_close_bf16(got, ref, what):
    err = float((got - ref).abs().max())
    scale = float(ref.abs().max())
    r = rel_l2(got, ref)
    assert err <= scale * 2 ** -6 and r < 6e-3, f"{what}: max err {err:.4g} (scale {scale:.4g}) rel-L2 {r:.3g}"


def assert_close_act(got, ref, what):
    """forward-format outputs: fp16 by default (2^-11 relative rounding), bf16 when S2S_ACT_DTYPE=bf16"""
    if K().ACT == K().FMT_BF16:
        return assert_close_bf16(got, ref, what)
    err = float((got - ref).abs().max())
    scale = float(ref.abs().max())
    r = rel_l2(got, ref)
    assert err <= scale * 2 ** -9 and r < 8e-4, f"{what}: max err {err:.4g} (scale {scale:.4g}) rel-L2 {r:.3g}"


def pack_fwd(k, weights, cout):
    """weights: list of (w fp32 OIHW, ci_begin, ci_count) sharing one accumulator."""
    ktot = sum(w[0, 0].numel() * ((cnt + 63) // 64 * 64) for w, _, cnt in weights)
    wp = torch.zeros((k.padded_rows(cout), ktot), dtype=k.T16, device=DEV)
    off = 0
    for w, beg, cnt in weights:
        k.pack_conv_weight(w, wp, k_off=off, ci_begin=beg, ci_count=cnt)
        off += w[0, 0].numel() * ((cnt + 63) // 64 * 64)
    return wp


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 32, 32, 64, 64), (1, 16, 48, 128, 128), (2, 32, 32, 256, 256),
                                            (1, 32, 32, 128, 512), (3, 24, 40, 64, 128)])
def test_conv3x3_s1(B, H, W, Cin, Cout):
    k = K()
    g = torch.Generator(device=DEV).manual_seed(1)
    x = rb(torch.randn(B, Cin, H, W, device=DEV, generator=g))
    w = rb(torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / math.sqrt(9 * Cin))
    b = torch.randn(Cout, device=DEV, generator=g)
    r = rb(torch.randn(B, Cout, H, W, device=DEV, generator=g))
    ref = F.conv2d(x, w, b, padding=1) + r
    wp = pack_fwd(k, [(w, 0, Cin)], Cout)
    out = k.conv_fwd([(nhwc(x), 9, 1)], wp, Cout, H, W, bias=b, residual=nhwc(r))
    assert_close_act(nchw(out), ref, "conv3x3 s1")


def test_conv1x1_and_multiseg():
    k = K()
    g = torch.Generator(device=DEV).manual_seed(2)
    B, H, W = 2, 32, 32
    xa = rb(torch.randn(B, 128, H, W, device=DEV, generator=g))
    xs0 = rb(torch.randn(B, 128, H, W, device=DEV, generator=g))
    xs1 = rb(torch.randn(B, 64, H, W, device=DEV, generator=g))
    w3 = rb(torch.randn(128, 128, 3, 3, device=DEV, generator=g) / math.sqrt(9 * 128))
    w1 = rb(torch.randn(128, 192, 1, 1, device=DEV, generator=g) / math.sqrt(192))
    b = torch.randn(128, device=DEV, generator=g)
    ref = F.conv2d(xa, w3, b, padding=1) + F.conv2d(torch.cat([xs0, xs1], 1), w1)
    wp = pack_fwd(k, [(w3, 0, 128), (w1, 0, 128), (w1, 128, 64)], 128)
    out = k.conv_fwd([(nhwc(xa), 9, 1), (nhwc(xs0), 1, 1), (nhwc(xs1), 1, 1)], wp, 128, H, W, bias=b)
    assert_close_act(nchw(out), ref, "3x3 + 1x1 skip over a concat")
    # plain 1x1
    ref1 = F.conv2d(xs0, w1[:, :128].contiguous())
    wp1 = pack_fwd(k, [(w1, 0, 128)], 128)
    out1 = k.conv_fwd([(nhwc(xs0), 1, 1)], wp1, 128, H, W)
    assert_close_act(nchw(out1), ref1, "1x1")


@pytest.mark.parametrize("B,H,W,C", [(2, 64, 64, 128), (1, 32, 64, 256)])
def test_conv3x3_s2(B, H, W, C):
    k = K()
    g = torch.Generator(device=DEV).manual_seed(3)
    x = rb(torch.randn(B, C, H, W, device=DEV, generator=g))
    w = rb(torch.randn(C, C, 3, 3, device=DEV, generator=g) / math.sqrt(9 * C))
    b = torch.randn(C, device=DEV, generator=g)
    ref = F.conv2d(x, w, b, stride=2, padding=1)
    wp = pack_fwd(k, [(w, 0, C)], C)
    out = k.conv_fwd([(nhwc(x), 9, 2)], wp, C, H // 2, W // 2, bias=b)
    assert_close_act(nchw(out), ref, "conv3x3 s2")


def test_conv_head_f32_axpy():
    k = K()
    g = torch.Generator(device=DEV).manual_seed(4)
    B, H, W, C = 2, 32, 48, 128
    x = rb(torch.randn(B, C, H, W, device=DEV, generator=g))
    w = rb(torch.randn(3, C, 3, 3, device=DEV, generator=g) / math.sqrt(9 * C))
    b = torch.randn(3, device=DEV, generator=g)
    x_state = torch.randn(B, 3, H, W, device=DEV, generator=g)
    ref = F.conv2d(x, w, b, padding=1)
    wp = pack_fwd(k, [(w, 0, C)], 3)
    out = k.conv_fwd([(nhwc(x), 9, 1)], wp, 3, H, W, bias=b, out_f32=True)
    assert rel_l2(out, ref) < 1e-4, rel_l2(out, ref)
    out2 = k.conv_fwd([(nhwc(x), 9, 1)], wp, 3, H, W, bias=b, out_f32=True, axpy_x=x_state, axpy_a=0.02)
    assert rel_l2(out2, x_state + 0.02 * ref) < 1e-5


def test_conv_dgrad_via_flipped_pack():
    k = K()
    g = torch.Generator(device=DEV).manual_seed(5)
    B, H, W, Cin, Cout = 2, 32, 32, 128, 256
    x = torch.randn(B, Cin, H, W, device=DEV, generator=g).requires_grad_(True)
    w = rb(torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / math.sqrt(9 * Cin), "grad")
    dy = rb(torch.randn(B, Cout, H, W, device=DEV, generator=g), "grad")
    F.conv2d(x, w, padding=1).backward(dy)
    wd = torch.zeros((Cin, 9 * Cout), dtype=k.T16, device=DEV)
    k.pack_conv_weight(w, wd, transpose_flip=True, fmt=k.GRAD)
    dx = k.conv_fwd([(nhwc(dy, "grad"), 9, 1)], wd, Cin, H, W, a_fmt=k.GRAD, w_fmt=k.GRAD, out_fmt=k.GRAD)
    assert_close_bf16(nchw(dx, "grad"), x.grad, "dgrad 3x3")
    # stride-2 dgrad = zero insertion + stride-1 conv with the flipped weights
    x2 = torch.randn(B, Cin, H, W, device=DEV, generator=g).requires_grad_(True)
    w2 = rb(torch.randn(Cin, Cin, 3, 3, device=DEV, generator=g) / math.sqrt(9 * Cin), "grad")
    dy2 = rb(torch.randn(B, Cin, H // 2, W // 2, device=DEV, generator=g), "grad")
    F.conv2d(x2, w2, stride=2, padding=1).backward(dy2)
    wd2 = torch.zeros((Cin, 9 * Cin), dtype=k.T16, device=DEV)
    k.pack_conv_weight(w2, wd2, transpose_flip=True, fmt=k.GRAD)
    dx2 = k.conv_fwd([(k.zero_insert2x(nhwc(dy2, "grad")), 9, 1)], wd2, Cin, H, W, a_fmt=k.GRAD, w_fmt=k.GRAD,
                     out_fmt=k.GRAD)
    assert_close_bf16(nchw(dx2, "grad"), x2.grad, "dgrad 3x3 s2")


@pytest.mark.parametrize("taps,stride,Cin,Cout,B,H,W", [(9, 1, 128, 128, 2, 32, 32), (9, 1, 64, 256, 1, 16, 48),
                                                       (1, 1, 256, 128, 2, 32, 32), (9, 2, 128, 128, 2, 32, 32),
                                                       (9, 1, 512, 256, 1, 16, 16), (1, 1, 64, 128, 2, 24, 40),
                                                       # 128 output channels, 3x3 stride 1: the tap-pair mode of the CTA-pair
                                                       # kernel (ragged tiles, Cin = 256 -> N tile 256, Cin = 384 -> 3 x 128)
                                                       (9, 1, 256, 128, 3, 24, 40), (9, 1, 384, 128, 1, 32, 32),
                                                       (9, 1, 128, 384, 2, 16, 16)])
def test_conv_wgrad(taps, stride, Cin, Cout, B, H, W):
    k = K()
    g = torch.Generator(device=DEV).manual_seed(6)
    ks = 3 if taps == 9 else 1
    x = rb(torch.randn(B, Cin, H * stride, W * stride, device=DEV, generator=g), "grad")
    w = torch.zeros(Cout, Cin, ks, ks, device=DEV, requires_grad=True)
    dy = rb(torch.randn(B, Cout, H, W, device=DEV, generator=g), "grad")
    F.conv2d(x, w, stride=stride, padding=ks // 2).backward(dy)
    dw = torch.zeros((taps, Cout, Cin), dtype=torch.float32, device=DEV)
    k.conv_wgrad(nhwc(dy, "grad"), nhwc(x, "grad"), taps, stride, dw)
    grad = torch.zeros_like(w)
    k.unpack_wgrad(dw, grad, 0, Cin, 0, 0.0)
    r = rel_l2(grad, w.grad)
    assert r < 2e-3, f"wgrad rel-L2 {r}"


def test_patch27_and_stem():
    k = K()
    g = torch.Generator(device=DEV).manual_seed(7)
    B, H, W = 2, 32, 48
    x0 = torch.rand(B, 3, H, W, device=DEV, generator=g) * 2 - 1
    x1 = torch.rand(B, 3, H, W, device=DEV, generator=g) * 2 - 1
    t = torch.rand(B, device=DEV, generator=g)
    w = rb(torch.randn(128, 3, 3, 3, device=DEV, generator=g) / math.sqrt(27))
    b = torch.randn(128, device=DEV, generator=g)
    xt_ref = (1 - t)[:, None, None, None] * x0 + t[:, None, None, None] * x1
    patches, xt = k.patch27_pack(x0, 1, x1, t, want_xt=True)
    assert rel_l2(xt, xt_ref) < 1e-6
    # stem conv as a 1x1 GEMM over the patches: weight [128][64] with column tap*3 + c
    wp = torch.zeros((128, 64), dtype=k.T16, device=DEV)
    k.pack_conv_weight(w, wp)
    out = k.conv_fwd([(patches, 1, 1)], wp, 128, H, W, bias=b)
    ref = F.conv2d(rb(xt_ref), w, b, padding=1)
    assert_close_act(nchw(out), ref, "stem via patches")
    # sgn = -1: adjoint patches, check against unfold of the flipped image
    xp = F.pad(x0, (1, 1, 1, 1))
    for sgn in (1, -1):
        pm = k.patch27_pack(x0, sgn)
        cols = []
        for tap in range(9):
            oy, ox = 1 + sgn * (tap // 3 - 1), 1 + sgn * (tap % 3 - 1)
            cols.append(xp[:, :, oy:oy + H, ox:ox + W])
        ref_m = torch.stack(cols, 1).permute(0, 3, 4, 1, 2).reshape(B, H, W, 27)
        pmf = k.to_float(pm, k.ACT)
        assert rel_l2(pmf[..., :27], rb(ref_m)) < 1e-6, sgn
        assert float(pmf[..., 27:].abs().max()) == 0.0


def _gn_ref(x, gamma, beta, film, silu, G=32):
    y = F.group_norm(x, G, gamma, beta, eps=1e-5)
    if film is not None:
        C = x.shape[1]
        y = y * (1 + film[:, :C, None, None]) + film[:, C:, None, None]
    return F.silu(y) if silu else y


@pytest.mark.parametrize("C0,C1,film,silu", [(128, 0, False, True), (256, 128, False, True), (512, 256, True, True),
                                             (64, 0, True, False), (1024, 0, False, True), (384, 0, True, True)])
def test_groupnorm_fwd_bwd(C0, C1, film, silu):
    k = K()
    g = torch.Generator(device=DEV).manual_seed(8)
    B, H, W = 3, 16, 24
    Ct = C0 + C1
    xs = [rb(torch.randn(B, c, H, W, device=DEV, generator=g) * 1.5 + 0.3) for c in (C0, C1) if c]
    gamma = (1 + 0.2 * torch.randn(Ct, device=DEV, generator=g)).requires_grad_(True)
    beta = (0.2 * torch.randn(Ct, device=DEV, generator=g)).requires_grad_(True)
    fl = (0.3 * torch.randn(B, 2 * Ct, device=DEV, generator=g)).requires_grad_(True) if film else None
    xcat = torch.cat(xs, 1).requires_grad_(True)
    ref = _gn_ref(xcat, gamma, beta, fl, silu)
    gy = rb(torch.randn(B, Ct, H, W, device=DEV, generator=g), "grad")
    ref.backward(gy)

    stats = k.gn_partial_buffer(B, H * W, Ct, DEV)
    xh = [nhwc(x) for x in xs]
    off = 0
    for x in xh:
        k.gn_stats(x, stats, off)
        off += x.shape[3]
    coef, mr = k.gn_coef(stats, gamma.detach(), beta.detach(), fl.detach() if film else None, H * W)
    y = torch.empty((B, H, W, Ct), dtype=k.T16, device=DEV)
    off = 0
    for x in xh:
        k.gn_apply(x, coef, y, off, silu)
        off += x.shape[3]
    assert_close_act(nchw(y), ref.detach(), "gn fwd")

    red = k.gn_partial_buffer(B, H * W, Ct, DEV)
    gh = nhwc(gy, "grad")
    off = 0
    for x in xh:
        k.gn_bwd_reduce(x, gh, coef, mr, red, off, silu)
        off += x.shape[3]
    dgamma = torch.zeros(Ct, device=DEV)
    dbeta = torch.zeros(Ct, device=DEV)
    pqr, dfilm = k.gn_bwd_coef(red, mr, gamma.detach(), beta.detach(), fl.detach() if film else None, H * W, dgamma,
                               dbeta, film)
    off = 0
    dxs = []
    for x in xh:
        dx = torch.empty_like(x)
        k.gn_bwd_apply(x, gh, coef, pqr, off, None, dx, silu)
        dxs.append(nchw(dx, "grad"))
        off += x.shape[3]
    dx = torch.cat(dxs, 1)
    assert rel_l2(dx, xcat.grad) < 1e-2, f"gn dx rel-L2 {rel_l2(dx, xcat.grad)}"
    assert rel_l2(dgamma, gamma.grad) < 5e-3, f"dgamma {rel_l2(dgamma, gamma.grad)}"
    assert rel_l2(dbeta, beta.grad) < 5e-3, f"dbeta {rel_l2(dbeta, beta.grad)}"
    if film:
        assert rel_l2(dfilm, fl.grad) < 5e-3, f"dfilm {rel_l2(dfilm, fl.grad)}"


def test_dropout_statistics_and_replay():
    k = K()
    B, H, W, Cc = 2, 32, 32, 128
    x = k.from_float(torch.ones((B, H, W, Cc), device=DEV), k.ACT)
    coef = torch.zeros((B, Cc, 2), device=DEV)
    coef[..., 0] = 1.0  # y = x
    y1 = torch.empty_like(x)
    y2 = torch.empty_like(x)
    k.gn_apply(x, coef, y1, 0, False, 0.1, 1234)
    k.gn_apply(x, coef, y2, 0, False, 0.1, 1234)
    assert torch.equal(y1, y2)
    y1f = k.to_float(y1, k.ACT)
    keep = float((y1f != 0).float().mean())
    assert abs(keep - 0.9) < 0.005, keep
    kept = y1f[y1f != 0]
    assert float((kept - 1 / 0.9).abs().max()) < 0.01
    y3 = torch.empty_like(x)
    k.gn_apply(x, coef, y3, 0, False, 0.1, 99)
    assert not torch.equal(y1, y3)


def test_resample_and_sums():
    k = K()
    g = torch.Generator(device=DEV).manual_seed(9)
    x = rb(torch.randn(2, 64, 16, 24, device=DEV, generator=g))
    up = k.upsample2x(nhwc(x))
    assert torch.equal(nchw(up), F.interpolate(x, scale_factor=2, mode="nearest"))
    big = rb(torch.randn(2, 64, 32, 48, device=DEV, generator=g), "grad")
    sp = k.sumpool2x(nhwc(big, "grad"))
    assert_close_bf16(nchw(sp, "grad"), F.avg_pool2d(big, 2) * 4, "sumpool")
    zi = nchw(k.zero_insert2x(nhwc(x)))
    assert torch.equal(zi[:, :, ::2, ::2], x)
    assert float(zi[:, :, 1::2, :].abs().max()) == 0.0 and float(zi[:, :, :, 1::2].abs().max()) == 0.0
    xr = rb(x)
    assert torch.equal(k.to_float(k.convert16(k.from_float(xr, k.FMT_F16), k.FMT_F16, k.FMT_BF16), k.FMT_BF16),
                       xr.to(torch.bfloat16).float())
    for C in (64, 384):
        big = rb(torch.randn(2, C, 32, 48, device=DEV, generator=g), "grad")
        cs = torch.zeros(C, device=DEV)
        k.channel_sum(nhwc(big, "grad"), cs)
        assert rel_l2(cs, big.sum(dim=(0, 2, 3))) < 1e-4
    for kind in ("act", "grad"):
        xr = rb(x, kind)
        assert torch.equal(nchw(nhwc(xr, kind), kind), xr)


def test_fm_loss():
    k = K()
    g = torch.Generator(device=DEV).manual_seed(10)
    v = torch.randn(3, 3, 32, 32, device=DEV, generator=g, requires_grad=True)
    x0 = torch.rand(3, 3, 32, 32, device=DEV, generator=g)
    x1 = torch.rand(3, 3, 32, 32, device=DEV, generator=g)
    ref = torch.mean((v - (x1 - x0)) ** 2)
    ref.backward()
    loss, dv = k.fm_loss(v.detach(), x0, x1, True)
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref)) + 1e-7
    assert rel_l2(dv, v.grad) < 1e-6


@pytest.mark.parametrize("wd", [0.0, 1e-5])
def test_fused_adam_matches_torch_adam(wd):
    """FusedAdam (one multi-tensor launch) vs torch.optim.Adam, same params / grads, 3 steps; state_dict layout equal."""
    from stain2stain_b200.optim import FusedAdam
    g = torch.Generator(device=DEV).manual_seed(3)
    shapes = [(128, 64, 3, 3), (257,), (3, 5, 7), (40000,), (1,)]
    pa = [torch.randn(s, device=DEV, generator=g).requires_grad_() for s in shapes]
    pb = [p.detach().clone().requires_grad_() for p in pa]
    oa = FusedAdam(pa, lr=1e-3, weight_decay=wd)
    ob = torch.optim.Adam(pb, lr=1e-3, weight_decay=wd)
    for it in range(3):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, device=DEV, generator=g)
            a.grad, b.grad = gr.clone(), gr.clone()
        oa.step(), ob.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), float((a - b).abs().max())
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa["state"].keys() == sb["state"].keys()
    for k in sa["state"]:
        assert set(sa["state"][k]) == set(sb["state"][k]) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(sa["state"][k]["step"]) == float(sb["state"][k]["step"]) == 3.0
        assert torch.allclose(sa["state"][k]["exp_avg_sq"], sb["state"][k]["exp_avg_sq"], rtol=1e-5, atol=1e-8)


def test_fused_adam_resume_uses_the_loaded_moments():
    """step -> load_state_dict (in-process resume swaps the moment tensors) -> step: the device pointer table must be
    rebuilt, the loaded moments used and the new ones written back (optim.py:37)."""
    from stain2stain_b200.optim import FusedAdam
    g = torch.Generator(device=DEV).manual_seed(4)
    shapes = [(64, 32, 3, 3), (513,), (20000,)]
    pa = [torch.randn(s, device=DEV, generator=g).requires_grad_() for s in shapes]
    pb = [p.detach().clone().requires_grad_() for p in pa]
    oa, ob = FusedAdam(pa, lr=1e-3, weight_decay=1e-5), torch.optim.Adam(pb, lr=1e-3, weight_decay=1e-5)

    def both_step():
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, device=DEV, generator=g)
            a.grad, b.grad = gr.clone(), gr.clone()
        oa.step(), ob.step()
    both_step()
    both_step()
    import copy
    ckpt_a, ckpt_b = copy.deepcopy(oa.state_dict()), copy.deepcopy(ob.state_dict())
    w_a = [p.detach().clone() for p in pa]
    both_step()  # moves on past the checkpoint ...
    with torch.no_grad():  # ... then resume from it in-process: same parameter / gradient storage, NEW moment tensors
        for a, b, w in zip(pa, pb, w_a):
            a.copy_(w), b.copy_(w)
    oa.load_state_dict(ckpt_a), ob.load_state_dict(ckpt_b)
    old_m = [oa.state[p]["exp_avg"] for p in pa]
    both_step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), float((a - b).abs().max())
    for p, q, m in zip(pa, pb, old_m):
        assert oa.state[p]["exp_avg"] is m and float(oa.state[p]["step"]) == 3.0
        assert torch.allclose(oa.state[p]["exp_avg"], ob.state[q]["exp_avg"], rtol=1e-5, atol=1e-7)
        assert torch.allclose(oa.state[p]["exp_avg_sq"], ob.state[q]["exp_avg_sq"], rtol=1e-5, atol=1e-9)


def test_stored_dropout_mask_equals_rehash():
    """The keep bits written by gn_apply (1 bit / element) are exactly the Philox mask, and the backward kernels give
    bit-identical results whether they read the stored bits or re-evaluate the hash."""
    k = K()
    g = torch.Generator(device=DEV).manual_seed(12)
    B, H, W, Cc = 2, 16, 24, 128
    x = nhwc(rb(torch.randn(B, Cc, H, W, device=DEV, generator=g)))
    gy = nhwc(rb(torch.randn(B, Cc, H, W, device=DEV, generator=g), "grad"), "grad")
    gamma, beta = torch.ones(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    stats = k.gn_partial_buffer(B, H * W, Cc, DEV)
    k.gn_stats(x, stats, 0)
    coef, mr = k.gn_coef(stats, gamma, beta, None, H * W)
    y, y2 = torch.empty_like(x), torch.empty_like(x)
    mask = torch.zeros((B, H, W, Cc // 8), dtype=torch.uint8, device=DEV)
    k.gn_apply(x, coef, y, 0, True, 0.1, 77, y2=y2, mask=mask)
    yf = k.to_float(y, k.ACT)
    bits = ((mask.unsqueeze(-1) >> torch.arange(8, device=DEV, dtype=torch.uint8)) & 1).reshape(B, H, W, Cc).bool()
    assert torch.equal(bits | (yf == 0), torch.ones_like(bits)) and torch.equal(yf[~bits], torch.zeros_like(yf[~bits]))
    assert abs(float(bits.float().mean()) - 0.9) < 0.01
    assert float((y2.float() - yf).abs().max()) <= 2 ** -8 * float(yf.abs().max())  # bf16 copy of the same values
    outs = []
    for m in (None, mask):
        red = k.gn_partial_buffer(B, H * W, Cc, DEV)
        k.gn_bwd_reduce(x, gy, coef, mr, red, 0, True, 0.1, 77, mask=m)
        dg, db = torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV)
        pqr, _ = k.gn_bwd_coef(red, mr, gamma, beta, None, H * W, dg, db, False)
        dx = torch.empty_like(x)
        k.gn_bwd_apply(x, gy, coef, pqr, 0, None, dx, True, 0.1, 77, mask=m)
        outs.append((red.clone(), dx.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("taps", [9, 1])
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 32, 32, 128, 128), (3, 16, 48, 64, 256), (2, 24, 20, 128, 128),
                                            (1, 8, 8, 256, 512), (3, 40, 40, 128, 128), (1, 128, 144, 64, 128)])
def test_conv_epilogue_statistics_match_a_stats_pass(B, H, W, Cin, Cout, taps):
    """GroupNorm statistics emitted by the CTA-pair conv epilogues (per 128-pixel sub-tile, from the staged 16-bit tile)
    give the same coefficients as a separate gn_stats pass over the stored tensor -- incl. ragged tiles.  taps = 9 runs
    the halo-tiled pair kernel (8 x 16 pixel tiles), taps = 1 the plain pair kernel (16 x 8 pixel tiles)."""
    k = K()
    g = torch.Generator(device=DEV).manual_seed(21)
    x = nhwc(rb(torch.randn(B, Cin, H, W, device=DEV, generator=g)))
    w = torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / (3 * Cin ** 0.5)
    if taps == 1:
        w = w[:, :, 1:2, 1:2].contiguous() * 3
    wp = pack_fwd(k, [(w, 0, Cin)], Cout)
    bias = torch.randn(Cout, device=DEV, generator=g)
    res = k.conv_fwd([(x, taps, 1)], wp, Cout, H, W, bias=bias, want_stats="force")  # (short-K convs skip them by default)
    y, st = res
    if st is None:
        pytest.skip("no CTA-pair path for this geometry")
    assert st.shape[0] == B and tuple(st.shape[2:]) == (Cout, 2) and st.shape[1] * 128 >= H * W
    gamma = 1 + 0.1 * torch.randn(Cout, device=DEV, generator=g)
    beta = 0.1 * torch.randn(Cout, device=DEV, generator=g)
    coef_e, mr_e = k.gn_coef_parts([st], gamma, beta, None, H * W)
    stats = k.gn_partial_buffer(B, H * W, Cout, DEV)
    k.gn_stats(y, stats, 0)
    coef_s, mr_s = k.gn_coef(stats, gamma, beta, None, H * W)
    assert torch.allclose(mr_e, mr_s, rtol=2e-4, atol=2e-5), float((mr_e - mr_s).abs().max())
    assert torch.allclose(coef_e, coef_s, rtol=2e-4, atol=2e-5)
    # two-source fold == fold of the concatenation
    y2, st2 = k.conv_fwd([(x, taps, 1)], wp, Cout, H, W, bias=bias * 0.5, want_stats="force")
    g2 = torch.cat([gamma, gamma]); b2 = torch.cat([beta, beta])
    coef_2, mr_2 = k.gn_coef_parts([st, st2], g2, b2, None, H * W)
    stats_c = k.gn_partial_buffer(B, H * W, 2 * Cout, DEV)
    k.gn_stats(y, stats_c, 0)
    k.gn_stats(y2, stats_c, Cout)
    coef_c, mr_c = k.gn_coef(stats_c, g2, b2, None, H * W)
    assert torch.allclose(mr_2, mr_c, rtol=2e-4, atol=2e-5) and torch.allclose(coef_2, coef_c, rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("B,H,W,C0,C1,Cout,skip", [(2, 32, 32, 128, 0, 128, False), (3, 16, 48, 128, 64, 256, True),
                                                   (1, 24, 40, 256, 128, 128, True), (3, 8, 8, 256, 0, 256, False),
                                                   (2, 64, 64, 64, 64, 128, False)])
def test_conv_with_fused_norm_prologue_matches_apply_then_conv(B, H, W, C0, C1, Cout, skip):
    """s2s_conv_fwd_norm (the conv applies silu(x*A + Bc) to its shared-memory tiles) against s2s_gn_apply followed by
    s2s_conv_fwd on the materialised tensor: one / two normalised 3x3 sources (channel concat), optional raw 1x1 skip
    segments, ragged tiles, odd tile counts, zero padding behind the activation."""
    k = K()
    g = torch.Generator(device=DEV).manual_seed(31)
    widths = [C0] + ([C1] if C1 else [])
    ctot = sum(widths)
    xs = [nhwc(rb(torch.randn(B, c, H, W, device=DEV, generator=g))) for c in widths]
    coef = torch.stack([1 + 0.3 * torch.randn(B, ctot, device=DEV, generator=g),
                        0.5 * torch.randn(B, ctot, device=DEV, generator=g)], dim=2).contiguous()
    w3 = torch.randn(Cout, ctot, 3, 3, device=DEV, generator=g) / (3 * ctot ** 0.5)
    bias = torch.randn(Cout, device=DEV, generator=g)
    segs3 = [(w3, off, c) for off, c in zip([0, C0], widths)]
    srcs = [(x, 9, 1) for x in xs]
    norms = [(coef, off) for off in [0, C0][:len(widths)]]
    weights = list(segs3)
    if skip:  # raw 1x1 skip segments over the same sources (ResBlock conv 2 shape)
        w1 = torch.randn(Cout, ctot, 1, 1, device=DEV, generator=g) / ctot ** 0.5
        weights += [(w1, off, c) for off, c in zip([0, C0], widths)][:3 - len(widths)]
        srcs += [(x, 1, 1) for x in xs][:3 - len(widths)]
        norms += [None] * (len(srcs) - len(norms))
    if not k.conv_norm_fusable(srcs, Cout, force=True):
        pytest.skip("halo-tiled pair kernel not selected (or bf16 forward format)")
    wp = pack_fwd(k, weights, Cout)
    got, st = k.conv_fwd(srcs, wp, Cout, H, W, bias=bias, want_stats="force", norms=norms, norm_act=1)
    # reference: materialise silu(x*A+Bc) per source, then the same conv
    acts = []
    off = 0
    for x, c in zip(xs, widths):
        y = torch.empty((B, H, W, c), dtype=k.T16, device=DEV)
        k.gn_apply(x, coef[:, off:off + c].contiguous(), y, 0, True)
        acts.append(y)
        off += c
    ref_srcs = [(a, 9, 1) for a in acts] + srcs[len(widths):]
    want, st_w = k.conv_fwd(ref_srcs, wp, Cout, H, W, bias=bias, want_stats="force")
    a, b = k.to_float(got, k.ACT), k.to_float(want, k.ACT)
    # the prologue's SiLU is the one-MUFU tanh.approx.f16x2 form (|error| <= |z|/2 * 2^-11 per activation), gn_apply's the
    # exact ex2 + rcp form: agreement to a few 1e-4 of the output scale, far inside the 1e-2 velocity gate
    assert rel_l2(a, b) <= 1e-3, rel_l2(a, b)
    assert float((a - b).abs().max()) <= 2e-2 * float(b.abs().max())
    assert torch.allclose(st, st_w, rtol=2e-2, atol=2e-2 * float(st_w.abs().max()))


def test_gn_bwd_reduce_bf16_side_copy_equals_convert16():
    """The norm backward reduce pass can store its input as bf16 on the way (the skip conv's weight-gradient operand):
    same bits as s2s_convert16, same reduction result as without the side product."""
    k = K()
    g = torch.Generator(device=DEV).manual_seed(41)
    B, H, W, C = 2, 24, 20, 192
    x = nhwc(rb(torch.randn(B, C, H, W, device=DEV, generator=g)))
    gy = nhwc(rb(torch.randn(B, C, H, W, device=DEV, generator=g), "grad"), "grad")
    gamma = 1 + 0.1 * torch.randn(C, device=DEV, generator=g)
    beta = 0.1 * torch.randn(C, device=DEV, generator=g)
    stats = k.gn_partial_buffer(B, H * W, C, DEV)
    k.gn_stats(x, stats, 0)
    coef, mr = k.gn_coef(stats, gamma, beta, None, H * W)
    red_a = k.gn_partial_buffer(B, H * W, C, DEV)
    red_b = k.gn_partial_buffer(B, H * W, C, DEV)
    k.gn_bwd_reduce(x, gy, coef, mr, red_a, 0, True)
    side = torch.full_like(x, float("nan"))
    k.gn_bwd_reduce(x, gy, coef, mr, red_b, 0, True, x_bf16_out=side)
    assert torch.equal(red_a, red_b)
    assert torch.equal(side.view(torch.int16), k.convert16(x, k.ACT, k.GRAD).view(torch.int16))


@pytest.mark.parametrize("B,H,W,C,Cout", [(2, 32, 32, 128, 3), (1, 20, 24, 64, 3), (3, 8, 16, 256, 5), (1, 64, 64, 128, 8)])
def test_head_conv_kernel_matches_conv2d_and_fused_euler_update(B, H, W, C, Cout):
    """s2s_head_conv: the C -> <=8 channel output projection from the fp32 weight, fp32 NCHW out, optional x + dt * v."""
    k = K()
    g = torch.Generator(device=DEV).manual_seed(C + H)
    a = rb(torch.randn(B, C, H, W, device=DEV, generator=g))
    w = torch.randn(Cout, C, 3, 3, device=DEV, generator=g) / math.sqrt(9 * C)
    b = torch.randn(Cout, device=DEV, generator=g) * 0.1
    ref = F.conv2d(a, rb(w), b, padding=1)  # the kernel rounds the weight to the activation format
    out = k.head_conv(nhwc(a), w.contiguous(), b)
    assert out.dtype == torch.float32 and tuple(out.shape) == (B, Cout, H, W)
    assert rel_l2(out, ref) < 1e-3, rel_l2(out, ref)
    assert float((out - ref).abs().max()) < 2e-3 * max(1.0, float(ref.abs().max()))
    x = torch.randn(B, Cout, H, W, device=DEV, generator=g)
    want = x + 0.02 * ref
    got = k.head_conv(nhwc(a), w.contiguous(), b, axpy_x=x, axpy_a=0.02)
    assert got.data_ptr() == x.data_ptr() and torch.allclose(got, want, atol=1e-4, rtol=1e-4)


def _expected_pack(k, w, ci_begin, ci_count, transpose_flip, fmt, mode):
    """The operand layouts of include/s2s_b200.h written with torch indexing (bit-exact: one rounding fp32 -> 16 bit)."""
    cout, cin = w.shape[:2]
    taps = w[0, 0].numel()
    w3 = w.reshape(cout, cin, taps)[:, ci_begin:ci_begin + ci_count]          # [co, ci, tap]
    if mode == 0:
        lt = w3.flip(2) if transpose_flip else w3
    else:
        ph = mode - 1
        rng = {(0, 0): [0], (0, 1): [1, 2], (1, 0): [0, 1], (1, 1): [2]}
        w9 = w3.reshape(cout, ci_count, 3, 3)
        cols = []
        for ti in range(4):
            ys, xs = rng[(ph >> 1, ti >> 1)], rng[(ph & 1, ti & 1)]
            acc = torch.zeros(cout, ci_count, device=w.device)
            for y in ys:             # same summation order as the kernel (dy outer, dx inner), fp32
                for x in xs:
                    acc = acc + w9[:, :, y, x]
            cols.append(acc)
        lt = torch.stack(cols, dim=2)                                          # [co, ci, 4]
    if transpose_flip:
        out = lt.permute(1, 2, 0).reshape(ci_count, -1)                       # [ci][tap * Cout + co]
    else:
        out = lt.permute(0, 2, 1).reshape(cout, -1)                           # [co][tap * ci_count + ci]
    return k.from_float(out.contiguous(), fmt)


@pytest.mark.parametrize("cout,cin,taps,ci_begin,ci_count", [(128, 128, 9, 0, 128), (96, 72, 9, 8, 40), (64, 200, 1, 64, 136),
                                                             (40, 24, 9, 0, 24), (256, 384, 9, 128, 256), (1024, 512, 1, 0, 512)])
def test_pack_conv_weight_layouts_bit_exact(cout, cin, taps, ci_begin, ci_count):
    """s2s_pack_conv_weight(_mode) and the multi-tensor launch: forward / transposed-flipped / phase-summed operands, ragged
    tile edges (sizes that are no multiples of the 16 x 64 tile), channel segments, non-zero column offsets."""
    k = K()
    g = torch.Generator(device=DEV).manual_seed(21)
    w = torch.randn((cout, cin, 3, 3) if taps == 9 else (cout, cin, 1, 1), device=DEV, generator=g)
    jobs, expect = [], []
    for tf in (False, True):
        for mode in ([0, 1, 2, 3, 4] if taps == 9 else [0]):
            fmt = k.GRAD if tf else k.ACT
            ltaps = taps if mode == 0 else 4
            rows, inner = (ci_count, cout) if tf else (cout, ci_count)
            k_off = 64
            single = torch.full((rows + 3, k_off + ltaps * inner + 6), 7, dtype=torch.int16, device=DEV).view(k.T16)
            multi = single.clone()
            k.pack_conv_weight(w, single, k_off=k_off, ci_begin=ci_begin, ci_count=ci_count, transpose_flip=tf, fmt=fmt, mode=mode)
            exp = _expected_pack(k, w, ci_begin, ci_count, tf, fmt, mode)
            got = single[:rows, k_off:k_off + ltaps * inner]
            assert torch.equal(got.view(torch.int16), exp.view(torch.int16)), (tf, mode)
            # nothing outside the operand is written
            pad = single.clone().view(torch.int16)
            pad[:rows, k_off:k_off + ltaps * inner] = 7
            assert bool((pad == 7).all()), (tf, mode)
            jobs.append((w, multi, k_off, ci_begin, ci_count, tf, fmt, mode))
            expect.append(single)
    k.pack_conv_weight_multi(jobs, {})
    for (_, multi, *_), single in zip(jobs, expect):
        assert torch.equal(multi.view(torch.int16), single.view(torch.int16))
