"""The attention core kernels (s2s_attn_fwd / s2s_attn_bwd) against the definition written with stock torch ops in fp32:
softmax(q k^T / sqrt(ch)) v per (sample, head), legacy and new channel order, ragged token counts, 32 / 64 head channels."""
import math

import pytest
import torch

from test_gpu_kernels import DEV, K, nchw, nhwc, rb, rel_l2

pytestmark = pytest.mark.gpu


def _ref(qkv, heads, new_order):
    """qkv fp32 [B, 3C, H, W] -> a [B, C, H, W]"""
    B, C3, H, W = qkv.shape
    C, T = C3 // 3, H * W
    ch = C // heads
    x = qkv.reshape(B, C3, T)
    if new_order:
        q, k, v = x.reshape(B, 3, heads, ch, T).unbind(1)
    else:
        q, k, v = x.reshape(B, heads, 3, ch, T).unbind(2)
    w = torch.softmax(torch.einsum("bhct,bhcs->bhts", q, k) / math.sqrt(ch), dim=-1)
    return torch.einsum("bhts,bhcs->bhct", w, v).reshape(B, C, H, W)


@pytest.mark.parametrize("B,H,W,heads,ch,new_order", [(2, 32, 32, 4, 32, False), (1, 16, 16, 8, 32, True), (2, 8, 8, 2, 64, False),
                                                     (1, 12, 12, 2, 32, False), (3, 16, 8, 4, 64, True), (1, 32, 32, 16, 32, False)])
def test_attention_forward_and_backward(B, H, W, heads, ch, new_order):
    k = K()
    assert k.attn_supported(ch)
    g = torch.Generator(device=DEV).manual_seed(H * 3 + heads)
    C = heads * ch
    qkv = rb(torch.randn(B, 3 * C, H, W, device=DEV, generator=g) * 1.5).requires_grad_()
    ref = _ref(qkv, heads, new_order)
    out, lse = k.attn_fwd(nhwc(qkv.detach()), heads, new_order, want_lse=True)
    got = nchw(out)
    assert rel_l2(got, ref.detach()) < 2e-3, rel_l2(got, ref.detach())
    assert float((got - ref.detach()).abs().max()) <= 2 ** -8 * float(ref.abs().max()) + 1e-3
    # lse: log2-domain logsumexp of the scaled scores
    x = qkv.detach().reshape(B, 3 * C, H * W)
    q, kk, _ = (x.reshape(B, 3, heads, ch, -1).unbind(1) if new_order else x.reshape(B, heads, 3, ch, -1).unbind(2))
    want_lse = torch.logsumexp(torch.einsum("bhct,bhcs->bhts", q, kk) / math.sqrt(ch), dim=-1) / math.log(2.0)
    assert torch.allclose(lse.view(B, heads, -1), want_lse, atol=2e-2, rtol=1e-3)
    d_out = rb(torch.randn(B, C, H, W, device=DEV, generator=g), "grad")
    ref.backward(d_out)
    d_qkv = nchw(k.attn_bwd(nhwc(qkv.detach()), out, nhwc(d_out, "grad"), lse, heads, new_order), "grad")
    r = rel_l2(d_qkv, qkv.grad)
    assert r < 1.5e-2, f"attention backward rel-L2 {r}"
    # q, k and v gradients separately (a slip in one of the three kernels must not hide in the norm of the others)
    gq = qkv.grad.reshape(B, 3, heads, ch, -1) if new_order else qkv.grad.reshape(B, heads, 3, ch, -1)
    dq = d_qkv.reshape(B, 3, heads, ch, -1) if new_order else d_qkv.reshape(B, heads, 3, ch, -1)
    for w in range(3):
        a, bref = (dq[:, w], gq[:, w]) if new_order else (dq[:, :, w], gq[:, :, w])
        assert rel_l2(a, bref) < 2e-2, (w, rel_l2(a, bref))


def test_attention_block_runs_on_the_own_kernel(monkeypatch):
    """No library attention on the path: torch SDPA must not be called for the reference's head size."""
    import torch.nn.functional as F
    from stain2stain_b200 import unet as punet

    def boom(*a, **kw):
        raise AssertionError("F.scaled_dot_product_attention was called")
    monkeypatch.setattr(F, "scaled_dot_product_attention", boom)
    blk = punet.AttentionBlock(128, num_head_channels=32).to(DEV)
    with torch.no_grad():
        for q in blk.parameters():
            q.normal_(0, 0.05)
    x = nhwc(rb(torch.randn(2, 128, 16, 16, device=DEV))).requires_grad_()
    y = blk([x])
    y.float().sum().backward()
    assert torch.isfinite(blk.qkv.weight.grad).all() and float(blk.qkv.weight.grad.abs().max()) > 0
