"""Thin functional wrappers: torch tensors in, C-ABI calls (stain2stain_b200/_lib.py) on the current CUDA stream.

Layouts: activations bf16 NHWC `[B, H, W, C]` contiguous; parameters fp32 in the reference's layouts.
No autograd here (see ops.py) and no CPU path: host tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ConvSrc, check, ptr, stream_ptr

BF16 = torch.bfloat16


def _L():
    return _lib.load()


def _nhwc_check(x: torch.Tensor):
    assert x.is_cuda and x.dtype == BF16 and x.dim() == 4 and x.is_contiguous(), \
        f"expected contiguous CUDA bf16 NHWC, got {x.dtype} {tuple(x.shape)} contiguous={x.is_contiguous()}"


def padded_rows(cout: int) -> int:
    if cout <= 16:
        return 16
    assert cout % 64 == 0, f"Cout = {cout}: need <= 16 or a multiple of 64"
    return cout


def pack_conv_weight(w: torch.Tensor, dst: torch.Tensor, k_off: int = 0, ci_begin: int = 0,
                     ci_count: Optional[int] = None, transpose_flip: bool = False):
    """w: fp32 [Cout, Cin, kh, kw] (or [Cout, Cin(,1)] for linear / Conv1d) -> rows of dst (bf16 [rows, ld_k])."""
    cout, cin = w.shape[0], w.shape[1]
    taps = w[0, 0].numel() if w.dim() > 2 else 1
    ci_count = cin - ci_begin if ci_count is None else ci_count
    assert w.is_contiguous() and w.dtype == torch.float32 and dst.dtype == BF16 and dst.is_contiguous()
    check(_L().s2s_pack_conv_weight(ptr(w), cout, cin, taps, ci_begin, ci_count, ptr(dst), dst.shape[1], k_off,
                                    int(transpose_flip), stream_ptr()), "pack_conv_weight")


def conv_fwd(srcs: Sequence[Tuple[torch.Tensor, int, int]], w_packed: torch.Tensor, cout: int, hout: int, wout: int,
             bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, out_f32: bool = False,
             axpy_x: Optional[torch.Tensor] = None, axpy_a: float = 0.0, out: Optional[torch.Tensor] = None):
    """srcs: [(x NHWC bf16, taps, stride)].  Returns bf16 NHWC [B,hout,wout,cout], or fp32 NCHW when out_f32."""
    arr = (ConvSrc * len(srcs))()
    B = srcs[0][0].shape[0]
    for i, (x, taps, stride) in enumerate(srcs):
        _nhwc_check(x)
        assert x.shape[0] == B and x.shape[1] == hout * stride and x.shape[2] == wout * stride, \
            f"conv src {i}: {tuple(x.shape)} vs out {hout}x{wout} stride {stride}"
        arr[i].x, arr[i].C, arr[i].taps, arr[i].stride = x.data_ptr(), x.shape[3], taps, stride
    dev = srcs[0][0].device
    if out_f32:
        if out is None:
            out = torch.empty((B, cout, hout, wout), dtype=torch.float32, device=dev)
        o_bf, o_f = None, ptr(out)
    else:
        if out is None:
            out = torch.empty((B, hout, wout, cout), dtype=BF16, device=dev)
        o_bf, o_f = ptr(out), None
    if residual is not None:
        _nhwc_check(residual)
        assert tuple(residual.shape) == (B, hout, wout, cout)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == cout
    assert w_packed.dtype == BF16 and w_packed.is_contiguous() and w_packed.shape[0] >= padded_rows(cout)
    check(_L().s2s_conv_fwd(arr, len(srcs), B, hout, wout, ptr(w_packed), w_packed.shape[1], cout, ptr(bias),
                            ptr(residual), o_bf, o_f, ptr(axpy_x), float(axpy_a), stream_ptr()), "conv_fwd")
    return out


def conv_wgrad(dy: torch.Tensor, x: torch.Tensor, taps: int, stride: int, dw: torch.Tensor, n_off: int = 0):
    """dw[tap][m][n_off+n] += sum dy[..., m] * x[shifted, n].  dw: fp32 [taps, Cm, ldn]."""
    _nhwc_check(dy)
    _nhwc_check(x)
    B, ho, wo, cm = dy.shape
    assert x.shape[1] == ho * stride and x.shape[2] == wo * stride
    assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.shape[0] == taps and dw.shape[1] == cm
    check(_L().s2s_conv_wgrad(ptr(dy), cm, ptr(x), x.shape[3], taps, stride, B, ho, wo, ptr(dw), dw.shape[2], n_off,
                              stream_ptr()), "conv_wgrad")


def unpack_wgrad(dw: torch.Tensor, grad: torch.Tensor, n_off: int, n_count: int, n_begin: int, beta: float):
    taps, m, ldn = dw.shape
    cin_total = grad.shape[1]
    assert grad.dtype == torch.float32 and grad.is_contiguous() and grad.shape[0] == m
    check(_L().s2s_unpack_wgrad(ptr(dw), taps, m, ldn, n_off, n_count, ptr(grad), cin_total, n_begin, float(beta),
                                stream_ptr()), "unpack_wgrad")


def patch27_pack(x0: torch.Tensor, sgn: int = 1, x1: Optional[torch.Tensor] = None, t: Optional[torch.Tensor] = None,
                 want_xt: bool = False):
    """fp32 NCHW [B,3,H,W] -> bf16 NHWC [B,H,W,64] 3x3 patches (optionally of the FM interpolant)."""
    assert x0.dtype == torch.float32 and x0.is_contiguous() and x0.shape[1] == 3
    B, _, H, W = x0.shape
    dst = torch.empty((B, H, W, 64), dtype=BF16, device=x0.device)
    xt = torch.empty_like(x0) if want_xt else None
    if x1 is not None:
        assert x1.shape == x0.shape and x1.is_contiguous() and t.dtype == torch.float32 and t.numel() == B
    check(_L().s2s_patch27_pack(ptr(x0), ptr(x1), ptr(t), B, H, W, sgn, ptr(dst), ptr(xt), stream_ptr()), "patch27_pack")
    return (dst, xt) if want_xt else dst


def gn_stats(x: torch.Tensor, stats: torch.Tensor, c_off: int = 0):
    _nhwc_check(x)
    B, H, W, Cc = x.shape
    check(_L().s2s_gn_stats(ptr(x), B, H * W, Cc, ptr(stats), stats.shape[1], c_off, stream_ptr()), "gn_stats")


def gn_coef(stats, gamma, beta, film, HW: int, G: int = 32, eps: float = 1e-5):
    B, Cc, _ = stats.shape
    coef = torch.empty((B, Cc, 2), dtype=torch.float32, device=stats.device)
    mr = torch.empty((B, G, 2), dtype=torch.float32, device=stats.device)
    check(_L().s2s_gn_coef(ptr(stats), ptr(gamma), ptr(beta), ptr(film), B, Cc, G, HW, eps, ptr(coef), ptr(mr),
                           stream_ptr()), "gn_coef")
    return coef, mr


def gn_apply(x, coef, y, c_off: int, silu: bool, drop_p: float = 0.0, seed: int = 0):
    _nhwc_check(x)
    B, H, W, Cc = x.shape
    check(_L().s2s_gn_apply(ptr(x), B, H * W, Cc, ptr(coef), coef.shape[1], c_off, ptr(y), y.shape[3], int(silu),
                            float(drop_p), int(seed), stream_ptr()), "gn_apply")


def gn_bwd_reduce(x, g, coef, mr, red, c_off: int, silu: bool, drop_p: float = 0.0, seed: int = 0):
    B, H, W, Cc = x.shape
    check(_L().s2s_gn_bwd_reduce(ptr(x), ptr(g), g.shape[3], B, H * W, Cc, ptr(coef), ptr(mr), mr.shape[1],
                                 coef.shape[1], c_off, ptr(red), int(silu), float(drop_p), int(seed), stream_ptr()),
          "gn_bwd_reduce")


def gn_bwd_coef(red, mr, gamma, beta, film, HW: int, dgamma, dbeta, want_dfilm: bool):
    B, Cc, _ = red.shape
    pqr = torch.empty((B, Cc, 4), dtype=torch.float32, device=red.device)
    dfilm = torch.empty((B, 2 * Cc), dtype=torch.float32, device=red.device) if want_dfilm else None
    check(_L().s2s_gn_bwd_coef(ptr(red), ptr(mr), ptr(gamma), ptr(beta), ptr(film), B, Cc, mr.shape[1], HW, ptr(pqr),
                               ptr(dgamma), ptr(dbeta), ptr(dfilm), stream_ptr()), "gn_bwd_coef")
    return pqr, dfilm


def gn_bwd_apply(x, g, coef, pqr, c_off: int, add, dx, silu: bool, drop_p: float = 0.0, seed: int = 0):
    B, H, W, Cc = x.shape
    check(_L().s2s_gn_bwd_apply(ptr(x), ptr(g), g.shape[3], B, H * W, Cc, ptr(coef), ptr(pqr), coef.shape[1], c_off,
                                ptr(add), ptr(dx), int(silu), float(drop_p), int(seed), stream_ptr()), "gn_bwd_apply")


def upsample2x(x):
    _nhwc_check(x)
    B, H, W, Cc = x.shape
    out = torch.empty((B, 2 * H, 2 * W, Cc), dtype=BF16, device=x.device)
    check(_L().s2s_upsample2x(ptr(x), ptr(out), B, H, W, Cc, stream_ptr()), "upsample2x")
    return out


def sumpool2x(x):
    _nhwc_check(x)
    B, H2, W2, Cc = x.shape
    out = torch.empty((B, H2 // 2, W2 // 2, Cc), dtype=BF16, device=x.device)
    check(_L().s2s_sumpool2x(ptr(x), ptr(out), B, H2 // 2, W2 // 2, Cc, stream_ptr()), "sumpool2x")
    return out


def zero_insert2x(x):
    _nhwc_check(x)
    B, H, W, Cc = x.shape
    out = torch.empty((B, 2 * H, 2 * W, Cc), dtype=BF16, device=x.device)
    check(_L().s2s_zero_insert2x(ptr(x), ptr(out), B, H, W, Cc, stream_ptr()), "zero_insert2x")
    return out


def channel_sum(x, out):
    _nhwc_check(x)
    npix = x.shape[0] * x.shape[1] * x.shape[2]
    check(_L().s2s_channel_sum(ptr(x), npix, x.shape[3], ptr(out), stream_ptr()), "channel_sum")


def fm_loss(v, x0, x1, want_grad: bool):
    assert v.dtype == torch.float32 and v.is_contiguous() and x0.is_contiguous() and x1.is_contiguous()
    loss = torch.zeros((), dtype=torch.float32, device=v.device)
    dv = torch.empty_like(v) if want_grad else None
    check(_L().s2s_fm_loss(ptr(v), ptr(x0), ptr(x1), v.numel(), ptr(loss), ptr(dv), stream_ptr()), "fm_loss")
    return loss, dv


def nchw_to_nhwc_bf16(x):
    assert x.dtype == torch.float32 and x.is_contiguous()
    B, Cc, H, W = x.shape
    out = torch.empty((B, H, W, Cc), dtype=BF16, device=x.device)
    check(_L().s2s_nchw_f32_to_nhwc_bf16(ptr(x), ptr(out), B, Cc, H * W, stream_ptr()), "nchw_to_nhwc")
    return out


def nhwc_to_nchw_f32(x):
    _nhwc_check(x)
    B, H, W, Cc = x.shape
    out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    check(_L().s2s_nhwc_bf16_to_nchw_f32(ptr(x), ptr(out), B, Cc, H * W, stream_ptr()), "nhwc_to_nchw")
    return out
