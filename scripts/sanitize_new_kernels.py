"""Small driver for `compute-sanitizer --tool memcheck`: one launch of every kernel rewritten in round 2's second session, on
shapes with ragged tails (odd pixel counts per CTA, partial bulk-copy tiles, partial pack tiles, channel-sliced gradients).
Usage (GPU box): compute-sanitizer --tool memcheck python scripts/sanitize_new_kernels.py
(compute-sanitizer is closed on this pool -- the plain run, every launch completing on the ragged shapes, is what was done here;
the layouts themselves are pinned bit-exactly, with canaries around the destinations, in tests/test_gpu_kernels.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from stain2stain_b200 import kernels as K  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(3)


def r16(shape, fmt):
    return K.from_float(torch.randn(shape, device=dev, generator=g), fmt)


# ---- streaming norm kernels: (B, H, W, C): bulk path (HW >= 16384, contiguous), ragged chunk tails, small levels, C = 384 (240-thread CTAs)
for B, H, W, C in [(1, 128, 137, 128), (2, 40, 24, 256), (1, 16, 16, 512), (1, 24, 40, 384), (1, 128, 128, 64)]:
    HW = H * W
    x, gr = r16((B, H, W, C), K.ACT), r16((B, H, W, C), K.GRAD)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    film = torch.randn(B, 2 * C, device=dev, generator=g) * 0.1
    stats = K.gn_partial_buffer(B, HW, C, dev)
    K.gn_stats(x, stats, 0)
    coef, mr = K.gn_coef(stats, gamma, beta, film, HW)
    y, y2, dx = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    mask = torch.empty((B, H, W, C // 8), dtype=torch.uint8, device=dev)
    for p, m in [(0.0, None), (0.1, None), (0.1, mask)]:
        K.gn_apply(x, coef, y, 0, True, p, 7, y2=y2, mask=m)
        red = K.gn_partial_buffer(B, HW, C, dev)
        K.gn_bwd_reduce(x, gr, coef, mr, red, 0, True, p, 7, mask=m, x_bf16_out=y2 if p == 0.0 else None)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        pqr, _ = K.gn_bwd_coef(red, mr, gamma, beta, film, HW, dg, db, True)
        K.gn_bwd_apply(x, gr, coef, pqr, 0, None, dx, True, p, 7, mask=m)
        K.gn_bwd_apply(x, gr, coef, pqr, 0, gr, dx, True, p, 7, mask=m)
    # channel slice of a wider gradient (the concat case: ld_g = 2C, c_off = C)
    gw = r16((B, H, W, 2 * C), K.GRAD)
    stats2 = K.gn_partial_buffer(B, HW, 2 * C, dev)
    K.gn_stats(x, stats2, 0)
    K.gn_stats(x, stats2, C)
    coef2, mr2 = K.gn_coef(stats2, torch.ones(2 * C, device=dev), torch.zeros(2 * C, device=dev), None, HW)
    red2 = K.gn_partial_buffer(B, HW, 2 * C, dev)
    for off in (0, C):
        K.gn_bwd_reduce(x, gw, coef2, mr2, red2, off, True)
    dg2, db2 = torch.zeros(2 * C, device=dev), torch.zeros(2 * C, device=dev)
    pqr2, _ = K.gn_bwd_coef(red2, mr2, torch.ones(2 * C, device=dev), torch.zeros(2 * C, device=dev), None, HW, dg2, db2, False)
    for off in (0, C):
        K.gn_bwd_apply(x, gw, coef2, pqr2, off, gr, dx, True)
# ---- weight packing (ragged tiles), unpack
for cout, cin, taps, cb, cc in [(96, 72, 9, 8, 40), (40, 24, 9, 0, 24), (64, 200, 1, 64, 136), (128, 128, 9, 0, 128)]:
    w = torch.randn((cout, cin, 3, 3) if taps == 9 else (cout, cin, 1, 1), device=dev, generator=g)
    jobs = []
    for tf in (False, True):
        for mode in ([0, 1, 4] if taps == 9 else [0]):
            lt = taps if mode == 0 else 4
            rows, inner = (cc, cout) if tf else (cout, cc)
            dst = torch.zeros((rows, 64 + lt * inner), dtype=K.T16, device=dev)
            K.pack_conv_weight(w, dst, k_off=64, ci_begin=cb, ci_count=cc, transpose_flip=tf, fmt=K.GRAD if tf else K.ACT, mode=mode)
            jobs.append((w, dst, 64, cb, cc, tf, K.GRAD if tf else K.ACT, mode))
    K.pack_conv_weight_multi(jobs, {})
    dw = torch.randn((taps, cout, cc), device=dev, generator=g)
    grad = torch.zeros_like(w)
    K.unpack_wgrad(dw, grad.view(cout, cin, -1), 0, cc, cb, 0.0)
# ---- multitask: folds, layout converter
parts = torch.randn(3, 700, 72, 2, device=dev, generator=g)
for sl in (1, 5, 8):
    K.bn_fold(parts, sl)
K.nchw_to_nhwc16_pad(torch.randn(2, 5, 24, 40, device=dev, generator=g), 64)
K.nchw_to_nhwc16_pad(torch.randn(1, 3, 16, 16, device=dev, generator=g), 8)
torch.cuda.synchronize()
print("sanitize_new_kernels: all launches completed")
