#!/usr/bin/env python
"""Per-kernel micro-benchmark of the C-ABI kernels on the config-A layer shapes (CUDA events on the launching stream,
inputs larger than L2).  Also the small, fast target for `ncu --set full` captures:

    python scripts/kbench.py [conv] [wgrad] [gn] [--batch 32] [--iters 5] [--json out.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from stain2stain_b200 import kernels as K  # noqa: E402

DEV = "cuda"
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


ONCE = False  # --once: every kernel is launched exactly once (the ncu capture target)


def timeit(fn, iters, warm=2):
    if ONCE:
        warm, iters = 0, 1
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def rnd16(shape, fmt, scale=1.0):
    x = torch.randn(shape, device=DEV) * scale
    return K.from_float(x, fmt)


# (Cin, Cout, H) 3x3 stride-1 shapes of config A, heaviest first
CONV_SHAPES = [(128, 128, 256), (256, 128, 256), (256, 256, 256), (256, 256, 128), (512, 256, 128), (256, 256, 64),
               (512, 512, 32), (1024, 512, 32)]


def bench_conv(B, iters, out):
    for cin, cout, H in CONV_SHAPES:
        x = rnd16((B, H, H, cin), K.ACT)
        w = torch.randn(cout, cin, 3, 3, device=DEV) / (3 * cin ** 0.5)
        wp = torch.zeros((K.padded_rows(cout), 9 * cin), dtype=K.T16, device=DEV)
        K.pack_conv_weight(w, wp)
        bias = torch.zeros(cout, device=DEV)
        y = torch.empty((B, H, H, cout), dtype=K.T16, device=DEV)
        ms = timeit(lambda: K.conv_fwd([(x, 9, 1)], wp, cout, H, H, bias=bias, out=y), iters)
        fl = 2.0 * B * H * H * cout * cin * 9
        tf = fl / ms / 1e9
        out.append(dict(kernel="conv_igemm", shape=f"{cin}->{cout}@{H} B{B}", ms=ms, tflops=tf,
                        frac_burst=tf / PEAKS["bf16_tflops"], frac_sustained=tf / PEAKS["bf16_tflops_sustained"]))
        print(out[-1], flush=True)
        del x, y


def bench_wgrad(B, iters, out):
    for cin, cout, H in CONV_SHAPES:
        x = rnd16((B, H, H, cin), K.GRAD)
        dy = rnd16((B, H, H, cout), K.GRAD)
        dw = torch.zeros((9, cout, cin), dtype=torch.float32, device=DEV)
        ms = timeit(lambda: K.conv_wgrad(dy, x, 9, 1, dw), iters)
        fl = 2.0 * B * H * H * cout * cin * 9
        tf = fl / ms / 1e9
        out.append(dict(kernel="conv_wgrad", shape=f"{cin}->{cout}@{H} B{B}", ms=ms, tflops=tf,
                        frac_burst=tf / PEAKS["bf16_tflops"], frac_sustained=tf / PEAKS["bf16_tflops_sustained"]))
        print(out[-1], flush=True)
        del x, dy


GN_SHAPES = [(128, 256), (256, 128), (384, 256), (512, 32)]


def bench_gn(B, iters, out):
    for C, H in GN_SHAPES:
        HW = H * H
        x = rnd16((B, H, H, C), K.ACT)
        g = rnd16((B, H, H, C), K.GRAD)
        gamma, beta = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
        film = torch.randn(B, 2 * C, device=DEV) * 0.1
        stats = K.gn_partial_buffer(B, HW, C, DEV)
        y = torch.empty_like(x)
        dx = torch.empty_like(x)
        n = x.numel()

        def rec(name, ms, nbytes):
            gbs = nbytes / ms / 1e6
            out.append(dict(kernel=name, shape=f"C{C}@{H} B{B}", ms=ms, gbs=gbs, frac_hbm=gbs / PEAKS["hbm_gbs"]))
            print(out[-1], flush=True)

        rec("gn_stats", timeit(lambda: K.gn_stats(x, stats, 0), iters), 2.0 * n)
        coef, mr = K.gn_coef(stats, gamma, beta, film, HW)
        rec("gn_coef", timeit(lambda: K.gn_coef(stats, gamma, beta, film, HW), iters), 8.0 * stats.numel())
        for p in (0.0, 0.1):
            tag = "_dropout" if p else ""
            rec("gn_apply" + tag, timeit(lambda: K.gn_apply(x, coef, y, 0, True, p, 1234), iters), 4.0 * n)
            red = K.gn_partial_buffer(B, HW, C, DEV)
            rec("gn_bwd_reduce" + tag, timeit(lambda: K.gn_bwd_reduce(x, g, coef, mr, red, 0, True, p, 1234), iters), 4.0 * n)
            dgamma, dbeta = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
            pqr, _ = K.gn_bwd_coef(red, mr, gamma, beta, film, HW, dgamma, dbeta, True)
            rec("gn_bwd_apply" + tag, timeit(lambda: K.gn_bwd_apply(x, g, coef, pqr, 0, None, dx, True, p, 1234), iters), 6.0 * n)
            rec("gn_bwd_apply_add" + tag, timeit(lambda: K.gn_bwd_apply(x, g, coef, pqr, 0, g, dx, True, p, 1234), iters), 8.0 * n)
            if p:
                mask = torch.empty((B, H, H, C // 8), dtype=torch.uint8, device=DEV)
                rec("gn_apply_dropout_dual_mask", timeit(lambda: K.gn_apply(x, coef, y, 0, True, p, 1234, y2=dx, mask=mask), iters), 6.0 * n)
                rec("gn_bwd_reduce_dropout_mask", timeit(lambda: K.gn_bwd_reduce(x, g, coef, mr, red, 0, True, p, 1234, mask=mask), iters), 4.0 * n)
                rec("gn_bwd_apply_dropout_mask", timeit(lambda: K.gn_bwd_apply(x, g, coef, pqr, 0, None, dx, True, p, 1234, mask=mask), iters), 6.0 * n)
        xb = torch.empty_like(x)
        rec("convert16", timeit(lambda: K.convert16(x, K.ACT, K.GRAD), iters), 4.0 * n)
        del x, g, y, dx, xb


def bench_attn(B, iters, out):
    for heads, ch, H in [(16, 32, 32), (4, 32, 16)]:
        C = heads * ch
        qkv = rnd16((B, H, H, 3 * C), K.ACT)
        d_out = rnd16((B, H, H, C), K.GRAD)
        o, lse = K.attn_fwd(qkv, heads, False, True)
        T = H * H
        ms = timeit(lambda: K.attn_fwd(qkv, heads, False, True), iters)
        out.append(dict(kernel="attn_fwd", shape=f"h{heads}x{ch} T{T} B{B}", ms=ms, tflops=4.0 * B * heads * T * T * ch / ms / 1e9))
        print(out[-1], flush=True)
        ms = timeit(lambda: K.attn_bwd(qkv, o, d_out, lse, heads, False), iters)
        out.append(dict(kernel="attn_bwd", shape=f"h{heads}x{ch} T{T} B{B}", ms=ms, tflops=14.0 * B * heads * T * T * ch / ms / 1e9))
        print(out[-1], flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="*", default=["conv", "wgrad", "gn"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--json", default=None)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--gn-shapes", default=None, help="CxH[,CxH...] for the gn group, e.g. 128x256,256x128")
    a = ap.parse_args()
    global ONCE, GN_SHAPES
    ONCE = a.once
    if a.gn_shapes:
        GN_SHAPES = [tuple(int(v) for v in t.split("x")) for t in a.gn_shapes.split(",")]
    out = []
    if "conv" in a.what:
        bench_conv(a.batch, a.iters, out)
    if "wgrad" in a.what:
        bench_wgrad(a.batch, a.iters, out)
    if "gn" in a.what:
        bench_gn(a.batch, a.iters, out)
    if "attn" in a.what:
        bench_attn(a.batch, a.iters, out)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
