"""Autograd operators of the hot path, each a fused forward/backward over the C-ABI kernels (kernels.py).

Activations between operators are bf16 NHWC CUDA tensors; parameters stay fp32 `nn.Parameter`s in the reference's
layouts (so Lightning checkpoints, torch optimizers and DDP see exactly what they see in the reference) and are packed
to bf16 GEMM operands on the fly (cached while the parameter is unchanged).
"""
from __future__ import annotations

import weakref
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import kernels as K

T16 = K.T16


# --------------------------------------------------------------------------------------------- packed-weight cache
class _PackCache:
    """16-bit GEMM operands derived from fp32 parameters, keyed by (purpose, parameter identity and version).

    `make(dst)` returns `(operand, jobs)`: the destination tensor (allocated zero-filled when `dst` is None, else the
    previous operand of the same parameters -- an optimizer step only changed the values, the zero padding is never
    written) and the pack jobs `(index of the weight in ws, k_off, ci_begin, ci_count, transpose_flip, fmt[, mode])` that fill it
    (mode: 0 plain, 1 + phase = tap-summed operand of the phase-decomposed Upsample conv).  When a stale
    operand is requested, EVERY stale operand of the cache is re-packed by one multi-tensor launch
    (`s2s_pack_conv_weight_multi`): after an optimizer step that is all 158 operands of the model at once."""

    def __init__(self):
        self._store: Dict[tuple, list] = {}   # key -> [sig, operand, weakrefs to the weights, jobs]
        self._tables: Dict[tuple, tuple] = {}  # job-set signature -> (device job table, device work list, n_work, keep-alive)

    @staticmethod
    def _sig(ws: Sequence[torch.Tensor]):
        return tuple((w.data_ptr(), w._version, tuple(w.shape)) for w in ws)

    def get(self, key: tuple, ws: Sequence[torch.Tensor], make):
        sig = self._sig(ws)
        hit = self._store.get(key)
        if hit is not None and not all(r() is w for r, w in zip(hit[2], ws)):
            # same key, other tensor objects: a freed model's id / allocator address / version count can all recur, so a
            # hit is only ever accepted for the very parameters it was packed from
            del self._store[key]
            hit = None
        if hit is not None and hit[0] == sig:
            return hit[1]
        if hit is not None and self._reusable(hit, sig, ws[0].device):
            self._refresh_stale(ws[0].device)
            hit = self._store.get(key)
            if hit is not None and hit[0] == sig:
                return hit[1]
        val, jobs = make(None)
        self._run_jobs(val, jobs, ws)
        self._store[key] = [sig, val, [weakref.ref(w) for w in ws], jobs]  # jobs name weights by index: no strong refs
        return val

    @staticmethod
    def _reusable(hit, sig, device) -> bool:
        return len(hit[0]) == len(sig) and all(a[0] == b[0] and a[2] == b[2] for a, b in zip(hit[0], sig)) and \
            hit[1].device == device

    @staticmethod
    def _run_jobs(dst, jobs, ws):
        for (wi, k_off, ci_begin, ci_count, tf, fmt, *mode) in jobs:
            K.pack_conv_weight(ws[wi].detach(), dst, k_off=k_off, ci_begin=ci_begin, ci_count=ci_count, transpose_flip=tf,
                               fmt=fmt, mode=mode[0] if mode else 0)

    def _refresh_stale(self, device):
        """Re-pack, in one launch, every operand on `device` whose parameters changed in place (same storage, new version)."""
        stale, dead = [], []
        for key, ent in self._store.items():
            ws = [r() for r in ent[2]]
            if any(w is None for w in ws):
                dead.append(key)
                continue
            sig = self._sig(ws)
            if sig != ent[0] and ent[1].device == device and self._reusable(ent, sig, device):
                stale.append((key, ent, ws, sig))
        for key in dead:
            del self._store[key]
        if not stale:
            return
        jobs = []
        for key, ent, ws, sig in stale:
            for (wi, k_off, ci_begin, ci_count, tf, fmt, *mode) in ent[3]:
                jobs.append((ws[wi].detach(), ent[1], k_off, ci_begin, ci_count, tf, fmt, mode[0] if mode else 0))
        K.pack_conv_weight_multi(jobs, self._tables)
        for key, ent, ws, sig in stale:
            ent[0] = sig

    def clear(self):
        self._store.clear()
        self._tables.clear()


PACK_CACHE = _PackCache()


@dataclass(frozen=True)
class Seg:
    """One GEMM segment of a fused conv: source index, weight index, input-channel slice, filter taps, stride."""
    src: int
    weight: int
    ci_begin: int
    ci_count: int
    taps: int
    stride: int = 1


@dataclass
class ConvPlan:
    segs: Tuple[Seg, ...]
    cout: int
    uid: int = field(default_factory=lambda: ConvPlan._next_uid())
    _counter = [0]

    @staticmethod
    def _next_uid():
        ConvPlan._counter[0] += 1
        return ConvPlan._counter[0]

    def ktot(self):
        return sum(s.taps * ((s.ci_count + 63) // 64 * 64) for s in self.segs)

    def packed_fwd(self, weights: Sequence[torch.Tensor]) -> torch.Tensor:
        def make(dst):
            wp = dst if dst is not None else \
                torch.zeros((K.padded_rows(self.cout), self.ktot()), dtype=T16, device=weights[0].device)
            jobs, off = [], 0
            for s in self.segs:
                jobs.append((s.weight, off, s.ci_begin, s.ci_count, False, K.ACT))
                off += s.taps * ((s.ci_count + 63) // 64 * 64)
            return wp, jobs
        return PACK_CACHE.get(("fwd", self.uid), weights, make)

    def packed_dgrad(self, si: int, weights: Sequence[torch.Tensor]) -> torch.Tensor:
        s = self.segs[si]
        w = weights[s.weight]

        def make(dst):
            wd = dst if dst is not None else torch.zeros((s.ci_count, s.taps * self.cout), dtype=T16, device=w.device)
            return wd, [(0, 0, s.ci_begin, s.ci_count, True, K.GRAD)]
        return PACK_CACHE.get(("dgrad", self.uid, si), [w], make)


# --------------------------------------------------------------------------------------------- fused convolution
def usable_stats(src_stats) -> bool:
    """Epilogue statistics of every source are present and cut the samples into the same number of sub-tiles."""
    return len(src_stats) <= 2 and all(st is not None for st in src_stats) and \
        all(st.shape[1] == src_stats[0].shape[1] for st in src_stats)


class _FusedConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: ConvPlan, n_src: int, n_w: int, residual, boxes, *tensors):
        stats_box, twins = boxes if isinstance(boxes, tuple) else (boxes, [None] * n_src)
        srcs = tensors[:n_src]
        weights = tensors[n_src:n_src + n_w]
        biases = [b for b in tensors[n_src + n_w:] if b is not None]
        B = srcs[0].shape[0]
        s0 = plan.segs[0]
        hout, wout = srcs[s0.src].shape[1] // s0.stride, srcs[s0.src].shape[2] // s0.stride
        wp = plan.packed_fwd(weights)
        bias = None
        if biases:
            bias = biases[0].detach().float()
            for b in biases[1:]:
                bias = bias + b.detach()
            bias = bias.contiguous()
        out, st = K.conv_fwd([(srcs[s.src], s.taps, s.stride) for s in plan.segs], wp, plan.cout, hout, wout, bias=bias,
                             residual=residual, want_stats=True)
        if stats_box is not None:
            stats_box.append(st)
        ctx.plan, ctx.n_src, ctx.n_w = plan, n_src, n_w
        ctx.has_res = residual is not None
        ctx.bias_present = [b is not None for b in tensors[n_src + n_w:]]
        # a source whose producer already wrote its bf16 twin is saved AS the twin (the fp16 tensor is not needed again:
        # dgrad reads only d_out and the weights), so backward has no conversion pass and memory does not grow
        ctx.twin = [tw is not None and K.ACT != K.GRAD for tw in twins]
        ctx.save_for_backward(*[tw if has else x for x, tw, has in zip(srcs, twins, ctx.twin)], *weights)
        return out

    @staticmethod
    def backward(ctx, d_out):
        plan, n_src, n_w = ctx.plan, ctx.n_src, ctx.n_w
        saved = ctx.saved_tensors
        srcs, weights = saved[:n_src], saved[n_src:]
        d_out = d_out.contiguous()
        B, hout, wout, cout = d_out.shape
        need = ctx.needs_input_grad  # (plan, n_src, n_w, residual, stats_box, *tensors)
        need = need[:4] + need[5:]
        d_res = d_out if (ctx.has_res and need[3]) else None
        d_srcs: List[Optional[torch.Tensor]] = [None] * n_src
        d_ws: List[Optional[torch.Tensor]] = [None] * n_w
        # bias gradient (shared by every bias folded into this accumulator)
        d_bias = None
        if any(ctx.bias_present) and any(need[4 + n_src + n_w:]):
            d_bias = torch.zeros(cout, dtype=torch.float32, device=d_out.device)
            K.channel_sum(d_out, d_bias)
        for si, s in enumerate(plan.segs):
            x = srcs[s.src]
            w = weights[s.weight]
            if need[4 + n_src + s.weight]:
                if d_ws[s.weight] is None:
                    d_ws[s.weight] = torch.empty_like(w, dtype=torch.float32)
                    if sum(t.ci_count for t in plan.segs if t.weight == s.weight) != w.shape[1]:
                        d_ws[s.weight].zero_()
                dw = torch.zeros((s.taps, cout, x.shape[3]), dtype=torch.float32, device=d_out.device)
                K.conv_wgrad(d_out, x if ctx.twin[s.src] else K.convert16(x, K.ACT, K.GRAD), s.taps, s.stride, dw)
                K.unpack_wgrad(dw, d_ws[s.weight].view(cout, w.shape[1], -1), 0, s.ci_count, s.ci_begin, 0.0)
            if need[4 + s.src]:
                wd = plan.packed_dgrad(si, weights)
                if s.stride == 2 and s.taps == 9 and K.upconv_supported(cout, s.ci_count):
                    # phase-decomposed transposed conv: no zero-inserted tensor, the 9 algorithmic taps exactly
                    dx = K.downconv_dgrad(d_out, wd, s.ci_count)
                    d_srcs[s.src] = dx if d_srcs[s.src] is None else d_srcs[s.src] + dx
                    continue
                g = d_out if s.stride == 1 else K.zero_insert2x(d_out)
                dx = K.conv_fwd([(g, s.taps, 1)], wd, s.ci_count, g.shape[1], g.shape[2], a_fmt=K.GRAD, w_fmt=K.GRAD,
                                out_fmt=K.GRAD, alg_macs=float(B) * hout * wout * cout * s.ci_count * s.taps)
                d_srcs[s.src] = dx if d_srcs[s.src] is None else d_srcs[s.src] + dx
        d_biases = [d_bias if (present and n) else None
                    for present, n in zip(ctx.bias_present, need[4 + n_src + n_w:])]
        return (None, None, None, d_res, None, *d_srcs, *d_ws, *d_biases)


def _tag(out, box):
    """Attach the conv epilogue's GroupNorm statistics to the Python tensor object that travels on (never keyed by
    address: a freed-and-reused buffer must not inherit stale statistics)."""
    out._s2s_stats = box[0] if box else None
    return out


def stats_of(t):
    return getattr(t, "_s2s_stats", None)


def twin_of(t):
    """bf16 copy of a forward-format activation written by its producer (a norm kernel's dual-format store): the operand the
    consuming conv's weight-gradient GEMM needs, so backward does not run a conversion pass over the tensor."""
    return getattr(t, "_s2s_g16", None)


def fused_conv(plan: ConvPlan, srcs: Sequence[torch.Tensor], weights: Sequence[torch.Tensor],
               biases: Sequence[Optional[torch.Tensor]], residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    box = []
    twins = [twin_of(t) for t in srcs] if torch.is_grad_enabled() else [None] * len(srcs)
    return _tag(_FusedConv.apply(plan, len(srcs), len(weights), residual, (box, twins), *srcs, *weights, *biases), box)


# --------------------------------------------------------------------------------------------- GroupNorm (+FiLM+SiLU+dropout, +concat)
class _GroupNormAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, n_src: int, groups: int, eps: float, silu: bool, drop_p: float, seed: int, src_stats, gamma, beta,
                film, *srcs):
        B, H, W, _ = srcs[0].shape
        ctot = sum(s.shape[3] for s in srcs)
        dev = srcs[0].device
        film_c = film.detach().float().contiguous() if film is not None else None
        if src_stats is not None and usable_stats(src_stats):
            coef, mr = K.gn_coef_parts(list(src_stats), gamma.detach(), beta.detach(), film_c, H * W, groups, eps)
        else:
            stats = K.gn_partial_buffer(B, H * W, ctot, dev)
            off = 0
            for s in srcs:
                K.gn_stats(s, stats, off)
                off += s.shape[3]
            coef, mr = K.gn_coef(stats, gamma.detach(), beta.detach(), film_c, H * W, groups, eps)
        y = torch.empty((B, H, W, ctot), dtype=T16, device=dev)
        off = 0
        for s in srcs:
            K.gn_apply(s, coef, y, off, silu, drop_p, seed)
            off += s.shape[3]
        ctx.cfg = (n_src, groups, silu, drop_p, seed, film is not None)
        ctx.save_for_backward(coef, mr, gamma, beta, film_c, *srcs)
        return y

    @staticmethod
    def backward(ctx, g):
        n_src, groups, silu, drop_p, seed, has_film = ctx.cfg
        coef, mr, gamma, beta, film_c, *srcs = ctx.saved_tensors
        g = g.contiguous()
        B, H, W, ctot = g.shape
        dev = g.device
        red = K.gn_partial_buffer(B, H * W, ctot, dev)
        off = 0
        for s in srcs:
            K.gn_bwd_reduce(s, g, coef, mr, red, off, silu, drop_p, seed)
            off += s.shape[3]
        dgamma = torch.zeros(ctot, dtype=torch.float32, device=dev)
        dbeta = torch.zeros(ctot, dtype=torch.float32, device=dev)
        pqr, dfilm = K.gn_bwd_coef(red, mr, gamma.detach(), beta.detach(), film_c, H * W, dgamma, dbeta, has_film)
        dxs = []
        off = 0
        for i, s in enumerate(srcs):
            if ctx.needs_input_grad[10 + i]:
                dx = torch.empty_like(s)
                K.gn_bwd_apply(s, g, coef, pqr, off, None, dx, silu, drop_p, seed)
                dxs.append(dx)
            else:
                dxs.append(None)
            off += s.shape[3]
        return (None, None, None, None, None, None, None, dgamma, dbeta, dfilm, *dxs)


def group_norm_act(srcs: Sequence[torch.Tensor], gamma, beta, film=None, silu=True, drop_p=0.0, seed=0, groups=32,
                   eps=1e-5) -> torch.Tensor:
    """GroupNorm over the channel-concatenation of `srcs`, optional FiLM `(1+scale), shift`, SiLU and dropout.
    Statistics come from the producing convs' epilogues when every source carries them."""
    src_stats = tuple(stats_of(t) for t in srcs)
    return _GroupNormAct.apply(len(srcs), groups, eps, bool(silu), float(drop_p), int(seed), src_stats, gamma, beta,
                               film, *srcs)


# --------------------------------------------------------------------------------------------- resampling
class _Upsample2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return K.upsample2x(x)

    @staticmethod
    def backward(ctx, g):
        return K.sumpool2x(g.contiguous())


def upsample2x(x):
    return _Upsample2x.apply(x)


class UpConvPlan:
    """Packed operands of one Upsample conv (nearest x2 -> conv3x3) in its phase-decomposed form."""
    _counter = [0]

    def __init__(self, cin: int, cout: int):
        UpConvPlan._counter[0] += 1
        self.uid, self.cin, self.cout = UpConvPlan._counter[0], cin, cout

    def packed_fwd(self, w):
        def make(dst):
            wp = dst if dst is not None else torch.zeros((self.cout, 16 * self.cin), dtype=T16, device=w.device)
            return wp, [(0, ph * 4 * self.cin, 0, self.cin, False, K.ACT, 1 + ph) for ph in range(4)]
        return PACK_CACHE.get(("upfwd", self.uid), [w], make)

    def packed_dgrad(self, w):
        def make(dst):
            wd = dst if dst is not None else torch.zeros((self.cin, 16 * self.cout), dtype=T16, device=w.device)
            return wd, [(0, ph * 4 * self.cout, 0, self.cin, True, K.GRAD, 1 + ph) for ph in range(4)]
        return PACK_CACHE.get(("updgrad", self.uid), [w], make)


class _UpConv(torch.autograd.Function):
    """torchcfm `Upsample.forward` with use_conv: F.interpolate(x, 2, "nearest") -> conv3x3, without the 4x tensor and at
    4/9 of its MACs (four 2x2 convs over the low-resolution input, one per output pixel phase)."""

    @staticmethod
    def forward(ctx, plan: UpConvPlan, stats_box, x, w, b):
        out, st = K.upconv_fwd(x, plan.packed_fwd(w), plan.cout, b.detach(), want_stats=True)
        stats_box.append(st)
        ctx.plan = plan
        ctx.save_for_backward(x, w)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, w = ctx.saved_tensors
        plan = ctx.plan
        d_out = d_out.contiguous()
        need = ctx.needs_input_grad
        d_x = d_w = d_b = None
        if need[4]:
            d_b = torch.zeros(plan.cout, dtype=torch.float32, device=d_out.device)
            K.channel_sum(d_out, d_b)
        if need[3]:
            d_w = K.upconv_wgrad(d_out, K.convert16(x, K.ACT, K.GRAD))
        if need[2]:
            d_x = K.upconv_dgrad(d_out, plan.packed_dgrad(w), plan.cin)
        return None, None, d_x, d_w, d_b


def upsample_conv(plan: UpConvPlan, x, w, b):
    box = []
    return _tag(_UpConv.apply(plan, box, x, w, b), box)


# --------------------------------------------------------------------------------------------- stem / head (3-channel ends)
class _Stem(torch.autograd.Function):
    """3x3 conv on the fp32 NCHW image (optionally the FM interpolant of x0, x1 at t; optionally with a condition
    channel `extra` appended, never materialised as a concat) -> 16-bit NHWC features."""

    @staticmethod
    def forward(ctx, x0, x1, t, w, b, stats_box, extra):
        cout, cin = w.shape[0], w.shape[1]
        x1c = None if x1 is None else x1.contiguous()
        tc = None if t is None else t.float().contiguous()
        if cin == 3 and extra is None:
            patches = K.patch27_pack(x0.contiguous(), 1, x1c, tc)
        else:
            assert x0.shape[1] + (extra is not None) == cin, "stem: input channels do not match the conv weight"
            patches = K.patch_pack(x0.contiguous(), x1c, tc, None if extra is None else extra.float().contiguous())

        def make(dst):
            wp = dst if dst is not None else torch.zeros((cout, 64), dtype=T16, device=w.device)
            return wp, [(0, 0, 0, w.shape[1], False, K.ACT)]
        wp = PACK_CACHE.get(("stem", id(w)), [w], make)
        B, _, H, W = x0.shape
        out, st = K.conv_fwd([(patches, 1, 1)], wp, cout, H, W, bias=b.detach(), want_stats=True,
                             alg_macs=float(B) * H * W * cout * 9 * cin)  # the operand is padded 9*cin -> 64 columns
        if stats_box is not None:
            stats_box.append(st)
        ctx.save_for_backward(patches)
        ctx.wshape = tuple(w.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        (patches,) = ctx.saved_tensors
        g = g.contiguous()
        cout, cin = g.shape[3], ctx.wshape[1]
        dw = torch.zeros((1, cout, 64), dtype=torch.float32, device=g.device)
        K.conv_wgrad(g, K.convert16(patches, K.ACT, K.GRAD), 1, 1, dw,
                     alg_macs=float(g.shape[0]) * g.shape[1] * g.shape[2] * cout * 9 * cin)
        d_w = dw[0, :, :9 * cin].reshape(cout, 9, cin).permute(0, 2, 1).reshape(ctx.wshape).contiguous()
        d_b = torch.zeros(cout, dtype=torch.float32, device=g.device)
        K.channel_sum(g, d_b)
        return None, None, None, d_w, d_b, None, None


def stem_conv(x0, w, b, x1=None, t=None, extra=None):
    assert w.shape[1] in (3, 4) and tuple(w.shape[2:]) == (3, 3), "stem expects a 3- or 4-channel 3x3 conv"
    box = []
    return _tag(_Stem.apply(x0, x1, t, w, b, box, extra), box)


class _HeadConv(torch.autograd.Function):
    """3x3 conv to <= 3 channels, fp32 NCHW output; optional fused `axpy_x + a * v` (Euler update, inference only)."""

    @staticmethod
    def forward(ctx, a, w, b, axpy_x, axpy_a):
        cout, cin = w.shape[0], w.shape[1]
        if K.head_conv_supported(cin, cout):  # dedicated kernel: the input is read once, not once per filter tap
            out = K.head_conv(a, w.detach(), b.detach(), axpy_x=axpy_x, axpy_a=axpy_a)
            ctx.save_for_backward(a, w)
            return out

        def make(dst):
            wp = dst if dst is not None else torch.zeros((16, 9 * cin), dtype=T16, device=w.device)
            return wp, [(0, 0, 0, cin, False, K.ACT)]
        wp = PACK_CACHE.get(("head", id(w)), [w], make)
        B, H, W, _ = a.shape
        out = K.conv_fwd([(a, 9, 1)], wp, cout, H, W, bias=b.detach(), out_f32=True, axpy_x=axpy_x, axpy_a=axpy_a,
                         out=axpy_x if axpy_x is not None else None)
        ctx.save_for_backward(a, w)
        return out

    @staticmethod
    def backward(ctx, dv):
        a, w = ctx.saved_tensors
        cout, cin = w.shape[0], w.shape[1]
        assert cout == 3
        dv = dv.contiguous().float()
        dcol = K.patch27_pack(dv, -1, fmt=K.GRAD)  # [B,H,W,64], column tap*3 + co = dv[q - tap_off][co]
        # dgrad: da[q][ci] = sum_j dcol[q][j] * Wd[ci][j], Wd[ci][tap*3+co] = w[co][ci][tap]
        wd = torch.zeros((cin, 64), dtype=torch.bfloat16, device=w.device)
        wd[:, :27] = w.detach().permute(1, 2, 3, 0).reshape(cin, 27).to(torch.bfloat16)
        B, H, W, _ = a.shape
        da = K.conv_fwd([(dcol, 1, 1)], wd, cin, H, W, a_fmt=K.GRAD, w_fmt=K.GRAD, out_fmt=K.GRAD,
                        alg_macs=float(B) * H * W * cin * 27)
        # wgrad: D[ci][tap*3+co] = sum_q a[q][ci] * dcol[q][tap*3+co]
        dw = torch.zeros((1, cin, 64), dtype=torch.float32, device=w.device)
        K.conv_wgrad(K.convert16(a, K.ACT, K.GRAD), dcol, 1, 1, dw, alg_macs=float(B) * H * W * cin * 27)
        d_w = dw[0, :, :27].reshape(cin, 9, 3).permute(2, 0, 1).reshape(cout, cin, 3, 3).contiguous()
        d_b = dv.sum(dim=(0, 2, 3))
        return da, d_w, d_b, None, None


def head_conv(a, w, b, axpy_x=None, axpy_a=0.0):
    return _HeadConv.apply(a, w, b, axpy_x, axpy_a)


# --------------------------------------------------------------------------------------------- loss
class _FMLoss(torch.autograd.Function):
    """mean((v - (x1 - x0))^2) with the gradient produced in the same pass."""

    @staticmethod
    def forward(ctx, v, x0, x1):
        loss, dv = K.fm_loss(v.contiguous(), x0.contiguous(), x1.contiguous(), True)
        ctx.save_for_backward(dv)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dv,) = ctx.saved_tensors
        return dv * g, None, None


def fm_loss(v, x0, x1):
    return _FMLoss.apply(v, x0, x1)


class _FMLossWeighted(torch.autograd.Function):
    """sum(w (v - (x1 - x0))^2) / (sum(w) + 1e-8), w = 1 + lam * mask (conditional_flow_matching_masked.py:76-92)."""

    @staticmethod
    def forward(ctx, v, x0, x1, mask, lam):
        sums, dv = K.fm_loss_weighted(v.contiguous(), x0.contiguous(), x1.contiguous(), mask.contiguous(), lam, True)
        inv = 1.0 / (sums[1] + 1e-8)
        ctx.save_for_backward(dv, inv)
        return sums[0] * inv

    @staticmethod
    def backward(ctx, g):
        dv, inv = ctx.saved_tensors
        return dv * (g * inv), None, None, None, None


def fm_loss_weighted(v, x0, x1, mask, lam: float):
    return _FMLossWeighted.apply(v, x0, x1, mask, float(lam))


def roi_charbonnier(x0, x1, t, mask, eps_charb: float = 1e-3, eps_area: float = 1e-8):
    """(charb(xt - x1) * m).sum() / (m.sum() * C + eps_area) (conditional_flow_matching_ROI_loss.py:80-92); no gradient:
    the term depends on the data only."""
    sums = K.roi_charbonnier(x0.contiguous(), x1.contiguous(), t.float().contiguous(), mask.float().contiguous(), eps_charb)
    return sums[0] / (sums[1] * x0.shape[1] + eps_area)


# --------------------------------------------------------------------------------------------- opaque <-> real dtype glue
class _ActToBF16(torch.autograd.Function):
    """Opaque forward-format activations -> a real torch.bfloat16 tensor (for library ops such as SDPA)."""

    @staticmethod
    def forward(ctx, x):
        return x if K.ACT == K.FMT_BF16 else x.view(torch.float16).to(torch.bfloat16)

    @staticmethod
    def backward(ctx, g):
        return g  # gradients are bf16 on both sides


class _BF16ToAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x if K.ACT == K.FMT_BF16 else x.to(torch.float16).view(T16)

    @staticmethod
    def backward(ctx, g):
        return g


def act_to_bf16(x):
    return _ActToBF16.apply(x)


def bf16_to_act(x):
    return _BF16ToAct.apply(x)


# ============================================================================================= multitask model (config M)
class _BatchNormRelu(torch.autograd.Function):
    """nn.BatchNorm2d (+ ReLU) on a 16-bit NHWC tensor: train mode = batch statistics (running stats updated in
    place), eval mode = running statistics.  The streaming passes are the GroupNorm kernels with act = ReLU, G = C."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training: bool, momentum: float, eps: float, relu: bool,
                src_stats=None, twin_box=None, sync_group=None):
        B, H, W, C = x.shape
        act = K.ACT_RELU if relu else K.ACT_NONE
        ctx.sync_group, ctx.count = None, B * H * W
        if training:
            stats = src_stats  # per-sub-tile partials from the producing conv's epilogue, same [B, n, C, 2] layout
            if stats is None:
                stats = K.gn_partial_buffer(B, H * W, C, x.device)
                K.gn_stats(x, stats, 0)
            if sync_group is not None:
                # torch.nn.SyncBatchNorm: statistics over the batches of ALL ranks.  One collective of 2C floats (the
                # per-channel sums); running statistics come out identical on every rank.  Every rank holds the same number
                # of samples (DistributedSampler pads to that), so the global count needs no second collective / host sync.
                import torch.distributed as dist
                sums = K.bn_fold(stats)
                dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=sync_group)
                ctx.sync_group, ctx.count = sync_group, B * H * W * dist.get_world_size(sync_group)
                coef, mr = K.bn_coef_sums(sums, ctx.count, B, gamma.detach(), beta.detach(), eps, momentum, running_mean,
                                          running_var)
            else:
                coef, mr = K.bn_coef(stats, gamma.detach(), beta.detach(), H * W, eps, momentum, running_mean, running_var)
        else:
            A = gamma.detach() * torch.rsqrt(running_var + eps)
            coef = torch.stack([A, beta.detach() - running_mean * A], dim=1).unsqueeze(0).expand(B, C, 2).contiguous()
            mr = None
        y = torch.empty_like(x)
        y2 = torch.empty_like(x) if (twin_box is not None and K.ACT != K.GRAD) else None
        K.gn_apply(x, coef, y, 0, act, y2=y2)  # y2: the same values as bf16, the consuming conv's wgrad operand
        if y2 is not None:
            twin_box.append(y2)
        ctx.act, ctx.training = act, training
        ctx.save_for_backward(x, coef, mr, gamma)
        return y

    @staticmethod
    def backward(ctx, g):
        if not ctx.training:
            raise NotImplementedError("backward through eval-mode BatchNorm is not on the reference's path")
        x, coef, mr, gamma = ctx.saved_tensors
        g = g.contiguous()
        B, H, W, C = x.shape
        red = K.gn_partial_buffer(B, H * W, C, x.device)
        K.gn_bwd_reduce(x, g, coef, mr, red, 0, ctx.act)
        if ctx.sync_group is not None:
            # (sum dz, sum dz*xhat) of all ranks enter dx; the parameter gradients stay local (DDP averages them), as in
            # torch's SyncBatchNorm backward
            import torch.distributed as dist
            local = K.bn_fold(red)  # [1, C, 2]
            total = local.clone()
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=ctx.sync_group)
            pqr = K.bn_bwd_coef_sums(total, ctx.count, B, mr, gamma.detach())
            dgamma, dbeta = local[0, :, 1].contiguous(), local[0, :, 0].contiguous()
        else:
            dgamma = torch.zeros(C, dtype=torch.float32, device=x.device)
            dbeta = torch.zeros(C, dtype=torch.float32, device=x.device)
            pqr = K.bn_bwd_coef(red, mr, gamma.detach(), H * W, dgamma, dbeta)
        dx = torch.empty_like(x)
        K.gn_bwd_apply(x, g, coef, pqr, 0, None, dx, ctx.act)
        return dx, dgamma, dbeta, None, None, None, None, None, None, None, None, None


def _sync_group_of(bn):
    """The process group to synchronise batch statistics over, or None.  A module is synchronised exactly when the reference's
    would be: after `torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)` (what Lightning's `sync_batchnorm: True` --
    configs/trainer/ddp.yaml:9 -- runs on the LightningModule) the BatchNorm2d children ARE SyncBatchNorm instances; like
    torch's, they only synchronise in training mode, with an initialised process group of more than one rank."""
    if not isinstance(bn, torch.nn.SyncBatchNorm) or not bn.training:
        return None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None
    group = bn.process_group if bn.process_group is not None else dist.group.WORLD
    return group if dist.get_world_size(group) > 1 else None


def batch_norm_relu(x, bn: "torch.nn.BatchNorm2d", relu: bool = True):
    if bn.training and bn.track_running_stats:
        bn.num_batches_tracked.add_(1)
    momentum = 0.1 if bn.momentum is None else bn.momentum
    box = [] if (bn.training and torch.is_grad_enabled()) else None
    y = _BatchNormRelu.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bool(bn.training),
                             float(momentum), float(bn.eps), relu, stats_of(x), box, _sync_group_of(bn))
    if box:
        y._s2s_g16 = box[0]
    return y


class _MaxPool2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return K.maxpool2x(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return K.maxpool2x_bwd(x, g.contiguous())


def maxpool2x(x):
    return _MaxPool2x.apply(x)


class _Bilinear2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return K.bilinear2x(x)

    @staticmethod
    def backward(ctx, g):
        return K.bilinear2x_bwd(g.contiguous())


def bilinear2x(x):
    return _Bilinear2x.apply(x)


class _ChannelBiasAdd(torch.autograd.Function):
    """x[b, :, :, c] + t[b, c] (the time conditioning of the flow decoder's bottleneck, task_decoders.py:119-125)."""

    @staticmethod
    def forward(ctx, x, t):
        B, H, W, C = x.shape
        coef = torch.stack([torch.ones_like(t), t], dim=2).float().contiguous()
        y = torch.empty_like(x)
        K.gn_apply(x, coef, y, 0, K.ACT_NONE)
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        B, H, W, C = g.shape
        stats = K.gn_partial_buffer(B, H * W, C, g.device)
        K.gn_stats(g, stats, 0, x_fmt=K.GRAD)
        return K.convert16(g, K.GRAD, K.GRAD), stats[..., 0].sum(dim=1)


def channel_bias_add(x, t):
    return _ChannelBiasAdd.apply(x, t)


class _Head1x1(torch.autograd.Function):
    """1x1 conv to <= 16 channels with fp32 NCHW output (the decoders' `outc`, task_decoders.py:100,170)."""

    @staticmethod
    def forward(ctx, a, w, b):
        cout, cin = w.shape[0], w.shape[1]

        def make(dst):
            wp = dst if dst is not None else torch.zeros((16, (cin + 63) // 64 * 64), dtype=T16, device=w.device)
            return wp, [(0, 0, 0, cin, False, K.ACT)]
        wp = PACK_CACHE.get(("head1x1", id(w)), [w], make)
        B, H, W, _ = a.shape
        out = K.conv_fwd([(a, 1, 1)], wp, cout, H, W, bias=b.detach(), out_f32=True)
        ctx.save_for_backward(a, w)
        return out

    @staticmethod
    def backward(ctx, dv):
        a, w = ctx.saved_tensors
        cout, cin = w.shape[0], w.shape[1]
        dv = dv.contiguous().float()
        dcol = K.nchw_to_nhwc16_pad(dv, 64, K.GRAD)  # [B,H,W,64], channels >= cout are zero
        wd = torch.zeros((cin, 64), dtype=torch.bfloat16, device=w.device)
        wd[:, :cout] = w.detach().reshape(cout, cin).t().to(torch.bfloat16)
        B, H, W, _ = a.shape
        da = K.conv_fwd([(dcol, 1, 1)], wd, cin, H, W, a_fmt=K.GRAD, w_fmt=K.GRAD, out_fmt=K.GRAD)
        dw = torch.zeros((1, cin, 64), dtype=torch.float32, device=w.device)
        K.conv_wgrad(K.convert16(a, K.ACT, K.GRAD), dcol, 1, 1, dw)
        d_w = dw[0, :, :cout].t().reshape(w.shape).contiguous()
        return da, d_w, dv.sum(dim=(0, 2, 3))


def head_conv1x1(a, w, b):
    return _Head1x1.apply(a, w, b)


class _SegLoss(torch.autograd.Function):
    """dice_weight * MulticlassDice + (1 - dice_weight) * CrossEntropy, one pass for the sums, one for the gradient."""

    @staticmethod
    def forward(ctx, logits, target, num_classes: int, ignore_index: int, dice_weight: float, smooth: float):
        logits = logits.contiguous()
        target = target.contiguous()
        C = logits.shape[1]
        assert C == num_classes
        sums = K.seg_loss_sums(logits, target, ignore_index)
        inter, psum, tsum = sums[:C], sums[C:2 * C], sums[2 * C:3 * C]
        dice = 1.0 - ((2.0 * inter + smooth) / (psum + tsum + smooth)).mean()
        ce = sums[3 * C] / sums[3 * C + 1]
        seg = dice_weight * dice + (1.0 - dice_weight) * ce
        ctx.cfg = (ignore_index, dice_weight, smooth)
        ctx.save_for_backward(logits, target, sums)
        dice, ce = dice.float(), ce.float()
        ctx.mark_non_differentiable(dice, ce)
        return seg.float(), dice, ce

    @staticmethod
    def backward(ctx, g_seg, _gd, _gc):
        logits, target, sums = ctx.saved_tensors
        ignore_index, dice_weight, smooth = ctx.cfg
        d = K.seg_loss_bwd(logits, target, ignore_index, sums, smooth, dice_weight, 1.0 - dice_weight,
                           g_seg.float().contiguous())
        return d, None, None, None, None, None


def seg_loss(logits, target, num_classes: int, ignore_index: int = -100, dice_weight: float = 0.5, smooth: float = 1.0):
    """-> (seg_total, dice, ce); target: int64 [B,H,W]."""
    return _SegLoss.apply(logits, target, num_classes, ignore_index, dice_weight, smooth)


# ============================================================================================= attention core (rows a14, f3)
class _Attention(torch.autograd.Function):
    """softmax(q k^T / sqrt(ch)) v per (sample, head) on the qkv conv's NHWC output (csrc/attention.cuh)."""

    @staticmethod
    def forward(ctx, qkv, heads: int, new_order: bool, grad_mode: bool):
        train = grad_mode and ctx.needs_input_grad[0]  # (inside forward() grad mode is always off: the caller passes it)
        out, lse = K.attn_fwd(qkv, heads, new_order, want_lse=train)
        ctx.cfg = (heads, new_order)
        if train:
            ctx.save_for_backward(qkv, out, lse)
        return out

    @staticmethod
    def backward(ctx, d_out):
        qkv, out, lse = ctx.saved_tensors
        heads, new_order = ctx.cfg
        return K.attn_bwd(qkv, out, d_out.contiguous(), lse, heads, new_order), None, None, None


def attention(qkv, heads: int, new_order: bool = False):
    return _Attention.apply(qkv, int(heads), bool(new_order), torch.is_grad_enabled())


# ============================================================================================= embedding path (row a9)
class _EmbedFilms(torch.autograd.Function):
    """timestep embedding -> time_embed MLP (+ label embedding) -> SiLU -> ALL FiLM projections of the UNet's ResBlocks.

    torchcfm evaluates `emb_layers = SiLU -> Linear(512, 2*Cout)` inside each of the 22 ResBlocks: 24 small
    GEMMs forward, 66 backward.  The SiLU'd embedding is the same for every block, so all projections run here as ONE
    multi-GEMM launch (fp32 FMA: the reference's arithmetic), and the backward of all of them -- dX partials, dW, db -- is
    one more.  Outputs: one contiguous fp32 [B, 2*Cout_i] tensor per block (`film` of the fused ResBlock node)."""

    @staticmethod
    def forward(ctx, t, model_channels: int, label_vec, w0, b0, w1, b1, *film_params):
        B = t.shape[0]
        dev = t.device
        f32 = dict(dtype=torch.float32, device=dev)
        temb = K.timestep_embedding(t.detach().float().contiguous(), model_channels)
        ted = w0.shape[0]
        h0, a0 = torch.empty((B, ted), **f32), torch.empty((B, ted), **f32)
        K.linear_multi([K.linear_fwd_job(temb, w0.detach(), b0.detach(), h0, act_out=a0)])
        e, ea = torch.empty((B, w1.shape[0]), **f32), torch.empty((B, w1.shape[0]), **f32)
        lab = None if label_vec is None else label_vec.detach().float().contiguous()
        K.linear_multi([K.linear_fwd_job(a0, w1.detach(), b1.detach(), e, act_out=ea, add=lab)])
        films, jobs = [], []
        for i in range(0, len(film_params), 2):
            w, b = film_params[i], film_params[i + 1]
            f = torch.empty((B, w.shape[0]), **f32)
            films.append(f)
            jobs.append(K.linear_fwd_job(ea, w.detach(), b.detach(), f))
        K.linear_multi(jobs)
        ctx.has_label = label_vec is not None
        ctx.save_for_backward(temb, h0, a0, e, ea, w0, w1, *film_params[0::2])
        return tuple(films)

    @staticmethod
    def backward(ctx, *dfilms):
        temb, h0, a0, e, ea, w0, w1, *fws = ctx.saved_tensors
        B, ted = ea.shape
        dev = ea.device
        f32 = dict(dtype=torch.float32, device=dev)
        ones = torch.ones(B, **f32)
        present = [i for i, d in enumerate(dfilms) if d is not None]
        parts = torch.empty((max(1, len(present)), B, ted), **f32)
        grads_f = [None] * (2 * len(fws))
        jobs = []
        for slot, i in enumerate(present):
            d = dfilms[i].float().contiguous()
            w = fws[i]
            dw, db = torch.empty_like(w, dtype=torch.float32), torch.empty(w.shape[0], **f32)
            grads_f[2 * i], grads_f[2 * i + 1] = dw, db
            jobs += [K.linear_dx_job(d, w.detach(), parts[slot]), K.linear_dw_job(d, ea, dw), K.linear_db_job(d, ones, db)]
        K.linear_multi(jobs)
        d_e = torch.empty((B, ted), **f32)
        if present:
            K.sum_parts_silu_bwd(parts, len(present), e, d_e)
        else:
            d_e.zero_()
        d_a0 = torch.empty((B, w1.shape[1]), **f32)
        d_w1, d_b1 = torch.empty_like(w1, dtype=torch.float32), torch.empty(w1.shape[0], **f32)
        K.linear_multi([K.linear_dx_job(d_e, w1.detach(), d_a0), K.linear_dw_job(d_e, a0, d_w1), K.linear_db_job(d_e, ones, d_b1)])
        d_h0 = torch.empty_like(d_a0)
        K.sum_parts_silu_bwd(d_a0, 1, h0, d_h0)
        d_w0, d_b0 = torch.empty_like(w0, dtype=torch.float32), torch.empty(w0.shape[0], **f32)
        K.linear_multi([K.linear_dw_job(d_h0, temb, d_w0), K.linear_db_job(d_h0, ones, d_b0)])
        return (None, None, d_e if ctx.has_label else None, d_w0, d_b0, d_w1, d_b1, *grads_f)


def embed_films(t, model_channels: int, label_vec, time_embed_params, film_params):
    """-> list of FiLM tensors, one per (weight, bias) pair of `film_params` (ResBlock execution order)."""
    flat = [p for wb in film_params for p in wb]
    return list(_EmbedFilms.apply(t, int(model_channels), label_vec, *time_embed_params, *flat))


class _FilmOne(torch.autograd.Function):
    """FiLM projection of ONE block from an already SiLU'd embedding (blocks used outside a UNet: tests, custom nets)."""

    @staticmethod
    def forward(ctx, emb_act, w, b):
        ea = emb_act.detach().float().contiguous()
        f = torch.empty((ea.shape[0], w.shape[0]), dtype=torch.float32, device=ea.device)
        K.linear_multi([K.linear_fwd_job(ea, w.detach(), b.detach(), f)])
        ctx.save_for_backward(ea, w)
        ctx.in_dtype = emb_act.dtype
        return f

    @staticmethod
    def backward(ctx, d):
        ea, w = ctx.saved_tensors
        d = d.float().contiguous()
        f32 = dict(dtype=torch.float32, device=d.device)
        d_ea, d_w, d_b = torch.empty_like(ea), torch.empty_like(w, dtype=torch.float32), torch.empty(w.shape[0], **f32)
        K.linear_multi([K.linear_dx_job(d, w.detach(), d_ea), K.linear_dw_job(d, ea, d_w),
                        K.linear_db_job(d, torch.ones(d.shape[0], **f32), d_b)])
        return d_ea.to(ctx.in_dtype), d_w, d_b


def film_one(emb_act, w, b):
    return _FilmOne.apply(emb_act, w, b)


# ============================================================================================= fused ResBlock
@dataclass
class ResBlockCfg:
    plan1: ConvPlan                 # 3x3 conv over the normalised (concatenated) input
    plan2: ConvPlan                 # 3x3 conv over the second normalised tensor (+ 1x1 skip segments over the raw inputs)
    has_skip_conv: bool
    groups: int = 32
    eps: float = 1e-5
    plan1_multi: Optional[ConvPlan] = None  # inference: conv 1 over the RAW sources, one 3x3 segment per source
    plan2_id: Optional[ConvPlan] = None     # identity skip as a 1x1 GEMM segment with the identity matrix (see _identity_weight)


# test hook: when a list, every training-mode ResBlock appends its dropout keep bits (uint8 [B,H,W,C/8], bit j of byte
# c8 = channel 8*c8 + j) in execution order, so a parity test can replay the SAME masks into the fp32 oracle
DROPOUT_TAP: Optional[list] = None


def _wgrad_to(d_out, x_g, taps: int, stride: int, d_w_view, n_begin: int):
    """d_w_view[m][n_begin + n][tap] = sum_pixels d_out[., m] * x_g[. + tap, n]   (x_g: bf16 NHWC, all its channels)."""
    cout, cq = d_out.shape[3], x_g.shape[3]
    dw = torch.zeros((taps, cout, cq), dtype=torch.float32, device=d_out.device)
    K.conv_wgrad(d_out, x_g, taps, stride, dw)
    K.unpack_wgrad(dw, d_w_view, 0, cq, n_begin, 0.0)


_EYES: Dict[tuple, torch.Tensor] = {}


def _identity_weight(c: int, device) -> torch.Tensor:
    """fp32 [C, C, 1, 1] identity "conv weight".  An identity skip `x + h` is run as one more 1x1 GEMM segment of conv 2 over the
    raw block input with this matrix: 1.0 and 0.0 are exact in the 16-bit operand format and the products accumulate in fp32,
    so the result equals adding x in the epilogue up to fp32 summation order -- but the epilogue's pixel-strided 16-byte residual loads made
    conv 2 of the 128-channel 256^2 blocks take 2.0-2.3 ms instead of 0.85 ms, while the extra segment costs +11 % MMAs."""
    key = (c, str(device))
    eye = _EYES.get(key)
    if eye is None:
        eye = _EYES[key] = torch.eye(c, dtype=torch.float32, device=device).reshape(c, c, 1, 1).contiguous()
    return eye


class _ResBlockFn(torch.autograd.Function):
    """One torchcfm ResBlock (GN-SiLU-conv3x3, FiLM'd GN-SiLU-dropout-conv3x3, identity / 1x1 skip) as ONE autograd node.

    Compared with chaining the per-op nodes this (a) lets the normalisation kernels emit the bf16 copy the weight-
    gradient GEMMs need next to the fp16 forward operand (no conversion pass in backward), and (b) folds the gradient
    of the residual / skip path into the first norm's backward kernel (`add`) instead of a separate accumulation pass
    over the block input."""

    @staticmethod
    def forward(ctx, cfg: ResBlockCfg, n_src: int, drop_p: float, seed: int, src_stats, stats_box, grad_mode: bool,
                *tensors):
        srcs = tensors[:n_src]
        film = tensors[n_src].detach()  # fp32 [B, 2*Cout] = Linear(SiLU(emb)) of this block (ops.embed_films)
        gn1w, gn1b, c1w, c1b, gn2w, gn2b, c2w, c2b = tensors[n_src + 1:n_src + 9]
        skip = tensors[n_src + 9:]
        B, H, W, _ = srcs[0].shape
        ctot = sum(s.shape[3] for s in srcs)
        cout = c1w.shape[0]
        dev = srcs[0].device
        # `needs_input_grad` is True for parameters even under torch.no_grad(); the caller passes the grad mode
        train = grad_mode and any(ctx.needs_input_grad)
        dual = train and K.ACT != K.GRAD
        mask = None
        # ---- norm 1 (+SiLU) over the concatenated input
        if usable_stats(src_stats):  # statistics written by the producing convs' epilogues: no pass over the inputs
            coef1, mr1 = K.gn_coef_parts(list(src_stats), gn1w.detach(), gn1b.detach(), None, H * W, cfg.groups, cfg.eps)
        else:
            stats = K.gn_partial_buffer(B, H * W, ctot, dev)
            off = 0
            for s in srcs:
                K.gn_stats(s, stats, off)
                off += s.shape[3]
            coef1, mr1 = K.gn_coef(stats, gn1w.detach(), gn1b.detach(), None, H * W, cfg.groups, cfg.eps)
        # inference: the norm-apply passes disappear -- the convs read the RAW tensors and apply silu(x * A + Bc) to
        # their shared-memory tiles (s2s_conv_fwd_norm); the concat is then one GEMM segment per source
        fuse = (not train) and drop_p == 0 and cfg.plan1_multi is not None and \
            K.conv_norm_fusable([(s, 9, 1) for s in srcs], cout)  # (the fused prologue has no dropout)
        if fuse:
            offs, off = [], 0
            for s in srcs:
                offs.append(off)
                off += s.shape[3]
            h, h_stats = K.conv_fwd([(s, 9, 1) for s in srcs], cfg.plan1_multi.packed_fwd([c1w]), cout, H, W,
                                    bias=c1b.detach(), want_stats=True, norms=[(coef1, o) for o in offs], norm_act=1)
        else:
            a1 = torch.empty((B, H, W, ctot), dtype=T16, device=dev)
            a1g = torch.empty_like(a1) if dual else None
            off = 0
            for s in srcs:
                K.gn_apply(s, coef1, a1, off, True, y2=a1g)
                off += s.shape[3]
            h, h_stats = K.conv_fwd([(a1, 9, 1)], cfg.plan1.packed_fwd([c1w]), cout, H, W, bias=c1b.detach(),
                                    want_stats=True)
        # ---- norm 2 with FiLM (+SiLU, dropout)
        assert film.dtype == torch.float32 and film.is_contiguous() and tuple(film.shape) == (B, 2 * cout)
        if h_stats is not None:
            coef2, mr2 = K.gn_coef_parts([h_stats], gn2w.detach(), gn2b.detach(), film, H * W, cfg.groups, cfg.eps)
            part2 = h_stats
        else:
            stats2 = K.gn_partial_buffer(B, H * W, cout, dev)
            K.gn_stats(h, stats2, 0)
            coef2, mr2 = K.gn_coef(stats2, gn2w.detach(), gn2b.detach(), film, H * W, cfg.groups, cfg.eps)
            part2 = stats2
        # sum over pixels of h per (sample, channel): lets backward write conv 1's bias gradient in closed form
        sum_h = part2[..., 0].sum(dim=1) if train else None
        if fuse:
            a2, norms2 = h, [(coef2, 0)]
        else:
            a2 = torch.empty((B, H, W, cout), dtype=T16, device=dev)
            a2g = torch.empty_like(a2) if dual else None
            mask = torch.empty((B, H, W, cout // 8), dtype=torch.uint8, device=dev) if (train and drop_p > 0) else None
            K.gn_apply(h, coef2, a2, 0, True, drop_p, seed, y2=a2g, mask=mask)
            if DROPOUT_TAP is not None and mask is not None:
                DROPOUT_TAP.append(mask)
            norms2 = None
        # ---- conv 2 with the skip path in the same accumulator
        if cfg.has_skip_conv:
            sw, sb = skip
            wp = cfg.plan2.packed_fwd([c2w, sw])
            bias = (c2b.detach() + sb.detach()).contiguous()
            out, out_stats = K.conv_fwd([(a2, 9, 1)] + [(s, 1, 1) for s in srcs], wp, cout, H, W, bias=bias,
                                        want_stats=True,
                                        norms=None if norms2 is None else norms2 + [None] * len(srcs))
        elif cfg.plan2_id is not None:
            assert n_src == 1 and ctot == cout
            wp = cfg.plan2_id.packed_fwd([c2w, _identity_weight(cout, dev)])
            out, out_stats = K.conv_fwd([(a2, 9, 1), (srcs[0], 1, 1)], wp, cout, H, W, bias=c2b.detach(), want_stats=True,
                                        norms=None if norms2 is None else norms2 + [None],
                                        alg_macs=float(B) * H * W * cout * cout * 9)
        else:
            assert n_src == 1 and ctot == cout
            out, out_stats = K.conv_fwd([(a2, 9, 1)], cfg.plan2.packed_fwd([c2w]), cout, H, W, bias=c2b.detach(),
                                        residual=srcs[0], want_stats=True, norms=norms2)
        stats_box.append(out_stats)
        if train:
            ctx.cfg, ctx.n_src, ctx.drop = cfg, n_src, (drop_p, seed)
            ctx.dual = dual
            ctx.mask = mask  # uint8 keep bits of the dropout (1 bit / element), read by the two norm-backward passes
            ctx.save_for_backward(*srcs, h, a1g if dual else a1, a2g if dual else a2, coef1, mr1, coef2, mr2,
                                  film, gn1w, gn1b, c1w, gn2w, gn2b, c2w, sum_h, *skip[:1])
        return out

    @staticmethod
    def backward(ctx, d_out):
        cfg, n_src = ctx.cfg, ctx.n_src
        drop_p, seed = ctx.drop
        sv = ctx.saved_tensors
        srcs = sv[:n_src]
        h, a1s, a2s, coef1, mr1, coef2, mr2, film, gn1w, gn1b, c1w, gn2w, gn2b, c2w, sum_h = sv[n_src:n_src + 15]
        sw = sv[n_src + 15] if cfg.has_skip_conv else None
        a1g = a1s if ctx.dual else K.convert16(a1s, K.ACT, K.GRAD)
        a2g = a2s if ctx.dual else K.convert16(a2s, K.ACT, K.GRAD)
        d_out = d_out.contiguous()
        B, H, W, cout = d_out.shape
        dev = d_out.device
        ctot = a1g.shape[3]
        f32 = dict(dtype=torch.float32, device=dev)
        # ---- conv 2 (+ skip): weight / bias gradients, data gradients
        d_c2w = torch.empty_like(c2w, dtype=torch.float32)
        _wgrad_to(d_out, a2g, 9, 1, d_c2w.view(cout, cout, -1), 0)
        d_b2 = torch.zeros(cout, **f32)
        K.channel_sum(d_out, d_b2)
        d_a2 = K.conv_fwd([(d_out, 9, 1)], cfg.plan2.packed_dgrad(0, [c2w] + ([sw] if sw is not None else [])), cout, H, W,
                          a_fmt=K.GRAD, w_fmt=K.GRAD, out_fmt=K.GRAD)
        d_skip: List[torch.Tensor] = []
        d_sw = None
        if cfg.has_skip_conv:
            # data gradient of the 1x1 skip conv now; its weight gradient needs the raw inputs in bf16, which the first
            # norm's backward reduce pass (it reads them anyway) writes as a side product further down
            for i, s in enumerate(srcs):
                d_skip.append(K.conv_fwd([(d_out, 1, 1)], cfg.plan2.packed_dgrad(1 + i, [c2w, sw]), s.shape[3], H, W,
                                         a_fmt=K.GRAD, w_fmt=K.GRAD, out_fmt=K.GRAD))
        else:
            d_skip.append(d_out)
        # ---- norm 2 backward
        red2 = K.gn_partial_buffer(B, H * W, cout, dev)
        K.gn_bwd_reduce(h, d_a2, coef2, mr2, red2, 0, True, drop_p, seed, mask=ctx.mask)
        d_gn2w, d_gn2b = torch.zeros(cout, **f32), torch.zeros(cout, **f32)
        pqr2, dfilm, red2f = K.gn_bwd_coef(red2, mr2, gn2w, gn2b, film, H * W, d_gn2w, d_gn2b, True, want_red=True)
        d_h = torch.empty_like(h)
        K.gn_bwd_apply(h, d_a2, coef2, pqr2, 0, None, d_h, True, drop_p, seed, mask=ctx.mask)
        # ---- conv 1  (dfilm goes back to ops.embed_films, which runs every block's FiLM backward as one multi-GEMM)
        d_c1w = torch.empty_like(c1w, dtype=torch.float32)
        _wgrad_to(d_h, a1g, 9, 1, d_c1w.view(cout, ctot, -1), 0)
        # bias gradient of conv 1 = sum over pixels of d_h = sum_b (P * sum dz + Q * sum h + HW * R): closed form from
        # per-(sample, channel) quantities that exist already -- no pass over d_h
        d_b1 = (pqr2[..., 0] * red2f[..., 0] + pqr2[..., 1] * sum_h + float(H * W) * pqr2[..., 2]).sum(dim=0)
        d_a1 = K.conv_fwd([(d_h, 9, 1)], cfg.plan1.packed_dgrad(0, [c1w]), ctot, H, W, a_fmt=K.GRAD, w_fmt=K.GRAD,
                          out_fmt=K.GRAD)
        # ---- norm 1 backward; the skip-path gradient is added in the same pass
        red1 = K.gn_partial_buffer(B, H * W, ctot, dev)
        off = 0
        need_raw = cfg.has_skip_conv and ctx.dual
        if cfg.has_skip_conv:
            d_sw = torch.empty_like(sw, dtype=torch.float32)
        for s in srcs:
            s_g = torch.empty_like(s) if need_raw else None
            K.gn_bwd_reduce(s, d_a1, coef1, mr1, red1, off, True, x_bf16_out=s_g)
            if cfg.has_skip_conv:
                _wgrad_to(d_out, s_g if need_raw else K.convert16(s, K.ACT, K.GRAD), 1, 1, d_sw.view(cout, ctot, -1), off)
            off += s.shape[3]
        d_gn1w, d_gn1b = torch.zeros(ctot, **f32), torch.zeros(ctot, **f32)
        pqr1, _ = K.gn_bwd_coef(red1, mr1, gn1w, gn1b, None, H * W, d_gn1w, d_gn1b, False)
        d_srcs = []
        off = 0
        for i, s in enumerate(srcs):
            dx = torch.empty_like(s)
            K.gn_bwd_apply(s, d_a1, coef1, pqr1, off, d_skip[i], dx, True)
            d_srcs.append(dx)
            off += s.shape[3]
        grads = [None, None, None, None, None, None, None, *d_srcs, dfilm, d_gn1w, d_gn1b, d_c1w, d_b1,
                 d_gn2w, d_gn2b, d_c2w, d_b2]
        if cfg.has_skip_conv:
            grads += [d_sw, d_b2]
        return tuple(grads)


def res_block(cfg: ResBlockCfg, srcs, film, params, drop_p: float, seed: int):
    """film: fp32 [B, 2*Cout] FiLM tensor of this block; params: gn1 w/b, conv1 w/b, gn2 w/b, conv2 w/b (+ skip w/b)."""
    box = []
    src_stats = tuple(stats_of(t) for t in srcs)
    return _tag(_ResBlockFn.apply(cfg, len(srcs), float(drop_p), int(seed), src_stats, box, torch.is_grad_enabled(),
                                  *srcs, film, *params), box)
