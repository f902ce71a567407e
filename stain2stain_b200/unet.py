"""B200-native velocity-field UNet: drop-in for `torchcfm.models.unet.UNetModel` (wrapper spelling) and
`torchcfm.models.unet.unet.UNetModel` (raw spelling) as the reference instantiates them:

  * configs/model/conditional_flow_matching.yaml:16-26, configs/model/class_conditional_flow_matching.yaml:16-28
  * src/models/components/unet_4to3.py:51-67 (raw ctor)
  * called as net(t, x) / net(t, x, y=y): src/models/conditional_flow_matching.py:51,
    src/models/class_conditional_flow_matching.py:47; handed to NeuralODE: conditional_flow_matching.py:157-167.

The module tree, parameter names, shapes, dtypes (fp32) and initialisation are those of the reference model, so
`state_dict()` / `load_state_dict(strict=True)` interoperate with the reference's checkpoints.  The
nn.Conv2d / nn.GroupNorm / nn.Linear children are parameter containers only: `forward` never calls them.  All heavy
math runs in the sm_100a kernels (bf16 NHWC activations, tcgen05 implicit-GEMM convs, fused GroupNorm/FiLM/SiLU).
There is no CPU path: calling the model on host tensors raises.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .ops import ConvPlan, Seg

NUM_CLASSES = 1000


def zero_module(m: nn.Module) -> nn.Module:
    for p in m.parameters():
        p.detach().zero_()
    return m


class GroupNorm32(nn.GroupNorm):
    """Parameter container (the reference normalises in fp32; so do the kernels)."""


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


_DROPOUT_CALLS = [0]


def _next_dropout_seed() -> int:
    _DROPOUT_CALLS[0] += 1
    base = torch.initial_seed() & 0xFFFFFFFF
    return ((base * 0x9E3779B1) ^ (_DROPOUT_CALLS[0] * 0x85EBCA77)) & 0xFFFFFFFFFFFFFFFF


class TimestepBlock(nn.Module):
    pass


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """Children take `(srcs, emb_act)`; `srcs` is the list of NHWC tensors whose channel concat is the block input."""

    def forward(self, srcs, emb_act):
        for layer in self:
            if isinstance(layer, TimestepBlock):
                h = layer(srcs, emb_act)
            elif isinstance(layer, nn.Conv2d):  # the stem is handled by UNetModel.forward
                raise RuntimeError("stem conv is dispatched by UNetModel.forward")
            else:
                h = layer(srcs)
            srcs = [h]
        return srcs[0]


def _single(srcs: List[torch.Tensor]) -> torch.Tensor:
    return srcs[0] if len(srcs) == 1 else torch.cat(srcs, dim=3)


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, out_channels=None):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        if use_conv:
            self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=1)
            self._plan = ConvPlan((Seg(0, 0, 0, channels, 9, 1),), self.out_channels)
            self._up_plan = ops.UpConvPlan(channels, self.out_channels)

    def forward(self, srcs):
        x = _single(srcs)
        if self.use_conv and ops.K.upconv_supported(self.channels, self.out_channels):
            # phase-decomposed: four 2x2 convs over the low-resolution tensor, the 4x tensor is never written
            return ops.upsample_conv(self._up_plan, x, self.conv.weight, self.conv.bias)
        u = ops.upsample2x(x)
        if not self.use_conv:
            return u
        return ops.fused_conv(self._plan, [u], [self.conv.weight], [self.conv.bias])


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, out_channels=None):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        if not use_conv:
            raise NotImplementedError("conv_resample=False (AvgPool downsample) is not used by the reference configs")
        self.op = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=1)
        self._plan = ConvPlan((Seg(0, 0, 0, channels, 9, 2),), self.out_channels)

    def forward(self, srcs):
        return ops.fused_conv(self._plan, [_single(srcs)], [self.op.weight], [self.op.bias])


class ResBlock(TimestepBlock):
    """GN-SiLU-conv3x3, FiLM'd GN-SiLU-dropout-conv3x3, plus skip (identity or 1x1) -- two fused convs, two fused norms.

    The second conv accumulates the 1x1 skip connection in the same TMEM accumulator (extra K-blocks over the raw
    block input); an identity skip is added in the conv epilogue.  A two-tensor input (UNet skip concat) is never
    materialised: the norm reads both tensors, the 1x1 skip reads both as separate GEMM segments.
    """

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False,
                 up=False, down=False, split: Optional[List[int]] = None):
        super().__init__()
        if up or down:
            raise NotImplementedError("resblock_updown=True is not used by the reference configs")
        if not use_scale_shift_norm:
            raise NotImplementedError("use_scale_shift_norm=False is not used by the reference configs")
        if use_conv:
            raise NotImplementedError("3x3 skip connections (use_conv=True) are not used by the reference configs")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.dropout = dropout
        oc = self.out_channels
        self.in_layers = nn.Sequential(GroupNorm32(32, channels), nn.SiLU(), nn.Conv2d(channels, oc, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, 2 * oc))
        self.out_layers = nn.Sequential(GroupNorm32(32, oc), nn.SiLU(), nn.Dropout(p=dropout),
                                        zero_module(nn.Conv2d(oc, oc, 3, padding=1)))
        self.skip_connection = nn.Identity() if oc == channels else nn.Conv2d(channels, oc, 1)
        self._plan1 = ConvPlan((Seg(0, 0, 0, channels, 9, 1),), oc)
        self._plan2 = ConvPlan((Seg(0, 0, 0, oc, 9, 1),), oc)
        # identity skip as a GEMM segment: weight 1 = the constant identity matrix (ops._identity_weight)
        self._plan2_id = ConvPlan((Seg(0, 0, 0, oc, 9, 1), Seg(1, 1, 0, oc, 1, 1)), oc) if oc == channels else None
        self._skip_plans = {}
        self._cfgs = {}

    def _plan2_skip(self, widths):
        key = tuple(widths)
        if key not in self._skip_plans:
            segs, off = [Seg(0, 0, 0, self.out_channels, 9, 1)], 0
            for i, w in enumerate(widths):
                segs.append(Seg(1 + i, 1, off, w, 1, 1))
                off += w
            self._skip_plans[key] = ConvPlan(tuple(segs), self.out_channels)
        return self._skip_plans[key]

    def forward(self, srcs, emb):
        """emb: the UNet's FiLM table {ResBlock: fp32 [B, 2*Cout]} (all blocks' projections computed in one launch by
        ops.embed_films), or -- for a block used on its own -- the SiLU'd embedding tensor [B, emb_channels]."""
        gn1, conv1 = self.in_layers[0], self.in_layers[2]
        gn2, conv2 = self.out_layers[0], self.out_layers[3]
        lin = self.emb_layers[1]
        p = self.dropout if (self.training and self.dropout > 0) else 0.0
        has_skip = not isinstance(self.skip_connection, nn.Identity)
        widths = tuple(s.shape[3] for s in srcs)
        cfg = self._cfgs.get(widths)
        if cfg is None:
            segs, off = [], 0
            for i, w in enumerate(widths):
                segs.append(Seg(i, 0, off, w, 9, 1))
                off += w
            cfg = self._cfgs[widths] = ops.ResBlockCfg(self._plan1, self._plan2_skip(widths) if has_skip else self._plan2,
                                                       has_skip, plan1_multi=ConvPlan(tuple(segs), self.out_channels),
                                                       plan2_id=None if has_skip else self._plan2_id)
        params = [gn1.weight, gn1.bias, conv1.weight, conv1.bias, gn2.weight, gn2.bias, conv2.weight, conv2.bias]
        if has_skip:
            params += [self.skip_connection.weight, self.skip_connection.bias]
        film = emb[self] if isinstance(emb, dict) else ops.film_one(emb, lin.weight, lin.bias)
        return ops.res_block(cfg, list(srcs), film, params, p, _next_dropout_seed() if p > 0 else 0)


class AttentionBlock(nn.Module):
    """GN -> qkv 1x1 -> multi-head softmax attention over H*W tokens -> proj 1x1 (+x).  qkv/proj are tcgen05 GEMMs; the
    (T x T) softmax core is this package's flash-style kernel (csrc/attention.cuh) reading q/k/v in place from the qkv conv's
    output, for 32 or 64 channels per head (every reference config: 32).  Other head sizes -- only reachable with
    constructor arguments no reference config uses -- go through torch SDPA."""

    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_new_attention_order=False):
        super().__init__()
        self.channels = channels
        if num_head_channels == -1:
            self.num_heads = num_heads
        else:
            assert channels % num_head_channels == 0
            self.num_heads = channels // num_head_channels
        self.new_order = use_new_attention_order
        self.norm = GroupNorm32(32, channels)
        self.qkv = nn.Conv1d(channels, channels * 3, 1)
        self.proj_out = zero_module(nn.Conv1d(channels, channels, 1))
        self._plan_qkv = ConvPlan((Seg(0, 0, 0, channels, 1, 1),), 3 * channels)
        self._plan_proj = ConvPlan((Seg(0, 0, 0, channels, 1, 1),), channels)

    def forward(self, srcs):
        x = _single(srcs)
        B, H, W, Cc = x.shape
        T, nh = H * W, self.num_heads
        ch = Cc // nh
        a = ops.group_norm_act([x], self.norm.weight, self.norm.bias, None, silu=False)
        qkv = ops.fused_conv(self._plan_qkv, [a], [self.qkv.weight], [self.qkv.bias])
        if ops.K.attn_supported(ch):
            o = ops.attention(qkv, nh, self.new_order)
            return ops.fused_conv(self._plan_proj, [o], [self.proj_out.weight], [self.proj_out.bias], residual=x)
        qkv = ops.act_to_bf16(qkv)
        if self.new_order:
            q, k, v = qkv.view(B, T, 3, nh, ch).unbind(2)
        else:  # legacy: channel = head*3*ch + {q,k,v}*ch + c
            q, k, v = qkv.view(B, T, nh, 3, ch).unbind(3)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        o = ops.bf16_to_act(o.transpose(1, 2).reshape(B, H, W, Cc).contiguous())
        return ops.fused_conv(self._plan_proj, [o], [self.proj_out.weight], [self.proj_out.bias], residual=x)


class RawUNetModel(nn.Module):
    """Same constructor as `torchcfm.models.unet.unet.UNetModel`."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False):
        super().__init__()
        if dims != 2:
            raise NotImplementedError("only 2-D UNets are on the reference's path")
        if use_fp16:
            raise NotImplementedError("use_fp16 is never set by the reference; this engine computes in bf16/fp32")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = tuple(attention_resolutions)
        self.dropout = dropout
        self.channel_mult = tuple(channel_mult)
        self.num_classes = num_classes
        self.dtype = torch.float32
        ssn = use_scale_shift_norm

        ted = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))
        if num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, ted)

        ch = input_ch = int(channel_mult[0] * model_channels)
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, ch, 3, padding=1))])
        chans = [ch]
        ds = 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, ted, dropout, out_channels=int(mult * model_channels), use_scale_shift_norm=ssn)]
                ch = int(mult * model_channels)
                if ds in self.attention_resolutions:
                    layers.append(AttentionBlock(ch, num_heads, num_head_channels, use_new_attention_order))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                if resblock_updown:
                    raise NotImplementedError("resblock_updown=True is not used by the reference configs")
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, out_channels=ch)))
                chans.append(ch)
                ds *= 2

        self.middle_block = TimestepEmbedSequential(
            ResBlock(ch, ted, dropout, use_scale_shift_norm=ssn),
            AttentionBlock(ch, num_heads, num_head_channels, use_new_attention_order),
            ResBlock(ch, ted, dropout, use_scale_shift_norm=ssn),
        )

        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [ResBlock(ch + ich, ted, dropout, out_channels=int(model_channels * mult),
                                   use_scale_shift_norm=ssn)]
                ch = int(model_channels * mult)
                if ds in self.attention_resolutions:
                    layers.append(AttentionBlock(ch, num_heads_upsample, num_head_channels, use_new_attention_order))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))

        self.out = nn.Sequential(GroupNorm32(32, ch), nn.SiLU(),
                                 zero_module(nn.Conv2d(input_ch, out_channels, 3, padding=1)))

    # ------------------------------------------------------------------------------------------ forward
    def _embed(self, t, x, y):
        assert (y is not None) == (self.num_classes is not None), "must specify y if and only if the model is class-conditional"
        while t.dim() > 1:
            t = t[:, 0]
        if t.dim() == 0:
            t = t.repeat(x.shape[0])
        label_vec = None
        if self.num_classes is not None:
            assert y.shape == (x.shape[0],)
            label_vec = self.label_emb(y)
        # every ResBlock's emb_layers starts with SiLU: it is applied once, and all FiLM projections run in one launch
        blocks = self._res_blocks()
        te = self.time_embed
        films = ops.embed_films(t.to(x.device).float().contiguous(), self.model_channels, label_vec,
                                [te[0].weight, te[0].bias, te[2].weight, te[2].bias],
                                [(b.emb_layers[1].weight, b.emb_layers[1].bias) for b in blocks])
        return dict(zip(blocks, films))

    def _res_blocks(self):
        """The ResBlocks in execution order (input, middle, output blocks)."""
        if getattr(self, "_res_block_list", None) is None:
            object.__setattr__(self, "_res_block_list", [m for m in self.modules() if isinstance(m, ResBlock)])
        return self._res_block_list

    def _trunk(self, h, emb_act):
        hs = [h]
        for m in list(self.input_blocks)[1:]:
            h = m([h], emb_act)
            hs.append(h)
        h = self.middle_block([h], emb_act)
        for m in self.output_blocks:
            h = m([h, hs.pop()], emb_act)
        gn = self.out[0]
        return ops.group_norm_act([h], gn.weight, gn.bias, None, silu=True)

    def forward(self, t, x, y=None):
        """x: fp32 NCHW image batch (CUDA).  Returns the velocity, fp32 NCHW, like the reference."""
        return self._run(t, x, y)

    def _run(self, t, x, y=None, x1=None, axpy_a: Optional[float] = None, extra=None):
        """`extra`: optional fp32 [B,1,H,W] condition channel appended behind x's channels inside the stem operand
        packing (the mask of the mask-conditioned variants; `torch.cat([x, mask], 1)` is never materialised)."""
        if not x.is_cuda:
            raise RuntimeError("stain2stain_b200.UNetModel runs on CUDA (sm_100a) only; there is no CPU fallback")
        if self.in_channels not in (3, 4):
            raise NotImplementedError("the B200 stem kernel takes 3- or 4-channel tiles (reference configs 1-4 and the "
                                      "mask-conditioned variants)")
        if x.shape[1] + (0 if extra is None else 1) != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {x.shape[1]}"
                             + ("" if extra is None else " + 1 condition channel"))
        in_dtype = x.dtype
        x = x.float().contiguous()
        emb_act = self._embed(t, x, y)
        stem = self.input_blocks[0][0]
        tt = None
        if x1 is not None:  # fused flow-matching interpolation: stem input = (1-t) x + t x1
            tt = t
            while tt.dim() > 1:
                tt = tt[:, 0]
            tt = tt.to(x.device).float()
        h = ops.stem_conv(x, stem.weight, stem.bias, x1=None if x1 is None else x1.float(), t=tt, extra=extra)
        a = self._trunk(h, emb_act)
        head = self.out[2]
        if axpy_a is not None:  # inference-only fused Euler update, in place on x
            return ops.head_conv(a, head.weight, head.bias, axpy_x=x, axpy_a=float(axpy_a))
        return ops.head_conv(a, head.weight, head.bias).to(in_dtype)

    @torch.no_grad()
    def euler_step_(self, t, x, dt: float, y=None, extra=None):
        """x <- x + dt * v(t, x) with the update fused into the head conv's epilogue (x: fp32 NCHW contiguous)."""
        assert x.dtype == torch.float32 and x.is_contiguous()
        assert x.shape[1] == self.out_channels, "the fused Euler update needs state channels == out_channels"
        return self._run(t, x, y, axpy_a=dt, extra=extra)

    def velocity_of_interpolant(self, t, x0, x1, y=None, extra=None):
        """v(t, (1-t) x0 + t x1) with the interpolation fused into the stem operand packing (training path)."""
        return self._run(t, x0, y, x1=x1, extra=extra)


def default_channel_mult(image_size: int):
    table = {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4),
             64: (1, 2, 3, 4), 32: (1, 2, 2, 2), 28: (1, 2, 2)}
    if image_size not in table:
        raise ValueError(f"unsupported image size: {image_size}")
    return table[image_size]


class UNetModel(RawUNetModel):
    """Same constructor as `torchcfm.models.unet.UNetModel` (the wrapper)."""

    def __init__(self, dim, num_channels, num_res_blocks, channel_mult=None, learn_sigma=False, class_cond=False,
                 num_classes=NUM_CLASSES, use_checkpoint=False, attention_resolutions="16", num_heads=1,
                 num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False, dropout=0,
                 resblock_updown=False, use_fp16=False, use_new_attention_order=False):
        image_size = dim[-1]
        channel_mult = default_channel_mult(image_size) if channel_mult is None else tuple(channel_mult)
        attention_ds = [image_size // int(r) for r in str(attention_resolutions).split(",")]
        if learn_sigma:
            raise NotImplementedError("learn_sigma is a diffusion option; the flow-matching configs do not set it")
        super().__init__(image_size=image_size, in_channels=dim[0], model_channels=num_channels, out_channels=dim[0],
                         num_res_blocks=num_res_blocks, attention_resolutions=tuple(attention_ds), dropout=dropout,
                         channel_mult=channel_mult, num_classes=(num_classes if class_cond else None),
                         use_checkpoint=use_checkpoint, use_fp16=use_fp16, num_heads=num_heads,
                         num_head_channels=num_head_channels, num_heads_upsample=num_heads_upsample,
                         use_scale_shift_norm=use_scale_shift_norm, resblock_updown=resblock_updown,
                         use_new_attention_order=use_new_attention_order)

    def forward(self, t, x, y=None, *args, **kwargs):
        return super().forward(t, x, y=y)
