"""ORACLE (test infrastructure only) -- fp32 PyTorch restatement of the torchcfm 1.0.7 UNet.

PARITY UNPINNED for the UNet internals (pinned only through the reference LitModule vectors in tests/golden, see
oracle/make_golden.py): torchcfm==1.0.7 (uv.lock:4618-4619 of the reference) is not vendored under
/root/reference and is not installable here, and the reference's tests hold no golden vector for
this path.  This file restates the published guided-diffusion style UNet that
torchcfm ships (`torchcfm.models.unet.UNetModel` == `UNetModelWrapper`) from its specification
 and is anchored on the reference's own call sites:

  * wrapper ctor kwargs ....... configs/model/conditional_flow_matching.yaml:16-26
  * raw ctor kwargs ........... src/models/components/unet_4to3.py:51-67
  * call signature net(t,x,y) . src/models/conditional_flow_matching.py:51,
                                src/models/class_conditional_flow_matching.py:47
  * state_dict key names ...... src/infer_simple_flowmatching.py:51 (strict load of `net.*`)

Known-answer checks that pin the structure (tests/test_oracle.py): parameter count 70 954 883 for
config A, 70 956 419 for the class-conditional config, 35 746 307 for torchcfm's CIFAR-10 config,
output == 0 at initialisation.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product package (stain2stain_b200/) never does.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

NUM_CLASSES = 1000


def zero_module(m: nn.Module) -> nn.Module:
    for p in m.parameters():
        p.detach().zero_()
    return m


class GroupNorm32(nn.GroupNorm):
    # A.2: normalisation is evaluated in fp32 and cast back.
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


def normalization(ch: int) -> nn.Module:
    return GroupNorm32(32, ch)


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    # A.2: cos first, then sin; freqs = exp(-ln(max_period) * i / half)
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


class TimestepBlock(nn.Module):
    pass


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    def forward(self, x, emb):
        for layer in self:
            x = layer(x, emb) if isinstance(layer, TimestepBlock) else layer(x)
        return x


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, out_channels=None):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        if use_conv:
            self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=1)

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        return self.conv(x) if self.use_conv else x


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, out_channels=None):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        if use_conv:
            self.op = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=1)
        else:
            self.op = nn.AvgPool2d(2)

    def forward(self, x):
        return self.op(x)


class ResBlock(TimestepBlock):
    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False,
                 use_scale_shift_norm=False, up=False, down=False):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_scale_shift_norm = use_scale_shift_norm
        oc = self.out_channels
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(), nn.Conv2d(channels, oc, 3, padding=1))
        self.updown = up or down
        if up:
            self.h_upd, self.x_upd = Upsample(channels, False), Upsample(channels, False)
        elif down:
            self.h_upd, self.x_upd = Downsample(channels, False), Downsample(channels, False)
        else:
            self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, 2 * oc if use_scale_shift_norm else oc))
        self.out_layers = nn.Sequential(normalization(oc), nn.SiLU(), nn.Dropout(p=dropout),
                                        zero_module(nn.Conv2d(oc, oc, 3, padding=1)))
        if oc == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = nn.Conv2d(channels, oc, 3, padding=1)
        else:
            self.skip_connection = nn.Conv2d(channels, oc, 1)

    def forward(self, x, emb):
        if self.updown:
            in_rest, in_conv = self.in_layers[:-1], self.in_layers[-1]
            h = in_conv(self.h_upd(in_rest(x)))
            x = self.x_upd(x)
        else:
            h = self.in_layers(x)
        emb_out = self.emb_layers(emb).type(h.dtype)
        while emb_out.dim() < h.dim():
            emb_out = emb_out[..., None]
        if self.use_scale_shift_norm:
            out_norm, out_rest = self.out_layers[0], self.out_layers[1:]
            scale, shift = torch.chunk(emb_out, 2, dim=1)
            h = out_rest(out_norm(h) * (1 + scale) + shift)
        else:
            h = self.out_layers(h + emb_out)
        return self.skip_connection(x) + h


class QKVAttentionLegacy(nn.Module):
    """Head-major interleave: channel = head*3*ch + {q,k,v}*ch + c (Appendix A.2)."""

    def __init__(self, n_heads):
        super().__init__()
        self.n_heads = n_heads

    def forward(self, qkv):
        bs, width, length = qkv.shape
        ch = width // (3 * self.n_heads)
        q, k, v = qkv.reshape(bs * self.n_heads, ch * 3, length).split(ch, dim=1)
        scale = 1.0 / math.sqrt(math.sqrt(ch))
        w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
        w = torch.softmax(w.float(), dim=-1).type(w.dtype)
        a = torch.einsum("bts,bcs->bct", w, v)
        return a.reshape(bs, -1, length)


class QKVAttention(nn.Module):
    """'new attention order': channel = {q,k,v}*heads*ch + head*ch + c."""

    def __init__(self, n_heads):
        super().__init__()
        self.n_heads = n_heads

    def forward(self, qkv):
        bs, width, length = qkv.shape
        ch = width // (3 * self.n_heads)
        q, k, v = qkv.chunk(3, dim=1)
        scale = 1.0 / math.sqrt(math.sqrt(ch))
        w = torch.einsum("bct,bcs->bts", (q * scale).view(bs * self.n_heads, ch, length),
                         (k * scale).view(bs * self.n_heads, ch, length))
        w = torch.softmax(w.float(), dim=-1).type(w.dtype)
        a = torch.einsum("bts,bcs->bct", w, v.reshape(bs * self.n_heads, ch, length))
        return a.reshape(bs, -1, length)


class AttentionBlock(nn.Module):
    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_new_attention_order=False):
        super().__init__()
        self.channels = channels
        if num_head_channels == -1:
            self.num_heads = num_heads
        else:
            assert channels % num_head_channels == 0
            self.num_heads = channels // num_head_channels
        self.norm = normalization(channels)
        self.qkv = nn.Conv1d(channels, channels * 3, 1)
        self.attention = QKVAttention(self.num_heads) if use_new_attention_order else QKVAttentionLegacy(self.num_heads)
        self.proj_out = zero_module(nn.Conv1d(channels, channels, 1))

    def forward(self, x):
        b, c, *spatial = x.shape
        x = x.reshape(b, c, -1)
        h = self.proj_out(self.attention(self.qkv(self.norm(x))))
        return (x + h).reshape(b, c, *spatial)


class RawUNetModel(nn.Module):
    """`torchcfm.models.unet.unet.UNetModel` (Appendix A.3)."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2,
                 num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False):
        super().__init__()
        assert dims == 2
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
        self.dtype = torch.float16 if use_fp16 else torch.float32

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
                layers = [ResBlock(ch, ted, dropout, out_channels=int(mult * model_channels),
                                   use_scale_shift_norm=use_scale_shift_norm)]
                ch = int(mult * model_channels)
                if ds in self.attention_resolutions:
                    layers.append(AttentionBlock(ch, num_heads, num_head_channels, use_new_attention_order))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                if resblock_updown:
                    blk = ResBlock(ch, ted, dropout, out_channels=ch, use_scale_shift_norm=use_scale_shift_norm, down=True)
                else:
                    blk = Downsample(ch, conv_resample, out_channels=ch)
                self.input_blocks.append(TimestepEmbedSequential(blk))
                chans.append(ch)
                ds *= 2

        self.middle_block = TimestepEmbedSequential(
            ResBlock(ch, ted, dropout, use_scale_shift_norm=use_scale_shift_norm),
            AttentionBlock(ch, num_heads, num_head_channels, use_new_attention_order),
            ResBlock(ch, ted, dropout, use_scale_shift_norm=use_scale_shift_norm),
        )

        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [ResBlock(ch + ich, ted, dropout, out_channels=int(model_channels * mult),
                                   use_scale_shift_norm=use_scale_shift_norm)]
                ch = int(model_channels * mult)
                if ds in self.attention_resolutions:
                    layers.append(AttentionBlock(ch, num_heads_upsample, num_head_channels, use_new_attention_order))
                if level and i == num_res_blocks:
                    if resblock_updown:
                        layers.append(ResBlock(ch, ted, dropout, out_channels=ch,
                                               use_scale_shift_norm=use_scale_shift_norm, up=True))
                    else:
                        layers.append(Upsample(ch, conv_resample, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))

        self.out = nn.Sequential(normalization(ch), nn.SiLU(),
                                 zero_module(nn.Conv2d(input_ch, out_channels, 3, padding=1)))

    def forward(self, t, x, y=None):
        assert (y is not None) == (self.num_classes is not None), "must specify y iff the model is class-conditional"
        while t.dim() > 1:
            t = t[:, 0]
        if t.dim() == 0:
            t = t.repeat(x.shape[0])
        emb = self.time_embed(timestep_embedding(t, self.model_channels))
        if self.num_classes is not None:
            assert y.shape == (x.shape[0],)
            emb = emb + self.label_emb(y)
        hs = []
        h = x.type(self.dtype)
        for m in self.input_blocks:
            h = m(h, emb)
            hs.append(h)
        h = self.middle_block(h, emb)
        for m in self.output_blocks:
            h = torch.cat([h, hs.pop()], dim=1)
            h = m(h, emb)
        return self.out(h.type(x.dtype))


def default_channel_mult(image_size: int):
    table = {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4),
             64: (1, 2, 3, 4), 32: (1, 2, 2, 2), 28: (1, 2, 2)}
    if image_size not in table:
        raise ValueError(f"unsupported image size: {image_size}")
    return table[image_size]


class UNetModel(RawUNetModel):
    """`torchcfm.models.unet.UNetModel` (the wrapper, Appendix A.1)."""

    def __init__(self, dim, num_channels, num_res_blocks, channel_mult=None, learn_sigma=False, class_cond=False,
                 num_classes=NUM_CLASSES, use_checkpoint=False, attention_resolutions="16", num_heads=1,
                 num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False, dropout=0,
                 resblock_updown=False, use_fp16=False, use_new_attention_order=False):
        image_size = dim[-1]
        if channel_mult is None:
            channel_mult = default_channel_mult(image_size)
        else:
            channel_mult = tuple(channel_mult)
        attention_ds = [image_size // int(r) for r in str(attention_resolutions).split(",")]
        super().__init__(image_size=image_size, in_channels=dim[0], model_channels=num_channels,
                         out_channels=(dim[0] if not learn_sigma else dim[0] * 2), num_res_blocks=num_res_blocks,
                         attention_resolutions=tuple(attention_ds), dropout=dropout, channel_mult=channel_mult,
                         num_classes=(num_classes if class_cond else None), use_checkpoint=use_checkpoint,
                         use_fp16=use_fp16, num_heads=num_heads, num_head_channels=num_head_channels,
                         num_heads_upsample=num_heads_upsample, use_scale_shift_norm=use_scale_shift_norm,
                         resblock_updown=resblock_updown, use_new_attention_order=use_new_attention_order)

    def forward(self, t, x, y=None, *args, **kwargs):
        return super().forward(t, x, y=y)


CONFIG_A = dict(dim=[3, 256, 256], num_channels=128, num_res_blocks=2, attention_resolutions="16,8", dropout=0.1,
                use_scale_shift_norm=True, num_heads=4, num_head_channels=32, channel_mult=[1, 2, 2, 4])
CONFIG_B = dict(CONFIG_A, class_cond=True, num_classes=3)


def dezero_(net: nn.Module, seed: int = 1984, std: float = 0.02) -> nn.Module:
    """Re-draw every zero-initialised tensor (SURVEY finding 7, section 8c) so parity tests are not vacuous.

    Weight tensors that are exactly zero get N(0, (std * gain)^2) values where gain makes conv outputs O(1);
    zero biases of those same layers get small non-zero values too.  Deterministic for a given seed.
    """
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if p.dim() >= 2 and float(p.abs().max()) == 0.0:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (1.0 / math.sqrt(fan_in)))
                parent = name.rsplit(".", 1)[0]
                b = dict(net.named_parameters()).get(parent + ".bias")
                if b is not None:
                    b.copy_(torch.randn(b.shape, generator=g) * std)
    return net
