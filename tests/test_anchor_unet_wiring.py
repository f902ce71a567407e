"""Independent anchor for the WIRING of the torchcfm / guided-diffusion UNet.

tests/test_anchors.py pins the blocks one by one; what it cannot see is how they are strung together: which blocks carry
attention, the order the skips are popped in, where the resampling layers sit, how `t` and `y` enter, which state_dict
key feeds which layer.  oracle/unet.py and stain2stain_b200/unet.py both build that tree as `nn.Module`s.  Here the same
network is evaluated a third way: ONE function that walks the published construction rule (level / res-block / `ds in
attention_resolutions` / `level and i == num_res_blocks`) and pulls every tensor out of a flat state_dict BY NAME with
stock `torch.nn.functional` ops.  Nothing is imported from oracle/ for the expected values, and no module tree is used:

  * a block wired differently, a skip popped in another order, attention in another set of blocks -> different output;
  * a key the rule expects but the tree lacks -> KeyError; a key the tree has but the rule never reads -> assertion.

The engine is held to the oracle by the `gpu` tests, so pinning the oracle's wiring (and the product's key scheme, whose
module tree builds on the CPU) pins the engine's.
"""
import math

import pytest
import torch
import torch.nn.functional as F


def _walk_unet(sd, t, x, y, *, in_ch, mc, out_ch, nrb, channel_mult, attn_ds, num_heads, head_ch, heads_up=-1,
               num_classes=None, new_order=False):
    """guided-diffusion UNet forward (torchcfm argument order `t, x, y`), scale-shift-norm ResBlocks, conv resampling,
    eval mode.  Returns (output, set of state_dict keys read)."""
    used = set()

    def P(key):
        used.add(key)
        return sd[key]

    def gn(h, pre):
        return F.group_norm(h.float(), 32, P(pre + ".weight"), P(pre + ".bias"), eps=1e-5).type(h.dtype)

    def conv(h, pre, stride=1, padding=1):
        return F.conv2d(h, P(pre + ".weight"), P(pre + ".bias"), stride=stride, padding=padding)

    def res(h, emb, pre, cin, cout):
        z = conv(F.silu(gn(h, pre + "in_layers.0")), pre + "in_layers.2")
        e = F.linear(F.silu(emb), P(pre + "emb_layers.1.weight"), P(pre + "emb_layers.1.bias"))[:, :, None, None]
        scale, shift = torch.chunk(e, 2, dim=1)
        z = gn(z, pre + "out_layers.0") * (1 + scale) + shift
        z = conv(F.silu(z), pre + "out_layers.3")  # dropout is the identity in eval mode
        if cin != cout:
            h = conv(h, pre + "skip_connection", padding=0)
        return h + z

    def attn(h, pre, heads):
        B, C, H, W = h.shape
        n = C // head_ch if head_ch != -1 else heads
        ch = C // n
        a = h.reshape(B, C, H * W)
        qkv = F.conv1d(gn(a, pre + "norm"), P(pre + "qkv.weight"), P(pre + "qkv.bias"))
        s = 1.0 / math.sqrt(math.sqrt(ch))
        if new_order:   # QKVAttention: [q | k | v] thirds, heads inside each third
            q, k, v = (part.reshape(B * n, ch, -1) for part in qkv.chunk(3, dim=1))
        else:           # QKVAttentionLegacy: per head [q | k | v]
            q, k, v = qkv.reshape(B * n, 3 * ch, -1).split(ch, dim=1)
        w = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s).float(), dim=-1)
        o = torch.einsum("bts,bcs->bct", w, v).reshape(B, C, -1)
        o = F.conv1d(o, P(pre + "proj_out.weight"), P(pre + "proj_out.bias"))
        return (a + o).reshape(B, C, H, W)

    # ---- embedding: t raw (no x1000), cos first
    while t.dim() > 1:
        t = t[:, 0]
    if t.dim() == 0:
        t = t.repeat(x.shape[0])
    half = mc // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    temb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    emb = F.linear(F.silu(F.linear(temb, P("time_embed.0.weight"), P("time_embed.0.bias"))),
                   P("time_embed.2.weight"), P("time_embed.2.bias"))
    assert (y is not None) == (num_classes is not None)
    if num_classes is not None:
        emb = emb + F.embedding(y, P("label_emb.weight"))
    if heads_up == -1:
        heads_up = num_heads

    # ---- down path
    hs, chans = [], []
    ch = int(channel_mult[0] * mc)
    h = conv(x, "input_blocks.0.0")
    hs.append(h), chans.append(ch)
    idx, ds = 0, 1
    for level, mult in enumerate(channel_mult):
        for _ in range(nrb):
            idx += 1
            cout = int(mult * mc)
            h = res(h, emb, f"input_blocks.{idx}.0.", ch, cout)
            ch = cout
            if ds in attn_ds:
                h = attn(h, f"input_blocks.{idx}.1.", num_heads)
            hs.append(h), chans.append(ch)
        if level != len(channel_mult) - 1:
            idx += 1
            h = conv(h, f"input_blocks.{idx}.0.op", stride=2)
            hs.append(h), chans.append(ch)
            ds *= 2
    # ---- middle
    h = res(h, emb, "middle_block.0.", ch, ch)
    h = attn(h, "middle_block.1.", num_heads)
    h = res(h, emb, "middle_block.2.", ch, ch)
    # ---- up path
    j = 0
    for level, mult in reversed(list(enumerate(channel_mult))):
        for i in range(nrb + 1):
            skip, ich = hs.pop(), chans.pop()
            h = torch.cat([h, skip], dim=1)
            cout = int(mc * mult)
            h = res(h, emb, f"output_blocks.{j}.0.", ch + ich, cout)
            ch = cout
            sub = 1
            if ds in attn_ds:
                h = attn(h, f"output_blocks.{j}.{sub}.", heads_up)
                sub += 1
            if level and i == nrb:
                h = conv(F.interpolate(h, scale_factor=2, mode="nearest"), f"output_blocks.{j}.{sub}.conv")
                ds //= 2
            j += 1
    assert not hs
    h = conv(F.silu(gn(h, "out.0")), "out.2")
    assert h.shape[1] == out_ch
    return h, used


def _dezero(sd, seed):
    """Every zero-initialised tensor (out.2, out_layers.3, proj_out) gets small random values so all paths contribute."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        v = v.detach().clone()
        if v.is_floating_point() and float(v.abs().max()) == 0.0:
            v = torch.randn(v.shape, generator=g) * (1.0 / math.sqrt(v[0].numel()) if v.dim() >= 2 else 0.05)
        elif k.endswith(".bias"):
            v = v + torch.randn(v.shape, generator=g) * 0.05
        out[k] = v
    return out


CASES = {
    # wrapper kwargs (configs/model/conditional_flow_matching.yaml:16-26 spelling), scaled down
    "attn_two_levels_legacy": dict(dim=[3, 32, 32], num_channels=32, num_res_blocks=2, channel_mult=[1, 2, 2],
                                   attention_resolutions="16,8", num_heads=4, num_head_channels=16,
                                   use_scale_shift_norm=True, dropout=0.1),
    "class_cond_heads_by_count": dict(dim=[3, 32, 32], num_channels=32, num_res_blocks=1, channel_mult=[1, 2, 4],
                                      attention_resolutions="8", num_heads=2, num_head_channels=-1,
                                      use_scale_shift_norm=True, dropout=0.0, class_cond=True, num_classes=3),
    "middle_attention_only_like_config_a": dict(dim=[3, 32, 32], num_channels=32, num_res_blocks=2,
                                                channel_mult=[1, 2, 2, 4], attention_resolutions="2,1", num_heads=4,
                                                num_head_channels=32, use_scale_shift_norm=True, dropout=0.1),
    "new_attention_order": dict(dim=[3, 16, 16], num_channels=32, num_res_blocks=1, channel_mult=[1, 2],
                                attention_resolutions="8", num_heads=2, num_head_channels=16,
                                use_scale_shift_norm=True, dropout=0.0, use_new_attention_order=True),
}


def _walker_kwargs(cfg):
    size = cfg["dim"][-1]
    return dict(in_ch=cfg["dim"][0], mc=cfg["num_channels"], out_ch=cfg["dim"][0], nrb=cfg["num_res_blocks"],
                channel_mult=cfg["channel_mult"],
                attn_ds={size // int(r) for r in cfg["attention_resolutions"].split(",")},  # image_size // res (A.1)
                num_heads=cfg["num_heads"], head_ch=cfg["num_head_channels"],
                num_classes=cfg["num_classes"] if cfg.get("class_cond") else None,
                new_order=cfg.get("use_new_attention_order", False))


def _inputs(cfg, seed=3, B=2):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, *cfg["dim"], generator=g) * 2 - 1
    t = torch.rand(B, generator=g)
    y = torch.randint(0, cfg["num_classes"], (B,), generator=g) if cfg.get("class_cond") else None
    return t, x, y


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_unet_equals_the_state_dict_walker(name):
    from oracle import unet as ounet
    cfg = CASES[name]
    torch.manual_seed(11)
    net = ounet.UNetModel(**cfg).eval()
    sd = _dezero(net.state_dict(), seed=5)
    net.load_state_dict(sd, strict=True)
    t, x, y = _inputs(cfg)
    want, used = _walk_unet(sd, t, x, y, **_walker_kwargs(cfg))
    assert used == set(sd), (sorted(set(sd) - used), sorted(used - set(sd)))
    with torch.no_grad():
        got = net(t, x, y=y) if y is not None else net(t, x)
    assert got.shape == want.shape == x.shape
    err = float((got - want).abs().max() / want.abs().max())
    assert err <= 2e-5, err
    assert float(want.abs().max()) > 1e-3  # the de-zeroed head really produces a signal
    # the three spellings of t the reference produces: [B], 0-dim (ODE solver), [B,1,1,1] (padded like x)
    with torch.no_grad():
        t0 = t[:1].reshape(())
        a = net(t0, x, y=y) if y is not None else net(t0, x)
        b, _ = _walk_unet(sd, t0.repeat(x.shape[0]), x, y, **_walker_kwargs(cfg))
        assert float((a - b).abs().max() / b.abs().max()) <= 2e-5
        c = net(t.reshape(-1, 1, 1, 1), x, y=y) if y is not None else net(t.reshape(-1, 1, 1, 1), x)
        assert torch.equal(c, got)


@pytest.mark.parametrize("name", sorted(CASES))
def test_product_state_dict_is_exactly_what_the_walker_reads(name):
    """The engine's module tree (buildable without a GPU) exposes the same keys and shapes the construction rule reads:
    a checkpoint written by torchcfm's UNet loads strictly, and every tensor lands in the layer the rule assigns it to."""
    from oracle import unet as ounet
    from stain2stain_b200.unet import UNetModel
    cfg = CASES[name]
    torch.manual_seed(11)
    prod = UNetModel(**cfg)
    sd = _dezero(prod.state_dict(), seed=5)
    t, x, y = _inputs(cfg)
    _, used = _walk_unet(sd, t, x, y, **_walker_kwargs(cfg))  # shape errors inside F.conv2d / F.linear would raise here
    assert used == set(sd), (sorted(set(sd) - used), sorted(used - set(sd)))
    ref_sd = ounet.UNetModel(**cfg).state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v.shape) for k, v in ref_sd.items()}


def test_walker_reads_every_tensor_of_config_a_and_counts_70_954_883():
    """Known answer: configs[1]'s UNet has 70 954 883 parameters, and with `attention_resolutions: "16,8"`
    at 256x256 (ds 16 and 32, never reached by four levels) attention sits in the middle block only.  The walker is run
    over tensors of config A's shapes at a 32x32 input (the wiring does not depend on the spatial size; zeros are enough,
    only the keys it reads matter) and must consume the whole state_dict."""
    from oracle import unet as ounet
    cfg = dict(dim=[3, 256, 256], num_channels=128, num_res_blocks=2, channel_mult=[1, 2, 2, 4],
               attention_resolutions="16,8", num_heads=4, num_head_channels=32, use_scale_shift_norm=True, dropout=0.1)
    with torch.device("meta"):
        net = ounet.UNetModel(**cfg)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = {k: torch.zeros(s) for k, s in shapes.items()}
    t, x, _ = _inputs(dict(dim=[3, 32, 32]), B=1)
    _, used = _walk_unet(sd, t, x, None, **_walker_kwargs(cfg))
    assert used == set(sd)
    assert [k for k in sd if ".qkv." in k] == ["middle_block.1.qkv.weight", "middle_block.1.qkv.bias"]
    assert sum(math.prod(s) for s in shapes.values()) == 70_954_883


@pytest.mark.gpu
def test_engine_unet_equals_the_state_dict_walker():
    """The B200 engine against the walker directly (not through the oracle): same gate as the oracle parity tests,
    velocity rel-L2 <= 1e-2 (16-bit tensor-core operands vs true fp32; TF32 is off, tests/conftest.py)."""
    from stain2stain_b200.unet import UNetModel
    cfg = dict(dim=[3, 64, 64], num_channels=64, num_res_blocks=1, attention_resolutions="16,8", dropout=0.0,
               use_scale_shift_norm=True, num_heads=4, num_head_channels=32, channel_mult=[1, 2, 2, 4])
    dev = torch.device("cuda", 0)
    torch.manual_seed(11)
    net = UNetModel(**cfg)
    sd = _dezero(net.state_dict(), seed=5)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    t, x, _ = _inputs(cfg, B=2)
    t, x = t.to(dev), x.to(dev)
    want, used = _walk_unet({k: v.to(dev) for k, v in sd.items()}, t, x, None, **_walker_kwargs(cfg))
    assert used == set(sd)
    with torch.no_grad():
        got = net(t, x)
    rel = float((got.double() - want.double()).norm() / want.double().norm())
    assert rel <= 1e-2, rel
