"""GPU parity of the BENCHMARKED configuration (config A: 256^2 tiles, 128 base channels, channel_mult [1,2,2,4]) against
the fp32 oracle, forward AND backward, eval and train mode ("What's weak" 2 and 3).

Gates (north_star / SURVEY 8c): velocity and loss rel-L2 <= 1e-2, whole-model gradient rel-L2 <= 2e-2, worst single
tensor <= 8e-2.  The oracle runs on the GPU in TRUE fp32 (TF32 is switched off in tests/conftest.py).

  * eval mode, B = 4: forward + backward through every layer shape of the benchmark (multi-segment concat convs at
    384 / 512 / 768 / 1024 input channels included);
  * train mode, p = 0.1: the engine's stored dropout keep bits (1 bit / element, ops.DROPOUT_TAP) are replayed into the
    oracle through a hook that replaces its nn.Dropout modules -- same masks on both sides, same gates;
  * B = 64 forward under no_grad: the batch of configs[1]; activations up to 3.2 GB (> 2^31 bytes) per tensor;
  * fp16 saturation: the forward storage format has 5 exponent bits; a channel driven past 65 504 must saturate
    (finite), never come back as inf / nan.
"""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _pair(cfg, seed=0):
    from oracle import unet as ounet
    from stain2stain_b200.unet import UNetModel
    torch.manual_seed(seed)
    ref = ounet.UNetModel(**cfg)
    ounet.dezero_(ref, seed=1984)
    net = UNetModel(**cfg)
    net.load_state_dict(ref.state_dict(), strict=True)
    return ref.to(DEV), net.to(DEV)


def _inputs(B, H, seed=1):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x0 = torch.rand(B, 3, H, H, device=DEV, generator=g) * 2 - 1
    x1 = torch.rand(B, 3, H, H, device=DEV, generator=g) * 2 - 1
    t = torch.rand(B, device=DEV, generator=g)
    return x0, x1, t


def _grad_gates(net, ref, total_gate=2e-2, worst_gate=8e-2):
    num = den = 0.0
    worst = (0.0, "")
    for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, n
        assert torch.isfinite(p.grad).all(), n
        d = float((p.grad.double() - q.grad.double()).norm())
        s = float(q.grad.double().norm())
        num += d * d
        den += s * s
        if s > 0 and d / s > worst[0]:
            worst = (d / s, n)
    total = (num / den) ** 0.5
    assert total <= total_gate, f"whole-model gradient rel-L2 {total}; worst tensor {worst}"
    assert worst[0] <= worst_gate, f"worst per-tensor gradient rel-L2 {worst} (whole model {total})"
    return total, worst


def _step_both(net, ref, x0, x1, t):
    from oracle.flow import rel_l2
    xt = (1 - t)[:, None, None, None] * x0 + t[:, None, None, None] * x1
    v = net(t, xt)
    v_ref = ref(t, xt)
    r = rel_l2(v, v_ref)
    assert r <= 1e-2, f"velocity rel-L2 {r}"
    loss = torch.mean((v - (x1 - x0)) ** 2)
    loss_ref = torch.mean((v_ref - (x1 - x0)) ** 2)
    assert abs(float(loss) - float(loss_ref)) <= 1e-2 * float(loss_ref), (float(loss), float(loss_ref))
    loss.backward()
    loss_ref.backward()
    return r


def test_config_a_forward_backward_parity_eval():
    from oracle import unet as ounet
    ref, net = _pair(ounet.CONFIG_A)
    ref.eval(), net.eval()
    x0, x1, t = _inputs(4, 256, seed=21)
    _step_both(net, ref, x0, x1, t)
    _grad_gates(net, ref)


class _ReplayDropout(nn.Module):
    """Stands in for the oracle's nn.Dropout: multiplies by the engine's keep mask / (1 - p)."""

    def __init__(self, p):
        super().__init__()
        self.p, self.keep = p, None

    def forward(self, x):
        assert self.keep is not None and self.keep.shape == x.shape, (None if self.keep is None else self.keep.shape, x.shape)
        return x * self.keep * (1.0 / (1.0 - self.p))


def _unpack_keep_bits(mask_u8):
    """uint8 [B,H,W,C/8] keep bits (bit j of byte c8 = channel 8*c8 + j) -> float NCHW [B,C,H,W]."""
    B, H, W, C8 = mask_u8.shape
    bits = (mask_u8.unsqueeze(-1) >> torch.arange(8, device=mask_u8.device, dtype=torch.uint8)) & 1
    return bits.reshape(B, H, W, C8 * 8).permute(0, 3, 1, 2).float().contiguous()


def test_config_a_train_mode_dropout_replayed_into_oracle():
    from oracle import unet as ounet
    from stain2stain_b200 import ops
    p = ounet.CONFIG_A["dropout"]
    assert p == 0.1
    ref, net = _pair(ounet.CONFIG_A)
    ref.train(), net.train()
    x0, x1, t = _inputs(4, 256, seed=22)
    xt = (1 - t)[:, None, None, None] * x0 + t[:, None, None, None] * x1
    # engine forward first: it draws the masks (Philox, one counter per 8 elements) and stores the keep bits
    ops.DROPOUT_TAP = []
    try:
        v = net(t, xt)
        masks = list(ops.DROPOUT_TAP)
    finally:
        ops.DROPOUT_TAP = None
    blocks = [m for m in ref.modules() if isinstance(m, ounet.ResBlock)]
    assert len(masks) == len(blocks) == 22
    kept = 0.0
    for blk, m in zip(blocks, masks):  # module order == execution order (input, middle, output blocks)
        rd = _ReplayDropout(p)
        rd.keep = _unpack_keep_bits(m)
        assert rd.keep.shape[1] == blk.out_channels
        kept += float(rd.keep.mean()) / len(blocks)
        blk.out_layers[2] = rd
    assert abs(kept - (1 - p)) < 2e-3, f"keep rate {kept}"  # statistical check of the generator itself
    v_ref = ref(t, xt)
    from oracle.flow import rel_l2
    r = rel_l2(v, v_ref)
    assert r <= 1e-2, f"train-mode velocity rel-L2 {r}"
    loss = torch.mean((v - (x1 - x0)) ** 2)
    loss_ref = torch.mean((v_ref - (x1 - x0)) ** 2)
    assert abs(float(loss) - float(loss_ref)) <= 1e-2 * float(loss_ref)
    loss.backward()
    loss_ref.backward()
    _grad_gates(net, ref)


def test_config_a_batch64_forward_parity():
    """configs[1]'s per-GPU batch: tensors beyond 2^31 bytes (the [64,256,256,384] concat input is 3.2 GB in 16 bit)."""
    from oracle import unet as ounet
    from oracle.flow import rel_l2
    ref, net = _pair(ounet.CONFIG_A)
    ref.eval(), net.eval()
    x0, x1, t = _inputs(64, 256, seed=23)
    xt = (1 - t)[:, None, None, None] * x0 + t[:, None, None, None] * x1
    with torch.no_grad():
        v = net(t, xt)
        v_ref = torch.cat([ref(t[i:i + 16], xt[i:i + 16]) for i in range(0, 64, 16)])
    assert torch.isfinite(v).all()
    r = rel_l2(v, v_ref)
    assert r <= 1e-2, f"B=64 velocity rel-L2 {r}"
    # per-sample: no sample may be off (an indexing overflow would corrupt the high samples only)
    per = ((v - v_ref).double().flatten(1).norm(dim=1) / v_ref.double().flatten(1).norm(dim=1)).max()
    assert float(per) <= 2e-2, f"worst sample rel-L2 {float(per)}"


def test_fp16_forward_saturates_instead_of_overflowing():
    """One channel's conv weights are scaled until the fp32 oracle's activation passes fp16's 65 504: the engine's
    16-bit stores saturate (cvt.rn.satfinite), so everything downstream stays finite -- never inf / nan."""
    from oracle import unet as ounet
    from stain2stain_b200 import kernels as K
    small = dict(dim=[3, 64, 64], num_channels=64, num_res_blocks=1, attention_resolutions="16,8", dropout=0.0,
                 use_scale_shift_norm=True, num_heads=4, num_head_channels=32, channel_mult=[1, 2, 2, 4])
    ref, net = _pair(small)
    ref.eval(), net.eval()
    x0, _, t = _inputs(2, 64, seed=24)
    seen = {}
    conv = ref.input_blocks[1][0].in_layers[2]
    conv.register_forward_hook(lambda m, i, o: seen.__setitem__("max", float(o.abs().max())))
    with torch.no_grad():
        for scale in (1e2, 1e3, 1e4, 1e5, 1e6):
            w = conv.weight.detach().clone()
            w[0] *= scale
            conv.weight.copy_(w)
            net.load_state_dict(ref.state_dict(), strict=True)
            ref(t, x0)
            if seen["max"] > 6.0e4:
                break
        assert seen["max"] > 6.0e4, seen
        v = net(t, x0)
    assert torch.isfinite(v).all(), "fp16 overflow leaked an inf/nan into the velocity"
    if K.ACT == K.FMT_F16:
        # the conv epilogue itself: accumulators far beyond the fp16 range come out as +-65504
        xs = K.from_float(torch.full((1, 16, 16, 64), 60.0, device=DEV), K.ACT)
        wp = K.from_float(torch.full((64, 64), 60.0, device=DEV), K.ACT)
        out = K.to_float(K.conv_fwd([(xs, 1, 1)], wp, 64, 16, 16), K.ACT)
        assert torch.isfinite(out).all() and float(out.max()) == 65504.0
