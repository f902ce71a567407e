"""Checkpoint interop: a Lightning-style checkpoint written by the reference's LitModule
(`{'state_dict': {'net.<torchcfm names>': fp32 OIHW ...}}`, loaded by src/infer_simple_flowmatching.py:21-22,51 with
`torch.load(..., weights_only=False)['state_dict']` + `load_state_dict`) loads strictly into the B200 modules and back.
Module construction and (de)serialisation are host logic: this runs on CPU; the numerics after loading are covered by
the GPU parity tests, which load oracle weights the same way."""
import functools
import io

import pytest
import torch

from oracle import flow as oflow
from oracle import multitask as omt
from oracle import unet as ounet

SMALL = dict(dim=[3, 64, 64], num_channels=64, num_res_blocks=1, attention_resolutions="16,8", dropout=0.1,
             use_scale_shift_norm=True, num_heads=4, num_head_channels=32, channel_mult=[1, 2, 2, 4])


def _roundtrip(obj):
    buf = io.BytesIO()
    torch.save(obj, buf)
    buf.seek(0)
    return torch.load(buf, map_location="cpu", weights_only=False)


def _reference_lit(cfg, class_cond):
    """The reference's own LitModule when /root/reference is present, else a stand-in with the same `net.` prefix."""
    from oracle import ref_bridge as rb
    net = ounet.dezero_(ounet.UNetModel(**cfg))
    if rb.available():
        name = "class_conditional_flow_matching" if class_cond else "conditional_flow_matching"
        mod = rb.reference_module(f"src.models.{name}")
        cls = mod.ClassConditionalFlowMatchingLitModule if class_cond else mod.ConditionalFlowMatchingLitModule
        kw = {} if class_cond else dict(log_images=False)
        return cls(net=net, flow_matcher=oflow.ConditionalFlowMatcher(0.0),
                   solver=functools.partial(oflow.NeuralODE, solver="dopri5"),
                   optimizer=functools.partial(torch.optim.Adam, lr=1e-4), scheduler=None, **kw)
    holder = torch.nn.Module()
    holder.net = net
    return holder


@pytest.mark.parametrize("class_cond", [False, True])
def test_reference_checkpoint_loads_strictly_and_back(class_cond):
    from stain2stain_b200.flow_matching import ConditionalFlowMatcher
    from stain2stain_b200.lit import ClassConditionalFlowMatchingLitModule, ConditionalFlowMatchingLitModule
    from stain2stain_b200.neural_ode import NeuralODE
    from stain2stain_b200.unet import UNetModel
    cfg = dict(SMALL, class_cond=True, num_classes=3) if class_cond else dict(SMALL)
    torch.manual_seed(3)
    ref = _reference_lit(cfg, class_cond)
    adam = torch.optim.Adam(ref.parameters(), lr=1e-4)
    ckpt = _roundtrip({"state_dict": ref.state_dict(), "optimizer_states": [adam.state_dict()], "epoch": 7})
    assert all(k.startswith("net.") for k in ckpt["state_dict"])
    cls = ClassConditionalFlowMatchingLitModule if class_cond else ConditionalFlowMatchingLitModule
    lit = cls(net=UNetModel(**cfg), flow_matcher=ConditionalFlowMatcher(0.0),
              solver=functools.partial(NeuralODE, solver="euler"), optimizer=functools.partial(torch.optim.Adam, lr=1e-4))
    res = lit.load_state_dict(ckpt["state_dict"], strict=True)  # infer_simple_flowmatching.py:51
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in ref.state_dict().items():
        assert torch.equal(lit.state_dict()[k], v) and lit.state_dict()[k].dtype == torch.float32
    # and back: a checkpoint written by the B200 module loads into the reference / oracle module
    back = _roundtrip({"state_dict": lit.state_dict()})
    res = ref.load_state_dict(back["state_dict"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    # the module (with its conv plans and caches) stays picklable: Lightning's save_hyperparameters pickles `net`
    clone = _roundtrip(lit.net)
    assert list(clone.state_dict()) == list(lit.net.state_dict())


def test_multitask_checkpoint_keys_include_batchnorm_buffers():
    from stain2stain_b200 import multitask as mt
    from stain2stain_b200.flow_matching import ConditionalFlowMatcher
    f = [64, 128, 256]
    torch.manual_seed(4)
    ref = omt.build(f, 5, 64)
    lit = mt.MultiTaskFlowMatchingLitModule(
        mt.SharedEncoder(3, f, True), mt.FlowMatchingDecoder(f[-1], f[:-1][::-1], 3, 64, True),
        mt.SegmentationDecoder(f[-1], f[:-1][::-1], 5, True), ConditionalFlowMatcher(0.0), num_classes=5,
        time_emb_dim=64, log_images=False)
    sd = _roundtrip({"state_dict": ref.state_dict()})["state_dict"]
    res = lit.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert "encoder.inc.double_conv.1.running_mean" in sd and "encoder.inc.double_conv.1.num_batches_tracked" in sd
    assert list(lit.state_dict()) == list(ref.state_dict())
