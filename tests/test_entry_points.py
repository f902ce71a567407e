"""Config-level boundary: the Hydra `_target_` / `_partial_` yamls instantiate the B200 drop-ins -- this
package's own configs/model/*.yaml and, when /root/reference is present, the REFERENCE'S yaml files unmodified through
the target remap.  CPU tier: construction only.  GPU tier: the train / infer entry points run end to end."""
import functools
import os

import pytest
import torch

from stain2stain_b200 import entry, hydra_lite

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "configs", "model")
REF = "/root/reference/configs/model"
SMALL_NET = dict(dim=[3, 64, 64], num_channels=64, num_res_blocks=1, attention_resolutions="16,8", dropout=0.1,
                 use_scale_shift_norm=True, num_heads=4, num_head_channels=32, channel_mult=[1, 2, 2, 4])


def _nparams(m):
    return sum(p.numel() for p in m.parameters())


def test_hydra_lite_semantics():
    cfg = {"a": 3, "n": {"_target_": "torch.nn.Linear", "in_features": "${a}", "out_features": 2},
           "opt": {"_target_": "torch.optim.Adam", "_partial_": True, "lr": "1e-4", "weight_decay": 0.0},
           "lst": [{"_target_": "torch.nn.ReLU"}, 5], "name": "x_${a}"}
    out = hydra_lite.instantiate(cfg)
    assert isinstance(out["n"], torch.nn.Linear) and out["n"].in_features == 3
    assert isinstance(out["opt"], functools.partial) and out["opt"].keywords["lr"] == 1e-4
    assert isinstance(out["lst"][0], torch.nn.ReLU) and out["lst"][1] == 5 and out["name"] == "x_3"
    opt = out["opt"](params=out["n"].parameters())
    assert isinstance(opt, torch.optim.Adam)
    from stain2stain_b200.optim import FusedAdam
    assert hydra_lite.instantiate(cfg, fused_optimizer=True)["opt"].func is FusedAdam


@pytest.mark.parametrize("name,params", [("conditional_flow_matching.yaml", 70_954_883),
                                         ("class_conditional_flow_matching.yaml", 70_956_419)])
def test_own_yaml_builds_the_dropins(name, params):
    with torch.device("meta"):
        model = hydra_lite.instantiate(hydra_lite.load_yaml(os.path.join(CFG, name)))
    assert _nparams(model) == params
    assert type(model.net).__module__ == "stain2stain_b200.unet"
    assert isinstance(model.solver, functools.partial) and isinstance(model.optimizer, functools.partial)
    assert set(model.configure_optimizers()) == {"optimizer", "lr_scheduler"}


def test_own_multitask_yaml():
    with torch.device("meta"):
        model = hydra_lite.instantiate(hydra_lite.load_yaml(os.path.join(CFG, "conditional_flow_matching_multitask_multiclass.yaml")))
    assert _nparams(model.encoder) == 18_851_136 and _nparams(model.flow_decoder) == 12_934_467
    assert _nparams(model.seg_decoder) == 12_539_845  # the reference components' parameter counts
    assert model.seg_decoder.outc.out_channels == 5 and model.num_classes == 5


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is only present in the build container")
@pytest.mark.parametrize("name,params", [("conditional_flow_matching.yaml", 70_954_883),
                                         ("class_conditional_flow_matching.yaml", 70_956_419),
                                         ("conditional_flow_matching_multitask_multiclass.yaml", 44_325_448)])
def test_reference_yaml_unmodified_builds_the_dropins(name, params):
    with torch.device("meta"):
        model = hydra_lite.instantiate(hydra_lite.load_yaml(os.path.join(REF, name)), remap=True, fused_optimizer=True)
    assert type(model).__module__.startswith("stain2stain_b200.")
    assert _nparams(model) == params
    from stain2stain_b200.optim import FusedAdam
    assert model.optimizer.func is FusedAdam and model.solver.func.__module__ == "stain2stain_b200.neural_ode"
    assert model.solver.keywords["solver"] == "dopri5"  # the reference's yaml value is kept


def test_denormalize_and_uint8():
    t = torch.tensor([[[[-1.0, 0.0]], [[1.0, 2.0]], [[0.5, -3.0]]]])
    u = entry.to_uint8_hwc(t)
    assert u.shape == (1, 1, 2, 3) and u.dtype == torch.uint8
    assert u[0, 0, 0].tolist() == [0, 255, 191] and u[0, 0, 1].tolist() == [128, 255, 0]


@pytest.mark.gpu
def test_train_and_infer_entry_points(tmp_path):
    cfg = hydra_lite.load_yaml(os.path.join(CFG, "conditional_flow_matching.yaml"))
    cfg["net"].update(SMALL_NET)
    ckpt = str(tmp_path / "last.ckpt")
    out = entry.train(cfg, steps=12, batch=4, device="cuda", ckpt_path=ckpt)
    assert out["kind"] == "simple" and out["logged"] == ["train/loss"]
    assert all(torch.isfinite(torch.tensor(out["losses"])))
    assert min(out["losses"][-3:]) < out["loss_first"], out["losses"]  # the zero-init net starts at mean(ut^2): it learns
    state = torch.load(ckpt, map_location="cpu", weights_only=False)
    assert set(state) >= {"state_dict", "optimizer_states", "epoch", "global_step"}
    assert all(k.startswith("net.") for k in state["state_dict"])
    src = torch.rand(3, 3, 64, 64) * 2 - 1
    gen, img = entry.infer_simple(cfg, ckpt, src, num_steps=6)
    assert gen.shape == (3, 3, 64, 64) and img.shape == (3, 64, 64, 3) and img.dtype == torch.uint8
    assert torch.isfinite(gen).all()
    # the same weights through the oracle: Euler grids agree (PSNR >= 40 dB)
    from oracle import flow as oflow
    from oracle import unet as ounet
    ref = ounet.UNetModel(**{k: v for k, v in cfg["net"].items() if k != "_target_"}).cuda()
    ref.load_state_dict({k[4:]: v for k, v in state["state_dict"].items()})
    want = oflow.generate(ref, src.cuda(), num_steps=6, solver="euler")
    assert oflow.psnr(gen, want) >= 40.0


@pytest.mark.gpu
def test_multitask_and_class_conditional_entry_points():
    cfg = hydra_lite.load_yaml(os.path.join(CFG, "conditional_flow_matching_multitask_multiclass.yaml"))
    cfg["encoder"]["features"] = [64, 128, 256]
    for d in ("flow_decoder", "seg_decoder"):
        cfg[d].update(bottleneck_channels=256, features=[128, 64])
    cfg["flow_decoder"]["time_emb_dim"] = cfg["time_emb_dim"] = 64
    out = entry.train(cfg, steps=4, batch=2, size=32, device="cuda")
    assert out["kind"] == "multitask" and "train/seg_dice" in out["logged"] and len(out["logged"]) == 5
    gen, img, mask = entry.infer_simple(cfg, None, torch.rand(2, 3, 32, 32) * 2 - 1, num_steps=3)
    assert gen.shape == (2, 3, 32, 32) and mask.shape == (2, 1, 32, 32) and mask.dtype == torch.int64
    cfg = hydra_lite.load_yaml(os.path.join(CFG, "class_conditional_flow_matching.yaml"))
    cfg["net"].update(SMALL_NET)
    out = entry.train(cfg, steps=3, batch=2, device="cuda")
    assert out["kind"] == "class_conditional"
    gen, img = entry.infer_simple(cfg, None, torch.rand(2, 3, 64, 64) * 2 - 1, num_steps=4, target_class=2)
    assert gen.shape == (2, 3, 64, 64)
