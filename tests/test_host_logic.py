"""CPU tests of the host side that mirrors the reference's interfaces (no kernels run here):
LitModule surface (src/models/conditional_flow_matching.py:9-170, class_conditional_flow_matching.py:8-190),
the `torchdyn.core.NeuralODE` stand-in on arbitrary vector fields, the flow matcher, module tree / state_dict keys
of the B200 UNet vs the oracle, and conv-plan bookkeeping."""
import functools
import pickle

import pytest
import torch
import torch.nn as nn

from oracle import flow as oflow
from oracle import unet as ounet
from stain2stain_b200.flow_matching import ConditionalFlowMatcher
from stain2stain_b200.lit import (ClassConditionalFlowMatchingLitModule, ConditionalFlowMatchingLitModule,
                                  ConditionalWrapper)
from stain2stain_b200.neural_ode import NeuralODE
from stain2stain_b200.ops import ConvPlan, Seg
from stain2stain_b200.unet import RawUNetModel, UNetModel


class TinyField(nn.Module):
    """A stand-in `net(t, x, y=None)` so the generic (reference-shaped) LitModule path runs on CPU."""

    def __init__(self, classes=0):
        super().__init__()
        self.conv = nn.Conv2d(3, 3, 3, padding=1)
        self.emb = nn.Embedding(classes, 3) if classes else None

    def forward(self, t, x, y=None):
        while t.dim() > 1:
            t = t[:, 0]
        if t.dim() == 0:
            t = t.repeat(x.shape[0])
        h = self.conv(x) * (1 + t.view(-1, 1, 1, 1))
        if self.emb is not None:
            h = h + self.emb(y).view(-1, 3, 1, 1)
        return h


def _lit(net=None, solver="euler", **kw):
    return ConditionalFlowMatchingLitModule(
        net=net or TinyField(), flow_matcher=ConditionalFlowMatcher(0.0),
        solver=functools.partial(NeuralODE, solver=solver, sensitivity="adjoint", atol=1e-4, rtol=1e-4),
        optimizer=functools.partial(torch.optim.Adam, lr=1e-3, weight_decay=0.0),
        scheduler=functools.partial(torch.optim.lr_scheduler.ReduceLROnPlateau, mode="min", factor=0.1, patience=10),
        **kw)


def test_unet_tree_matches_oracle_for_every_reference_spelling():
    with torch.device("meta"):
        for cfg in (ounet.CONFIG_A, ounet.CONFIG_B):
            a, b = ounet.UNetModel(**cfg).state_dict(), UNetModel(**cfg).state_dict()
            assert list(a) == list(b) and all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
        raw = dict(image_size=256, in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2,
                   attention_resolutions=[16, 8], dropout=0.1, channel_mult=(1, 2, 2, 4), num_head_channels=32,
                   use_scale_shift_norm=True)  # raw ctor spelling: src/models/components/unet_4to3.py:51-67
        a, b = ounet.RawUNetModel(**raw).state_dict(), RawUNetModel(**raw).state_dict()
        assert list(a) == list(b) and sum(v.numel() for v in b.values()) == 76_214_275 - 128 * 9  # 3-ch stem
    net = UNetModel(**dict(ounet.CONFIG_A, num_channels=32))
    assert float(net.out[2].weight.abs().max()) == 0.0  # zero_module init like the reference


def test_unsupported_reference_options_raise_instead_of_falling_back():
    with pytest.raises(NotImplementedError):
        UNetModel(dim=[3, 32, 32], num_channels=32, num_res_blocks=1, use_scale_shift_norm=False)
    with pytest.raises(NotImplementedError):
        UNetModel(dim=[3, 32, 32], num_channels=32, num_res_blocks=1, use_scale_shift_norm=True, use_fp16=True)
    with pytest.raises(ValueError):
        UNetModel(dim=[3, 48, 48], num_channels=32, num_res_blocks=1, use_scale_shift_norm=True)


def test_litmodule_surface_matches_reference():
    lit = _lit()
    x0, x1 = torch.rand(2, 3, 8, 8), torch.rand(2, 3, 8, 8)
    t = torch.tensor([0.2, 0.9])
    loss = lit.model_step((x0, x1), t=t)
    want = oflow.model_step(lit.net, oflow.ConditionalFlowMatcher(0.0), (x0, x1), t=t)
    assert torch.allclose(loss, want)
    assert lit.training_step((x0, x1), 0).dim() == 0 and "train/loss" in lit.logged
    lit.validation_step((x0, x1), 0), lit.test_step((x0, x1), 0)
    assert {"val/loss", "test/loss"} <= set(lit.logged)
    cfg = lit.configure_optimizers()
    assert isinstance(cfg["optimizer"], torch.optim.Adam)
    assert cfg["lr_scheduler"]["monitor"] == "val/loss" and cfg["lr_scheduler"]["interval"] == "epoch"
    assert cfg["lr_scheduler"]["frequency"] == 1
    assert all(k.startswith("net.") for k in lit.state_dict())
    pickle.loads(pickle.dumps(TinyField()))  # nets must stay picklable (save_hyperparameters pickles `net`)


def test_generate_semantics():
    lit = _lit()
    x = torch.rand(3, 8, 8)
    out = lit.generate(x, num_steps=6)  # (C,H,W) accepted; eval mode; num_steps time points = 5 Euler steps
    assert out.shape == (1, 3, 8, 8) and not lit.training
    want = oflow.generate(lit.net, x, num_steps=6, solver="euler")
    assert torch.allclose(out, want, atol=1e-6)
    lit_none = ConditionalFlowMatchingLitModule(net=TinyField(), flow_matcher=ConditionalFlowMatcher(0.0), solver=None)
    with pytest.raises(ValueError, match="Solver is not initialized"):
        lit_none.generate(x)
    # literal reference behaviour: a functools.partial has no .solver/.atol attributes -> dopri5 @ 1e-4 always
    ref_like = _lit(solver="euler", reference_solver_defaults=True)
    node = ref_like._make_node(ref_like.net)
    assert (node.solver, node.atol, node.rtol, node.sensitivity) == ("dopri5", 1e-4, 1e-4, "adjoint")
    out_d = ref_like.generate(x, num_steps=2)
    want_d = oflow.generate(ref_like.net, x, num_steps=2, solver="dopri5")
    assert torch.allclose(out_d, want_d, atol=1e-5)


def test_class_conditional_module():
    net = TinyField(classes=3)
    lit = ClassConditionalFlowMatchingLitModule(
        net=net, flow_matcher=ConditionalFlowMatcher(0.0),
        solver=functools.partial(NeuralODE, solver="euler"), optimizer=functools.partial(torch.optim.Adam, lr=1e-3))
    x0, x1 = torch.rand(2, 3, 8, 8), torch.rand(2, 3, 8, 8)
    y = torch.tensor([0, 2])
    t = torch.tensor([0.1, 0.6])
    loss = lit.model_step((x0, x1, y.float()), t=t)  # labels arrive as any dtype; `.long()` like the reference
    tt = t.view(-1, 1, 1, 1)
    assert torch.allclose(loss, torch.mean((net(t, tt * x1 + (1 - tt) * x0, y=y) - (x1 - x0)) ** 2))
    a = lit.generate(x0, 2, num_steps=4)
    b = lit.generate(x0, torch.tensor(2), num_steps=4)
    c = lit.generate(x0, torch.tensor([2, 2]), num_steps=4)
    assert torch.allclose(a, b) and torch.allclose(a, c)
    w = ConditionalWrapper(net, y)
    assert torch.allclose(w(t, x0, args=None), net(t, x0, y=y))  # absorbs torchdyn's `args=` keyword


def test_neural_ode_generic_paths_match_oracle():
    f = TinyField().eval()
    x = torch.rand(2, 3, 8, 8)
    for solver in ("euler", "midpoint", "rk4", "dopri5"):
        ts = torch.linspace(0, 1, 4)
        with torch.no_grad():
            got = NeuralODE(f, solver=solver, atol=1e-4, rtol=1e-4).trajectory(x, ts)
            want = oflow.NeuralODE(f, solver=solver, atol=1e-4, rtol=1e-4).trajectory(x, ts)
        assert got.shape == (4, 2, 3, 8, 8)
        assert torch.allclose(got, want, atol=1e-5), solver
    t_eval, sol = NeuralODE(f, solver="euler")(x, torch.linspace(0, 1, 3))
    assert sol.shape[0] == 3 and len(t_eval) == 3
    with pytest.raises(NotImplementedError):
        NeuralODE(f, solver="tsit5").trajectory(x, torch.linspace(0, 1, 3))


def test_flow_matcher_semantics():
    fm = ConditionalFlowMatcher(0.5)
    x0, x1 = torch.zeros(4, 3, 4, 4), torch.ones(4, 3, 4, 4)
    torch.manual_seed(3)
    t, xt, ut, eps = fm.sample_location_and_conditional_flow(x0, x1, return_noise=True)
    torch.manual_seed(3)
    t_o, xt_o, ut_o, eps_o = oflow.ConditionalFlowMatcher(0.5).sample_location_and_conditional_flow(
        x0, x1, return_noise=True)
    assert torch.equal(t, t_o) and torch.equal(xt, xt_o) and torch.equal(ut, ut_o) and torch.equal(eps, eps_o)


def test_conv_plan_bookkeeping():
    p = ConvPlan((Seg(0, 0, 0, 128, 9, 1), Seg(1, 1, 0, 256, 1, 1), Seg(2, 1, 256, 96, 1, 1)), 128)
    assert p.ktot() == 9 * 128 + 256 + 128  # channel counts are padded to 64-wide k-blocks per segment
    q = ConvPlan((Seg(0, 0, 0, 64, 9, 2),), 64)
    assert q.uid != p.uid and q.ktot() == 576


def test_pack_cache_refreshes_every_stale_operand_in_one_launch(monkeypatch):
    """The packed-operand cache: first use packs one operand, an in-place parameter update (optimizer step: same storage,
    new version) makes ALL operands of those parameters stale and the next request re-packs them with ONE multi-tensor
    launch into the SAME buffers; dead parameters drop out.  (A cache that missed the version bump served the initial
    weights for ever -- the bug the entry-point test found.)"""
    import gc

    import torch

    from stain2stain_b200 import kernels as K
    from stain2stain_b200 import ops
    calls = []
    monkeypatch.setattr(K, "pack_conv_weight", lambda w, dst, **kw: calls.append(("one", w.data_ptr())))
    monkeypatch.setattr(K, "pack_conv_weight_multi", lambda jobs, tc: calls.append(("multi", sorted(j[0].data_ptr() for j in jobs))))
    pc = ops._PackCache()
    w1 = torch.nn.Parameter(torch.zeros(8, 4, 3, 3))
    w2 = torch.nn.Parameter(torch.zeros(8, 4, 3, 3))

    def make(dst):
        return (dst if dst is not None else torch.zeros(8, 64, dtype=K.T16)), [(0, 0, 0, 4, False, K.ACT)]
    a = pc.get(("a",), [w1], make)
    b = pc.get(("b",), [w2], make)
    assert pc.get(("a",), [w1], make) is a and len(calls) == 2          # fresh hit: nothing packed
    with torch.no_grad():
        w1.add_(1)
        w2.add_(1)
    assert pc.get(("a",), [w1], make) is a                                # same buffer, refreshed in place ...
    assert calls[-1] == ("multi", sorted([w1.data_ptr(), w2.data_ptr()]))  # ... together with every other stale operand
    assert pc.get(("b",), [w2], make) is b and len(calls) == 3            # already fresh
    torch.autograd.graph.increment_version(w1)                            # what FusedAdam does after its raw-pointer update
    assert pc.get(("a",), [w1], make) is a and calls[-1] == ("multi", [w1.data_ptr()])
    w3 = torch.nn.Parameter(torch.zeros(8, 4, 3, 3))                      # replaced parameter (other storage): new operand
    a3 = pc.get(("a",), [w3], make)
    assert a3 is not a and calls[-1] == ("one", w3.data_ptr())
    del w2, b
    gc.collect()
    with torch.no_grad():
        w3.add_(1)
    pc.get(("a",), [w3], make)
    assert ("b",) not in pc._store                                        # entries of dead parameters are dropped


def test_pack_cache_never_serves_another_tensors_operand(monkeypatch):
    """ADVICE round 1 (ops.py:329): stem / head operands are keyed by a Python id.  A model built after another was freed
    can reuse the same id, the same caching-allocator address and the same construction version count; the hit must
    still be refused because the weak references of the entry do not point at THIS tensor."""
    import torch

    from stain2stain_b200 import kernels as K
    from stain2stain_b200 import ops
    calls = []
    monkeypatch.setattr(K, "pack_conv_weight", lambda w, dst, **kw: calls.append(id(w)))
    pc = ops._PackCache()

    def make(dst):
        return (dst if dst is not None else torch.zeros(8, 64, dtype=K.T16)), [(0, 0, 0, 4, False, K.ACT)]
    w1 = torch.nn.Parameter(torch.zeros(8, 4, 3, 3))
    a = pc.get(("stem", 1234), [w1], make)
    # an impostor with the SAME storage address, version and shape under the same key (what id()/address reuse produces)
    w2 = torch.nn.Parameter(w1.detach())
    assert w2.data_ptr() == w1.data_ptr() and w2._version == w1._version and w2 is not w1
    b = pc.get(("stem", 1234), [w2], make)
    assert b is not a and len(calls) == 2, "the cache served an operand packed from a different tensor object"
    assert pc.get(("stem", 1234), [w2], make) is b and len(calls) == 2


def test_euler_graph_registry_is_weak():
    import gc

    import torch

    from stain2stain_b200 import neural_ode as node
    net = torch.nn.Linear(2, 2)
    node._GRAPHS.setdefault(net, {})["k"] = object()
    assert len(node._GRAPHS) >= 1
    n0 = len(node._GRAPHS)
    del net
    gc.collect()
    assert len(node._GRAPHS) == n0 - 1
    node.clear_graphs()
    assert len(node._GRAPHS) == 0


def _write(path, text):
    import os
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(text)


def test_hydra_lite_composes_defaults_lists_experiments_and_overrides(tmp_path):
    """The defaults-list composition the reference's `python src/train.py experiment=...` relies on (configs/train.yaml:5-31,
    configs/experiment/gray_matter/simple_flow_matching.yaml:4-8), on a miniature config tree."""
    from stain2stain_b200 import hydra_lite
    d = str(tmp_path)
    _write(f"{d}/train.yaml", "defaults:\n  - _self_\n  - data: mnist\n  - model: mnist\n  - trainer: gpu\n  - experiment: null\n"
                              "  - optional local: default\n  - debug: null\nseed: null\ntags: [dev]\ntask_name: train\n")
    _write(f"{d}/data/mnist.yaml", "batch_size: 128\nname: mnist\n")
    _write(f"{d}/data/paired.yaml", "batch_size: 6\ndata_dir: ${paths.data_dir}\nname: paired\n")
    _write(f"{d}/model/mnist.yaml", "_target_: collections.OrderedDict\nlr: 0.1\n")
    _write(f"{d}/model/cfm.yaml", "_target_: collections.OrderedDict\noptimizer:\n  _target_: collections.OrderedDict\n"
                                  "  _partial_: true\n  lr: 0.001\n  weight_decay: 0.0\nnet:\n  dim: [3, 256, 256]\n")
    _write(f"{d}/trainer/default.yaml", "max_epochs: 10\naccelerator: cpu\ndevices: 1\n")
    _write(f"{d}/trainer/gpu.yaml", "defaults:\n  - default\naccelerator: gpu\n")
    _write(f"{d}/trainer/ddp.yaml", "defaults:\n  - default\nstrategy: ddp\naccelerator: gpu\ndevices: 4\nsync_batchnorm: true\n")
    _write(f"{d}/experiment/gm/simple.yaml", "# @package _global_\n\ndefaults:\n  - override /data: paired\n  - override /model: cfm\n"
                                             "  - override /trainer: default.yaml\nseed: 1984\nbatch_size: 32\ntrainer:\n  max_epochs: 200\n"
                                             "  devices: 4\nmodel:\n  optimizer:\n    lr: 1e-4\n    weight_decay: 1e-5\ndata:\n  batch_size: ${batch_size}\n"
                                             "tags: [simple]\n")
    base = hydra_lite.compose(d, "train")
    assert base["data"]["name"] == "mnist" and base["trainer"] == {"max_epochs": 10, "accelerator": "gpu", "devices": 1}
    assert base["seed"] is None and "experiment" not in base and "local" not in base
    cfg = hydra_lite.compose(d, "train", ["experiment=gm/simple", "model.optimizer.lr=0.002", "trainer.devices=8"])
    assert cfg["data"]["name"] == "paired" and cfg["model"]["net"]["dim"] == [3, 256, 256]       # groups re-selected
    assert cfg["trainer"]["accelerator"] == "cpu" and cfg["trainer"]["max_epochs"] == 200        # experiment body on top
    assert cfg["trainer"]["devices"] == 8 and cfg["model"]["optimizer"]["lr"] == 0.002           # value overrides last
    assert cfg["seed"] == 1984 and cfg["tags"] == ["simple"]                                     # lists are replaced
    res = hydra_lite.resolve(cfg)
    assert res["data"]["batch_size"] == 32                                                       # ${batch_size}
    assert res["data"]["data_dir"] == "${paths.data_dir}"                                        # group not composed: left alone
    assert res["model"]["optimizer"]["weight_decay"] == "1e-5" or res["model"]["optimizer"]["weight_decay"] == 1e-5
    cfg2 = hydra_lite.compose(d, "train", ["experiment=gm/simple", "trainer=ddp"])               # command line beats the experiment
    assert cfg2["trainer"]["strategy"] == "ddp" and cfg2["trainer"]["max_epochs"] == 200 and cfg2["trainer"]["devices"] == 4
    model = hydra_lite.instantiate(res["model"])
    assert model["optimizer"].keywords["weight_decay"] == 1e-5 and model["optimizer"].keywords["lr"] == 0.002


def test_hydra_lite_composes_the_reference_experiment_tree():
    """`experiment=gray_matter/simple_flow_matching` over the reference's UNMODIFIED configs/ (build container only)."""
    import os

    import pytest
    from stain2stain_b200 import hydra_lite
    cdir = "/root/reference/configs"
    if not os.path.isdir(cdir):
        pytest.skip("the reference tree only exists in the build container")
    cfg = hydra_lite.resolve(hydra_lite.compose(cdir, "train", ["experiment=gray_matter/simple_flow_matching", "trainer=ddp"]))
    m = cfg["model"]
    assert m["_target_"] == "src.models.conditional_flow_matching.ConditionalFlowMatchingLitModule"
    assert m["net"]["_target_"] == "torchcfm.models.unet.UNetModel" and m["net"]["channel_mult"] == [1, 2, 2, 4]
    assert float(m["optimizer"]["weight_decay"]) == 1e-5 and float(m["optimizer"]["lr"]) == 1e-4   # the experiment's overrides
    assert cfg["seed"] == 1984 and cfg["data"]["batch_size"] == 32 and cfg["data"]["image_size"] == 256
    assert cfg["trainer"]["strategy"] == "ddp" and cfg["trainer"]["max_epochs"] == 200
    # the composed model node builds the B200 drop-ins (meta device: no GPU needed for construction)
    import torch
    with torch.device("meta"):
        lit = hydra_lite.instantiate(m, remap=True, fused_optimizer=True)
    from stain2stain_b200.lit import ConditionalFlowMatchingLitModule
    from stain2stain_b200.unet import UNetModel
    assert isinstance(lit, ConditionalFlowMatchingLitModule) and isinstance(lit.net, UNetModel)
    assert sum(p.numel() for p in lit.net.parameters()) == 70_954_883


def test_invalidate_caches_clears_operands_and_graphs():
    import stain2stain_b200
    from stain2stain_b200 import neural_ode, ops
    ops.PACK_CACHE._store[("x",)] = [None, None, [], []]
    stain2stain_b200.invalidate_caches()
    assert not ops.PACK_CACHE._store and len(neural_ode._GRAPHS) == 0


def test_conditional_wrapper_label_broadcast_and_truncation():
    """reference class_conditional_flow_matching.py:168-173"""
    import torch
    from stain2stain_b200.lit import ConditionalWrapper
    seen = {}

    class Net(torch.nn.Module):
        def forward(self, t, x, y=None):
            seen["y"] = y
            return x
    w = ConditionalWrapper(Net(), torch.tensor(2))
    w(torch.zeros(()), torch.zeros(3, 1))
    assert seen["y"].tolist() == [2, 2, 2]
    w = ConditionalWrapper(Net(), torch.tensor([0, 1, 2, 1]))
    w(torch.zeros(()), torch.zeros(2, 1))
    assert seen["y"].tolist() == [0, 1]
