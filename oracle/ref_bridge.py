"""ORACLE (test infrastructure only) -- runs the REFERENCE'S OWN LitModule code in this container.

/root/reference/src/models/*.py import `lightning`, `torchcfm` and `torchdyn`, none of which is installed or
installable here (SURVEY.md finding 3).  This bridge injects minimal stand-ins for exactly those three imports into
`sys.modules` -- `lightning.LightningModule` (an nn.Module with save_hyperparameters/log), and the oracle's
restatements for `torchcfm.conditional_flow_matching.ConditionalFlowMatcher` / `torchdyn.core.NeuralODE` -- and then
imports the reference's modules unmodified from /root/reference.  Everything the reference itself implements
(model_step, compute_segmentation_loss, MulticlassDiceLoss, generate + FlowWrapper / ConditionalWrapper, the in-repo
encoder / decoders) therefore executes for real; only the two absent third-party packages are restated.

Used by oracle/make_golden.py (fixture generation) and by tests that skip when /root/reference is absent (GPU box).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("S2S_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models"))


class _HParams(dict):
    __getattr__ = dict.get


class _LightningModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.hparams = _HParams()
        self.logged = {}
        self.current_epoch = 0
        self.trainer = None

    def save_hyperparameters(self, *a, **k):
        return None

    def log(self, name, value, **kw):
        self.logged[name] = value.detach() if torch.is_tensor(value) else value

    @property
    def device(self):
        return next(self.parameters()).device


class _LightningDataModule:
    def __init__(self):
        self.hparams = _HParams()
        self.trainer = None

    def save_hyperparameters(self, *a, logger=True, **k):
        import inspect
        frame = inspect.currentframe().f_back
        self.hparams.update({k2: v for k2, v in frame.f_locals.items() if k2 not in ("self", "__class__")})


def install_stubs():
    from . import flow as oflow
    from . import unet as ounet
    if "lightning" not in sys.modules:
        m = types.ModuleType("lightning")
        m.LightningModule = _LightningModule
        m.LightningDataModule = _LightningDataModule
        sys.modules["lightning"] = m
    if "torchcfm" not in sys.modules:
        pkg = types.ModuleType("torchcfm")
        cfm = types.ModuleType("torchcfm.conditional_flow_matching")
        cfm.ConditionalFlowMatcher = oflow.ConditionalFlowMatcher
        models = types.ModuleType("torchcfm.models")
        unet_pkg = types.ModuleType("torchcfm.models.unet")
        unet_pkg.UNetModel = ounet.UNetModel
        unet_mod = types.ModuleType("torchcfm.models.unet.unet")
        unet_mod.UNetModel = ounet.RawUNetModel
        pkg.conditional_flow_matching, pkg.models = cfm, models
        models.unet, unet_pkg.unet = unet_pkg, unet_mod
        sys.modules.update({"torchcfm": pkg, "torchcfm.conditional_flow_matching": cfm, "torchcfm.models": models,
                            "torchcfm.models.unet": unet_pkg, "torchcfm.models.unet.unet": unet_mod})
    if "torchdyn" not in sys.modules:
        pkg = types.ModuleType("torchdyn")
        core = types.ModuleType("torchdyn.core")
        core.NeuralODE = oflow.NeuralODE
        pkg.core = core
        sys.modules.update({"torchdyn": pkg, "torchdyn.core": core})
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:  # noqa: BLE001
            sys.modules["wandb"] = types.ModuleType("wandb")


def reference_module(name: str):
    """Import `src.models.<name>` (or `src.models.components.<name>`) from the reference tree, unmodified."""
    if not available():
        raise RuntimeError(f"{REFERENCE_ROOT} is not present (the reference only exists in the build container)")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module(name)
