"""The slice of Hydra / OmegaConf the reference's entry points use on this path, for images without hydra-core:
`OmegaConf.load(yaml)` + `hydra.utils.instantiate(cfg)` (src/train.py:59-71, src/infer_simple_flowmatching.py:48-49).

Supported: nested `_target_` instantiation (recursive, children first), `_partial_: true` -> functools.partial,
`${key}` / `${a.b}` interpolation against the root config, keyword overrides, and `compose()`: Hydra's defaults-list
composition as the reference's configs use it (configs/train.yaml:5-31 + `experiment=<group file>` with
`# @package _global_` and `override /group: name` entries, e.g. configs/experiment/gray_matter/simple_flow_matching.yaml:4-8)
plus `group=name` / `a.b.c=value` command-line overrides.  `remap=True` additionally swaps the
reference's `_target_` strings for this package's drop-ins, so the reference's OWN yaml files (unmodified) build the
B200 modules: that is the config-level boundary of SURVEY.md 8(b).
"""
from __future__ import annotations

import copy
import functools
import importlib
import os
import re
from typing import Any, Dict, List, Optional, Sequence

import yaml

TARGET_MAP = {
    "torchcfm.models.unet.UNetModel": "stain2stain_b200.unet.UNetModel",
    "torchcfm.models.unet.unet.UNetModel": "stain2stain_b200.unet.RawUNetModel",
    "torchcfm.conditional_flow_matching.ConditionalFlowMatcher": "stain2stain_b200.flow_matching.ConditionalFlowMatcher",
    "torchdyn.core.NeuralODE": "stain2stain_b200.neural_ode.NeuralODE",
    "src.models.conditional_flow_matching.ConditionalFlowMatchingLitModule":
        "stain2stain_b200.lit.ConditionalFlowMatchingLitModule",
    "src.models.class_conditional_flow_matching.ClassConditionalFlowMatchingLitModule":
        "stain2stain_b200.lit.ClassConditionalFlowMatchingLitModule",
    "src.models.conditional_flow_matching_multitask_multiclassloss.MultiTaskFlowMatchingLitModule":
        "stain2stain_b200.multitask.MultiTaskFlowMatchingLitModule",
    "src.models.conditional_flow_matching_masked.ConditionalFlowMatchingLitModule":
        "stain2stain_b200.lit_masked.MaskWeightedFlowMatchingLitModule",
    "src.models.conditional_flow_matching_ROI_loss.ConditionalFlowMatchingLitModule":
        "stain2stain_b200.lit_masked.ROILossFlowMatchingLitModule",
    "src.models.conditional_flow_matching_conditional_mask.ConditionalFlowMatchingLitModule":
        "stain2stain_b200.lit_masked.MaskConditionedFlowMatchingLitModule",
    "src.models.conditional_flow_matching_conditional_toggle_mask.ConditionalFlowMatchingLitModule":
        "stain2stain_b200.lit_masked.MaskToggleFlowMatchingLitModule",
    "src.models.components.shared_encoder.SharedEncoder": "stain2stain_b200.multitask.SharedEncoder",
    "src.models.components.task_decoders.FlowMatchingDecoder": "stain2stain_b200.multitask.FlowMatchingDecoder",
    "src.models.components.task_decoders.SegmentationDecoder": "stain2stain_b200.multitask.SegmentationDecoder",
}
FUSED_OPTIMIZER_MAP = {"torch.optim.Adam": "stain2stain_b200.optim.FusedAdam"}

_INTERP = re.compile(r"\$\{([^}]+)\}")


def load_yaml(path: str) -> Dict[str, Any]:
    with open(path) as f:
        return yaml.safe_load(f)


def _lookup(root, dotted: str):
    cur = root
    for part in dotted.split("."):
        cur = cur[int(part)] if isinstance(cur, list) else cur[part]
    return cur


def _resolve(node, root):
    if isinstance(node, dict):
        return {k: _resolve(v, root) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root) for v in node]
    if isinstance(node, str):
        def value(ref: str, whole: str):
            if ":" in ref:  # resolver syntax (${hydra:...}, ${oc.env:...}): Hydra's runtime, not part of this slice
                if ref.startswith("oc.env:"):
                    name, _, default = ref[len("oc.env:"):].partition(",")
                    return os.environ.get(name.strip(), default.strip() or whole)
                return whole
            try:
                return _resolve(_lookup(root, ref), root)
            except (KeyError, IndexError, TypeError):
                return whole  # leave what cannot be resolved here (keys of config groups that were not composed)
        m = _INTERP.fullmatch(node)
        if m:
            return value(m.group(1), node)
        return _INTERP.sub(lambda mm: str(value(mm.group(1), mm.group(0))), node)
    return node


def _locate(target: str):
    module, _, name = target.rpartition(".")
    return getattr(importlib.import_module(module), name)


def _coerce(v):
    # YAML 1.1 reads `1e-4` as a string; Hydra/OmegaConf read it as a float.
    if isinstance(v, str):
        try:
            return float(v) if re.fullmatch(r"[+-]?\d+(\.\d*)?[eE][+-]?\d+", v) else v
        except ValueError:
            return v
    return v


def _build(node, remap: bool, fused_optimizer: bool):
    if isinstance(node, list):
        return [_build(v, remap, fused_optimizer) for v in node]
    if not isinstance(node, dict):
        return _coerce(node)
    if "_target_" not in node:
        return {k: _build(v, remap, fused_optimizer) for k, v in node.items()}
    target = node["_target_"]
    if remap:
        target = TARGET_MAP.get(target, target)
    if fused_optimizer:
        target = FUSED_OPTIMIZER_MAP.get(target, target)
    partial = bool(node.get("_partial_", False))
    kwargs = {k: _build(v, remap, fused_optimizer) for k, v in node.items() if not (k.startswith("_") and k.endswith("_"))}
    fn = _locate(target)
    return functools.partial(fn, **kwargs) if partial else fn(**kwargs)


def instantiate(cfg: Dict[str, Any], remap: bool = False, fused_optimizer: bool = False, **overrides):
    """`hydra.utils.instantiate(cfg, **overrides)` for the subset described in the module docstring."""
    cfg = dict(cfg)
    cfg.update(overrides)
    return _build(_resolve(cfg, cfg), remap, fused_optimizer)


# ------------------------------------------------------------------------------------------------ defaults-list composition
def _merge(dst: Dict[str, Any], src: Dict[str, Any]) -> Dict[str, Any]:
    """OmegaConf.merge for plain dicts: dicts merge key by key, everything else (lists included) is replaced."""
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


def _read_group_file(config_dir: str, group: str, name: str):
    """-> (body without `defaults`, its defaults list, is `# @package _global_`)."""
    name = name[:-5] if name.endswith(".yaml") else name
    path = os.path.join(config_dir, group, name + ".yaml") if group else os.path.join(config_dir, name + ".yaml")
    with open(path) as f:
        text = f.read()
    body = yaml.safe_load(text) or {}
    is_global = bool(re.search(r"^#\s*@package\s+_global_", text, flags=re.M))
    defaults = body.pop("defaults", []) or []
    return body, defaults, is_global


def _parse_default(entry):
    """One defaults-list entry -> (group, name, is_override, is_optional) or ('_self_', ...)."""
    if entry == "_self_":
        return "_self_", None, False, False
    if isinstance(entry, str):  # a file of the same group, e.g. `- default`
        return None, entry, False, False
    (key, name), = entry.items()
    toks = key.split()
    group = toks[-1].lstrip("/")
    return group, name, "override" in toks, "optional" in toks


def _set_dotted(cfg: Dict[str, Any], dotted: str, value):
    cur = cfg
    parts = dotted.split(".")
    for part in parts[:-1]:
        cur = cur.setdefault(part, {})
    cur[parts[-1]] = value


def compose(config_dir: str, config_name: str = "train", overrides: Sequence[str] = (), skip_groups=("hydra",)) -> Dict[str, Any]:
    """`hydra.compose(config_name, overrides)` for the reference's layout: the primary config's defaults list selects one
    file per config group; `experiment=<file>` (a `# @package _global_` file listed last) may re-select groups with
    `override /group: name` and then overwrites individual keys; `group=name` overrides re-select groups from the command
    line and `a.b.c=value` overrides set values (YAML-typed) after composition.  Returns the merged, UNRESOLVED config
    (interpolations are resolved by `instantiate` / `resolve`)."""
    body, defaults, _ = _read_group_file(config_dir, "", config_name)
    order: List[str] = []                        # composition order: '_self_' and group names
    choice: Dict[str, Optional[str]] = {}
    optional: Dict[str, bool] = {}
    for e in defaults:
        g, n, _, opt = _parse_default(e)
        if g == "_self_":
            order.append("_self_")
            continue
        order.append(g)
        choice[g], optional[g] = n, opt
    if "_self_" not in order:
        order.append("_self_")  # Hydra >= 1.1: the primary config is merged last unless it says otherwise
    values: List[str] = []
    for ov in overrides:
        key, _, val = ov.partition("=")
        key = key.lstrip("+")
        if "." not in key and (key in choice or os.path.isdir(os.path.join(config_dir, key))):
            if key not in choice:
                order.append(key)
            choice[key] = None if val in ("null", "") else val
        else:
            values.append(ov.lstrip("+"))
    # `override /group: name` entries of the selected global files (experiment, debug, ...) re-select groups first
    for g in list(order):
        if g == "_self_" or not choice.get(g) or g in skip_groups:
            continue
        path = os.path.join(config_dir, g, str(choice[g]).replace(".yaml", "") + ".yaml")
        if not os.path.exists(path):
            continue
        _, sub_defaults, is_global = _read_group_file(config_dir, g, choice[g])
        if is_global:
            for e in sub_defaults:
                sg, sn, is_override, _ = _parse_default(e)
                if is_override and sg:
                    if sg not in choice:
                        order.insert(order.index(g), sg)
                    choice[sg] = sn
    # command-line group selections win over the experiment's
    for ov in overrides:
        key, _, val = ov.partition("=")
        if "." not in key and key in choice and key.lstrip("+") == key:
            choice[key] = None if val in ("null", "") else val
    cfg: Dict[str, Any] = {}

    def load_group(group: str, name: str) -> Dict[str, Any]:
        sub, sub_defaults, is_global = _read_group_file(config_dir, group, name)
        merged: Dict[str, Any] = {}
        self_done = False
        for e in sub_defaults:
            sg, sn, is_override, _ = _parse_default(e)
            if sg == "_self_":
                _merge(merged, sub)
                self_done = True
            elif sg is None:              # sibling file of the same group
                _merge(merged, load_group(group, sn)[0])
            elif not is_override and not is_global:
                _merge(merged.setdefault(sg, {}), load_group(sg, sn)[0])
        if not self_done:
            _merge(merged, sub)
        return merged, is_global

    for g in order:
        if g == "_self_":
            _merge(cfg, body)
            continue
        name = choice.get(g)
        if name is None or g in skip_groups:
            continue
        path = os.path.join(config_dir, g, str(name).replace(".yaml", "") + ".yaml")
        if not os.path.exists(path):
            if optional.get(g):
                continue
            raise FileNotFoundError(f"config group '{g}' has no option '{name}' ({path})")
        merged, is_global = load_group(g, name)
        if is_global:
            _merge(cfg, merged)
        else:
            _merge(cfg.setdefault(g, {}), merged)
    for ov in values:
        key, _, val = ov.partition("=")
        _set_dotted(cfg, key, _coerce(yaml.safe_load(val)))
    return cfg


def resolve(cfg: Dict[str, Any]) -> Dict[str, Any]:
    """OmegaConf.to_container(cfg, resolve=True) for what can be resolved here."""
    return _resolve(cfg, cfg)
