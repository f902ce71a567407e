"""Entry points mirroring the reference's scripts for images without hydra / lightning:

  train(...)         <- src/train.py:43-134    (seed -> instantiate model -> fit loop -> checkpoint)
  infer_simple(...)  <- src/infer_simple_flowmatching.py:40-127 (yaml -> instantiate -> load ckpt['state_dict'] ->
                        generate -> denormalize), batched instead of one tile per call

With real Lightning/Hydra installed the reference's own scripts drive the same modules; these loops
do what `Trainer.fit` does on this path -- configure_optimizers, training_step, backward, optimizer step, a Lightning-
style checkpoint dict -- on synthetic tile pairs (the data modules are host I/O and out of scope).

    python -m stain2stain_b200.entry train --config configs/model/conditional_flow_matching.yaml --steps 20
    python -m stain2stain_b200.entry infer --config ... --ckpt last.ckpt --tiles 8 --num-steps 51
"""
from __future__ import annotations

import argparse
import json
import time
from typing import Any, Dict, Optional

import torch

from . import hydra_lite


def denormalize(t: torch.Tensor) -> torch.Tensor:
    """[-1, 1] -> [0, 1] (src/infer_simple_flowmatching.py:37-38)."""
    return (t * 0.5 + 0.5).clamp(0, 1)


def to_uint8_hwc(t: torch.Tensor) -> torch.Tensor:
    """Normalised NCHW tiles -> uint8 NHWC (what the reference hands to matplotlib / W&B).  Device tensors go through
    the fused `s2s_denorm_u8` kernel (data.denormalize_to_uint8); host tensors through the same formula in torch."""
    if t.is_cuda:
        from .data import denormalize_to_uint8
        return denormalize_to_uint8(t)
    return (denormalize(t) * 255.0 + 0.5).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def synthetic_batch(kind: str, batch: int, size: int, device, generator, num_classes: int = 3):
    """Tile pairs in the datamodule's range after Normalize(0.5, 0.5) (src/data/paired_data_module.py:145,198-199)."""
    x0 = torch.rand(batch, 3, size, size, device=device, generator=generator) * 2 - 1
    x1 = torch.rand(batch, 3, size, size, device=device, generator=generator) * 2 - 1
    if kind == "class_conditional":
        return x0, x1, torch.randint(0, num_classes, (batch,), device=device, generator=generator)
    if kind == "multitask":
        return x0, x1, torch.randint(0, num_classes, (batch, 1, size, size), device=device, generator=generator).float()
    return x0, x1


def model_kind(model) -> str:
    name = type(model).__name__
    return "multitask" if "MultiTask" in name else ("class_conditional" if "ClassConditional" in name else "simple")


def build_model(config, remap: bool = True, fused_optimizer: bool = True, **overrides):
    cfg = hydra_lite.load_yaml(config) if isinstance(config, str) else config
    return hydra_lite.instantiate(cfg, remap=remap, fused_optimizer=fused_optimizer, **overrides)


def compose_experiment(config_dir: str, overrides=()) -> Dict[str, Any]:
    """What `@hydra.main(config_path="../configs", config_name="train.yaml")` hands to `train(cfg)` (src/train.py:137-154):
    the primary config composed with `experiment=...` and the other command-line overrides."""
    return hydra_lite.compose(config_dir, "train", list(overrides))


def train_experiment(config_dir: str, overrides=(), steps: int = 10, batch: Optional[int] = None, device: str = "cuda",
                     **kw) -> Dict[str, Any]:
    """`python src/train.py experiment=gray_matter/simple_flow_matching ...` on synthetic tiles: composes the reference's
    config tree (unmodified), seeds from `cfg.seed`, instantiates `cfg.model` with the B200 drop-ins and runs `steps`
    training steps at the per-device batch `cfg.data.batch_size // world` (src/data/paired_data_module.py:273-278)."""
    cfg = hydra_lite.resolve(compose_experiment(config_dir, overrides))
    if batch is None:
        from .parallel import per_rank_batch
        gb = (cfg.get("data") or {}).get("batch_size", 4)
        batch = per_rank_batch(int(gb), int(kw.pop("world_size", 1)))
    seed = cfg.get("seed")
    kw.setdefault("sync_batchnorm", bool((cfg.get("trainer") or {}).get("sync_batchnorm", False)))
    out = train(cfg["model"], steps=steps, batch=batch, device=device, seed=1984 if seed is None else int(seed), **kw)
    out["cfg"] = cfg
    return out


def train(config, steps: int = 10, batch: int = 4, size: Optional[int] = None, device: str = "cuda", seed: int = 1984,
          ckpt_path: Optional[str] = None, sync_batchnorm: bool = False, **overrides) -> Dict[str, Any]:
    torch.manual_seed(seed)  # L.seed_everything(cfg.seed) (src/train.py:55-56)
    model = build_model(config, **overrides).to(device)
    if sync_batchnorm and torch.distributed.is_available() and torch.distributed.is_initialized() \
            and torch.distributed.get_world_size() > 1:
        # Trainer(sync_batchnorm=True) (configs/trainer/ddp.yaml:9): BatchNorm2d -> SyncBatchNorm; ops.batch_norm_relu then
        # folds the batch statistics over all ranks.  A no-op for the GroupNorm UNets and in a single process.
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    kind = model_kind(model)
    if size is None:
        size = model.net.image_size if hasattr(model, "net") else 256
    nc = getattr(model, "num_classes", None) or getattr(getattr(model, "net", None), "num_classes", None) or 3
    opt_cfg = model.configure_optimizers()
    optimizer = opt_cfg["optimizer"]
    model.train()
    g = torch.Generator(device=device).manual_seed(seed)
    losses = []
    t0 = time.perf_counter()
    for step in range(steps):
        batch_t = synthetic_batch(kind, batch, size, device, g, nc)
        optimizer.zero_grad(set_to_none=True)
        loss = model.training_step(batch_t, step)
        loss.backward()
        optimizer.step()
        losses.append(loss.detach())
    losses = [float(v) for v in torch.stack(losses).cpu()]
    dt = time.perf_counter() - t0
    out = {"kind": kind, "steps": steps, "loss_first": losses[0], "loss_last": losses[-1], "losses": losses,
           "tiles_per_s": steps * batch / dt, "logged": sorted(getattr(model, "logged", {}))}
    if ckpt_path:
        torch.save({"state_dict": model.state_dict(), "optimizer_states": [optimizer.state_dict()], "epoch": 0,
                    "global_step": steps}, ckpt_path)
        out["ckpt_path"] = ckpt_path
    out["model"] = model
    return out


@torch.no_grad()
def infer_simple(config, ckpt_path: Optional[str], source: torch.Tensor, num_steps: int = 2, device: str = "cuda",
                 target_class=None, **overrides):
    """-> (generated tiles in [-1, 1], uint8 NHWC images); multitask models also return the predicted mask."""
    model = build_model(config, **overrides).to(device)
    if ckpt_path:
        checkpoint = torch.load(ckpt_path, map_location=device, weights_only=False)
        model.load_state_dict(checkpoint["state_dict"])
    model.eval()
    kind = model_kind(model)
    src = source.to(device)
    if kind == "class_conditional":
        gen = model.generate(src, 0 if target_class is None else target_class, num_steps=num_steps)
    elif kind == "multitask":
        gen, mask = model.generate(src, num_steps=num_steps)
        return gen, to_uint8_hwc(gen), mask
    else:
        gen = model.generate(src, num_steps=num_steps)
    return gen, to_uint8_hwc(gen)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["train", "infer"])
    ap.add_argument("--config", required=True)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--ckpt", default=None)
    ap.add_argument("--tiles", type=int, default=4)
    ap.add_argument("--num-steps", type=int, default=51)
    a = ap.parse_args()
    if a.cmd == "train":
        out = train(a.config, a.steps, a.batch, a.size, ckpt_path=a.ckpt)
        out.pop("model")
        print(json.dumps(out))
    else:
        size = a.size or 256
        src = torch.rand(a.tiles, 3, size, size) * 2 - 1
        res = infer_simple(a.config, a.ckpt, src, a.num_steps)
        print(json.dumps({"generated": list(res[0].shape), "uint8": list(res[1].shape)}))


if __name__ == "__main__":
    main()
