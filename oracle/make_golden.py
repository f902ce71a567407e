#!/usr/bin/env python
"""Generates tests/golden/*.pt by running the REFERENCE'S OWN code (through oracle/ref_bridge.py) in the build
container:  python -m oracle.make_golden

Every fixture holds seeds + small inputs + the reference's outputs; weights are NOT stored -- they are re-drawn from
the recorded seed by a construction that consumes the RNG exactly like the reference's (per-parameter checksums are
stored so a drift is caught).  What executes for real: src/models/conditional_flow_matching.py,
src/models/class_conditional_flow_matching.py, src/models/conditional_flow_matching_multitask_multiclassloss.py,
src/models/components/shared_encoder.py, src/models/components/task_decoders.py.  What is restated (absent third-party
packages): the torchcfm UNet / matcher and the torchdyn solver (oracle/unet.py, oracle/flow.py).
"""
from __future__ import annotations

import functools
import os

import torch

from . import flow as oflow
from . import ref_bridge as rb
from . import unet as ounet

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SMALL = dict(dim=[3, 64, 64], num_channels=64, num_res_blocks=1, attention_resolutions="16,8", dropout=0.0,
             use_scale_shift_norm=True, num_heads=4, num_head_channels=32, channel_mult=[1, 2, 2, 4])
MT = dict(features=[64, 128, 256], num_classes=5, time_emb_dim=64, size=32, batch=2)


def checksums(module):
    return {k: float(v.double().sum()) for k, v in module.state_dict().items() if v.dtype.is_floating_point}


def inputs(seed, B, H, classes=0, mask_classes=0):
    g = torch.Generator().manual_seed(seed)
    d = dict(x0=torch.rand(B, 3, H, H, generator=g) * 2 - 1, x1=torch.rand(B, 3, H, H, generator=g) * 2 - 1,
             t=torch.rand(B, generator=g))
    if classes:
        d["y"] = torch.randint(0, classes, (B,), generator=g)
    if mask_classes:
        d["mask"] = torch.randint(0, mask_classes, (B, 1, H, H), generator=g).float()
    return d


def simple_fm(class_cond: bool):
    name = "class_conditional_flow_matching" if class_cond else "conditional_flow_matching"
    mod = rb.reference_module(f"src.models.{name}")
    cls = mod.ClassConditionalFlowMatchingLitModule if class_cond else mod.ConditionalFlowMatchingLitModule
    cfg = dict(SMALL, class_cond=True, num_classes=3) if class_cond else dict(SMALL)
    torch.manual_seed(0)
    net = ounet.dezero_(ounet.UNetModel(**cfg), seed=1984)
    kw = {} if class_cond else dict(log_images=False)
    lit = cls(net=net, flow_matcher=oflow.ConditionalFlowMatcher(0.0),
              solver=functools.partial(oflow.NeuralODE, solver="dopri5", sensitivity="adjoint", atol=1e-4, rtol=1e-4),
              optimizer=functools.partial(torch.optim.Adam, lr=1e-4), scheduler=None, **kw)
    lit.eval()
    inp = inputs(21, 2, 64, classes=3 if class_cond else 0)
    x0, x1, t = inp["x0"], inp["x1"], inp["t"]
    out = dict(config=cfg, weight_seed=0, dezero_seed=1984, input_seed=21, checksums=checksums(lit))
    with torch.no_grad():
        out["forward"] = lit(t, x0, inp["y"]) if class_cond else lit(t, x0)
    # model_step draws t from the CPU default generator (torchcfm semantics): seed it and record the draw
    torch.manual_seed(77)
    t_drawn = torch.rand(2)
    torch.manual_seed(77)
    batch = (x0, x1, inp["y"]) if class_cond else (x0, x1)
    loss = lit.model_step(batch)
    lit.zero_grad()
    loss.backward()
    out["model_step"] = dict(rng_seed=77, t=t_drawn, loss=loss.detach(),
                             grad_norms={k: float(p.grad.double().norm()) for k, p in lit.named_parameters()})
    # generate(): the reference always ends up with dopri5 @ 1e-4 (functools.partial has no .solver attribute)
    gen = lit.generate(x0[:1], 1, num_steps=2) if class_cond else lit.generate(x0[:1], num_steps=2)
    out["generate_num_steps2"] = gen
    torch.save(out, os.path.join(OUT, ("class_cond" if class_cond else "simple_fm") + "_small.pt"))
    return out


def multitask():
    m = rb.reference_module("src.models.conditional_flow_matching_multitask_multiclassloss")
    se = rb.reference_module("src.models.components.shared_encoder")
    td = rb.reference_module("src.models.components.task_decoders")
    f = MT["features"]
    torch.manual_seed(5)
    enc = se.SharedEncoder(3, f, True)
    fd = td.FlowMatchingDecoder(f[-1], f[:-1][::-1], 3, MT["time_emb_dim"], True)
    sd = td.SegmentationDecoder(f[-1], f[:-1][::-1], MT["num_classes"], True)
    lit = m.MultiTaskFlowMatchingLitModule(
        enc, fd, sd, oflow.ConditionalFlowMatcher(0.0), num_classes=MT["num_classes"],
        solver=functools.partial(oflow.NeuralODE, solver="dopri5", sensitivity="adjoint", atol=1e-4, rtol=1e-4),
        optimizer=functools.partial(torch.optim.Adam, lr=1e-4, weight_decay=1e-5), scheduler=None,
        time_emb_dim=MT["time_emb_dim"], log_images=False)
    inp = inputs(31, MT["batch"], MT["size"], mask_classes=MT["num_classes"])
    x0, x1, t, mask = inp["x0"], inp["x1"], inp["t"], inp["mask"]
    out = dict(config=MT, weight_seed=5, input_seed=31, checksums=checksums(lit))
    lit.train()  # BatchNorm in batch-statistics mode, as in training_step
    torch.manual_seed(78)
    t_drawn = torch.rand(MT["batch"])
    torch.manual_seed(78)
    total, d = lit.model_step((x0, x1, mask))
    lit.zero_grad()
    total.backward()
    out["model_step_train"] = dict(rng_seed=78, t=t_drawn, losses={k: v.detach() for k, v in d.items()},
                                   grad_norms={k: float(p.grad.double().norm()) for k, p in lit.named_parameters()},
                                   running_mean_inc=lit.encoder.inc.double_conv[1].running_mean.clone(),
                                   running_var_inc=lit.encoder.inc.double_conv[1].running_var.clone(),
                                   num_batches_tracked=int(lit.encoder.inc.double_conv[1].num_batches_tracked))
    lit.eval()  # running statistics (after exactly one training step = two encoder passes)
    with torch.no_grad():
        out["forward_flow_eval"] = lit.forward_flow(t, x0)
        out["forward_segmentation_eval"] = lit.forward_segmentation(x0)
    img, pm = lit.generate(x0, num_steps=3)
    out["generate_num_steps3"] = dict(image=img, mask=pm)
    torch.save(out, os.path.join(OUT, "multitask_small.pt"))
    return out


RAW4 = dict(image_size=64, in_channels=4, model_channels=64, out_channels=3, num_res_blocks=1, attention_resolutions=[8, 4],
            dropout=0.0, channel_mult=[1, 2, 2, 4], use_scale_shift_norm=True, num_heads=4, num_head_channels=32)
MASK_VARIANTS = {  # name -> (reference module, 4-channel raw UNet?)
    "mask_weighted": ("conditional_flow_matching_masked", False),
    "roi_loss": ("conditional_flow_matching_ROI_loss", False),
    "mask_conditioned": ("conditional_flow_matching_conditional_mask", True),
    "mask_toggle": ("conditional_flow_matching_conditional_toggle_mask", True),
}


def mask_variants():
    """The four mask / ROI LitModules of the reference, each executed unmodified."""
    out = dict(configs=dict(simple=SMALL, raw4=RAW4), weight_seed=0, dezero_seed=1984, input_seed=41)
    inp = inputs(41, 2, 64, mask_classes=2)  # binary ROI mask
    x0, x1, mask = inp["x0"], inp["x1"], inp["mask"]
    out["inputs_check"] = dict(x0=float(x0.double().sum()), mask=float(mask.sum()))
    for name, (modname, raw4) in MASK_VARIANTS.items():
        mod = rb.reference_module(f"src.models.{modname}")
        torch.manual_seed(0)
        net = ounet.dezero_(ounet.RawUNetModel(**RAW4) if raw4 else ounet.UNetModel(**SMALL), seed=1984)
        lit = mod.ConditionalFlowMatchingLitModule(
            net=net, flow_matcher=oflow.ConditionalFlowMatcher(0.0),
            solver=functools.partial(oflow.NeuralODE, solver="dopri5", sensitivity="adjoint", atol=1e-4, rtol=1e-4),
            optimizer=functools.partial(torch.optim.Adam, lr=1e-4), scheduler=None, log_images=False)
        lit.eval()
        rec = dict(checksums=checksums(lit))
        torch.manual_seed(79)
        t_drawn = torch.rand(2)
        torch.manual_seed(79)
        loss = lit.model_step((x0, x1, mask))
        lit.zero_grad()
        loss.backward()
        rec["model_step"] = dict(rng_seed=79, t=t_drawn, loss=loss.detach(),
                                 grad_norms={k: float(p.grad.double().norm()) for k, p in lit.named_parameters()})
        if name == "mask_toggle":  # training_step toggles: rand(1) is drawn BEFORE t; record one step of each kind
            rec["training_step"] = []
            seen, seed = set(), 80
            while len(seen) < 2:
                torch.manual_seed(seed)
                toggled = bool(torch.rand(1).item() < 0.5)
                t2 = torch.rand(2)
                if toggled not in seen:
                    seen.add(toggled)
                    torch.manual_seed(seed)
                    rec["training_step"].append(dict(rng_seed=seed, toggled=toggled, t=t2,
                                                     loss=lit.training_step((x0, x1, mask), 0).detach()))
                seed += 1
        with torch.no_grad():
            rec["forward"] = lit(inp["t"], x0, mask) if raw4 else lit(inp["t"], x0)
        rec["generate_num_steps2"] = lit.generate(x0[:1], mask[:1], num_steps=2) if raw4 else lit.generate(x0[:1], num_steps=2)
        out[name] = rec
    torch.save(out, os.path.join(OUT, "mask_variants_small.pt"))
    return out


def paired_dataset():
    """The reference's PairedDataset / PairedDataModule (src/data/paired_data_module.py), unmodified, on a tiny synthetic
    PNG dataset.  The fixture carries the raw images so the test can rebuild the PNG files."""
    import random
    import tempfile

    import cv2
    import numpy as np
    pdm = rb.reference_module("src.data.paired_data_module")
    rng = np.random.RandomState(7)
    n, hw = 3, 96
    # smooth-ish random RGB images (so the antialiased resize has structure to average)
    imgs = {f"img{k}_{kind}.png": rng.randint(0, 256, (hw, hw, 3)).astype(np.uint8) for k in range(n) for kind in ("he", "ihc")}
    out = dict(images=imgs, rows=[(f"img{k}_he.png", f"img{k}_ihc.png", "train" if k < 2 else "test") for k in range(n)])
    with tempfile.TemporaryDirectory() as d:
        for split in ("train", "test"):
            os.makedirs(os.path.join(d, split))
        with open(os.path.join(d, "meta.csv"), "w") as f:
            f.write("image_id,he_filepath,ihc_filepath,split\n")
            for k, (a, b, split) in enumerate(out["rows"]):
                f.write(f"{k},{a},{b},{split}\n")
                for name in (a, b):
                    cv2.imwrite(os.path.join(d, split, name), cv2.cvtColor(imgs[name], cv2.COLOR_RGB2BGR))
        kw = dict(data_dir=d, csv_file_name="meta.csv", source_column="he_filepath", target_column="ihc_filepath")
        # train: random crop 64 + flips, drawn from the torch / python default generators
        ds = pdm.PairedDataset(folder="train", image_size=64, use_augmentation=True, **kw)
        torch.manual_seed(123)
        random.seed(123)
        out["train_aug"] = dict(torch_seed=123, python_seed=123, image_size=64, items=[ds[0], ds[1], ds[0]])
        # eval: antialiased resize to 48 (and to 96 = identity), T2S direction, filenames returned
        ds = pdm.PairedDataset(folder="test", image_size=48, use_augmentation=False, direction="T2S", return_filename=True, **kw)
        out["eval_resize48_T2S"] = ds[0]
        ds = pdm.PairedDataset(folder="train", image_size=96, use_augmentation=False, **kw)
        out["eval_identity96"] = ds[1]
        dm = pdm.PairedDataModule(batch_size=8, image_size=64, **kw)

        class _T:
            world_size = 4
        dm.trainer = _T()
        dm.setup()
        out["per_device_batch_8_over_4"] = dm.batch_size_per_device
        _T.world_size = 3
        try:
            dm.setup()
            out["indivisible_raises"] = None
        except RuntimeError as e:
            out["indivisible_raises"] = str(e)
    torch.save(out, os.path.join(OUT, "paired_dataset_small.pt"))
    return out


def any2any_dataset():
    """The reference's PairedAnyToAnyDataset / ClassConditionalAnyToAnyDataModule (src/data/class_conditional_he_amyloid.py),
    unmodified, on a tiny synthetic three-domain PNG tree."""
    import random
    import tempfile

    import numpy as np
    from PIL import Image
    mod = rb.reference_module("src.data.class_conditional_he_amyloid")
    rng = np.random.RandomState(9)
    mapping = {0: "he", 1: "amyloid", 2: "tau"}
    names = [f"t{k}.png" for k in range(5)]
    imgs = {(c, n): rng.randint(0, 256, (80, 88, 3)).astype(np.uint8) for c in mapping for n in names}
    del imgs[(2, "t4.png")]  # one file is missing in one domain: intersection drops it, union keeps it
    out = dict(mapping=mapping, names=names, images={f"{c}/{n}": v for (c, n), v in imgs.items()})
    with tempfile.TemporaryDirectory() as d:
        for c, folder in mapping.items():
            os.makedirs(os.path.join(d, folder))
        for (c, n), v in imgs.items():
            Image.fromarray(v).save(os.path.join(d, mapping[c], n))
        for tag, kw in (("intersection_same_crop", dict()),
                        ("union_separate_crops_fixed_source", dict(filename_mode="union", same_crop_for_pair=False,
                                                                   source_domain_mode=1))):
            ds = mod.PairedAnyToAnyDataset(d, mapping, crop_size=64, **kw)
            torch.manual_seed(321)
            random.seed(321)
            items = [ds[i] for i in (0, 3, len(ds) - 1, 1)]
            out[tag] = dict(kwargs=kw, torch_seed=321, python_seed=321, filenames=list(ds.filenames), indices=(0, 3, len(ds) - 1, 1),
                            items=[(a, b, int(c)) for a, b, c in items])
        dm = mod.ClassConditionalAnyToAnyDataModule(d, mapping, crop_size=64, batch_size=6, val_split=0.4, split_seed=7)
        dm.prepare_data()
        with open(os.path.join(d, "train_val_split.json")) as f:
            out["split_json"] = f.read()

        class _T:
            world_size = 3
        dm.trainer = _T()
        dm.setup()
        out["per_device_batch_6_over_3"] = dm.batch_size_per_device
        out["train_len"], out["val_len"] = len(dm.data_train), len(dm.data_val)
    torch.save(out, os.path.join(OUT, "any2any_dataset_small.pt"))
    return out


def mask_datasets():
    """The reference's mask-carrying datasets, unmodified: src/data/paired_data_multiclassmask.py (config M's data: class-id
    mask, same crop / flips as the tiles, Pillow NEAREST on the eval path) and src/data/paired_data_mask_he_amyloid.py
    (binarised mask, cv2 INTER_NEAREST)."""
    import random
    import tempfile

    import cv2
    import numpy as np
    mc = rb.reference_module("src.data.paired_data_multiclassmask")
    ha = rb.reference_module("src.data.paired_data_mask_he_amyloid")
    rng = np.random.RandomState(13)
    n, h, w = 2, 90, 100
    imgs = {}
    rows = []
    for k in range(n):
        imgs[f"i{k}_he.png"] = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        imgs[f"i{k}_ihc.png"] = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        imgs[f"i{k}_gw.png"] = rng.randint(0, 5, (h, w)).astype(np.uint8)          # class ids 0..4
        imgs[f"i{k}_am.png"] = (rng.randint(0, 4, (h, w)) * 85).astype(np.uint8)   # 0, 85, 170, 255
        rows.append((f"i{k}_he.png", f"i{k}_ihc.png", f"i{k}_gw.png", f"i{k}_am.png", "train"))
    out = dict(images=imgs, rows=rows)
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "train"))
        with open(os.path.join(d, "meta.csv"), "w") as f:
            f.write("image_id,he_filepath,ihc_filepath,graywhite_filepath,amyloid_filepath,split\n")
            for k, r in enumerate(rows):
                f.write(f"{k},{','.join(r)}\n")
        for name, v in imgs.items():
            cv2.imwrite(os.path.join(d, "train", name), v if v.ndim == 2 else cv2.cvtColor(v, cv2.COLOR_RGB2BGR))
        kw = dict(data_dir=d, csv_file_name="meta.csv", source_column="he_filepath", target_column="ihc_filepath", folder="train")
        ds = mc.PairedDataset(image_size=64, use_augmentation=True, **kw)
        torch.manual_seed(55)
        random.seed(55)
        out["multiclass_train_aug"] = dict(torch_seed=55, python_seed=55, image_size=64, items=[ds[0], ds[1], ds[1]])
        ds = mc.PairedDataset(image_size=48, use_augmentation=False, direction="T2S", **kw)
        out["multiclass_eval48_T2S"] = ds[1]
        ds = ha.PairedHEIHCDataset(image_size=48, **kw)
        out["he_amyloid_eval48"] = ds[0]
        ds = ha.PairedHEIHCDataset(image_size=64, direction="IHC_to_HE", **kw)
        out["he_amyloid_eval64_reverse"] = ds[1]
    torch.save(out, os.path.join(OUT, "mask_datasets_small.pt"))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # fixed reduction order on the CPU
    for fn in (lambda: simple_fm(False), lambda: simple_fm(True), multitask, mask_variants, paired_dataset, any2any_dataset, mask_datasets):
        o = fn()
        print({k: (tuple(v.shape) if torch.is_tensor(v) else type(v).__name__) for k, v in o.items()})
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
