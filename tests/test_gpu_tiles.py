"""GPU parity of the image-space kernels either side of the UNet (csrc/tiles.cuh).

Integer / byte work is held bit-exact: `tile_prep` against the torchvision chain the reference's dataset runs (and against
the fixture its unmodified PairedDataset produced), `resample_u8` against real Pillow, `denorm_u8` against the torch
formula.  The floating-point loss kernels are held to 1e-5 relative (fp32 sums in another order)."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold():
    return torch.load(os.path.join(GOLD, "paired_dataset_small.pt"), map_location="cpu", weights_only=False)


@pytest.mark.parametrize("cx,with_extra,interp", [(3, False, False), (3, True, False), (3, True, True), (4, False, False)])
def test_patch_pack_matches_unfold(cx, with_extra, interp):
    from stain2stain_b200 import kernels as K
    g = torch.Generator(device=DEV).manual_seed(cx * 10 + with_extra)
    B, H, W = 3, 20, 24
    x0 = torch.rand(B, cx, H, W, device=DEV, generator=g) * 2 - 1
    x1 = torch.rand(B, cx, H, W, device=DEV, generator=g) * 2 - 1 if interp else None
    t = torch.rand(B, device=DEV, generator=g) if interp else None
    extra = (torch.rand(B, 1, H, W, device=DEV, generator=g) > 0.5).float() if with_extra else None
    got = K.to_float(K.patch_pack(x0, x1, t, extra), K.ACT)
    src = x0 if not interp else (1 - t)[:, None, None, None] * x0 + t[:, None, None, None] * x1
    if extra is not None:
        src = torch.cat([src, extra], 1)
    ct = src.shape[1]
    cols = torch.nn.functional.unfold(src, 3, padding=1).view(B, ct, 9, H, W)  # [B, c, tap, H, W]
    want = cols.permute(0, 3, 4, 2, 1).reshape(B, H, W, 9 * ct)               # column tap*ct + c
    assert got.shape == (B, H, W, 64)
    assert torch.allclose(got[..., :9 * ct], want, atol=2e-3, rtol=2e-3)       # 16-bit storage
    assert float(got[..., 9 * ct:].abs().max()) == 0.0


def test_fm_loss_weighted_fwd_bwd():
    from stain2stain_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(3)
    B, H = 3, 40
    v = (torch.randn(B, 3, H, H, device=DEV, generator=g)).requires_grad_(True)
    x0 = torch.rand(B, 3, H, H, device=DEV, generator=g) * 2 - 1
    x1 = torch.rand(B, 3, H, H, device=DEV, generator=g) * 2 - 1
    mask = (torch.rand(B, 1, H, H, device=DEV, generator=g) > 0.7).float()
    loss = ops.fm_loss_weighted(v, x0, x1, mask, 10.0)
    loss.backward()
    v2 = v.detach().clone().requires_grad_(True)
    w = (1.0 + 10.0 * mask).expand_as(v2)                      # conditional_flow_matching_masked.py:84-90
    want = (w * (v2 - (x1 - x0)) ** 2).sum() / (w.sum() + 1e-8)
    want.backward()
    assert abs(float(loss) - float(want)) <= 1e-5 * abs(float(want))
    assert torch.allclose(v.grad, v2.grad, rtol=1e-5, atol=1e-9)


def test_roi_charbonnier_matches_reference_formula():
    from stain2stain_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(4)
    B, H = 2, 48
    x0 = torch.rand(B, 3, H, H, device=DEV, generator=g) * 2 - 1
    x1 = torch.rand(B, 3, H, H, device=DEV, generator=g) * 2 - 1
    t = torch.rand(B, device=DEV, generator=g)
    for mask in ((torch.rand(B, 1, H, H, device=DEV, generator=g) > 0.5).float(), torch.zeros(B, 1, H, H, device=DEV)):
        xt = t[:, None, None, None] * x1 + (1 - t[:, None, None, None]) * x0
        diff = xt - x1                                          # conditional_flow_matching_ROI_loss.py:80-92
        charb = torch.sqrt(diff * diff + 1e-3 * 1e-3)
        want = (charb * mask).sum() / (mask.sum() * 3 + 1e-8)
        got = ops.roi_charbonnier(x0, x1, t, mask)
        assert abs(float(got) - float(want)) <= 1e-5 * max(abs(float(want)), 1e-12)


def test_denorm_u8_is_bit_exact():
    from stain2stain_b200 import data as D
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.rand(2, 3, 33, 47, device=DEV, generator=g) * 2.6 - 1.3  # includes values that clamp on both sides
    x[0, 0, 0, :6] = torch.tensor([-1.0, 1.0, 0.0, -0.999999, 0.999999, 1e-8], device=DEV)
    want = ((x * 0.5 + 0.5).clamp(0, 1) * 255.0 + 0.5).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    got = D.denormalize_to_uint8(x)
    assert got.dtype == torch.uint8 and torch.equal(got, want)
    assert torch.equal(D.denormalize_to_uint8(x[0]), want[:1])


def test_tile_prep_is_bit_exact_with_the_torchvision_chain():
    """The chain of paired_data_module.py:171-199 on PIL images vs one kernel on the raw BGR bytes."""
    import torchvision.transforms.functional as TF
    from PIL import Image
    from torchvision import transforms
    from stain2stain_b200 import kernels as K
    rng = np.random.RandomState(11)
    B, Hs, Ws, S = 5, 80, 112, 64
    rgb = rng.randint(0, 256, (2, B, Hs, Ws, 3)).astype(np.uint8)
    msk = (rng.randint(0, 2, (B, Hs, Ws)) * 255).astype(np.uint8)
    params = torch.tensor([[0, 0, 0, 0], [16, 48, 1, 0], [3, 7, 0, 1], [16, 0, 1, 1], [9, 48, 1, 0]], dtype=torch.int32)
    norm = transforms.Normalize(mean=[0.5, 0.5, 0.5], std=[0.5, 0.5, 0.5])
    want = [[], [], []]
    for b in range(B):
        i, j, hf, vf = params[b].tolist()
        for k in range(3):
            img = Image.fromarray(rgb[k][b]) if k < 2 else Image.fromarray(msk[b])
            img = TF.crop(img, i, j, S, S)
            if hf:
                img = TF.hflip(img)
            if vf:
                img = TF.vflip(img)
            ten = TF.to_tensor(img)
            want[k].append(norm(ten) if k < 2 else ten)
    bgr = np.ascontiguousarray(rgb[:, :, :, :, ::-1])  # what cv2.imread hands over
    x0, x1, m = K.tile_prep(torch.from_numpy(bgr[0]).to(DEV), torch.from_numpy(bgr[1]).to(DEV), params.to(DEV), S, bgr=True,
                            mask_u8=torch.from_numpy(msk).to(DEV))
    assert torch.equal(x0.cpu(), torch.stack(want[0])) and torch.equal(x1.cpu(), torch.stack(want[1]))
    assert torch.equal(m.cpu(), torch.stack(want[2]))
    y0, _, _ = K.tile_prep(torch.from_numpy(rgb[0]).to(DEV), None, params.to(DEV), S, bgr=False)
    assert torch.equal(y0, x0)


@pytest.mark.parametrize("src,dst", [((512, 512), (256, 256)), ((96, 96), (48, 48)), ((200, 120), (64, 64)),
                                     ((100, 100), (37, 53)), ((64, 64), (128, 128))])
def test_resize_u8_is_bit_exact_with_pillow(src, dst):
    from PIL import Image
    from stain2stain_b200 import data as D
    rng = np.random.RandomState(src[0] + dst[1])
    imgs = rng.randint(0, 256, (3, src[0], src[1], 3)).astype(np.uint8)
    want = np.stack([np.asarray(Image.fromarray(a).resize((dst[1], dst[0]), Image.BILINEAR)) for a in imgs])
    got = D.resize_u8(torch.from_numpy(imgs).to(DEV), dst[0], dst[1])
    assert np.array_equal(got.cpu().numpy(), want)


def test_paired_dataset_and_loader_reproduce_the_reference_items(tmp_path):
    """PNG files -> cv2 decode -> uint8 H2D -> kernels, against the items the reference's PairedDataset returned."""
    import cv2
    from stain2stain_b200 import data as D
    g = _gold()
    for split in ("train", "test"):
        os.makedirs(tmp_path / split)
    with open(tmp_path / "meta.csv", "w") as f:
        f.write("image_id,he_filepath,ihc_filepath,split\n")
        for k, (a, b, split) in enumerate(g["rows"]):
            f.write(f"{k},{a},{b},{split}\n")
            for name in (a, b):
                cv2.imwrite(str(tmp_path / split / name), cv2.cvtColor(g["images"][name], cv2.COLOR_RGB2BGR))
    kw = dict(data_dir=str(tmp_path), csv_file_name="meta.csv", source_column="he_filepath", target_column="ihc_filepath")
    rec = g["train_aug"]
    ds = D.PairedDataset(folder="train", image_size=rec["image_size"], use_augmentation=True, **kw)
    torch.manual_seed(rec["torch_seed"])
    random.seed(rec["python_seed"])
    for idx, want in zip((0, 1, 0), rec["items"]):
        s, t = ds.get_reference_item(idx)
        assert torch.equal(s.cpu(), want[0]) and torch.equal(t.cpu(), want[1])
    ds = D.PairedDataset(folder="test", image_size=48, use_augmentation=False, direction="T2S", return_filename=True, **kw)
    a, b, fa, fb = ds.get_reference_item(0)
    wa, wb, wfa, wfb = g["eval_resize48_T2S"]
    assert (fa, fb) == (wfa, wfb) and torch.equal(a.cpu(), wa) and torch.equal(b.cpu(), wb)
    # batched loader: same RNG protocol, whole batch in one kernel
    dm = D.PairedDataModule(batch_size=2, num_workers=2, image_size=rec["image_size"], use_augmentation=True, **kw)
    dm.setup()
    loader = dm.test_dataloader()   # no shuffle
    loader.ds = D.PairedDataset(folder="train", image_size=rec["image_size"], use_augmentation=True, **kw)
    torch.manual_seed(rec["torch_seed"])
    random.seed(rec["python_seed"])
    batches = list(loader)
    assert len(batches) == 1 and batches[0][0].shape == (2, 3, 64, 64) and batches[0][0].is_cuda
    assert torch.equal(batches[0][0][0].cpu(), rec["items"][0][0]) and torch.equal(batches[0][1][1].cpu(), rec["items"][1][1])


@pytest.mark.parametrize("tag", ["intersection_same_crop", "union_separate_crops_fixed_source"])
def test_any2any_dataset_reproduces_the_reference_items(tmp_path, tag):
    """src/data/class_conditional_he_amyloid.py through the device kernel: bit-identical tiles, same labels."""
    from PIL import Image
    from stain2stain_b200.data_any2any import AnyToAnyBatchLoader, PairedAnyToAnyDataset
    g = torch.load(os.path.join(GOLD, "any2any_dataset_small.pt"), map_location="cpu", weights_only=False)
    for c, folder in g["mapping"].items():
        os.makedirs(tmp_path / folder, exist_ok=True)
    for key, v in g["images"].items():
        c, n = key.split("/")
        Image.fromarray(v).save(str(tmp_path / g["mapping"][int(c)] / n))
    rec = g[tag]
    ds = PairedAnyToAnyDataset(str(tmp_path), g["mapping"], crop_size=64, **rec["kwargs"])
    torch.manual_seed(rec["torch_seed"])
    random.seed(rec["python_seed"])
    for idx, (want_s, want_t, want_label) in zip(rec["indices"], rec["items"]):
        s, t, label = ds.get_reference_item(idx)
        assert label == want_label and torch.equal(s.cpu(), want_s) and torch.equal(t.cpu(), want_t)
    loader = AnyToAnyBatchLoader(ds, batch_size=3, shuffle=False, num_workers=2)
    batches = list(loader)
    assert len(batches) == len(loader) == (len(ds) + 2) // 3
    x0, x1, y = batches[0]
    assert x0.shape == (3, 3, 64, 64) and x1.shape == x0.shape and x0.is_cuda and y.dtype == torch.int64 and y.shape == (3,)
    assert float(x0.min()) >= -1.0 and float(x0.max()) <= 1.0


def test_mask_datasets_reproduce_the_reference_items(tmp_path):
    """src/data/paired_data_multiclassmask.py and paired_data_mask_he_amyloid.py through the device kernels (raw class-id
    mask path of tile_prep included): bit-identical to the items the reference's datasets returned."""
    import cv2
    from stain2stain_b200 import data as D
    g = torch.load(os.path.join(GOLD, "mask_datasets_small.pt"), map_location="cpu", weights_only=False)
    os.makedirs(tmp_path / "train", exist_ok=True)
    with open(tmp_path / "meta.csv", "w") as f:
        f.write("image_id,he_filepath,ihc_filepath,graywhite_filepath,amyloid_filepath,split\n")
        for k, r in enumerate(g["rows"]):
            f.write(f"{k},{','.join(r)}\n")
    for name, v in g["images"].items():
        cv2.imwrite(str(tmp_path / "train" / name), v if v.ndim == 2 else cv2.cvtColor(v, cv2.COLOR_RGB2BGR))
    kw = dict(data_dir=str(tmp_path), csv_file_name="meta.csv", source_column="he_filepath", target_column="ihc_filepath",
              folder="train")
    rec = g["multiclass_train_aug"]
    ds = D.PairedMulticlassMaskDataset(image_size=rec["image_size"], use_augmentation=True, **kw)
    torch.manual_seed(rec["torch_seed"])
    random.seed(rec["python_seed"])
    for idx, want in zip((0, 1, 1), rec["items"]):
        got = ds.get_reference_item(idx)
        assert all(torch.equal(a.cpu(), b) for a, b in zip(got, want))
    ds = D.PairedMulticlassMaskDataset(image_size=48, use_augmentation=False, direction="T2S", **kw)
    assert all(torch.equal(a.cpu(), b) for a, b in zip(ds.get_reference_item(1), g["multiclass_eval48_T2S"]))
    ds = D.PairedHEIHCMaskDataset(image_size=48, **kw)
    got = ds.get_reference_item(0)
    assert got[2].dtype == torch.uint8 and all(torch.equal(a.cpu(), b) for a, b in zip(got, g["he_amyloid_eval48"]))
    ds = D.PairedHEIHCMaskDataset(image_size=64, direction="IHC_to_HE", **kw)
    assert all(torch.equal(a.cpu(), b) for a, b in zip(ds.get_reference_item(1), g["he_amyloid_eval64_reverse"]))
