"""Host side of the input / output pipeline (stain2stain_b200/data.py), CPU tier.

The Pillow resampling tables are pinned against REAL Pillow (the library `TF.resize` calls in the reference,
src/data/paired_data_module.py:201-203), bit for bit; crop / flip / normalise semantics and the datamodule's batch rule
are pinned against tests/golden/paired_dataset_small.pt, which the reference's own PairedDataset / PairedDataModule
produced unmodified (oracle/make_golden.py::paired_dataset)."""
import os
import random

import numpy as np
import pytest
import torch

from stain2stain_b200 import data as D

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold():
    return torch.load(os.path.join(GOLD, "paired_dataset_small.pt"), map_location="cpu", weights_only=False)


def _resize_numpy(img, oh, ow):
    h, w = img.shape[:2]
    if w != ow:
        img = D.resample_pass_numpy(img, *D.pillow_coeffs(w, ow), axis=1)
    if h != oh:
        img = D.resample_pass_numpy(img, *D.pillow_coeffs(h, oh), axis=0)
    return img


def _to_tensor_norm(img_u8_hwc):
    t = torch.from_numpy(np.ascontiguousarray(img_u8_hwc)).permute(2, 0, 1).float().div(255)
    return (t - 0.5) / 0.5


@pytest.mark.parametrize("src,dst", [((512, 512), (256, 256)), ((96, 96), (48, 48)), ((200, 120), (64, 64)),
                                     ((100, 100), (37, 53)), ((64, 64), (128, 128)), ((97, 131), (97, 40))])
def test_pillow_tables_match_real_pillow_bit_for_bit(src, dst):
    from PIL import Image
    rng = np.random.RandomState(src[0] * 7 + dst[1])
    img = rng.randint(0, 256, (src[0], src[1], 3)).astype(np.uint8)
    want = np.asarray(Image.fromarray(img).resize((dst[1], dst[0]), Image.BILINEAR))
    got = _resize_numpy(img, dst[0], dst[1])
    assert got.shape == want.shape and np.array_equal(got, want)


def test_exact_factor_two_table_is_the_known_answer():
    b, k = D.pillow_coeffs(8, 4)
    assert b.tolist() == [[0, 3], [1, 4], [3, 4], [5, 3]]
    one = 1 << D.PRECISION_BITS
    assert k[1].tolist()[:4] == [one // 8, 3 * one // 8, 3 * one // 8, one // 8]  # [1, 3, 3, 1] / 8


def test_eval_path_matches_the_reference_dataset():
    g = _gold()
    tgt, src, tf, sf = g["eval_resize48_T2S"]  # direction T2S: (target, source, target_filename, source_filename)
    assert (tf, sf) == ("img2_ihc.png", "img2_he.png")
    assert torch.equal(_to_tensor_norm(_resize_numpy(g["images"]["img2_he.png"], 48, 48)), src)
    assert torch.equal(_to_tensor_norm(_resize_numpy(g["images"]["img2_ihc.png"], 48, 48)), tgt)
    s, t = g["eval_identity96"]
    assert torch.equal(_to_tensor_norm(g["images"]["img1_he.png"]), s)


def test_augment_parameters_follow_the_reference_rng_calls():
    g = _gold()
    rec = g["train_aug"]
    torch.manual_seed(rec["torch_seed"])
    random.seed(rec["python_seed"])
    params = D.draw_augment_params(3, 96, 96, rec["image_size"])
    S = rec["image_size"]
    for (top, left, hf, vf), idx, item in zip(params.tolist(), (0, 1, 0), rec["items"]):
        for name, want in zip((f"img{idx}_he.png", f"img{idx}_ihc.png"), item):
            a = g["images"][name][top:top + S, left:left + S]
            if hf:
                a = a[:, ::-1]
            if vf:
                a = a[::-1]
            assert torch.equal(_to_tensor_norm(a), want)
    # image already at the crop size: RandomCrop.get_params draws nothing (only the two flips consume randomness)
    torch.manual_seed(5)
    before = torch.get_rng_state()
    D.draw_augment_params(2, 64, 64, 64)
    assert torch.equal(before, torch.get_rng_state())
    with pytest.raises(ValueError):
        D.draw_augment_params(1, 32, 64, 64)


def test_datamodule_setup_rule(tmp_path):
    g = _gold()
    dm = D.PairedDataModule(data_dir=str(tmp_path), batch_size=8, image_size=64)

    class T:
        world_size = 4
    dm.trainer = T()
    dm.setup()
    assert dm.batch_size_per_device == g["per_device_batch_8_over_4"] == 2
    T.world_size = 3
    with pytest.raises(RuntimeError) as e:
        dm.setup()
    assert str(e.value) == g["indivisible_raises"]


def test_paired_dataset_reads_the_csv_like_the_reference(tmp_path):
    import cv2
    g = _gold()
    for split in ("train", "test"):
        os.makedirs(tmp_path / split)
    with open(tmp_path / "meta.csv", "w") as f:
        f.write("image_id,he_filepath,ihc_filepath,split\n")
        for k, (a, b, split) in enumerate(g["rows"]):
            f.write(f"{k},{a},{b},{split}\n")
            for name in (a, b):
                cv2.imwrite(str(tmp_path / split / name), cv2.cvtColor(g["images"][name], cv2.COLOR_RGB2BGR))
    ds = D.PairedDataset(str(tmp_path), "meta.csv", "he_filepath", "ihc_filepath", "train", image_size=64)
    assert len(ds) == 2
    s, t, sf, tf = ds[1]
    assert (sf, tf) == ("img1_he.png", "img1_ihc.png")
    assert np.array_equal(s[:, :, ::-1], g["images"]["img1_he.png"]) and np.array_equal(t[:, :, ::-1], g["images"]["img1_ihc.png"])
    with pytest.raises(AssertionError):
        D.PairedDataset(str(tmp_path), "missing.csv", "he_filepath", "ihc_filepath", "train")


def test_png_writer_round_trip(tmp_path):
    import cv2
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (37, 53, 3)).astype(np.uint8)
    D.write_png(str(tmp_path / "a.png"), img)
    assert np.array_equal(cv2.imread(str(tmp_path / "a.png"))[:, :, ::-1], img)
    grey = rng.randint(0, 256, (16, 9)).astype(np.uint8)
    D.write_png(str(tmp_path / "g.png"), grey)
    assert np.array_equal(cv2.imread(str(tmp_path / "g.png"), cv2.IMREAD_GRAYSCALE), grey)
    w = D.AsyncPngWriter()
    w.submit([str(tmp_path / "b0.png"), str(tmp_path / "b1.png")], torch.from_numpy(np.stack([img, img[::-1].copy()])))
    w.wait()
    assert np.array_equal(cv2.imread(str(tmp_path / "b1.png"))[:, :, ::-1], img[::-1])


# ---- any-to-any (class-conditional) data module: src/data/class_conditional_he_amyloid.py ------------------------------
def _any2any_tree(tmp_path, g):
    from PIL import Image
    for c, folder in g["mapping"].items():
        os.makedirs(tmp_path / folder, exist_ok=True)
    for key, v in g["images"].items():
        c, n = key.split("/")
        Image.fromarray(v).save(str(tmp_path / g["mapping"][int(c)] / n))


@pytest.mark.parametrize("tag", ["intersection_same_crop", "union_separate_crops_fixed_source"])
def test_any2any_dataset_follows_the_reference(tmp_path, tag):
    """Same domain choices, same crops (RNG calls in the reference's order), same normalised tiles (host statement of the
    crop + to_tensor + normalise kernel) as the reference's PairedAnyToAnyDataset produced."""
    from stain2stain_b200.data_any2any import PairedAnyToAnyDataset
    g = torch.load(os.path.join(GOLD, "any2any_dataset_small.pt"), map_location="cpu", weights_only=False)
    _any2any_tree(tmp_path, g)
    rec = g[tag]
    ds = PairedAnyToAnyDataset(str(tmp_path), g["mapping"], crop_size=64, **rec["kwargs"])
    assert ds.filenames == rec["filenames"] and ds.num_classes == 3
    torch.manual_seed(rec["torch_seed"])
    random.seed(rec["python_seed"])
    for idx, (want_s, want_t, want_label) in zip(rec["indices"], rec["items"]):
        src, tgt, params, label = ds[idx]
        assert label == want_label
        (i, j, _, _), (i2, j2, _, _) = params.tolist()
        assert torch.equal(_to_tensor_norm(src[i:i + 64, j:j + 64]), want_s)
        assert torch.equal(_to_tensor_norm(tgt[i2:i2 + 64, j2:j2 + 64]), want_t)
    with pytest.raises(ValueError):
        PairedAnyToAnyDataset(str(tmp_path), {0: "he", 1: "missing"}, crop_size=64)
    with pytest.raises(ValueError):
        PairedAnyToAnyDataset(str(tmp_path), g["mapping"], crop_size=64, filename_mode="bogus")


def test_any2any_datamodule_split_and_batch_rule(tmp_path):
    from stain2stain_b200.data_any2any import ClassConditionalAnyToAnyDataModule
    g = torch.load(os.path.join(GOLD, "any2any_dataset_small.pt"), map_location="cpu", weights_only=False)
    _any2any_tree(tmp_path, g)
    dm = ClassConditionalAnyToAnyDataModule(str(tmp_path), g["mapping"], crop_size=64, batch_size=6, val_split=0.4, split_seed=7)
    with pytest.raises(RuntimeError):
        dm.setup()  # no split file yet
    dm.prepare_data()
    assert open(tmp_path / "train_val_split.json").read() == g["split_json"]

    class T:
        world_size = 3
    dm.trainer = T()
    dm.setup()
    assert dm.batch_size_per_device == g["per_device_batch_6_over_3"] == 2
    assert (len(dm.data_train), len(dm.data_val)) == (g["train_len"], g["val_len"])
    T.world_size = 4
    with pytest.raises(RuntimeError):
        dm.setup()


# ---- mask-carrying datasets: src/data/paired_data_multiclassmask.py, src/data/paired_data_mask_he_amyloid.py -------------
@pytest.mark.parametrize("src,dst", [((96, 96), (48, 48)), ((90, 100), (64, 64)), ((50, 70), (64, 64)), ((100, 100), (37, 53)),
                                     ((333, 217), (100, 300)), ((7, 5), (64, 33))])
def test_nearest_mask_resize_rules_match_pillow_and_opencv(src, dst):
    import cv2
    from PIL import Image
    m = np.random.RandomState(src[0] + dst[0]).randint(0, 6, src).astype(np.uint8)
    assert np.array_equal(D.nearest_resize_pil(m, *dst), np.asarray(Image.fromarray(m).resize((dst[1], dst[0]), Image.NEAREST)))
    assert np.array_equal(D.nearest_resize_cv2(m, *dst), cv2.resize(m, (dst[1], dst[0]), interpolation=cv2.INTER_NEAREST))


def _mask_tree(tmp_path, g):
    import cv2
    os.makedirs(tmp_path / "train", exist_ok=True)
    with open(tmp_path / "meta.csv", "w") as f:
        f.write("image_id,he_filepath,ihc_filepath,graywhite_filepath,amyloid_filepath,split\n")
        for k, r in enumerate(g["rows"]):
            f.write(f"{k},{','.join(r)}\n")
    for name, v in g["images"].items():
        cv2.imwrite(str(tmp_path / "train" / name), v if v.ndim == 2 else cv2.cvtColor(v, cv2.COLOR_RGB2BGR))
    return dict(data_dir=str(tmp_path), csv_file_name="meta.csv", source_column="he_filepath", target_column="ihc_filepath",
                folder="train")


def test_mask_datasets_follow_the_reference_host_side(tmp_path):
    """Decode + mask rules + RNG protocol of the two mask datasets against the items the reference's own datasets returned
    (the tile arithmetic is the host statement of the device kernel; the GPU tier runs the kernel itself)."""
    g = torch.load(os.path.join(GOLD, "mask_datasets_small.pt"), map_location="cpu", weights_only=False)
    kw = _mask_tree(tmp_path, g)
    rec = g["multiclass_train_aug"]
    ds = D.PairedMulticlassMaskDataset(image_size=rec["image_size"], use_augmentation=True, **kw)
    torch.manual_seed(rec["torch_seed"])
    random.seed(rec["python_seed"])
    params = D.draw_augment_params(3, 90, 100, rec["image_size"])
    S = rec["image_size"]
    for (top, left, hf, vf), idx, want in zip(params.tolist(), (0, 1, 1), rec["items"]):
        s, t, m = ds[idx]
        outs = []
        for a in (s[:, :, ::-1], t[:, :, ::-1], m):
            a = a[top:top + S, left:left + S]
            a = a[:, ::-1] if hf else a
            a = a[::-1] if vf else a
            outs.append(np.ascontiguousarray(a))
        assert torch.equal(_to_tensor_norm(outs[0]), want[0]) and torch.equal(_to_tensor_norm(outs[1]), want[1])
        assert torch.equal(torch.from_numpy(outs[2]).float().unsqueeze(0), want[2])
    # eval: Pillow NEAREST for the class-id mask
    ds = D.PairedMulticlassMaskDataset(image_size=48, use_augmentation=False, direction="T2S", **kw)
    s, t, m = ds[1]
    want = g["multiclass_eval48_T2S"]
    assert torch.equal(torch.from_numpy(D.nearest_resize_pil(m, 48, 48)).float().unsqueeze(0), want[2])
    assert torch.equal(_to_tensor_norm(_resize_numpy(t[:, :, ::-1], 48, 48)), want[0])  # T2S: target first
    # binarised amyloid mask: cv2 INTER_NEAREST, then > 1
    ds = D.PairedHEIHCMaskDataset(image_size=48, **kw)
    s, t, m = ds[0]
    want = g["he_amyloid_eval48"]
    assert m.dtype == np.uint8 and torch.equal(torch.from_numpy(m).unsqueeze(0), want[2])
    assert torch.equal(_to_tensor_norm(_resize_numpy(s[:, :, ::-1], 48, 48)), want[0])
