"""Input and output side of the hot path, B200-first:

  reference (src/data/paired_data_module.py:149-223), per item, on CPU workers:
      cv2.imread -> BGR2RGB -> PIL -> [RandomCrop + hflip + vflip | TF.resize] -> to_tensor -> Normalize(0.5, 0.5)
      -> fp32 CHW, collated and copied to the device as fp32 (2 x 3 MB per 512^2 pair)
  here:
      workers only DECODE (cv2.imread, bytes stay BGR uint8 HWC) into pinned staging, the batch crosses PCIe as uint8
      (4x fewer bytes), and ONE kernel (`s2s_tile_prep`) does crop + flips + channel swap + to_tensor + normalise into
      the fp32 NCHW tensors `model_step` takes -- bit-identical to the torchvision chain.  The eval path's `TF.resize`
      (Pillow's antialiased 8-bit resampling) runs as two uint8 passes (`s2s_resample_u8`) with integer coefficient
      tables built on the host exactly like Pillow's `precompute_coeffs` / `normalize_coeffs_8bpc`.

`PairedDataset` / `PairedDataModule` keep the reference's constructor keywords, CSV conventions, `setup` rule
(per-device batch = global batch // world size, RuntimeError if not divisible; :262-278) and `direction` semantics.
Output side: `denormalize_to_uint8` (src/infer_simple_flowmatching.py:37-38) on the device, `write_png` off the
critical path.  PNG *decode* stays on the host: nvJPEG/nvPNG are not in the image.
"""
from __future__ import annotations

import math
import os
import random
import struct
import threading
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2  # Pillow, src/libImaging/Resample.c


# ------------------------------------------------------------------------------------------------ Pillow resampling tables
def _bilinear_filter(x: float) -> float:
    x = -x if x < 0.0 else x
    return 1.0 - x if x < 1.0 else 0.0


def pillow_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Pillow's `precompute_coeffs` (bilinear, support 1.0) + `normalize_coeffs_8bpc` for a full-image box.
    -> (bounds int32 [out, 2] = (first, count), kk int32 [out, ksize])."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bilinear_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def resample_pass_numpy(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray, axis: int) -> np.ndarray:
    """Host statement of one `s2s_resample_u8` pass (uint8 [H,W,C]); used by the CPU tests against real Pillow."""
    src = np.moveaxis(img.astype(np.int64), axis, 0)
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for o in range(bounds.shape[0]):
        first, cnt = int(bounds[o, 0]), int(bounds[o, 1])
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[o, :cnt].astype(np.int64), src[first:first + cnt], axes=(0, 0))
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


_COEFF_CACHE: Dict[Tuple[int, int, str], Tuple[torch.Tensor, torch.Tensor]] = {}


def _coeffs_on(device, in_size: int, out_size: int):
    key = (in_size, out_size, str(device))
    hit = _COEFF_CACHE.get(key)
    if hit is None:
        b, k = pillow_coeffs(in_size, out_size)
        hit = _COEFF_CACHE[key] = (torch.from_numpy(b).to(device), torch.from_numpy(k).to(device))
    return hit


def resize_u8(x: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """`TF.resize(pil, (out_h, out_w))` (bilinear, antialiased) on a uint8 NHWC device batch: horizontal pass, then
    vertical pass, each rounded to uint8 like Pillow's two-pass ImagingResample.  Equal sizes skip the pass."""
    from . import kernels as K
    B, H, W, _ = x.shape
    if W != out_w:
        x = K.resample_u8(x, *_coeffs_on(x.device, W, out_w), vertical=False)
    if H != out_h:
        x = K.resample_u8(x, *_coeffs_on(x.device, H, out_h), vertical=True)
    return x


# ------------------------------------------------------------------------------------------------ augmentation parameters
def draw_augment_params(n: int, src_h: int, src_w: int, size: int) -> torch.Tensor:
    """(top, left, hflip, vflip) per sample, drawn with the SAME calls in the SAME order as the reference's
    `__getitem__` (paired_data_module.py:171-191): `transforms.RandomCrop.get_params` = two `torch.randint(...).item()`
    draws from the torch default generator (none if the image already has the crop size), then two `random.random() > 0.5`.
    -> int32 [n, 4] on the host."""
    out = torch.zeros((n, 4), dtype=torch.int32)
    for k in range(n):
        if src_h < size or src_w < size:
            raise ValueError(f"Required crop size {(size, size)} is larger than input image size {(src_h, src_w)}")
        if src_w == size and src_h == size:
            i = j = 0
        else:
            i = int(torch.randint(0, src_h - size + 1, size=(1,)).item())
            j = int(torch.randint(0, src_w - size + 1, size=(1,)).item())
        hf = random.random() > 0.5
        vf = random.random() > 0.5
        out[k] = torch.tensor([i, j, int(hf), int(vf)], dtype=torch.int32)
    return out


def prepare_tiles(src_u8: torch.Tensor, tgt_u8: Optional[torch.Tensor], size: int, use_augmentation: bool,
                  bgr: bool = False, params: Optional[torch.Tensor] = None, mask_u8: Optional[torch.Tensor] = None,
                  mask_raw: bool = False):
    """uint8 NHWC device batches -> fp32 NCHW normalised tiles (x0, x1[, mask]) exactly as the reference's dataset
    produces them: train = crop/flip, eval = antialiased resize (paired_data_module.py:171-210)."""
    from . import kernels as K
    B, H, W, _ = src_u8.shape
    dev = src_u8.device
    if use_augmentation:
        if params is None:
            params = draw_augment_params(B, H, W, size)
        params = params.to(dev, non_blocking=True)
    else:
        src_u8 = resize_u8(src_u8, size, size)
        tgt_u8 = None if tgt_u8 is None else resize_u8(tgt_u8, size, size)
        if mask_u8 is not None and tuple(mask_u8.shape[1:]) != (size, size):
            raise ValueError("eval path: resize the mask on the host first (nearest_resize_pil / nearest_resize_cv2): it "
                             "follows another interpolation rule than the images")
        params = torch.zeros((B, 4), dtype=torch.int32, device=dev)
    x0, x1, m = K.tile_prep(src_u8, tgt_u8, params, size, bgr=bgr, mask_u8=mask_u8, mask_raw=mask_raw)
    return (x0, x1) if mask_u8 is None else (x0, x1, m)


# ------------------------------------------------------------------------------------------------ nearest-neighbour mask resizing
def _pil_nearest_index(in_size: int, out_size: int) -> np.ndarray:
    """Pillow's ImagingScaleAffine: xo = a/2, then `xo += a` per output pixel (ACCUMULATED in double, which is what
    decides the ties of non-integer ratios), index = (int) xo."""
    a = float(in_size) / out_size
    idx = np.empty(out_size, dtype=np.int64)
    xo = a * 0.5
    for o in range(out_size):
        idx[o] = min(int(xo), in_size - 1)
        xo += a
    return idx


def nearest_resize_pil(mask: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """`TF.resize(pil_mask, (h, w), interpolation=NEAREST)` (src/data/paired_data_multiclassmask.py:121): Pillow samples the
    pixel under the centre of each output pixel."""
    h, w = mask.shape[:2]
    return mask[_pil_nearest_index(h, out_h)][:, _pil_nearest_index(w, out_w)]


def nearest_resize_cv2(mask: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """`cv2.resize(mask, (w, h), interpolation=cv2.INTER_NEAREST)` (src/data/paired_data_mask_he_amyloid.py:89): OpenCV
    samples at the output pixel's ORIGIN, index = min(floor(o * in / out), in - 1)."""
    h, w = mask.shape[:2]
    ys = np.minimum(np.floor(np.arange(out_h) * (h / out_h)).astype(np.int64), h - 1)
    xs = np.minimum(np.floor(np.arange(out_w) * (w / out_w)).astype(np.int64), w - 1)
    return mask[ys][:, xs]


# ------------------------------------------------------------------------------------------------ dataset / datamodule mirrors
class PairedDataset:
    """`src.data.paired_data_module.PairedDataset` with the decode kept on the host and everything else moved to the
    device: `__getitem__` returns the RAW uint8 BGR arrays (plus filenames if asked); `PairedDataModule` batches them
    and calls `prepare_tiles`.  `get_reference_item(idx)` reproduces the reference's per-item output through the same
    kernels (used by the parity tests and by `infer`-style one-tile loops)."""

    def __init__(self, data_dir, csv_file_name, source_column, target_column, folder, image_size=512, direction="S2T",
                 use_augmentation=False, return_filename=False):
        import pandas as pd
        self.source_dir = os.path.join(data_dir, folder)
        self.target_dir = os.path.join(data_dir, folder)
        self.image_size = image_size
        self.direction = direction
        self.source_column = source_column
        self.target_column = target_column
        self.return_filename = return_filename
        csv_path = os.path.join(data_dir, csv_file_name)
        assert os.path.exists(csv_path), 'csv not exists'
        self.metadata = pd.read_csv(csv_path)
        self.metadata = self.metadata[self.metadata['split'] == folder].reset_index(drop=True)
        self.use_augmentation = use_augmentation

    def __len__(self):
        return len(self.metadata)

    def filenames(self, idx) -> Tuple[str, str]:
        row = self.metadata.iloc[idx]
        return row[self.source_column], row[self.target_column]

    def __getitem__(self, idx):
        import cv2
        sf, tf = self.filenames(idx)
        sp, tp = os.path.join(self.source_dir, sf), os.path.join(self.target_dir, tf)
        assert os.path.exists(sp), f"Source image not found: {sp}"
        assert os.path.exists(tp), f"Target image not found: {tp}"
        return cv2.imread(sp), cv2.imread(tp), sf, tf  # BGR uint8 HWC; the kernel swaps channels

    def get_reference_item(self, idx, device="cuda"):
        s, t, sf, tf = self[idx]
        su = torch.from_numpy(s).unsqueeze(0).to(device)
        tu = torch.from_numpy(t).unsqueeze(0).to(device)
        x0, x1 = prepare_tiles(su, tu, self.image_size, self.use_augmentation, bgr=True)
        a, b = (x0[0], x1[0]) if self.direction == "S2T" else (x1[0], x0[0])
        if self.return_filename:
            return (a, b, sf, tf) if self.direction == "S2T" else (a, b, tf, sf)
        return a, b


class PairedMulticlassMaskDataset(PairedDataset):
    """`src.data.paired_data_multiclassmask.PairedDataset` (the multitask model's data, config M): tile pair + a
    class-id mask.  Train: the SAME crop / flips for all three (one kernel); eval: antialiased resize for the images,
    Pillow NEAREST for the mask (:119-123).  The mask comes back as fp32 [1, S, S] holding the raw class ids."""

    def __init__(self, data_dir, csv_file_name, source_column, target_column, folder, mask_column='graywhite_filepath',
                 image_size=512, direction="S2T", use_augmentation=False):
        super().__init__(data_dir, csv_file_name, source_column, target_column, folder, image_size, direction,
                         use_augmentation)
        self.mask_dir = os.path.join(data_dir, folder)
        self.mask_column = mask_column

    def __getitem__(self, idx):
        import cv2
        s, t, sf, tf = super().__getitem__(idx)
        mp = os.path.join(self.mask_dir, self.metadata.iloc[idx][self.mask_column])
        assert os.path.exists(mp), f"Mask image not found: {mp}"
        return s, t, cv2.imread(mp, cv2.IMREAD_GRAYSCALE)

    def get_reference_item(self, idx, device="cuda"):
        s, t, m = self[idx]
        S = self.image_size
        if not self.use_augmentation:
            m = nearest_resize_pil(m, S, S)
        su = torch.from_numpy(s).unsqueeze(0).to(device)
        tu = torch.from_numpy(t).unsqueeze(0).to(device)
        mu = torch.from_numpy(np.ascontiguousarray(m)).unsqueeze(0).to(device)
        x0, x1, mask = prepare_tiles(su, tu, S, self.use_augmentation, bgr=True, mask_u8=mu, mask_raw=True)
        return (x0[0], x1[0], mask[0]) if self.direction == "S2T" else (x1[0], x0[0], mask[0])


class PairedHEIHCMaskDataset(PairedDataset):
    """`src.data.paired_data_mask_he_amyloid.PairedHEIHCDataset` (the mask / ROI LitModules' data): resized tile pair +
    a binarised uint8 mask [1, S, S]: `cv2.resize(..., INTER_NEAREST)` then `> 1` (:88-91)."""

    def __init__(self, data_dir, csv_file_name, source_column, target_column, folder, image_size=512,
                 direction="HE_to_IHC"):
        super().__init__(data_dir, csv_file_name, source_column, target_column, folder, image_size, "S2T", False)
        self.he_ihc_direction = direction
        self.mask_dir = os.path.join(data_dir, folder)
        self.mask_column = 'amyloid_filepath'

    def __getitem__(self, idx):
        import cv2
        s, t, sf, tf = super().__getitem__(idx)
        mp = os.path.join(self.mask_dir, self.metadata.iloc[idx][self.mask_column])
        assert os.path.exists(mp), f"Mask image not found: {mp}"
        m = nearest_resize_cv2(cv2.imread(mp, cv2.IMREAD_GRAYSCALE), self.image_size, self.image_size)
        return s, t, np.where(m > 1, 1, 0).astype(np.uint8)

    def get_reference_item(self, idx, device="cuda"):
        s, t, m = self[idx]
        su = torch.from_numpy(s).unsqueeze(0).to(device)
        tu = torch.from_numpy(t).unsqueeze(0).to(device)
        x0, x1 = prepare_tiles(su, tu, self.image_size, False, bgr=True)
        mask = torch.from_numpy(m).unsqueeze(0).to(device)
        return (x0[0], x1[0], mask) if self.he_ihc_direction == "HE_to_IHC" else (x1[0], x0[0], mask)


class TileBatchLoader:
    """Iterates device-resident `(x0, x1)` batches: a thread pool decodes PNGs into a pinned uint8 staging buffer, the
    copy and the prep kernel run on a side stream one batch ahead of the consumer (double buffering)."""

    def __init__(self, dataset: PairedDataset, batch_size: int, shuffle: bool, num_workers: int = 4, device="cuda",
                 drop_last: bool = False, rank: int = 0, world_size: int = 1, seed: int = 0):
        self.ds, self.bs, self.shuffle, self.device = dataset, batch_size, shuffle, torch.device(device)
        self.drop_last, self.rank, self.world, self.seed, self.epoch = drop_last, rank, world_size, seed, 0
        self.pool = ThreadPoolExecutor(max_workers=max(1, num_workers))
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self._staging: List[Optional[Tuple[torch.Tensor, torch.Tensor]]] = [None, None]

    def __len__(self):
        n = len(self._indices())
        return n // self.bs if self.drop_last else (n + self.bs - 1) // self.bs

    def _indices(self) -> List[int]:
        idx = list(range(len(self.ds)))
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            idx = [idx[i] for i in torch.randperm(len(idx), generator=g).tolist()]
        return idx[self.rank::self.world]  # static partition across ranks, no collective

    def _decode(self, ids: Sequence[int], slot: int):
        items = list(self.pool.map(self.ds.__getitem__, ids))
        h, w = items[0][0].shape[:2]
        st = self._staging[slot]
        if st is None or st[0].shape[1:3] != (h, w) or st[0].shape[0] < len(ids):
            st = self._staging[slot] = tuple(torch.empty((self.bs, h, w, 3), dtype=torch.uint8).pin_memory() for _ in range(2))
        for k, (s, t, _, _) in enumerate(items):
            st[0][k].copy_(torch.from_numpy(s))
            st[1][k].copy_(torch.from_numpy(t))
        return st[0][:len(ids)], st[1][:len(ids)]

    def _stage(self, ids, slot):
        s_host, t_host = self._decode(ids, slot)
        params = draw_augment_params(len(ids), s_host.shape[1], s_host.shape[2], self.ds.image_size) \
            if self.ds.use_augmentation else None
        with torch.cuda.stream(self.stream):
            su = s_host.to(self.device, non_blocking=True)
            tu = t_host.to(self.device, non_blocking=True)
            x0, x1 = prepare_tiles(su, tu, self.ds.image_size, self.ds.use_augmentation, bgr=True, params=params)
            done = torch.cuda.Event()
            done.record(self.stream)
        return x0, x1, done

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        idx = self._indices()
        batches = [idx[i:i + self.bs] for i in range(0, len(idx), self.bs)]
        if self.drop_last and batches and len(batches[-1]) < self.bs:
            batches.pop()
        self.epoch += 1
        nxt = self._stage(batches[0], 0) if batches else None
        for k in range(len(batches)):
            x0, x1, done = nxt
            nxt = self._stage(batches[k + 1], (k + 1) & 1) if k + 1 < len(batches) else None
            torch.cuda.current_stream().wait_event(done)
            x0.record_stream(torch.cuda.current_stream())
            x1.record_stream(torch.cuda.current_stream())
            yield (x0, x1) if self.ds.direction == "S2T" else (x1, x0)


class PairedDataModule:
    """`src.data.paired_data_module.PairedDataModule`: same keywords; dataloaders yield device-resident batches."""

    def __init__(self, data_dir: str = "data/", csv_file_name: str = "dataset_nirschl_et_al_2026_metadata.csv",
                 source_column: str = 'he_filepath', target_column: str = 'ihc_filepath', batch_size: int = 8,
                 num_workers: int = 4, image_size: int = 512, direction: str = "S2T", pin_memory: bool = True,
                 use_augmentation: bool = False, device="cuda") -> None:
        self.hparams = dict(batch_size=batch_size, num_workers=num_workers, pin_memory=pin_memory)
        self.data_dir, self.csv_file_name = data_dir, csv_file_name
        self.source_column, self.target_column = source_column, target_column
        self.batch_size_per_device = batch_size
        self.num_workers, self.image_size, self.direction = num_workers, image_size, direction
        self.use_augmentation, self.device = use_augmentation, device
        self.trainer = None

    def prepare_data(self) -> None:
        pass

    def setup(self, stage: Optional[str] = None) -> None:
        # paired_data_module.py:273-278
        if self.trainer is not None:
            if self.hparams["batch_size"] % self.trainer.world_size != 0:
                raise RuntimeError(f"Batch size ({self.hparams['batch_size']}) is not divisible by the number of devices "
                                   f"({self.trainer.world_size}).")
            self.batch_size_per_device = self.hparams["batch_size"] // self.trainer.world_size

    def _loader(self, folder: str, shuffle: bool) -> TileBatchLoader:
        ds = PairedDataset(self.data_dir, self.csv_file_name, self.source_column, self.target_column, folder,
                           self.image_size, self.direction, self.use_augmentation)
        rank = getattr(self.trainer, "global_rank", 0) if self.trainer is not None else 0
        world = getattr(self.trainer, "world_size", 1) if self.trainer is not None else 1
        return TileBatchLoader(ds, self.batch_size_per_device, shuffle, self.num_workers, self.device, rank=rank,
                               world_size=world)

    def train_dataloader(self):
        self.data_train = self._loader("train", True)
        return self.data_train

    def val_dataloader(self):
        self.data_val = self._loader("val", True)  # the reference shuffles the validation loader too (:325)
        return self.data_val

    def test_dataloader(self):
        self.data_test = self._loader("test", False)
        return self.data_test

    def teardown(self, stage: Optional[str] = None) -> None:
        pass

    def state_dict(self) -> Dict[Any, Any]:
        return {}

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        pass


# ------------------------------------------------------------------------------------------------ output side
def denormalize_to_uint8(x: torch.Tensor) -> torch.Tensor:
    """fp32 NCHW tiles in [-1, 1] on the device -> uint8 NHWC (`denormalize` of src/infer_simple_flowmatching.py:37-38,
    then 8-bit quantisation) in one kernel."""
    from . import kernels as K
    if x.dim() == 3:
        x = x.unsqueeze(0)
    return K.denorm_u8(x.float().contiguous())


def write_png(path: str, img_hwc_u8: np.ndarray, level: int = 3) -> None:
    """Minimal PNG writer (8-bit RGB / grey, no interlace) so that results can be saved without matplotlib / PIL."""
    a = np.ascontiguousarray(img_hwc_u8)
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, c = a.shape
    assert a.dtype == np.uint8 and c in (1, 3)
    raw = np.concatenate([np.zeros((h, 1), dtype=np.uint8), a.reshape(h, w * c)], axis=1).tobytes()

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2 if c == 3 else 0, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, level)) + chunk(b"IEND", b""))


class AsyncPngWriter:
    """Saves uint8 tiles on a background thread (device -> pinned host copy is stream-ordered, the encode is not on the
    critical path of the sampler)."""

    def __init__(self, workers: int = 2):
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.pending = []
        self.lock = threading.Lock()

    def submit(self, paths: Sequence[str], tiles_u8: torch.Tensor):
        host = torch.empty(tiles_u8.shape, dtype=torch.uint8).pin_memory() if tiles_u8.is_cuda else tiles_u8
        if tiles_u8.is_cuda:
            host.copy_(tiles_u8, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        else:
            ev = None

        def work():
            if ev is not None:
                ev.synchronize()
            arr = host.numpy()
            for p, a in zip(paths, arr):
                write_png(p, a)
        with self.lock:
            self.pending.append(self.pool.submit(work))

    def wait(self):
        with self.lock:
            pend, self.pending = self.pending, []
        for f in pend:
            f.result()
