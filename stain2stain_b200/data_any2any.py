"""Any-to-any (class-conditional) input pipeline: mirror of `src/data/class_conditional_he_amyloid.py` (SURVEY 8f row
f1, "any2any variant") on the same device kernel as `data.py`.

Reference per item (`PairedAnyToAnyDataset.__getitem__`, :105-145): pick a source domain (`random.choice` unless fixed),
pick a target domain (`random.choice`), in `union` mode resample until both domains have the file, open both images
(`PIL ... .convert("RGB")`), draw ONE crop (`RandomCrop.get_params`: two `torch.randint(...).item()`) for the pair -- or one
per image with `same_crop_for_pair=False` --, `ToTensor`, `Normalize(0.5, 0.5)`; returns `(src, tgt, target_label)`.

Here the host keeps exactly those RNG calls in that order (so a seeded run selects the same domains and crops) and the
decode; crop + to_tensor + normalise run in `s2s_tile_prep` on the uint8 batch (bit-identical, tests/test_gpu_tiles.py).
`ClassConditionalAnyToAnyDataModule` keeps the reference's keywords, its reproducible `train_val_split.json`
(`prepare_data`, :190-240: `random.Random(split_seed).shuffle` over the first domain's sorted files) and the per-device
batch rule of `setup` (:282-288).
"""
from __future__ import annotations

import json
import os
import random
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Any, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .data import prepare_tiles


def _crop_params(h: int, w: int, size: int) -> Tuple[int, int]:
    """transforms.RandomCrop.get_params: no draw when the image already has the crop size."""
    if h < size or w < size:
        raise ValueError(f"Required crop size {(size, size)} is larger than input image size {(h, w)}")
    if w == size and h == size:
        return 0, 0
    i = int(torch.randint(0, h - size + 1, size=(1,)).item())
    j = int(torch.randint(0, w - size + 1, size=(1,)).item())
    return i, j


class _DomainTable:
    """Which tile names exist in which stain domain: one directory scan per domain, then name sets combined by
    `intersection` (tiles present in every domain) or `union` (present in at least one)."""

    def __init__(self, root: str, folders: Dict[int, str], exts: Sequence[str]):
        self.dirs: Dict[int, Path] = {label: Path(root) / sub for label, sub in folders.items()}
        exts = tuple(e.lower() for e in exts)
        self.names: Dict[int, frozenset] = {}
        for label, d in self.dirs.items():
            if not d.is_dir():
                raise ValueError(f"Folder not found: {d}")
            self.names[label] = frozenset(e.name for e in os.scandir(d) if e.name.lower().endswith(exts))

    def combined(self, mode: str) -> List[str]:
        if mode not in ("intersection", "union"):
            raise ValueError("filename_mode must be 'intersection' or 'union'")
        pools = list(self.names.values())
        if not pools:
            return []
        acc = pools[0]
        for pool in pools[1:]:
            acc = (acc & pool) if mode == "intersection" else (acc | pool)
        return sorted(acc)

    def has(self, label: int, name: str) -> bool:
        return name in self.names[label]


class PairedAnyToAnyDataset:
    """Mirror of the reference's any-to-any dataset (`src/data/class_conditional_he_amyloid.py:12-145`): same keywords,
    same item semantics, same consumption of the Python / torch RNG streams (so a seeded run picks the same domains and
    crops), with decode on the host and crop + to_tensor + normalise on the device."""

    _MAX_UNION_RETRIES = 50  # the reference gives up after 50 resamples in `union` mode (:121-131)

    def __init__(self, root_dir, class_folder_mapping, crop_size=256, transform=None, same_crop_for_pair=True,
                 source_domain_mode="random", filename_mode="intersection",
                 allowed_exts=(".png", ".jpg", ".jpeg", ".tif", ".tiff"), valid_filenames: Optional[List[str]] = None):
        if transform is not None:
            raise NotImplementedError("custom post-crop transforms run on the host; the B200 pipeline fuses the default one")
        self.root_dir, self.crop_size, self.same_crop_for_pair = root_dir, crop_size, same_crop_for_pair
        self.source_domain_mode, self.filename_mode = source_domain_mode, filename_mode
        self.class_folder_mapping = dict(class_folder_mapping)
        self.allowed_exts = tuple(allowed_exts)
        self.class_indices = sorted(self.class_folder_mapping)
        self.num_classes = len(self.class_indices)
        self._table = _DomainTable(root_dir, self.class_folder_mapping, self.allowed_exts)
        names = self._table.combined(filename_mode)
        if valid_filenames is not None:  # train / val split: keep the listed names only (order stays sorted)
            keep = set(valid_filenames)
            names = [n for n in names if n in keep]
        if not names:
            raise ValueError("No filenames found (check folders / extensions).")
        self.filenames = names
        # reference-named views of the same table (read by code written against the reference class)
        self.class_to_dir = {c: str(d) for c, d in self._table.dirs.items()}
        self.class_to_filenames = {c: set(v) for c, v in self._table.names.items()}

    def __len__(self):
        return len(self.filenames)

    def _load_rgb(self, class_idx, filename) -> np.ndarray:
        from PIL import Image
        return np.asarray(Image.open(self._table.dirs[class_idx] / filename).convert("RGB"))

    def _pick_source(self) -> int:
        mode = self.source_domain_mode
        if mode == "random":
            return random.choice(self.class_indices)
        if isinstance(mode, int):
            return mode
        raise ValueError("source_domain_mode must be 'random' or an int class index")

    def draw(self, idx) -> Tuple[str, int, int]:
        """(file name, source domain, target domain) of item `idx`.  RNG protocol of the reference's `__getitem__`
        (:109-131): one `random.choice` for the source (random mode only), one for the target; in `union` mode the pair is
        redrawn -- source first, then target -- until both domains hold the file, at most 50 times."""
        name = self.filenames[idx]
        src = self._pick_source()
        tgt = random.choice(self.class_indices)
        if self.filename_mode != "union":
            return name, src, tgt
        for _ in range(self._MAX_UNION_RETRIES + 1):
            if self._table.has(src, name) and self._table.has(tgt, name):
                return name, src, tgt
            src = self._pick_source()
            tgt = random.choice(self.class_indices)
        raise RuntimeError(f"Could not find paired file '{name}' across sampled domains. "
                           "Consider using intersection mode.")

    def __getitem__(self, idx):
        """-> (src uint8 HWC RGB, tgt uint8 HWC RGB, crop params int32 [2, 4] (src, tgt), target_label)."""
        fname, source_label, target_label = self.draw(idx)
        src = self._load_rgb(source_label, fname)
        tgt = self._load_rgb(target_label, fname)
        i, j = _crop_params(src.shape[0], src.shape[1], self.crop_size)
        if self.same_crop_for_pair:
            i2, j2 = i, j
        else:
            i2, j2 = _crop_params(tgt.shape[0], tgt.shape[1], self.crop_size)
        params = torch.tensor([[i, j, 0, 0], [i2, j2, 0, 0]], dtype=torch.int32)
        return src, tgt, params, target_label

    def get_reference_item(self, idx, device="cuda"):
        """What the reference's `__getitem__` returns, through the device kernel: (src, tgt, target_label)."""
        src, tgt, params, label = self[idx]
        su = torch.from_numpy(np.ascontiguousarray(src)).unsqueeze(0).to(device)
        tu = torch.from_numpy(np.ascontiguousarray(tgt)).unsqueeze(0).to(device)
        if self.same_crop_for_pair:
            x0, x1 = prepare_tiles(su, tu, self.crop_size, True, bgr=False, params=params[:1])
        else:
            x0, _ = prepare_tiles(su, None, self.crop_size, True, bgr=False, params=params[:1])
            x1, _ = prepare_tiles(tu, None, self.crop_size, True, bgr=False, params=params[1:])
        return x0[0], x1[0], label


class AnyToAnyBatchLoader:
    """Device-resident `(x0, x1, target_label)` batches: thread-pool decode, uint8 H2D, one prep kernel per batch."""

    def __init__(self, dataset: PairedAnyToAnyDataset, batch_size: int, shuffle: bool, num_workers: int = 4,
                 device="cuda", rank: int = 0, world_size: int = 1, seed: int = 0):
        self.ds, self.bs, self.shuffle, self.device = dataset, batch_size, shuffle, torch.device(device)
        self.rank, self.world, self.seed, self.epoch = rank, world_size, seed, 0
        self.pool = ThreadPoolExecutor(max_workers=max(1, num_workers))

    def _indices(self) -> List[int]:
        idx = list(range(len(self.ds)))
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            idx = [idx[i] for i in torch.randperm(len(idx), generator=g).tolist()]
        return idx[self.rank::self.world]

    def __len__(self):
        return (len(self._indices()) + self.bs - 1) // self.bs

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        idx = self._indices()
        self.epoch += 1
        for k in range(0, len(idx), self.bs):
            ids = idx[k:k + self.bs]
            draws = [self.ds.draw(i) for i in ids]  # RNG calls stay on this thread, in item order
            imgs = list(self.pool.map(lambda d: (self.ds._load_rgb(d[1], d[0]), self.ds._load_rgb(d[2], d[0])), draws))
            S = self.ds.crop_size
            p_src, p_tgt = [], []
            for s, t in imgs:
                i, j = _crop_params(s.shape[0], s.shape[1], S)
                p_src.append([i, j, 0, 0])
                p_tgt.append([i, j, 0, 0] if self.ds.same_crop_for_pair else [*_crop_params(t.shape[0], t.shape[1], S), 0, 0])
            su = torch.from_numpy(np.stack([s for s, _ in imgs])).pin_memory().to(self.device, non_blocking=True)
            tu = torch.from_numpy(np.stack([t for _, t in imgs])).pin_memory().to(self.device, non_blocking=True)
            ps = torch.tensor(p_src, dtype=torch.int32)
            if self.ds.same_crop_for_pair:
                x0, x1 = prepare_tiles(su, tu, S, True, bgr=False, params=ps)
            else:
                x0, _ = prepare_tiles(su, None, S, True, bgr=False, params=ps)
                x1, _ = prepare_tiles(tu, None, S, True, bgr=False, params=torch.tensor(p_tgt, dtype=torch.int32))
            yield x0, x1, torch.tensor([d[2] for d in draws], dtype=torch.int64).to(self.device, non_blocking=True)


class ClassConditionalAnyToAnyDataModule:
    def __init__(self, data_dir: str, class_folder_mapping: dict, crop_size: int = 256, same_crop_for_pair: bool = True,
                 batch_size: int = 32, num_workers: int = 4, pin_memory: bool = True, source_domain_mode: str = "random",
                 filename_mode: str = "intersection", allowed_exts: tuple = (".png", ".jpg", ".jpeg", ".tif", ".tiff"),
                 val_split: float = 0.2, split_seed: int = 42, device="cuda"):
        self.hparams = dict(batch_size=batch_size, num_workers=num_workers, pin_memory=pin_memory)
        self.data_dir, self.class_folder_mapping, self.crop_size = data_dir, class_folder_mapping, crop_size
        self.same_crop_for_pair, self.batch_size_per_device, self.num_workers = same_crop_for_pair, batch_size, num_workers
        self.source_domain_mode, self.filename_mode, self.allowed_exts = source_domain_mode, filename_mode, allowed_exts
        self.val_split, self.split_seed, self.device = val_split, split_seed, device
        self.train_filenames = self.val_filenames = None
        self.split_file = Path(data_dir) / "train_val_split.json"
        self.trainer = None

    def prepare_data(self) -> None:
        if self.split_file.exists():
            return
        first_class = list(self.class_folder_mapping.keys())[0]
        folder_path = os.path.join(self.data_dir, self.class_folder_mapping[first_class])
        if not os.path.isdir(folder_path):
            raise ValueError(f"Folder not found: {folder_path}")
        all_files = sorted(f for f in os.listdir(folder_path) if f.lower().endswith(self.allowed_exts))
        if len(all_files) == 0:
            raise ValueError(f"No files found in {folder_path}")
        rng = random.Random(self.split_seed)
        rng.shuffle(all_files)
        n_val = int(len(all_files) * self.val_split)
        val_files, train_files = all_files[:n_val], all_files[n_val:]
        with open(self.split_file, "w") as f:
            json.dump({"train": train_files, "val": val_files, "split_seed": self.split_seed, "val_split": self.val_split,
                       "total_files": len(all_files), "train_files": len(train_files), "val_files": len(val_files)}, f,
                      indent=2)

    def setup(self, stage: Optional[str] = None) -> None:
        if not self.split_file.exists():
            raise RuntimeError(f"Split file not found: {self.split_file}. Make sure prepare_data() was called.")
        with open(self.split_file) as f:
            split_data = json.load(f)
        self.train_filenames, self.val_filenames = split_data["train"], split_data["val"]
        kw = dict(root_dir=self.data_dir, class_folder_mapping=self.class_folder_mapping, crop_size=self.crop_size,
                  same_crop_for_pair=self.same_crop_for_pair, source_domain_mode=self.source_domain_mode,
                  filename_mode=self.filename_mode, allowed_exts=self.allowed_exts)
        self.data_train = PairedAnyToAnyDataset(valid_filenames=self.train_filenames, **kw)
        self.data_val = PairedAnyToAnyDataset(valid_filenames=self.val_filenames, **kw)
        if self.trainer is not None:
            if self.hparams["batch_size"] % self.trainer.world_size != 0:
                raise RuntimeError(f"Batch size ({self.hparams['batch_size']}) is not divisible by the number of devices "
                                   f"({self.trainer.world_size}).")
            self.batch_size_per_device = self.hparams["batch_size"] // self.trainer.world_size

    def _loader(self, ds, shuffle):
        rank = getattr(self.trainer, "global_rank", 0) if self.trainer is not None else 0
        world = getattr(self.trainer, "world_size", 1) if self.trainer is not None else 1
        return AnyToAnyBatchLoader(ds, self.batch_size_per_device, shuffle, self.num_workers, self.device, rank, world)

    def train_dataloader(self):
        return self._loader(self.data_train, True)

    def val_dataloader(self):
        return self._loader(self.data_val, False)

    def teardown(self, stage: Optional[str] = None) -> None:
        pass

    def state_dict(self) -> Dict[Any, Any]:
        return {}

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        pass
