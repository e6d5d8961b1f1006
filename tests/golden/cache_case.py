"""Shared fixture of the cache-writer tests: a small map-style dataset in the shape of the reference's CSArrayDataset
(data/dataloader.py:363-435): uint8 HWC arrays -> PIL -> transform -> (image, label) or (image, label, metadata dict).
Used by tests/golden/make_golden_cache.py (with the REFERENCE transform and writers) and by the tests (with ours)."""
from __future__ import annotations

import numpy as np
import torch
from PIL import Image

N_IMAGES, SIDE, BATCH = 10, 80, 4
WORDS = ["bog", "fen", "heath", "scrub", "dune"]


def case_images():
    from aihab_clip_b200.weights import synthetic_images_u8
    return np.concatenate([synthetic_images_u8(N_IMAGES // 2, SIDE, seed=2024),
                           synthetic_images_u8(N_IMAGES - N_IMAGES // 2, SIDE, seed=2024, start=N_IMAGES // 2, smooth=True)])


def case_labels():
    return [(7 * i + 3) % 20 for i in range(N_IMAGES)]


class CacheCaseDataset(torch.utils.data.Dataset):
    def __init__(self, transform, with_metadata: bool, raw_u8: bool = False):
        self.u8, self.labels = case_images(), case_labels()
        self.transform, self.with_metadata, self.raw_u8 = transform, with_metadata, raw_u8

    def __len__(self):
        return len(self.labels)

    def __getitem__(self, i):
        img = torch.from_numpy(self.u8[i]) if self.raw_u8 else self.transform(Image.fromarray(self.u8[i]))
        if not self.with_metadata:
            return img, self.labels[i]
        meta = {"file_name": f"img_{i:03d}.jpg", "plot_word_label": WORDS[i % len(WORDS)], "l2_label": self.labels[i] % 11}
        return img, self.labels[i], meta


def case_loader(transform, with_metadata: bool, raw_u8: bool = False, shuffle_seed=None):
    ds = CacheCaseDataset(transform, with_metadata, raw_u8)
    gen = None
    if shuffle_seed is not None:
        gen = torch.Generator().manual_seed(shuffle_seed)
    return torch.utils.data.DataLoader(ds, batch_size=BATCH, shuffle=shuffle_seed is not None, generator=gen, num_workers=0)


CFG = {"root_path": None, "backbone": "ViT-B/16", "dataset": "cs", "shots": 4, "seed": 3, "aug_views": 2,
       "finetune": {"cache_embeddings_dir": "emb_cache", "cache_embeddings_normalize": True}}
