"""world_size-2 / 3 gloo tests (CPU) of the SHARDED cache writers: aihab_clip_b200.feature_cache routed through
extraction.extract_loader — batches of a (shuffling) DataLoader split by image batch across ranks, rows gathered in ONE
all_gather_into_tensor, files written by rank 0.  The CUDA encoder is replaced by a deterministic per-image function;
what is checked is that a G-rank run writes byte-identical tensors / the same CSV as the 1-rank run, in loader order."""
import json
import os
import socket
import sys
from pathlib import Path

import pandas as pd
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
import cache_case as CC  # noqa: E402

from aihab_clip_b200 import feature_cache as FC  # noqa: E402
from aihab_clip_b200.extraction import LoaderShards, extract_loader  # noqa: E402

W = 24


def fake_encode(images: torch.Tensor) -> torch.Tensor:
    """Deterministic per-image 'features' in fp16 (the reference caches features in the model dtype)."""
    x = images.reshape(images.shape[0], -1).to(torch.float32)
    return torch.stack([x[:, i::W].mean(dim=1) for i in range(W)], dim=1).to(torch.float16)


def fake_normalize(f):
    return torch.nn.functional.normalize(f.float(), dim=-1).to(f.dtype)


def to_tensor(img):
    import numpy as np
    return torch.from_numpy(np.asarray(img, dtype=np.float32).transpose(2, 0, 1) / 255.0)


def run_writers(root, shuffle_seed):
    cfg = dict(CC.CFG, root_path=str(root))
    FC.cache_preprojection_features(cfg, {"clip_model": torch.nn.Identity()}, CC.case_loader(to_tensor, False, shuffle_seed=shuffle_seed),
                                    {"train_size": CC.N_IMAGES}, _encode_fn=fake_encode)
    FC.cache_openclip_embeddings(cfg, torch.nn.Identity(), CC.case_loader(to_tensor, True, shuffle_seed=shuffle_seed),
                                 split="Test", checkpoint_path="c.pt", _encode_fn=fake_encode, _normalize_fn=fake_normalize)
    # iter mode: a plain list of batches instead of a DataLoader
    batches = list(CC.case_loader(to_tensor, True))
    x, y, rows = extract_loader(None, batches, encode_fn=fake_encode, to_cpu=True, want_metadata=True, device="cpu")
    return x, y, rows


def _worker(rank, world, port, root, shuffle_seed):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, y, rows = run_writers(root, shuffle_seed)
        torch.save({"x": x, "y": y, "rows": rows}, os.path.join(root, f"iter_r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,shuffle_seed", [(2, None), (2, 11), (3, 5)])
def test_sharded_writers_equal_single_rank(tmp_path, world, shuffle_seed):
    multi, single = tmp_path / "multi", tmp_path / "single"
    multi.mkdir()
    single.mkdir()
    mp.spawn(_worker, args=(world, _free_port(), str(multi), shuffle_seed), nprocs=world, join=True)
    x1, y1, rows1 = run_writers(single, shuffle_seed)
    fdir = "features_ViTB16_cs/4_shot/seed3"
    for name in ("f0.pth", "f1.pth", "label.pth"):
        a = torch.load(multi / fdir / name, weights_only=True)
        b = torch.load(single / fdir / name, weights_only=True)
        assert a.dtype == b.dtype and torch.equal(a, b), name
    f0 = torch.load(single / fdir / "f0.pth", weights_only=True)
    lab = torch.load(single / fdir / "label.pth", weights_only=True)
    assert f0.dtype == torch.float16 and tuple(f0.shape) == (CC.N_IMAGES, W) and lab.dtype == torch.int64
    if shuffle_seed is None:  # loader order = dataset order
        assert lab.tolist() == CC.case_labels()
    else:
        assert sorted(lab.tolist()) == sorted(CC.case_labels()) and lab.tolist() != CC.case_labels()
    edir = "emb_cache/ViTB16_cs/test/seed3"
    for name in ("embeddings.pt", "labels.pt"):
        assert torch.equal(torch.load(multi / edir / name, weights_only=True), torch.load(single / edir / name, weights_only=True))
    assert (multi / edir / "metadata.csv").read_text() == (single / edir / "metadata.csv").read_text()
    df = pd.read_csv(multi / edir / "metadata.csv")
    assert list(df.columns) == ["file_name", "ground_truth_num_label", "ground_truth_word_label", "ground_truth_L2_num_label"]
    labs = torch.load(multi / edir / "labels.pt", weights_only=True).tolist()
    assert df["ground_truth_num_label"].tolist() == labs                       # rows stay aligned with the tensors
    assert all(CC.case_labels()[int(n[4:7])] == v for n, v in zip(df["file_name"], labs))
    info = json.loads((multi / edir / "meta.json").read_text())
    assert info["num_samples"] == CC.N_IMAGES and info["dim"] == W and info["normalized"] is True
    emb = torch.load(multi / edir / "embeddings.pt", weights_only=True).float()
    assert torch.allclose(emb.norm(dim=-1), torch.ones(CC.N_IMAGES), atol=2e-3)
    # iter mode (round-robin ownership): identical on every rank and to the 1-rank pass
    for r in range(world):
        got = torch.load(multi / f"iter_r{r}.pt", weights_only=False)
        assert torch.equal(got["x"], x1) and torch.equal(got["y"], y1) and got["rows"] == rows1


def test_loader_shards_cover_every_batch_once():
    class FakeDist:
        def __init__(self, batches):
            self.batches = batches

        def broadcast_object_list(self, box, src=0):
            box[0] = self.batches
    loader = CC.case_loader(to_tensor, False)
    idx = [list(map(int, b)) for b in loader.batch_sampler]
    for world in (2, 3, 4, 8):
        seen = []
        for r in range(world):
            sh = LoaderShards(loader, r, world, FakeDist(idx))
            assert sh.mode == "index" and sh.sizes == [len(b) for b in idx]
            seen += [b for b, o in enumerate(sh.owner) if o == r]
            per = max(sh.local_rows(q) for q in range(world))
            gi = sh.gather_index(per)
            assert gi.numel() == CC.N_IMAGES and gi.unique().numel() == CC.N_IMAGES
        assert sorted(seen) == list(range(len(idx)))


def test_extract_loader_requires_cuda_without_encode_fn():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        extract_loader(torch.nn.Linear(2, 2), [], device="cpu")
