"""world_size-2 (and 3) gloo tests of the N>1 path on CPU: shard assignment, packed [n_pad, E+1] rows, the single
all-gather and the invariant the GPU path relies on — the gathered G-rank result is bit-identical to the 1-rank
result.  The CUDA compute is replaced by a deterministic per-image function (step_fn)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aihab_clip_b200.extraction import ShardedExtractor, ZeroShotHead, array_source

E, C = 16, 5


def fake_step(images: torch.Tensor):
    """Per-image deterministic 'embedding': depends only on that image's pixels (batch-composition invariant)."""
    x = images.reshape(images.shape[0], -1).to(torch.float32)
    emb = torch.stack([x[:, i::E].sum(dim=1) for i in range(E)], dim=1)
    emb = emb / emb.norm(dim=1, keepdim=True)
    return emb, (x.sum(dim=1).to(torch.int64) % C).reshape(-1, 1)


def make_images(n):
    return torch.from_numpy(np.random.default_rng(5).integers(0, 256, (n, 4, 4, 3), dtype=np.uint8))


def run_single(n, batch):
    head = ZeroShotHead(torch.zeros(8, E), torch.zeros(E, C))
    ext = ShardedExtractor(None, head, batch_size=batch, device="cpu", rank=0, world_size=1, step_fn=fake_step)
    return ext.run(array_source(make_images(n), pin=False), n)


def _worker(rank, world, port, n, batch, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        head = ZeroShotHead(torch.zeros(8, E), torch.zeros(E, C))
        ext = ShardedExtractor(None, head, batch_size=batch, device="cpu", step_fn=fake_step)
        assert (ext.rank, ext.world) == (rank, world)
        res = ext.run(array_source(make_images(n), pin=False), n)
        torch.save({"features": res["features"].clone(), "preds": res["preds"].clone(), "range": res["range"]},
                   os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n,batch", [(2, 37, 8), (2, 16, 16), (3, 10, 4)])
def test_gathered_result_equals_single_rank(tmp_path, world, n, batch):
    mp.spawn(_worker, args=(world, _free_port(), n, batch, str(tmp_path)), nprocs=world, join=True)
    ref = run_single(n, batch)
    covered = []
    for r in range(world):
        got = torch.load(tmp_path / f"r{r}.pt")
        assert torch.equal(got["features"], ref["features"]), f"rank {r}: features differ from the 1-rank run"
        assert torch.equal(got["preds"], ref["preds"])
        covered += list(range(*got["range"]))
    assert covered == list(range(n))
    assert ref["features"].shape == (n, E) and ref["preds"].dtype == torch.int64


def test_extractor_requires_cuda_without_step_fn():
    head = ZeroShotHead(torch.zeros(8, E), torch.zeros(E, C))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ShardedExtractor(None, head, device="cpu")
