"""Sharded feature extraction + zero-shot scoring: the loop the reference runs per batch on one device
(methods/utils.py:142-189, aihab_utils/feature_cache.py:114-142) re-designed for 1/2/4/8 B200s of one box.

* one process per GPU (torch.distributed, NCCL); rank r owns the contiguous block
  ``[r * ceil(N/G), min(N, (r+1) * ceil(N/G)))`` of global image indices, so gathered output order equals the
  single-GPU order and per-image results are bit-identical for every G (no BatchNorm, LayerNorm per token,
  attention per image — SURVEY.md §8e);
* host -> device copies run on a side stream from pinned buffers, double buffered against compute;
* results accumulate in HBM (no per-batch ``.to('cpu')`` sync as in methods/utils.py:164);
* ONE collective at the end: a single ``all_gather_into_tensor`` of ``[n_pad, E + 1]`` fp32 rows (the normalised
  embedding and, in the last column, the argmax class index, exact in fp32 for C < 2**24).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np
import torch

from . import ops


@dataclass
class ZeroShotHead:
    """``visual.proj`` [D, E] and prompt-ensembled class text embeddings [E, C] (utils.py:31-57), fp32 on device."""
    proj: torch.Tensor
    text_weights: torch.Tensor
    scale: float = 100.0  # the reference uses the literal 100., not logit_scale.exp() (methods/utils.py:185)

    @classmethod
    def from_model(cls, model, text_weights: torch.Tensor, device=None) -> "ZeroShotHead":
        device = device or model.visual.proj.device
        return cls(model.visual.proj.detach().to(device=device, dtype=torch.float32).contiguous(),
                   text_weights.detach().to(device=device, dtype=torch.float32).contiguous())


def shard_range(n_total: int, rank: int, world_size: int):
    per = -(-n_total // world_size)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per), per


def encode_and_score(model, images: torch.Tensor, head: ZeroShotHead, k: int = 1):
    """One pass of the hot path over one device-resident batch: uint8 HWC (preprocessing fused) or preprocessed
    float NCHW -> encode_image -> proj -> normalise -> 100 * emb @ W -> top-k.  Returns (emb, logits, topk_idx)."""
    if images.dtype == torch.uint8:
        feats = model.visual.forward_u8(images, torch.float32)
    else:
        feats = model.visual(images.float() if images.dtype == torch.float64 else images).float()
    emb, logits, idx, _ = ops.score(feats, head.proj, head.text_weights, head.scale, k)
    return emb, logits, idx


class ShardedExtractor:
    """Runs ``encode_and_score`` over this rank's shard of an image source and gathers the result.

    ``source(lo, hi)`` returns images ``[hi-lo, ...]`` for GLOBAL indices lo..hi-1 as a CPU tensor (uint8 HWC or
    float NCHW; pinned memory is used as is) or as a CUDA tensor already on this rank's device.
    """

    def __init__(self, model, head: ZeroShotHead, batch_size: int = 128, device: Optional[torch.device] = None,
                 rank: Optional[int] = None, world_size: Optional[int] = None, copy_results_to_host: bool = False,
                 step_fn: Optional[Callable] = None):
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.rank = rank if rank is not None else (self.dist.get_rank() if self.dist else 0)
        self.world = world_size if world_size is not None else (self.dist.get_world_size() if self.dist else 1)
        self.model, self.head, self.batch = model, head, int(batch_size)
        self.device = torch.device(device) if device is not None else model.visual.proj.device
        self.copy_results_to_host = copy_results_to_host
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        # step_fn(images) -> (emb [n,E] fp32, idx [n,1] int64) replaces the CUDA path; it exists so that the
        # sharding / packing / gather logic can be exercised by world_size-2 gloo tests on a CPU-only box
        self._step_fn = step_fn
        if self.device.type != "cuda":
            if step_fn is None:
                raise RuntimeError("ShardedExtractor needs the model on a CUDA device (no CPU fallback)")
            self._copy_stream = self._out_stream = None
        else:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._out_stream = torch.cuda.Stream(self.device)

    def _step(self, images):
        if self._step_fn is not None:
            return self._step_fn(images)
        emb, _, idx = encode_and_score(self.model, images, self.head, 1)
        return emb, idx

    def _gather(self, packed, per, n_total):
        full = torch.empty(per * self.world, packed.shape[1], dtype=torch.float32, device=packed.device)
        self.dist.all_gather_into_tensor(full, packed)  # the ONLY collective of the path (NCCL over NVLink)
        return full[:n_total]

    def _run_host(self, source, n_total, gather):
        """Synchronous CPU variant of run() used with step_fn (tests)."""
        lo, hi, per = shard_range(n_total, self.rank, self.world)
        E = self.head.proj.shape[1]
        packed = torch.zeros(per, E + 1, dtype=torch.float32)
        for b0 in range(lo, hi, self.batch):
            emb, idx = self._step(source(b0, min(hi, b0 + self.batch)))
            packed[b0 - lo:b0 - lo + emb.shape[0], :E] = emb
            packed[b0 - lo:b0 - lo + emb.shape[0], E] = idx[:, 0].to(torch.float32)
        full = self._gather(packed, per, n_total) if (gather and self.world > 1) else packed[:hi - lo]
        return {"features": full[:, :E], "preds": full[:, E].to(torch.int64), "host_copy": None, "range": (lo, hi)}

    def run(self, source: Callable[[int, int], torch.Tensor], n_total: int, gather: bool = True):
        """Returns dict(features [N,E] fp32 normalised, preds [N] int64) — gathered over all ranks when ``gather``
        (identical on every rank), else this rank's shard only."""
        if self.device.type != "cuda":
            return self._run_host(source, n_total, gather)
        lo, hi, per = shard_range(n_total, self.rank, self.world)
        E = self.head.proj.shape[1]
        dev = self.device
        packed = torch.zeros(per, E + 1, dtype=torch.float32, device=dev)  # rows beyond the shard stay zero (padding)
        compute = torch.cuda.current_stream(dev)
        starts = list(range(lo, hi, self.batch))
        host_out = None
        if self.copy_results_to_host:  # pinned result buffer is cached: cudaHostAlloc is slow and must not sit in the loop
            cached = getattr(self, "_host_out", None)
            if cached is None or cached.shape[0] < per or cached.shape[1] != E + 1:  # grow-only cache
                self._host_out = cached = torch.empty(per, E + 1, dtype=torch.float32, pin_memory=True)
            host_out = cached[:per]

        def stage(i):
            b0 = starts[i]
            b1 = min(hi, b0 + self.batch)
            src = source(b0, b1)
            if src.is_cuda:
                return src, None
            with torch.cuda.stream(self._copy_stream):
                d = src.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self.h2d_bytes += src.numel() * src.element_size()
            return d, ev

        nxt = stage(0) if starts else None
        for i, b0 in enumerate(starts):
            cur, ev = nxt
            nxt = stage(i + 1) if i + 1 < len(starts) else None  # copy of batch i+1 overlaps compute of batch i
            if ev is not None:
                compute.wait_event(ev)
            emb, idx = self._step(cur)
            cur.record_stream(compute)
            n = emb.shape[0]
            row0 = b0 - lo
            packed[row0:row0 + n, :E] = emb
            packed[row0:row0 + n, E] = idx[:, 0].to(torch.float32)
            if host_out is not None:  # reference-style per-batch D2H (methods/utils.py:164), but asynchronous
                done = torch.cuda.Event()
                done.record(compute)
                with torch.cuda.stream(self._out_stream):
                    self._out_stream.wait_event(done)
                    host_out[row0:row0 + n].copy_(packed[row0:row0 + n], non_blocking=True)
                self.d2h_bytes += n * (E + 1) * 4
        if host_out is not None:
            compute.wait_stream(self._out_stream)
        if gather and self.world > 1:
            if self.dist is None:
                raise RuntimeError("world_size > 1 needs an initialised torch.distributed process group")
            full = self._gather(packed, per, n_total)
        else:
            full = packed[:hi - lo]
        return {"features": full[:, :E], "preds": full[:, E].to(torch.int64), "host_copy": host_out,
                "range": (lo, hi)}


def array_source(images, pin: bool = True) -> Callable[[int, int], torch.Tensor]:
    """Image source over an in-memory array (numpy or torch, uint8 HWC or float NCHW) indexed by global image id."""
    t = torch.from_numpy(images) if isinstance(images, np.ndarray) else images
    if pin and not t.is_cuda and torch.cuda.is_available():
        t = t.pin_memory()
    return lambda lo, hi: t[lo:hi]
