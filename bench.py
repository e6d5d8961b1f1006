#!/usr/bin/env python
"""bench.py — headline benchmark of the aihab-clip hot path on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--arch ViT-B/16] [--batch 256]
                    [--dtype fp16|bf16]

Workload (BASELINE.json configs[1]): CLIP ViT-B/16 feature-cache extraction over synthetic 224 px uint8 images —
one STEP = one pass of the hot path over one batch of `--batch` images per GPU: uint8 -> normalise -> patch embed
-> 12 transformer blocks -> ln_post -> visual.proj -> L2-norm -> x100 logits vs the 20-class text head -> argmax.
Shards are independent (weak scaling); for N > 1 the timed region ends with the path's single all-gather of the
normalised features + predictions of ALL its steps.  The K-step block is repeated until >= 2 s have been measured and
the MEDIAN block is reported (the step runs at the 1000 W power cap, so one 0.2 s block depends on the clock history).
One JSON line is printed by rank 0; besides the contract keys it carries
  config2_strong   BASELINE.json configs[1] literally: 100 000 images strong-scaled over the N ranks, full all-gather
  other_configs    one compact entry per other BASELINE config (ViT-B/32, ViT-L/14, ViT-L/14@336px, 1 M-row scoring)
  gather_check     N > 1: rank 0 recomputes a foreign shard block locally and compares it bit for bit with the gathered rows
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "images/sec CLIP ViT-B/16 encode_image+logits at 1/2/4/8 B200; % bf16 TC peak"
UNIT = "images/s"
TEXT_HEAD_20 = ("20 shipped classes x 1 shipped template through the model's own text tower "
                "(utils.py:31-57 clip_classifier; data/templates.py:204-226)")
TEXT_HEAD_RANDOM = "seeded unit-norm random head (PCG64 seed 7)"
WEIGHTS = "random-init (aihab_clip_b200.weights seed 0)"


def flops_per_image(geom, n_classes: int) -> float:
    """SURVEY.md §8d / BASELINE.md §3 algorithmic FLOPs (2*MAC) per image."""
    g2, L, D, p = geom.grid ** 2, geom.tokens, geom.vision_width, geom.vision_patch_size
    layers, E = geom.vision_layers, geom.embed_dim
    return (2.0 * g2 * 3 * p * p * D + layers * L * (2.0 * D * 3 * D + 2.0 * D * D + 4.0 * D * 4 * D)
            + layers * 4.0 * L * L * D + 2.0 * D * E + 2.0 * E * n_classes)


def measured_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.is_file():
        d = json.loads(p.read_text())
        return {"tensor": float(d["bf16_tflops_sustained"]), "tensor_burst": float(d["bf16_tflops"]),
                "hbm": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json, sustained bf16 cuBLAS / copy)"}
    return {"tensor": 1400.0, "tensor_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v == "Active":
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def workload_name(arch: str, R: int, classes: int) -> str:
    return (f"CLIP {arch} feature_cache extraction: uint8 {R}px -> encode_image -> proj -> L2-norm -> x100 logits "
            f"({classes} classes) -> argmax")


def make_config(args, geom, world: int) -> dict:
    """`config` of the JSON line — IDENTICAL for the b200 and the reference arm (arm-specific facts live in `arm`)."""
    return {"workload": workload_name(args.arch, geom.image_resolution, args.classes), "arch": args.arch,
            "batch_per_gpu": args.batch, "global_batch": args.batch * world, "classes": args.classes,
            "text_head": TEXT_HEAD_20 if args.classes == 20 else TEXT_HEAD_RANDOM, "weights": WEIGHTS,
            "parallelism": f"dp{world} (independent shards, one all-gather)"}


# ------------------------------------------------------------------------------------------------ CPU arms
def random_text_head(n_classes: int, embed_dim: int) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(7))
    w = rng.standard_normal((embed_dim, n_classes)).astype(np.float32)
    return w / np.linalg.norm(w, axis=0, keepdims=True)


class RealReference:
    """The UNMODIFIED reference on the host cores: its files travel in oracle/_ref (oracle/build_ref.py).
    clip.load(path, 'cpu') (clip/clip.py:89-137) -> build_clip_transforms(is_train=False) on PIL images
    (data/clip_transforms.py:50-56) -> encode_image -> proj -> F.normalize -> 100 * f @ W -> argmax
    (methods/utils.py:181-187), text head from utils.clip_classifier (utils.py:31-57)."""
    kind = "reference"
    note = "the unmodified reference (oracle/_ref: clip.load -> build_clip_transforms -> encode_image -> proj/normalise/logits)"

    def __init__(self, geom, classes: int):
        import torch
        from oracle import build_ref
        from aihab_clip_b200.weights import make_state_dict
        clip = build_ref.import_ref()
        import utils as ref_utils
        from data.clip_transforms import build_clip_transforms
        from data.templates import CS_CLASSNAMES, CS_TEMPLATES
        torch.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core
        self.cores = torch.get_num_threads()
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            torch.save(make_state_dict(geom, 0), f.name)
            state, self.model, _ = clip.load(f.name, device="cpu")
        self.tf = build_clip_transforms({}, is_train=False, resolution=geom.image_resolution)
        self.proj = state["visual.proj"].float()
        if classes == 20:
            import contextlib
            import io
            with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                _, _, self.text_w = ref_utils.clip_classifier(CS_CLASSNAMES, CS_TEMPLATES, self.model)
        else:
            self.text_w = torch.from_numpy(random_text_head(classes, geom.embed_dim))
        self.torch = torch

    def __call__(self, images_u8):
        from PIL import Image
        torch = self.torch
        x = torch.stack([self.tf(Image.fromarray(im)) for im in images_u8])
        with torch.no_grad():
            f = self.model.encode_image(x)
            e = torch.nn.functional.normalize(f @ self.proj, dim=-1)
            logits = 100. * e @ self.text_w
        return logits


class PortReference:
    """Fallback when oracle/_ref was not built: oracle/clip_oracle_torch.py, the torch operators the reference itself
    calls on CPU, flat over a state_dict."""
    kind = "port"
    note = "torch-operator restatement of the reference CPU path (oracle/clip_oracle_torch.py); oracle/_ref not built"

    def __init__(self, geom, classes: int):
        import torch
        from oracle import clip_oracle_torch as OT
        from aihab_clip_b200.weights import make_state_dict_np
        torch.set_num_threads(os.cpu_count() or 1)
        self.cores = torch.get_num_threads()
        self.OT, self.R = OT, geom.image_resolution
        self.sd = OT.to_torch_state(make_state_dict_np(geom, 0, with_text=False))
        self.text_w = torch.from_numpy(random_text_head(classes, geom.embed_dim))

    def __call__(self, images_u8):
        return self.OT.reference_pass(self.sd, self.text_w, images_u8, self.R)[1]  # (emb, logits, idx) -> logits


def cpu_reference(geom, classes):
    from oracle import build_ref
    return RealReference(geom, classes) if build_ref.available() else PortReference(geom, classes)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from aihab_clip_b200.weights import GEOMETRIES, synthetic_images_u8
    world = int(os.environ.get("WORLD_SIZE", "1"))
    geom = GEOMETRIES[args.arch]
    ref = cpu_reference(geom, args.classes)
    n = args.ref_batch if args.ref_batch > 0 else args.batch
    imgs = synthetic_images_u8(n, geom.image_resolution)
    for _ in range(args.warmup):
        ref(imgs[:max(1, n // 8)])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref(imgs)
    dt = time.perf_counter() - t0
    val = args.steps * n / dt
    sample = (f"{n} images per step x {args.steps} steps, all {ref.cores} host threads, fp32; warm-up steps use {max(1, n // 8)} images; "
              f"{ref.note}")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(args, geom, world),
            "arm": {"images_per_step": n, "operands": "fp32 on the host CPU", "implementation": ref.note},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": ref.cores, "kind": ref.kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ B200 arm
def other_configs(dev, dtype: str, peaks: dict, gpu_index: int, quick: bool) -> dict:
    """One compact entry per other BASELINE.json config, measured on this GPU with device-resident uint8 inputs
    (rank 0, N = 1): value, ms per step, roofline fraction, clocks during the measurement."""
    import torch
    from aihab_clip_b200 import ops
    from aihab_clip_b200.clip.model import build_model
    from aihab_clip_b200.extraction import ZeroShotHead, encode_and_score
    from aihab_clip_b200.weights import GEOMETRIES, make_state_dict

    def timed(fn, min_s=1.0, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(gpu_index)
        n, total = 0, 0.0
        while total < min_s * 1e3 and n < 200:
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            total += a.elapsed_time(b)
            n += 1
        return total / n, n, sampler.stop()

    out = {}
    # (key, arch, images per step, per-call chunk bound, classes)
    cases = [("config1_vitb32_batch64", "ViT-B/32", 64, 64, 18), ("config1_vitb32_batch757", "ViT-B/32", 757, 1024, 18),
             ("config3_vitl14_batch512", "ViT-L/14", 512, 128, 20), ("config4_vitl14_336_batch128", "ViT-L/14@336px", 128, 48, 20)]
    for key, arch, n_img, bound, classes in cases:
        geom = GEOMETRIES[arch]
        model = build_model(make_state_dict(geom, 0, with_text=False) | _text_stub(geom)).to(dev).float()
        model.visual.compute_dtype = dtype
        model.visual.max_batch = min(n_img, model.visual.preferred_batch(dev, bound))
        tw = torch.nn.functional.normalize(torch.randn(classes, geom.embed_dim, device=dev), dim=1).t().contiguous()
        head = ZeroShotHead.from_model(model, tw, dev)
        R = geom.image_resolution
        imgs = [torch.randint(0, 256, (n_img, R, R, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
        i = [0]

        def step():
            encode_and_score(model, imgs[i[0] % 2], head, 1)
            i[0] += 1
        ms, n, clocks = timed(step, 0.5 if quick else 1.5)
        ips = n_img / ms * 1e3
        tf = ips * flops_per_image(geom, classes) / 1e12
        out[key] = {"value": ips, "unit": "images/s", "ms_per_step": ms, "images_per_step": n_img,
                    "chunk": model.visual.max_batch, "tflops": tf, "frac_of_tensor_peak": tf / peaks["tensor"],
                    "steps_timed": n, "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons")}
        del model, head, imgs
        torch.cuda.empty_cache()
    # config 5: ProLIP / linear-probe scoring over 1 M cached fp16 ViT-B/16 features vs 1000 classes, top-5
    n = 250_000 if quick else 1_000_000
    g = torch.Generator(device=dev).manual_seed(11)
    feats = torch.randn(n, 768, device=dev, generator=g).half()
    proj = (torch.randn(768, 512, device=dev, generator=g) * 768 ** -0.5).half()
    tw = torch.nn.functional.normalize(torch.randn(1000, 512, device=dev, generator=g), dim=1).t().contiguous()
    ms, k, clocks = timed(lambda: ops.score16(feats, proj, tw, 100.0, 5), 0.5)
    fl = 2.0 * n * (768 * 512 + 512 * 1000)
    out["config5_scoring_1M_x_1000"] = {"value": n / ms * 1e3, "unit": "rows/s", "ms_per_step": ms, "rows": n,
                                        "tflops_algorithmic": fl / ms / 1e9, "frac_of_tensor_peak": fl / ms / 1e9 / peaks["tensor"],
                                        "steps_timed": k, "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons"),
                                        "path": "aihab_score16 (tcgen05, exact 16-bit products + fp16 hi/lo split logits, top-5)"}
    # preprocess with a real resize (configs/cs.yaml:23: 439 px arrays -> 224), standalone kernel, HBM roofline
    S, R, nb = 439, 224, 256
    u8 = [torch.randint(0, 256, (nb, S, S, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    j = [0]

    def pre():
        ops.preprocess_u8(u8[j[0] % 2], R, torch.float16)
        j[0] += 1
    ms, k, _ = timed(pre, 0.3)
    by = nb * (S * S * 3 + R * R * 3 * 2)
    out["preprocess_resize_439_to_224"] = {"value": nb / ms * 1e3, "unit": "images/s", "ms_per_step": ms, "gbs": by / ms / 1e6,
                                           "frac_of_hbm_peak": by / ms / 1e6 / peaks["hbm"],
                                           "algorithmic_bytes_per_image": S * S * 3 + R * R * 3 * 2}
    return out


def _text_stub(geom):
    """Smallest text tower the state_dict loader accepts (the image-side configs do not use it)."""
    import torch
    w = geom.transformer_width
    sd = {"positional_embedding": torch.zeros(geom.context_length, w), "text_projection": torch.zeros(w, geom.embed_dim),
          "logit_scale": torch.tensor(2.6593), "token_embedding.weight": torch.zeros(8, w),
          "ln_final.weight": torch.ones(w), "ln_final.bias": torch.zeros(w)}
    p = "transformer.resblocks.0."
    sd.update({p + "ln_1.weight": torch.ones(w), p + "ln_1.bias": torch.zeros(w), p + "ln_2.weight": torch.ones(w),
               p + "ln_2.bias": torch.zeros(w), p + "attn.in_proj_weight": torch.zeros(3 * w, w),
               p + "attn.in_proj_bias": torch.zeros(3 * w), p + "attn.out_proj.weight": torch.zeros(w, w),
               p + "attn.out_proj.bias": torch.zeros(w), p + "mlp.c_fc.weight": torch.zeros(4 * w, w),
               p + "mlp.c_fc.bias": torch.zeros(4 * w), p + "mlp.c_proj.weight": torch.zeros(w, 4 * w),
               p + "mlp.c_proj.bias": torch.zeros(w)})
    return sd


def run_b200(args):
    import torch
    import torch.distributed as dist
    from aihab_clip_b200 import _lib
    from aihab_clip_b200.clip.model import build_model
    from aihab_clip_b200.extraction import ShardedExtractor, ZeroShotHead, encode_and_score
    from aihab_clip_b200.weights import GEOMETRIES, make_state_dict, synthetic_images_u8

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: whatever NCCL logs (version banner at NCCL_DEBUG >= VERSION, INFO
        # lines when the caller asks for them) goes to stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        # ... and the "NCCL version" banner, which is written to fd 1 at communicator creation regardless of
        # NCCL_DEBUG: fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    if args.gpus != world and rank == 0:
        print(f"[bench] --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    geom = GEOMETRIES[args.arch]
    B, R, K, W = args.batch, geom.image_resolution, args.steps, max(args.warmup, 3)
    sd = make_state_dict(geom, 0)
    model = build_model(sd).to(dev).float()
    model.visual.compute_dtype = args.dtype
    model.visual.max_batch = B
    # text head from the model's own text tower (one-time, PyTorch) on the golden prompt tokens: 20 classes x 1
    # template as shipped (data/templates.py:204-226) — the same head utils.clip_classifier builds in the reference arm
    gpath = REPO / "tests" / "golden" / "reference_outputs.npz"
    with torch.no_grad():
        if args.classes == 20:
            tok = torch.from_numpy(np.load(gpath)["tok_tokens"][:20]).to(dev)
            _, te = model.encode_text(tok)
            te = te / te.norm(dim=-1, keepdim=True)
            text_w = te.t().contiguous().float()
        else:
            text_w = torch.from_numpy(random_text_head(args.classes, geom.embed_dim)).to(dev)
    head = ZeroShotHead.from_model(model, text_w, dev)

    # synthetic uint8 inputs resident in HBM: NB distinct batches keyed by global image index; NB * B * R*R*3 > L2
    img_bytes = R * R * 3
    NB = max(2, -(-160 * 2 ** 20 // (B * img_bytes)))
    gen = torch.Generator(device=dev)
    batches = []
    for j in range(NB):
        gen.manual_seed(1234 + rank * 100003 + j)
        batches.append(torch.randint(0, 256, (B, R, R, 3), dtype=torch.uint8, device=dev, generator=gen))
    E = geom.embed_dim
    out_emb = torch.empty(K * B, E + 1, dtype=torch.float32, device=dev)

    def step(i):
        emb, _, idx = encode_and_score(model, batches[i % NB], head, 1)
        r0 = (i % K) * B
        out_emb[r0:r0 + B, :E] = emb
        out_emb[r0:r0 + B, E] = idx[:, 0].float()

    for i in range(W):
        step(i)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- timed region 1: device-resident inputs (value) ----------------
    gather_buf = torch.empty(world * K * B, E + 1, dtype=torch.float32, device=dev) if world > 1 else None

    def timed_block(per_launch_events: bool):
        """EXACTLY K steps (+ the path's single all-gather of all K*B rows for N > 1) between CUDA events, barrier +
        synchronize on both sides; max over ranks.  With per_launch_events the library also records an event pair
        around every kernel launch on the launch stream (roofline / share evidence); those records cost a few percent,
        so `value` comes from the clean blocks."""
        _lib.profile_enable(per_launch_events)
        for c in _lib.PROFILE_CLASSES:
            _lib.profile_read(c, reset=True)
        barrier()
        n0 = _lib.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            step(i)
        if world > 1:  # the path's single collective: normalised features + predictions of every step of the block
            dist.all_gather_into_tensor(gather_buf, out_emb)
        e1.record()
        barrier()
        t_ms = allmax(e0.elapsed_time(e1))
        n_launch = _lib.kernel_launches() - n0
        _lib.profile_enable(False)
        sites_ = _lib.profile_sites("gemm") if per_launch_events else []
        prof_ = {c: _lib.profile_read(c, reset=True) for c in _lib.PROFILE_CLASSES}
        prof_["gemm"]["sites"] = sites_
        return t_ms, n_launch, prof_

    sampler = ClockSampler(local) if rank == 0 else None
    first, launches, _ = timed_block(False)
    n_blocks = int(min(64, max(3, -(-args.min_seconds * 1e3 // max(first, 1e-3)))))   # identical on every rank (allmax'ed time)
    block_ms = [first] + [timed_block(False)[0] for _ in range(n_blocks - 1)]
    ms = statistics.median(block_ms)
    ms_prof, _, prof = timed_block(not args.no_kernel_profile)
    clocks = sampler.stop() if sampler else None
    value = world * K * B / (ms / 1e3)

    # ---------------- timed region 2: end to end from pinned HOST buffers through the public extractor ----------
    pool = torch.from_numpy(synthetic_images_u8(min(4, NB) * B, R, seed=1234, start=rank * 1000003)).pin_memory()
    n_local = K * B

    def source(lo, hi):  # global index -> pinned host rows (pool is cycled; every step copies B fresh rows H2D)
        s = (lo - rank * n_local) % (pool.shape[0] - B + 1)
        return pool[s:s + (hi - lo)]

    ext = ShardedExtractor(model, head, batch_size=B, device=dev, rank=rank, world_size=world,
                           copy_results_to_host=True)
    ext.run(source, n_local * world)  # warm-up of the copy pipeline at the timed size (pinned result buffer, allocator)
    e2e_ms = []
    res = None
    for _ in range(max(3, min(9, n_blocks))):
        ext.h2d_bytes = ext.d2h_bytes = 0
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        res = ext.run(source, n_local * world)
        f1.record()
        barrier()
        e2e_ms.append(allmax(f0.elapsed_time(f1)))
    ms_e2e = statistics.median(e2e_ms)
    e2e_value = world * n_local / (ms_e2e / 1e3)
    assert res["features"].shape[0] == n_local * world

    # ---------------- hardware 1-vs-G bit identity: rank 0 recomputes a block of rank 1's shard ----------
    gather_check = None
    if world > 1:
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        if rank == 0:
            foreign = torch.from_numpy(synthetic_images_u8(B, R, seed=1234, start=1 * 1000003)).to(dev)  # rank 1, step 0
            emb1, _, idx1 = encode_and_score(model, foreign, head, 1)
            same = torch.equal(emb1, res["features"][n_local:n_local + B]) and \
                torch.equal(idx1[:, 0], res["preds"][n_local:n_local + B])
            ok[0] = 1 if same else 0
        dist.broadcast(ok, src=0)
        gather_check = bool(ok.item())

    # ---------------- BASELINE.json configs[1] literally: 100 000 images, strong-scaled over the ranks ----------------
    strong = None
    if args.images > 0:
        n_total = args.images
        dev_pool = torch.cat(batches[:min(NB, 4)], dim=0)

        def dev_source(lo, hi):  # device-resident synthetic images, cycled
            s = lo % (dev_pool.shape[0] - B + 1)
            return dev_pool[s:s + (hi - lo)]
        ext2 = ShardedExtractor(model, head, batch_size=B, device=dev, rank=rank, world_size=world)
        ext2.run(dev_source, min(n_total, 4 * B * world))
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        res2 = ext2.run(dev_source, n_total)
        g1.record()
        barrier()
        total_ms = allmax(g0.elapsed_time(g1))
        comp = torch.tensor([ext2.last_compute_ms], dtype=torch.float64, device=dev)
        comp_all = [torch.zeros_like(comp) for _ in range(world)]
        if world > 1:
            dist.all_gather(comp_all, comp)
        else:
            comp_all = [comp]
        comp_ms = [float(c.item()) for c in comp_all]
        strong = {"images": n_total, "value": n_total / (total_ms / 1e3), "unit": UNIT, "ms_total": total_ms,
                  "images_per_rank": -(-n_total // world), "scaling": "strong",
                  "gather_ms": allmax(ext2.last_gather_ms), "gather_bytes_total": int(res2["features"].shape[0]) * (E + 1) * 4,
                  "rank_compute_ms_min": min(comp_ms), "rank_compute_ms_max": max(comp_ms),
                  "tail_imbalance": (max(comp_ms) - min(comp_ms)) / max(comp_ms)}
        del res2

    if rank == 0:
        peaks = measured_peaks()
        g = prof["gemm"]
        gemm_tflops = g["work"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
        traffic = None
        tp = REPO / "profiles" / "roofline_traffic.json"
        if tp.is_file():
            traffic = json.loads(tp.read_text()).get("gemm_dram_bytes_per_launch")
        kernels = {}
        for c, r in prof.items():
            if r["ms"] <= 0:
                continue
            rate = r["work"] / (r["ms"] * 1e-3)
            if c in ("gemm", "attention", "score"):
                kernels[c] = {"share_of_step": r["ms"] / ms_prof, "launches": r["launches"], "tflops": rate / 1e12,
                              "frac_of_tensor_peak": rate / 1e12 / peaks["tensor"]}
            else:
                kernels[c] = {"share_of_step": r["ms"] / ms_prof, "launches": r["launches"], "gbs": rate / 1e9,
                              "frac_of_hbm_peak": rate / 1e9 / peaks["hbm"]}
        if g.get("sites"):
            # one entry per GEMM shape of the step (work per launch = 2*M*N*K identifies the site)
            M_tok, Dw, pp = B * geom.tokens, geom.vision_width, geom.vision_patch_size
            names = {(2.0 * M_tok * 3 * Dw * Dw, 3 * Dw): "qkv", (2.0 * M_tok * Dw * Dw, Dw): "out_proj",
                     (2.0 * M_tok * 4 * Dw * Dw, 4 * Dw): "c_fc", (2.0 * M_tok * 4 * Dw * Dw, Dw): "c_proj",
                     (2.0 * B * geom.grid ** 2 * Dw * (-(-3 * pp * pp // 64) * 64), Dw): "patch_embed",
                     # opt-in experiments (AIHAB_MLP_PIPE / AIHAB_MLP_FUSED): both MLP GEMMs timed as one site
                     (2.0 * M_tok * 4 * Dw * Dw * 2.0, -4 * Dw): "c_fc+c_proj"}
            kernels["gemm"]["sites"] = {
                names.get((sr["work"], sr["tag"]), "N=%d work=%.4g" % (sr["tag"], sr["work"])): {
                    "launches": sr["launches"], "avg_ms": sr["ms"] / sr["launches"],
                    "tflops": sr["work"] * sr["launches"] / (sr["ms"] * 1e-3) / 1e12}
                for sr in g["sites"] if sr["ms"] > 0}
        total_tflops = value / world * flops_per_image(geom, args.classes) / 1e12
        ws_mib = model.visual.engine(dev)._lib.aihab_vit_workspace_bytes(model.visual.engine(dev).handle) / 2 ** 20
        # CPU baseline on a bounded sample (rank 0, N = 1 only) + a parity spot check of the GPU path against it
        cpu, spot = None, None
        if world == 1 and not args.no_cpu_baseline:
            ref = cpu_reference(geom, args.classes)
            n_cpu = args.cpu_images
            imgs = synthetic_images_u8(n_cpu, R)
            ref_logits = ref(imgs)
            t0 = time.perf_counter()
            reps = 0
            while reps < 1 or (time.perf_counter() - t0 < 10.0 and reps < 64):
                ref(imgs)
                reps += 1
            dt = time.perf_counter() - t0
            cpu = {"value": reps * n_cpu / dt, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
                   "sample": f"{reps} x {n_cpu} images of the same workload ({dt:.1f} s), fp32, all host threads; {ref.note}"}
            if ref.kind == "reference" and args.classes == 20:
                _, lg, idx = encode_and_score(model, torch.from_numpy(imgs).to(dev), head, 1)
                rl = ref_logits.numpy() if hasattr(ref_logits, "numpy") else np.asarray(ref_logits)
                spot = {"images": n_cpu, "max_abs_dlogit": float(np.abs(lg.cpu().numpy() - rl).max()),
                        "argmax_agree": float((idx[:, 0].cpu().numpy() == rl.argmax(1)).mean()),
                        "against": "the unmodified reference's logits for the same images (oracle/_ref, CPU fp32)"}
        others = None
        if world == 1 and not args.no_other_configs:
            del batches
            torch.cuda.empty_cache()
            others = other_configs(dev, args.dtype, peaks, local, args.quick)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": make_config(args, geom, world),
            "arm": {"operands": f"{args.dtype} tensor-core operands, fp32 accumulate / residual / LN / softmax / scoring",
                    "l2": f"inputs rotate over {NB} distinct batches ({NB * B * img_bytes / 2**20:.0f} MiB > 126 MiB L2); "
                          f"activation workspace {ws_mib:.0f} MiB",
                    "timed_blocks": {"blocks_of_K_steps": len(block_ms), "reported": "median block", "block_ms": [round(x, 3) for x in block_ms],
                                     "min_seconds": args.min_seconds},
                    "collective": (f"one all_gather_into_tensor of {K * B} x {E + 1} fp32 rows per rank "
                                   f"({K * B * (E + 1) * 4 / 2**20:.1f} MiB) inside every timed block" if world > 1 else "none (N = 1)")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": ext.h2d_bytes // max(1, K),
                    "d2h_bytes_per_step": ext.d2h_bytes // max(1, K), "ms_per_step": ms_e2e / K,
                    "runs": len(e2e_ms), "reported": "median run",
                    "api": "aihab_clip_b200.extraction.ShardedExtractor.run (pinned host uint8 in, pinned host rows out)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "aihab::gemm_kernel (tcgen05, all GEMM sites)",
                         "achieved": gemm_tflops, "peak": peaks["tensor"], "unit": "TFLOP/s",
                         "frac": gemm_tflops / peaks["tensor"], "traffic": traffic, "peak_source": peaks["source"],
                         "launches": g["launches"], "avg_launch_ms": g["ms"] / max(1, g["launches"]),
                         "share_of_step": g["ms"] / ms_prof,
                         "measured_in": "one more block of the same K steps with a CUDA-event pair recorded on the "
                                        "launch stream around every kernel launch",
                         "ms_per_step_with_events": ms_prof / K},
            "whole_step": {"tflops": total_tflops, "frac_of_tensor_peak": total_tflops / peaks["tensor"],
                           "flops_per_image": flops_per_image(geom, args.classes)},
            "kernels": kernels,
            "cpu_baseline": cpu,
        }
        if spot is not None:
            line["parity_spot_check"] = spot
        if gather_check is not None:
            line["gather_check"] = gather_check
        if strong is not None:
            line["config2_strong"] = strong
        if others is not None:
            line["other_configs"] = others
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if gather_check is False:
        raise SystemExit("gather_check failed: the gathered rows of a foreign shard differ from a local recomputation")
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arch", default="ViT-B/16")
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--classes", type=int, default=20)
    ap.add_argument("--images", type=int, default=100000,
                    help="config2_strong: total images strong-scaled over the ranks (0 = skip)")
    ap.add_argument("--min-seconds", type=float, default=2.0,
                    help="repeat the K-step timed block until this much time has been measured; report the median block")
    ap.add_argument("--ref-batch", type=int, default=0,
                    help="--impl reference: images per step (0 = --batch, the GPU arm's per-step batch)")
    ap.add_argument("--cpu-images", type=int, default=16, help="cpu_baseline sample size per repetition")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--quick", action="store_true", help="shorter other_configs measurements")
    ap.add_argument("--no-kernel-profile", action="store_true",
                    help="do not record per-launch CUDA events in the timed region (roofline becomes 0)")
    args = ap.parse_args()
    return run_reference(args) if args.impl == "reference" else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
