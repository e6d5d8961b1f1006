#!/usr/bin/env python
"""bench.py — headline benchmark of the aihab-clip hot path on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--arch ViT-B/16] [--batch 128]
                    [--dtype fp16|bf16]

Workload (BASELINE.json configs[1]): CLIP ViT-B/16 feature-cache extraction over synthetic 224 px uint8 images —
one STEP = one pass of the hot path over one batch of `--batch` images per GPU: uint8 -> normalise -> patch embed
-> 12 transformer blocks -> ln_post -> visual.proj -> L2-norm -> x100 logits vs the 20-class text head -> argmax.
Shards are independent (weak scaling); for N > 1 the timed region ends with the path's single all-gather of the
normalised features + predictions.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "images/sec CLIP ViT-B/16 encode_image+logits at 1/2/4/8 B200; % bf16 TC peak"
UNIT = "images/s"


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


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_text_head(sd_np, n_classes: int, embed_dim: int) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(7))
    w = rng.standard_normal((embed_dim, n_classes)).astype(np.float32)
    return w / np.linalg.norm(w, axis=0, keepdims=True)


class CpuReference:
    """The reference's CPU fp32 path for this workload: oracle/clip_oracle_torch.py, i.e. the torch operators the
    reference itself calls on CPU (PIL / torchvision preprocessing per image, conv2d, LayerNorm,
    multi_head_attention_forward, Linear, QuickGELU, normalize, argmax), on all host threads."""

    def __init__(self, geom, sd_np, text_w):
        import torch
        from oracle import clip_oracle_torch as OT
        # torchrun exports OMP_NUM_THREADS=1; the CPU arms are meant to use every host core
        torch.set_num_threads(os.cpu_count() or 1)
        self.cores = torch.get_num_threads()
        self.OT, self.R = OT, geom.image_resolution
        self.sd = OT.to_torch_state(sd_np)
        self.text_w = torch.as_tensor(np.asarray(text_w), dtype=torch.float32)

    def __call__(self, images_u8):
        return self.OT.reference_pass(self.sd, self.text_w, images_u8, self.R)


CPU_KIND_NOTE = ("torch-operator restatement of the reference CPU path (oracle/clip_oracle_torch.py; the reference "
                 "itself is Python under /root/reference and cannot travel to the GPU box)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from aihab_clip_b200.weights import GEOMETRIES, make_state_dict_np, synthetic_images_u8
    geom = GEOMETRIES[args.arch]
    sd = make_state_dict_np(geom, 0, with_text=False)
    text_w = cpu_text_head(sd, args.classes, geom.embed_dim)
    ref = CpuReference(geom, sd, text_w)
    cores = ref.cores
    n = args.ref_batch
    imgs = synthetic_images_u8(n, geom.image_resolution)
    for _ in range(args.warmup):
        ref(imgs[:max(1, n // 4)])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref(imgs)
    dt = time.perf_counter() - t0
    val = args.steps * n / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.arch, geom.image_resolution, args.classes), "arch": args.arch,
                       "classes": args.classes, "weights": "random-init (aihab_clip_b200.weights seed 0)",
                       "images_per_step": n, "operands": "fp32, " + CPU_KIND_NOTE},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n} images per step x {args.steps} steps (the reference's shipped extraction "
                                       f"batch is 16, methods/utils.py:142-173); {CPU_KIND_NOTE}"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from aihab_clip_b200 import _lib
    from aihab_clip_b200.clip.model import build_model
    from aihab_clip_b200.extraction import ShardedExtractor, ZeroShotHead, encode_and_score
    from aihab_clip_b200.weights import GEOMETRIES, make_state_dict, make_state_dict_np, synthetic_images_u8

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
    B, R, K, W = args.batch, geom.image_resolution, args.steps, args.warmup
    sd = make_state_dict(geom, 0)
    model = build_model(sd).to(dev).float()
    model.visual.compute_dtype = args.dtype
    model.visual.max_batch = B
    # text head from the model's own text tower (one-time, PyTorch) on the golden prompt tokens: 20 classes x 1
    # template as shipped (data/templates.py:204-226); falls back to a seeded unit-norm head if fixtures are absent
    gpath = REPO / "tests" / "golden" / "reference_outputs.npz"
    with torch.no_grad():
        if gpath.is_file() and args.classes == 20:
            tok = torch.from_numpy(np.load(gpath)["tok_tokens"][:20]).to(dev)
            _, te = model.encode_text(tok)
            te = te / te.norm(dim=-1, keepdim=True)
            text_w = te.t().contiguous().float()
            head_src = "model text tower on the 20 shipped class prompts (golden tokens)"
        else:
            text_w = torch.from_numpy(cpu_text_head(None, args.classes, geom.embed_dim)).to(dev)
            head_src = "seeded unit-norm random head"
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
    out_emb = torch.empty(K * B if K * B <= 1 << 18 else 1 << 18, E + 1, dtype=torch.float32, device=dev)

    def step(i):
        emb, _, idx = encode_and_score(model, batches[i % NB], head, 1)
        r0 = (i * B) % (out_emb.shape[0] - B + 1)
        out_emb[r0:r0 + B, :E] = emb
        out_emb[r0:r0 + B, E] = idx[:, 0].float()

    for i in range(max(W, 3)):
        step(i)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- timed region 1: device-resident inputs (value) ----------------
    gather_buf = torch.empty(world * B, E + 1, dtype=torch.float32, device=dev) if world > 1 else None

    def timed_region(per_launch_events: bool):
        """K steps (+ the single all-gather for N > 1) between CUDA events; max over ranks.  With
        per_launch_events the library also records an event pair around every kernel launch on the launch stream
        (roofline / share evidence); those records cost a few percent, so `value` comes from the clean pass."""
        _lib.profile_enable(per_launch_events)
        for c in _lib.PROFILE_CLASSES:
            _lib.profile_read(c, reset=True)
        barrier()
        n0 = _lib.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            step(i)
        if world > 1:  # the path's single collective: gather normalised features + predictions of the last block
            dist.all_gather_into_tensor(gather_buf, out_emb[:B])
        e1.record()
        barrier()
        t_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        n_launch = _lib.kernel_launches() - n0
        _lib.profile_enable(False)
        sites_ = _lib.profile_sites("gemm") if per_launch_events else []
        prof_ = {c: _lib.profile_read(c, reset=True) for c in _lib.PROFILE_CLASSES}
        prof_["gemm"]["sites"] = sites_
        return float(t_ms.item()), n_launch, prof_

    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches, _ = timed_region(False)
    ms_prof, _, prof = timed_region(not args.no_kernel_profile)
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
    ext.h2d_bytes = ext.d2h_bytes = 0
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    res = ext.run(source, n_local * world)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    e2e_value = world * n_local / (ms_e2e / 1e3)
    assert res["features"].shape[0] == n_local * world

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
                     (2.0 * B * geom.grid ** 2 * Dw * (-(-3 * pp * pp // 64) * 64), Dw): "patch_embed"}
            kernels["gemm"]["sites"] = {
                names.get((sr["work"], sr["tag"]), "N=%d work=%.4g" % (sr["tag"], sr["work"])): {
                    "launches": sr["launches"], "avg_ms": sr["ms"] / sr["launches"],
                    "tflops": sr["work"] * sr["launches"] / (sr["ms"] * 1e-3) / 1e12}
                for sr in g["sites"] if sr["ms"] > 0}
        total_tflops = value / world * flops_per_image(geom, args.classes) / 1e12
        # CPU baseline on a bounded sample (rank 0, N = 1 only)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sd_np = make_state_dict_np(geom, 0, with_text=False)
            ref = CpuReference(geom, sd_np, text_w.cpu().numpy())
            n_cpu = args.cpu_images
            imgs = synthetic_images_u8(n_cpu, R)
            ref(imgs[:2])
            t0 = time.perf_counter()
            reps = 0
            while reps < 1 or (time.perf_counter() - t0 < 10.0 and reps < 64):
                ref(imgs)
                reps += 1
            dt = time.perf_counter() - t0
            cpu = {"value": reps * n_cpu / dt, "unit": UNIT, "cores": ref.cores, "kind": "port",
                   "sample": f"{reps} x {n_cpu} images of the same workload ({dt:.1f} s); {CPU_KIND_NOTE}"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload_name(args.arch, R, args.classes), "arch": args.arch, "batch_per_gpu": B, "global_batch": B * world, "classes": args.classes,
                       "text_head": head_src, "weights": "random-init (aihab_clip_b200.weights seed 0)",
                       "operands": f"{args.dtype} tensor-core operands, fp32 accumulate / residual / LN / softmax / scoring",
                       "l2": f"inputs rotate over {NB} distinct batches ({NB * B * img_bytes / 2**20:.0f} MiB > 126 MiB L2); "
                             f"activation workspace {model.visual.engine(dev)._lib.aihab_vit_workspace_bytes(model.visual.engine(dev).handle) / 2**20:.0f} MiB",
                       "parallelism": f"dp{world} (independent shards, one all-gather)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": ext.h2d_bytes // max(1, K),
                    "d2h_bytes_per_step": ext.d2h_bytes // max(1, K), "ms_per_step": ms_e2e / K,
                    "api": "aihab_clip_b200.extraction.ShardedExtractor.run (pinned host uint8 in, pinned host rows out)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "aihab::gemm_kernel (tcgen05, all GEMM sites)",
                         "achieved": gemm_tflops, "peak": peaks["tensor"], "unit": "TFLOP/s",
                         "frac": gemm_tflops / peaks["tensor"], "traffic": traffic, "peak_source": peaks["source"],
                         "launches": g["launches"], "avg_launch_ms": g["ms"] / max(1, g["launches"]),
                         "share_of_step": g["ms"] / ms_prof,
                         "measured_in": "second timed region of the same K steps with a CUDA-event pair recorded on the "
                                        "launch stream around every kernel launch",
                         "ms_per_step_with_events": ms_prof / K},
            "whole_step": {"tflops": total_tflops, "frac_of_tensor_peak": total_tflops / peaks["tensor"],
                           "flops_per_image": flops_per_image(geom, args.classes)},
            "kernels": kernels,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
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
    ap.add_argument("--ref-batch", type=int, default=16, help="--impl reference: images per step (bounded sample)")
    ap.add_argument("--cpu-images", type=int, default=16, help="cpu_baseline sample size per repetition")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-profile", action="store_true",
                    help="do not record per-launch CUDA events in the timed region (roofline becomes 0)")
    args = ap.parse_args()
    return run_reference(args) if args.impl == "reference" else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
