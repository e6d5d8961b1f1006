"""Energy per unit of work of the hot kernels on one B200 (NVML total-energy counter around ~2.5 s loops).

The step runs at the 1000 W software power cap (bench.py `clocks`), so its time is total energy / cap: a kernel that
needs fewer cycles but the same joules does not make the step faster, one that needs fewer joules does.  This probe
prints, per kernel: achieved rate, mean power, SM clock, and joules per TFLOP (or per GB) - next to cuBLAS on the same
shape as the yardstick.   python tools/energy_probe.py [--batch 256] [--seconds 2.5]
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aihab_clip_b200 import _lib, ops  # noqa: E402

import pynvml  # noqa: E402


def measure(fn, seconds, h):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    # calibrate iterations for ~`seconds`
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        fn()
    e.record()
    torch.cuda.synchronize()
    per = s.elapsed_time(e) / 20
    iters = max(20, int(seconds * 1e3 / per))
    clocks = []
    e0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    t0 = time.perf_counter()
    s.record()
    done = 0
    while done < iters:
        for _ in range(min(200, iters - done)):
            fn()
        done += min(200, iters - done)
        clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
    e.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    ms = s.elapsed_time(e) / iters
    joules = (e1 - e0) / 1e3
    clocks.sort()
    return {"ms": ms, "watts": joules / (t1 - t0), "joules_per_iter": joules / iters, "sm_mhz": clocks[len(clocks) // 2],
            "iters": iters}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--seconds", type=float, default=2.5)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    dev = torch.device("cuda:0")
    dt = torch.float16
    L, D = 197, 768
    M, H = args.batch * L, D // 64
    out = {}

    def report(name, r, flops=None, bytes_=None):
        if flops:
            r["tflops"] = flops / r["ms"] / 1e9
            r["joules_per_tflop"] = r["joules_per_iter"] / (flops / 1e12)
        if bytes_:
            r["gbs"] = bytes_ / r["ms"] / 1e6
            r["joules_per_gb"] = r["joules_per_iter"] / (bytes_ / 1e9)
        out[name] = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}
        print(name, json.dumps(out[name]), flush=True)

    want = set(args.only.split(",")) if args.only else None

    def on(name):
        return want is None or name in want

    if on("idle"):
        report("idle", measure(lambda: time.sleep(0.01), 1.0, h))
    if on("cublas_8192"):
        a = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)
        b = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)
        report("cublas_8192", measure(lambda: torch.matmul(a, b), args.seconds, h), flops=2 * 8192 ** 3)
        del a, b
    shapes = {"qkv": (M, 3 * D, D, _lib.EPI_BIAS_16), "out_proj": (M, D, D, _lib.EPI_BIAS_RES_32),
              "fc_gelu": (M, 4 * D, D, _lib.EPI_BIAS_GELU_16), "mlp_proj": (M, D, 4 * D, _lib.EPI_BIAS_RES_32)}
    for name, (m, n, k, epi) in shapes.items():
        if not on(name):
            continue
        a = torch.randn(m, k, device=dev).to(dt)
        w = (torch.randn(n, k, device=dev) * k ** -0.5).to(dt)
        bias = torch.randn(n, device=dev)
        o16 = torch.empty(m, n, device=dev, dtype=dt) if epi in (_lib.EPI_BIAS_16, _lib.EPI_BIAS_GELU_16) else None
        o32 = torch.zeros(m, n, device=dev) if o16 is None else None
        report(name, measure(lambda: ops.gemm16(a, w, epi, bias=bias, out16=o16, out32=o32), args.seconds, h), flops=2.0 * m * n * k)
        if name in ("qkv", "fc_gelu", "mlp_proj"):
            wt = w.t().contiguous()
            report("cublas_" + name, measure(lambda: torch.matmul(a, wt), args.seconds, h), flops=2.0 * m * n * k)
    if on("attention"):
        qkv = torch.randn(M, 3 * D, device=dev).to(dt)
        report("attention", measure(lambda: ops.attention(qkv, args.batch, L, H), args.seconds, h),
               flops=4.0 * args.batch * L * L * D)
    if on("layernorm"):
        x = torch.randn(M, D, device=dev)
        g = torch.ones(D, device=dev)
        report("layernorm", measure(lambda: ops.layernorm(x, g, g, dt), args.seconds, h), bytes_=M * D * 6.0)
    if on("copy"):
        x = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        y = torch.empty_like(x)
        report("copy_1GiB", measure(lambda: y.copy_(x), args.seconds, h), bytes_=2.0 * (1 << 30))
    if on("step"):
        import numpy as np
        from aihab_clip_b200.clip.model import build_model
        from aihab_clip_b200.extraction import ZeroShotHead, encode_and_score
        from aihab_clip_b200.weights import GEOMETRIES, make_state_dict
        geom = GEOMETRIES["ViT-B/16"]
        model = build_model(make_state_dict(geom, 0)).to(dev).float()
        model.visual.max_batch = args.batch
        tw = torch.nn.functional.normalize(torch.randn(512, 20, device=dev), dim=0)
        head = ZeroShotHead.from_model(model, tw, dev)
        imgs = [torch.randint(0, 256, (args.batch, 224, 224, 3), dtype=torch.uint8, device=dev) for _ in range(5)]
        i = [0]

        def step():
            encode_and_score(model, imgs[i[0] % 5], head, 1)
            i[0] += 1
        r = measure(step, max(args.seconds, 4.0), h)
        report("step_vitb16", r, flops=args.batch * 35126927360.0)
    (Path("gpurun_out") / "energy_probe.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
