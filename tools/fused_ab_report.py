"""Prints the per-site times of the bench lines tools/run_fused_ab.sh left under gpurun_out/fused/."""
import glob
import json
import os
import sys

for f in sorted(glob.glob((sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/fused") + "/bench_*.json"), key=os.path.getmtime):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable", e)
        continue
    sites = d["kernels"]["gemm"]["sites"]
    row = " ".join("%s=%.4f" % (k.split(" ")[0], v["avg_ms"]) for k, v in sites.items() if k != "patch_embed")
    att = d["kernels"]["attention"]
    print("%-10s %8.0f img/s %7.3f ms  clk %6.1f  attn %.4f  %s" % (
        os.path.basename(f)[6:-5], d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"],
        att["share_of_step"] * d["roofline"]["ms_per_step_with_events"] / 12, row))
