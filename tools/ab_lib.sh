#!/bin/bash
# same-box A/B of two builds of libaihab_clip.so: tools/ab_lib.sh ab/lib_old.so ab/lib_new.so [rounds]
A=$1; B=$2; R=${3:-3}
for r in $(seq 1 $R); do
  for L in $A $B; do
    AIHAB_CLIP_LIB=$PWD/$L timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null > /tmp/ab.json
    python - <<PY
import json
d=json.load(open("/tmp/ab.json"))
k=d["kernels"]
print("$L", "img/s", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "sm_mhz", d["clocks"]["sm_mhz"], "e2e", round(d["e2e"]["value"]),
      "| gemm TF", round(k["gemm"]["tflops"]), "share", round(k["gemm"]["share_of_step"],3), "attn TF", round(k["attention"]["tflops"]), "ms_prof", round(d["roofline"]["ms_per_step_with_events"],3))
PY
  done
done
