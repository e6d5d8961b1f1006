#!/bin/bash
# per-variant kernel timings + step time: tools/probe_run.sh ab/lib_a.so ab/lib_b.so ...
for L in "$@"; do
  echo "== $L"
  AIHAB_CLIP_LIB=$PWD/$L timeout -s KILL 100 python tools/bench_kernels.py --batch 256 2>&1 | tr -d "\n " | cut -c1-330; echo
  AIHAB_CLIP_LIB=$PWD/$L timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-kernel-profile 2>/dev/null > /tmp/ab.json
  python - <<PY
import json
d=json.load(open("/tmp/ab.json"))
print("   bench img/s", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "sm_mhz", d["clocks"]["sm_mhz"], "e2e", round(d["e2e"]["value"]))
PY
done
