#!/bin/bash
# A/B helper: tools/ab.sh <ENVVAR> <batch>   -> kernel timings and bench value for ENVVAR=0 and ENVVAR=1
V=$1; B=${2:-128}
for pm in 0 1; do
  echo "== $V=$pm kernels (batch $B)"
  env $V=$pm timeout -s KILL 100 python tools/bench_kernels.py --batch $B 2>&1 | tr -d "\n " | cut -c1-260; echo
done
for pm in 0 1; do
  env $V=$pm timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --batch $B --no-cpu-baseline 2>/dev/null > /tmp/ab_$pm.json
  python - <<PY
import json
d=json.load(open("/tmp/ab_$pm.json"))
print("== $V=$pm bench: img/s", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "gemm TF", round(d["roofline"]["achieved"]), "sm_mhz", d["clocks"]["sm_mhz"])
PY
done
