#!/bin/bash
# compute-sanitizer over the small kernel / model tests (SURVEY.md §5 "race detection"), on a B200:
#   gpurun -- bash tools/sanitize.sh            -> gpurun_out/sanitize/{memcheck,racecheck,debug_hang}.log + summary.txt
# 1. the parity tests of the tiny geometries and of the GEMM / attention / scoring kernels under --tool memcheck
# 2. the same under --tool racecheck (shared-memory hazards; mbarrier / TMA traffic is not tracked by the tool)
# 3. the same tests, no sanitizer, against the -DAIHAB_DEBUG_HANG build (bounded mbarrier waits that print the barrier
#    and trap instead of hanging; built here by `make DEBUG_HANG=1`)
# Each leg runs under its own timeout so that a hang cannot take the box down.
cd "$(dirname "$0")/.."
OUT=gpurun_out/sanitize; mkdir -p $OUT
TESTS="tests/test_gpu_kernels.py tests/test_gpu_model.py::test_encode_image_matches_reference_golden[tiny16] tests/test_gpu_model.py::test_encode_image_matches_reference_golden[tiny14] tests/test_gpu_model.py::test_batch_composition_invariance_and_chunking tests/test_gpu_cache_writers.py::test_l2_normalize_kernel"
KSEL=${SANITIZE_K:-"not 577 and not 1000 and not 4096 and not large and not score16"}
for tool in memcheck racecheck; do
  timeout -s KILL ${SANITIZE_TIMEOUT:-900} compute-sanitizer --tool $tool --error-exitcode 99 --launch-timeout 120 \
    python -m pytest $TESTS -m gpu -x -q -k "$KSEL" > $OUT/$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $OUT/summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT/$tool.log | tail -3 | tee -a $OUT/summary.txt
done
if [ -f ab/lib_debug_hang.so ]; then
  AIHAB_CLIP_LIB=$PWD/ab/lib_debug_hang.so timeout -s KILL 600 python -m pytest $TESTS -m gpu -x -q > $OUT/debug_hang.log 2>&1
  echo "debug_hang rc=$?" | tee -a $OUT/summary.txt
  tail -1 $OUT/debug_hang.log | tee -a $OUT/summary.txt
fi
