# A/B runs of the fused MLP kernel on one box (results under gpurun_out/fused/)
set -x
mkdir -p gpurun_out/fused
AIHAB_CLIP_LIB=$PWD/ab/lib_debug_hang.so timeout 300 python -m pytest tests/test_gpu_model.py -x -q -k pipelined_mlp > gpurun_out/fused/test_dbg.log 2>&1
echo "dbg rc=$?" > gpurun_out/fused/summary.txt
tail -5 gpurun_out/fused/test_dbg.log
if grep -q "passed" gpurun_out/fused/test_dbg.log && ! grep -q failed gpurun_out/fused/test_dbg.log; then
  run() {  # name, env...
    name=$1; shift
    env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-other-configs > gpurun_out/fused/bench_$name.json 2> gpurun_out/fused/bench_$name.err
    echo "bench $name rc=$?" >> gpurun_out/fused/summary.txt
  }
  run base AIHAB_MLP_FUSED=0
  run fused AIHAB_MLP_FUSED=1 AIHAB_MLP_RING_PAIRS=40
  run nowait AIHAB_MLP_FUSED=1 AIHAB_MLP_RING_PAIRS=40 AIHAB_MLP_NOWAIT=1
  run lag5 AIHAB_MLP_FUSED=1 AIHAB_MLP_RING_PAIRS=48 AIHAB_MLP_LAG=5
  # 3 operand stages + 4-box rings: bash tools/build_variant.sh s3r4 -DAIHAB_FUSED_STAGES=3 -DAIHAB_FUSED_RING=4
  [ -f ab/lib_s3r4.so ] && run s3r4 AIHAB_MLP_FUSED=1 AIHAB_MLP_RING_PAIRS=40 AIHAB_CLIP_LIB=$PWD/ab/lib_s3r4.so
  run base2 AIHAB_MLP_FUSED=0
fi
cat gpurun_out/fused/summary.txt
