# tools/run_ab.sh <out-subdir> <name>=<lib or -> [ENV=VAL ...] -- ...   (one bench per group, groups separated by --)
# Same-box A/B of bench.py with different libraries / environments; results under gpurun_out/<out-subdir>/bench_<name>.json
out=gpurun_out/$1; shift
mkdir -p $out
: > $out/summary.txt
while [ $# -gt 0 ]; do
  spec=$1; shift
  name=${spec%%=*}; lib=${spec#*=}
  envs=()
  while [ $# -gt 0 ] && [ "$1" != "--" ]; do envs+=("$1"); shift; done
  [ "$1" = "--" ] && shift
  [ "$lib" != "-" ] && envs+=("AIHAB_CLIP_LIB=$PWD/$lib")
  env "${envs[@]}" timeout 300 python bench.py --steps 20 --warmup 5 --no-other-configs > $out/bench_$name.json 2> $out/bench_$name.err
  echo "bench $name rc=$?" >> $out/summary.txt
done
cat $out/summary.txt
