# round-end verification set (run under gpurun): GPU tests, smoke, both bench arms, ncu launch list of this library's kernels
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b21.json 2>/dev/null && \
  ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:aihab -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1; echo "ncu rc=$?"
) > gpurun_out/final.log 2>&1
cat gpurun_out/final.log
cut -c1-300 gpurun_out/bench_ref.json
