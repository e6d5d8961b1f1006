# round-end verification set (run under gpurun): GPU tests, smoke, both bench arms; results under gpurun_out/final/
# (the ncu launch list of the step is a separate call: see profiles/README.md)
mkdir -p gpurun_out/final
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/final/gpu_tests.txt 2>&1; echo "tests rc=$?" > gpurun_out/final/summary.txt
tail -3 gpurun_out/final/gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/final/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err; echo "bench rc=$?" >> gpurun_out/final/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err; echo "ref rc=$?" >> gpurun_out/final/summary.txt
cat gpurun_out/final/summary.txt
