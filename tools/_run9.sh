python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/t22_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t22_smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/t22_bench.json 2> gpurun_out/t22_bench.err
cat gpurun_out/t22_tests.log gpurun_out/t22_smoke.log; tail -n 2 gpurun_out/t22_bench.err
