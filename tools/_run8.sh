python -m pytest tests/test_search_gpu.py -x -q 2>&1 | tail -4 > gpurun_out/t21_tests.log
run() { name=$1; shift; env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t21_$name.json 2> gpurun_out/t21_$name.err; }
run par B2H_X=1
run ser B2H_SERIAL_BIAS=1
run par2 B2H_X=1
cat gpurun_out/t21_tests.log
