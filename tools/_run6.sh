for cfg in "2 1.0" "3 0.6" "4 0.6" "3 0.45" "5 0.6"; do
  set -- $cfg
  B2H_WAVES=$1 B2H_WAVE_RATIO=$2 B2H_TRACE=1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/t7_bench_$1_$2.json 2> gpurun_out/t7_bench_$1_$2.err
done
CUDA_DEVICE_MAX_CONNECTIONS=8 B2H_WAVES=3 B2H_WAVE_RATIO=0.6 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/t7_bench_conn8.json 2> gpurun_out/t7_bench_conn8.err
