python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/t19_tests.log
B2H_TRACE=1 python bench.py --steps 5 --warmup 3 > gpurun_out/t19_bench.json 2> gpurun_out/t19_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/t19_ref.json 2> gpurun_out/t19_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/t19_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/t19_ncu1.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -k regex:ssv_kernel -c 400 --csv --log-file gpurun_out/t19_ssv_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/t19_ncu2.log 2>&1
ncu --set full --clock-control none --cache-control none -k regex:"ssv_kernel<8, 13>|rvit2_kernel<8>|rmsv_kernel<8, 13>|rfwd_kernel<8, 1>|bias_kernel" -s 20 -c 8 -o /tmp/t19_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/t19_ncu3.log 2>&1
ncu -i /tmp/t19_full.ncu-rep --page raw --csv > gpurun_out/t19_full_raw.csv
cat gpurun_out/t19_tests.log
