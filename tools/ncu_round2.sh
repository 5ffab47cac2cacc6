#!/bin/bash
# ncu evidence of round 2 (run on the GPU box): launch list of a bench step, full sets of the SSV tiles and the packed Viterbi
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r02b_launches.csv $B > /dev/null 2> gpurun_out/r02b_launches.err
ncu --set full --clock-control none --cache-control none --import-source on -k regex:ssv_kernel -s 150 -c 44 -o gpurun_out/r02b_ssv -f $B > /dev/null 2> gpurun_out/r02b_ssv.err
ncu -i gpurun_out/r02b_ssv.ncu-rep --page raw --csv > gpurun_out/r02b_ssv_raw.csv 2>/dev/null
ncu --set full --clock-control none --cache-control none --import-source on -k regex:rvit2_kernel -s 45 -c 15 -o gpurun_out/r02b_vit -f $B > /dev/null 2> gpurun_out/r02b_vit.err
ncu -i gpurun_out/r02b_vit.ncu-rep --page raw --csv > gpurun_out/r02b_vit_raw.csv 2>/dev/null
ls -la gpurun_out/r02b_* | head; rm -f gpurun_out/r02b_ssv.ncu-rep gpurun_out/r02b_vit.ncu-rep
wc -l gpurun_out/r02b_launches.csv gpurun_out/r02b_ssv_raw.csv gpurun_out/r02b_vit_raw.csv
