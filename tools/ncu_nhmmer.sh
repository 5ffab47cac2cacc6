#!/bin/bash
# ncu evidence for the long-target path (run on the GPU box): launch list of ONE nhmmer search of bench.py's configs[4] inputs
# (M = 1000 vs 100 Mb, both strands) and the full set of the first-pass lt_ssv_kernel launch.
P="python tools/nhmmer_probe.py 1000 100 0 bench"
PROBE_REPS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02c_nh_launches.csv $P > gpurun_out/r02c_nh_launches.out 2> gpurun_out/r02c_nh_launches.err
PROBE_REPS=1 ncu --set full --clock-control none --cache-control none --import-source on -k regex:lt_ssv_kernel -c 2 -o gpurun_out/r02c_ltssv -f $P > /dev/null 2> gpurun_out/r02c_ltssv.err
ncu -i gpurun_out/r02c_ltssv.ncu-rep --page raw --csv > gpurun_out/r02c_ltssv_raw.csv 2>/dev/null
rm -f gpurun_out/r02c_ltssv.ncu-rep
wc -l gpurun_out/r02c_nh_launches.csv gpurun_out/r02c_ltssv_raw.csv
