#!/bin/bash
# wave structure sweep with the wave-by-wave result path (tools/exp_bench.sh prints one line per run)
for cfg in "4 0.6" "4 0.5" "5 0.6" "5 0.7" "6 0.7" "3 0.5" "8 0.8"; do
  set -- $cfg
  tools/exp_bench.sh w$1_$2 B2H_WAVES=$1 B2H_WAVE_RATIO=$2
done
