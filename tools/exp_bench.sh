#!/bin/bash
# usage: tools/exp_bench.sh <tag> [ENV=VAL ...]   -- one short bench run under the given environment; one summary line
tag=$1; shift
env "$@" python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err
python - "$tag" "$*" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/exp_%s.json"%sys.argv[1]))
    st=d["config"]["stage_ms_per_step"]
    print("%-10s %-60s ms/step %.2f  e2e %.2f  stages %s"%(sys.argv[1],sys.argv[2],d["ms_per_step"],d["e2e"]["ms_per_step"],{k:round(v,1) for k,v in st.items() if v}))
except Exception as e:
    print(sys.argv[1],"FAILED",e)
PY
