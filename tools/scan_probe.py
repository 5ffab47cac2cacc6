"""Stage timings of one hmmscan (BASELINE configs[3]): python tools/scan_probe.py [n_profiles] [L]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_inputs
from pyhmmer_b200 import _lib, plan7, easel

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
amino = easel.Alphabet.amino()
models, cal = bench_inputs.pfam_like_models(n)
hmms = bench_inputs.to_hmms(models, amino)
ctx = _lib.context(0)
pli = plan7.Pipeline(amino)
oms = pli._optimized_many(hmms, 100)
q = easel.DigitalSequence(amino, name=b"q", sequence=bench_inputs.c4_query(models, L))
block = easel.DigitalSequenceBlock(amino, [q])
for rep in range(4):
    if rep == 3:
        os.environ["B2H_TRACE"] = "1"
    ctx.set_profiling(True); ctx.stage_ms(reset=True)
    t0 = time.perf_counter()
    hits, doms, text, counters = pli._run(oms, block, seq_counters=True)
    dt = time.perf_counter() - t0
    st = ctx.stage_ms(reset=True); ctx.set_profiling(False)
    print("rep %d: %.2f ms, counters %s, stages %s, host split %s" % (rep, dt * 1e3, counters[0].tolist(), {k: round(v, 2) for k, v in st.items() if v},
                                                                      [round(v * 1e3, 2) for v in pli._last_run_s]), flush=True)
