"""Fit the E-value statistics of the synthetic benchmark sets on a B200 (synth.calibrate: the procedure of p7_Calibrate with
numpy's random sequences and the GPU filters) and leave them under gpurun_out/ for committing to tests/golden/:

    bench_pfam_like_stats.npy   [20000][6]  the Pfam-A-sized set of BASELINE configs[2] / [3]  (bench_inputs.pfam_like_models)
    bench_dna_stats.json        the DNA model of configs[4] (bench_inputs.c5_inputs): evparam + max_length (p7_Builder_MaxLength)

    python tools/calibrate_sets.py [n_models]
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_inputs
from pyhmmer_b200 import _lib, easel, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else bench_inputs.PFAM_N
out = os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)
ctx = _lib.context(0)
t0 = time.time()
models, _ = bench_inputs.pfam_like_models(n)
print("generated %d models in %.1f s" % (n, time.time() - t0), flush=True)
abc = easel.Alphabet.amino()
ev = np.zeros((n, 6), np.float32)
t0 = time.time()
for i0 in range(0, n, 500):
    hmms = bench_inputs.to_hmms(models[i0:i0 + 500], abc)
    synth.calibrate(hmms, ctx)
    for j, h in enumerate(hmms):
        ev[i0 + j] = h._evparam
    del hmms
    print("calibrated %d / %d (%.1f s)" % (min(n, i0 + 500), n, time.time() - t0), flush=True)
np.save(os.path.join(out, "bench_pfam_like_stats.npy"), ev)
dna = easel.Alphabet.dna()
res = {}
for M in (1000,):
    model, _, _, _ = bench_inputs.c5_inputs(M, megabases=0.01)
    h = synth.hmm_from_arrays(dna, model)
    synth.calibrate([h], ctx)
    res[str(M)] = {"evparam": [float(v) for v in h._evparam], "max_length": int(h.compute_max_length())}
json.dump(res, open(os.path.join(out, "bench_dna_stats.json"), "w"))
print(res)
