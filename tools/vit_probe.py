"""Dense ViterbiFilter timing / redo statistics on the bench inputs (debugging aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pyhmmer_b200 import _lib, plan7

abc, hmms, seqs = bench.build_inputs(0, 1)
bench.apply_stats(hmms)
ctx = _lib.context(0)
pli = plan7.Pipeline(abc)
db = plan7.SequenceDatabase.of(ctx, seqs)
n = len(seqs)
sc = np.empty(n, np.float32); st = np.empty(n, np.int32)
tot = 0.0; cells = 0.0
for h in hmms[:int(os.environ.get("VP_N", "24"))]:
    om = pli._optimized(h, 350)
    dev = om._device(ctx)
    for rep in range(3):
        t0 = time.perf_counter()
        _lib.check(_lib.lib.b2h_viterbi_filter(ctx.handle, dev, db.handle, _lib.ptr(sc), _lib.ptr(st)), "vit", ctx.handle)
        dt = time.perf_counter() - t0
    tot += dt; cells += float(h.M) * seqs.total_residues
    print("M=%4d  %.2f ms  %.0f GCUPS  redo=%d  inf=%d" % (h.M, dt * 1e3, h.M * seqs.total_residues / dt / 1e9, int((st == 0x7e00d0).sum()), int(np.isinf(sc).sum())))
print("TOTAL %.1f ms, %.0f GCUPS" % (tot * 1e3, cells / tot / 1e9))
