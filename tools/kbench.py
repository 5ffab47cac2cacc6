"""Kernel-level timing on the GPU box (not the bench contract): SSV/MSV cell rate for one profile."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyhmmer_b200 import _lib, easel, plan7, synth

abc = easel.Alphabet.amino()
rng = np.random.default_rng(0)
nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
ctx = _lib.context(0)
seqs = synth.random_sequences(abc, nseq, rng)
db = plan7.SequenceDatabase.of(ctx, seqs)
nres = seqs.total_residues
sc = np.empty(nseq, np.float32); st = np.empty(nseq, np.int32)
for M in (60, 120, 200, 250, 380, 500, 1000, 2000):
    hmm = synth.random_hmm(abc, M, rng)
    om = plan7.Profile(M, abc).configure(hmm, plan7.Background(abc), 400).to_optimized()
    h = om._device(ctx)
    for fn, name in ((_lib.lib.b2h_ssv_filter, "ssv"), (_lib.lib.b2h_msv_filter, "msv")):
        fn(ctx.handle, h, db.handle, _lib.ptr(sc), _lib.ptr(st))
        ctx.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            fn(ctx.handle, h, db.handle, _lib.ptr(sc), _lib.ptr(st))
            ctx.synchronize()
            best = min(best, time.perf_counter() - t0)
        print("M=%4d %s: %.3f ms  %.1f GCUPS (wall incl. D2H of %d scores)  redo=%d inf=%d" % (
            M, name, best * 1e3, M * nres / best / 1e9, nseq, int((st == 19).sum()), int(np.isinf(sc).sum())), flush=True)
