"""Larger configurations than the bench line (BASELINE configs[2] and [3] scaled to one GPU): robustness + throughput.
   python tools/scale_probe.py search P N      # P synthetic profiles x N sequences through hmmsearch's engine
   python tools/scale_probe.py scan P L        # one query of length L against P profiles (hmmscan)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pyhmmer_b200 import _lib, plan7, easel, synth, hmmer

mode, P, N = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
abc = easel.Alphabet.amino()
rng = np.random.default_rng(5)
t0 = time.perf_counter()
Ms = np.clip(np.rint(np.exp(rng.normal(np.log(150.0), 0.6, P))), 20, 2300).astype(int)
hmms = [synth.random_hmm(abc, int(M), rng, name="syn%06d" % i) for i, M in enumerate(Ms)]
ev = np.array([-8.5, 0.71, -9.5, 0.71, -4.0, 0.71], np.float32)          # plausible statistics (not fitted: throughput probe)
for h in hmms:
    h._evparam[:] = ev + np.array([-(np.log2(h.M) - 7.0), 0, -(np.log2(h.M) - 7.0), 0, -(np.log2(h.M) - 7.0) * 0.7, 0], np.float32)
print("built %d HMMs (sum M = %d) in %.1f s" % (P, int(Ms.sum()), time.perf_counter() - t0), flush=True)
ctx = _lib.context(0)
pli = plan7.Pipeline(abc)
if mode == "search":
    seqs = synth.random_sequences(abc, N, rng)
    for j in range(N // 200):
        s = seqs[int(rng.integers(0, N))]
        s.sequence = np.concatenate([s.sequence[:50], synth.emit_sequence(hmms[j % P], rng), s.sequence[50:]])[:1500]
    seqs._cache = {}
    t0 = time.perf_counter()
    oms = [pli._optimized(h, 350) for h in hmms]
    t1 = time.perf_counter()
    plan7.SequenceDatabase.of(ctx, seqs); plan7.OptimizedProfile._device_many(ctx, oms); ctx.synchronize() if hasattr(ctx, "synchronize") else None
    t2 = time.perf_counter()
    for rep in range(2):
        t3 = time.perf_counter()
        hits, doms, text, counters = pli._run(oms, seqs)
        dt = time.perf_counter() - t3
        cells = float(Ms.sum()) * seqs.total_residues
        print("search rep %d: %.3f s, %.0f GCUPS, %d comparisons scored to completion, counters %s; host convert %.1f s, upload %.2f s"
              % (rep, dt, cells / dt / 1e9, len(hits), counters.sum(0).tolist(), t1 - t0, t2 - t1), flush=True)
else:
    q = easel.DigitalSequence(abc, name=b"query", sequence=np.concatenate(
        [rng.integers(0, abc.K, N).astype(np.uint8)[: N // 2], synth.emit_sequence(hmms[3], rng), synth.emit_sequence(hmms[7], rng), rng.integers(0, abc.K, N).astype(np.uint8)])[:N])
    t0 = time.perf_counter()
    block = plan7.OptimizedProfileBlock(abc, [pli._optimized(h, 350) for h in hmms])
    t1 = time.perf_counter()
    for rep in range(2):
        t3 = time.perf_counter()
        th = list(hmmer.hmmscan([q], block))[0]
        dt = time.perf_counter() - t3
        print("scan rep %d: %.3f s (%.0f GCUPS), %d hits, Z = %g; host convert %.1f s" % (rep, dt, float(Ms.sum()) * len(q) / dt / 1e9, len(th), th.Z, t1 - t0), flush=True)
