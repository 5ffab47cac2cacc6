"""Whole nhmmer search on a BASELINE configs[4]-like input: DNA profile (M ~ 1000) against a synthetic genome with planted
homologs on both strands, through `plan7.LongTargetsPipeline` -- wall time per stage (`longtarget` timings), hits found, and
the reference's own loop (oracle/_ref, single thread) on a bounded sample of the genome for comparison.

    python tools/nhmmer_probe.py <M> <megabases> [reference sample in megabases, default 2] [bench]

With `bench` as the fourth argument the inputs are bench.py's configs[4] block (bench_inputs.c5_inputs: 1 planted homolog per Mb,
calibrated statistics) instead of the hit-dense genome above.
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from pyhmmer_b200 import plan7, easel, synth, longtarget

M, MB = int(sys.argv[1]), float(sys.argv[2])
REF_MB = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
dna = easel.Alphabet.dna()
rng = np.random.default_rng(11)
h = synth.random_hmm(dna, M, rng, name="probe")
h.max_length = h.compute_max_length()
h._evparam[:] = np.array([-8.0 - np.log2(M) * 0.3, 0.70, -9.0, 0.70, -4.0, 0.70], np.float32)
genome = rng.integers(0, 4, int(MB * 1000000)).astype(np.uint8)
nplant = max(4, int(20 * MB))
for j in range(nplant):
    dom = synth.emit_sequence(h, rng)
    if j % 2:
        dom = longtarget.reverse_complement(dna, dom)
    pos = int(rng.integers(0, len(genome) - len(dom)))
    genome[pos:pos + len(dom)] = dom
if len(sys.argv) > 4 and sys.argv[4] == "bench":
    import bench_inputs
    model, genome, nplant, cal = bench_inputs.c5_inputs(M, MB)
    h = synth.hmm_from_arrays(dna, model)
    if h.max_length is None or h.max_length <= 0:
        h.max_length = h.compute_max_length()
block = easel.DigitalSequenceBlock(dna, [easel.DigitalSequence(dna, name=b"genome", sequence=genome)])
pli = plan7.LongTargetsPipeline(dna)
print("M=%d max_length=%d, %.1f Mb x 2 strands, %d planted" % (M, h.max_length, MB, nplant), flush=True)
for rep in range(int(os.environ.get("PROBE_REPS", "3"))):
    om = pli._optimized(h, 100)
    tm = {}
    t0 = time.perf_counter()
    res = longtarget.search(om, block, F1=pli.F1, F2=pli.F2, F3=pli.F3, block_length=pli.block_length, timings=tm)
    dt = time.perf_counter() - t0
    hits, doms, text, dup, stats = res
    cells = float(M) * stats["nres"]
    print("run %d: %.1f ms total = %.0f GCUPS (cells = M x residues searched); %d hits (%d duplicates); past msv/bias/vit/fwd = %d/%d/%d/%d residues"
          % (rep, dt * 1e3, cells / dt / 1e9, len(hits), sum(dup), stats["pos_past_msv"], stats["pos_past_bias"], stats["pos_past_vit"], stats["pos_past_fwd"]))
    print("       " + ", ".join("%s %.1f ms" % (k, v * 1e3) for k, v in tm.items()), flush=True)
if REF_MB > 0:
    from conftest import ModelPair
    pair = ModelPair(h)
    sample = genome[:int(REF_MB * 1000000)]
    t0 = time.perf_counter()
    rhits, rstats = pair.ref.nhmmer([sample], evalue_window=h.max_length)
    dt = time.perf_counter() - t0
    print("reference loop (1 thread) on the first %.1f Mb: %.2f s = %.2f GCUPS, %d hits" % (REF_MB, dt, float(M) * rstats[0] / dt / 1e9, len(rhits)))
