"""Throughput of the long-target first stage on a BASELINE configs[4]-like input: DNA profile (M ~ 1000) against a synthetic
genome cut into 262 144-residue chunks with overlap (both strands = twice the chunks)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhmmer_b200 import _lib, plan7, easel, synth

M, MB = int(sys.argv[1]), int(sys.argv[2])
dna = easel.Alphabet.dna()
rng = np.random.default_rng(11)
h = synth.random_hmm(dna, M, rng, name="probe")
h.max_length = 2 * M
h._evparam[:] = np.array([-8.0 - np.log2(M) * 0.3, 0.70, -9.0, 0.70, -4.0, 0.70], np.float32)
om = plan7.Profile(M, dna).configure(h, plan7.Background(dna), 400).to_optimized()
W = 262144
genome = rng.integers(0, 4, MB * 1000000).astype(np.uint8)
for j in range(200):
    dom = synth.emit_sequence(h, rng); pos = int(rng.integers(0, len(genome) - len(dom))); genome[pos:pos + len(dom)] = dom
chunks = []
step = W - h.max_length
for strand in range(2):
    g = genome if strand == 0 else (3 - genome[::-1])
    for lo in range(0, len(g), step):
        chunks.append(easel.DigitalSequence(dna, name=b"c%d_%d" % (strand, lo), sequence=g[lo:lo + W]))
block = easel.DigitalSequenceBlock(dna, chunks)
ctx = _lib.context(0)
plan7.SequenceDatabase.of(ctx, block); om._device(ctx)
for rep in range(3):
    t0 = time.perf_counter()
    raw, merged = plan7.long_target_windows(om, block)
    dt = time.perf_counter() - t0
    cells = float(M) * sum(len(c) for c in chunks)
    print("M=%d, %d chunks (%.0f Mb x 2 strands): %.1f ms, %.0f GCUPS, %d diagonals -> %d windows" % (M, len(chunks), MB, dt * 1e3, cells / dt / 1e9, len(raw), len(merged)), flush=True)
