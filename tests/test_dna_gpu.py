"""The search path on a nucleotide alphabet (K = 4, Kp = 18): the standard p7_Pipeline as pyhmmer runs it for
`Pipeline(Alphabet.dna())` on ordinary (not long-target) sequences.  Same bars as for proteins: SSV/MSV/Viterbi bit-exact,
Forward/Backward <= 1e-4 nats, null1 exact, hits identical to the reference loop."""
import os
import tempfile

import numpy as np
import pytest

from pyhmmer_b200 import _lib, easel, plan7, synth
from oracle import refshim
from test_msv_gpu import _targets, _run
from test_search_gpu import _compare_with_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M", [12, 90, 333, 1000])
def test_dna_filters(ctx, make_pair, M):
    dna = easel.Alphabet.dna()
    rng = np.random.default_rng(6000 + M)
    pair = make_pair(synth.random_hmm(dna, M, rng))
    block = _targets(dna, pair.hmm, rng, n_random=80, n_homolog=16)
    msv, mst = _run(_lib.lib.b2h_msv_filter, ctx, pair.om, block)
    vit, vst = _run(_lib.lib.b2h_viterbi_filter, ctx, pair.om, block)
    fwd, fst = _run(_lib.lib.b2h_forward_parser, ctx, pair.om, block)
    bck, bst = _run(_lib.lib.b2h_backward_parser, ctx, pair.om, block)
    for i, s in enumerate(block):
        assert pair.ref.msv(s.sequence) == (msv[i], mst[i]), (M, s.name)
        assert pair.ref.vit(s.sequence) == (vit[i], vst[i]), (M, s.name)
        rf, rb, rst = pair.ref.fwdbck(s.sequence)
        assert rst == 0 and fst[i] == 0 and (bst[i] & 0xff) == 0
        assert abs(rf - fwd[i]) <= 1e-4 + 2e-7 * abs(rf) and abs(rb - bck[i]) <= 1e-4 + 2e-7 * abs(rb), (M, s.name, rf, fwd[i], rb, bck[i])


def test_dna_search_matches_reference_loop():
    dna = easel.Alphabet.dna()
    rng = np.random.default_rng(61)
    hmms = [synth.random_hmm(dna, M, rng, name="dna%d" % i) for i, M in enumerate((60, 240))]
    synth.calibrate(hmms)
    seqs = synth.random_sequences(dna, 1500, rng)
    for i in range(40):
        h = hmms[i % 2]
        s = seqs[int(rng.integers(0, len(seqs)))]
        cut = int(rng.integers(0, len(s)))
        s.sequence = np.concatenate([s.sequence[:cut], synth.emit_sequence(h, rng), s.sequence[cut:]])[:1500]
    seqs._cache = {}
    pli = plan7.Pipeline(dna)
    nhit = 0
    with tempfile.TemporaryDirectory() as td:
        for i, h in enumerate(hmms):
            path = os.path.join(td, "m%d.hmm" % i)
            with open(path, "wb") as f:
                h.write(f)
            with plan7.HMMFile(path) as f:
                h2 = f.read()
            raw = pli._run([pli._optimized(h2, len(seqs[0]))], seqs)
            out = refshim.RefModel(path, 0, 400).search([s.sequence for s in seqs])
            _compare_with_ref((raw[0], raw[1], raw[2], raw[3][0]), out)
            nhit += len(out[0])
    assert nhit >= 20
