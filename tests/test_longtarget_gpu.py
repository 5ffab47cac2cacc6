"""GPU parity, first stage of the long-target (nhmmer) pipeline: p7_SSVFilter_longtarget + p7_pli_ExtendAndMergeWindows
 against the reference's own functions, chunk by chunk: identical diagonals (start, model end, length,
score as float32) in identical order, identical merged windows."""
import numpy as np
import pytest

from pyhmmer_b200 import easel, plan7, synth

pytestmark = pytest.mark.gpu


def _chunks(dna, hmm, rng, sizes, nplant):
    out = []
    for ci, L in enumerate(sizes):
        seq = rng.integers(0, dna.K, L).astype(np.uint8)
        for _ in range(nplant):
            dom = synth.emit_sequence(hmm, rng)
            if len(dom) < L:
                pos = int(rng.integers(0, L - len(dom)))
                seq[pos:pos + len(dom)] = dom
        out.append(easel.DigitalSequence(dna, name=b"chunk%d" % ci, sequence=seq))
    return easel.DigitalSequenceBlock(dna, out)


@pytest.mark.parametrize("M", [9, 60, 121, 333, 600, 1100])      # SSV tile families G = 8 / 16 / 32, leftover words 0..3
def test_ssv_longtarget_windows(make_pair, M):
    dna = easel.Alphabet.dna()
    rng = np.random.default_rng(7000 + M)
    h = synth.random_hmm(dna, M, rng, name="lt%d" % M)
    h.max_length = 2 * M + 50
    h._evparam[:] = np.array([-8.0 - np.log2(M) * 0.3, 0.70, -9.0, 0.70, -4.0, 0.70], np.float32)
    pair = make_pair(h)
    block = _chunks(dna, pair.hmm, rng, [40000, 7, 26214, 1, 15000, 3000, 262144 // 4], nplant=6)
    raw, merged = plan7.long_target_windows(pair.om, block, F1=0.02)
    nraw = nmer = 0
    for ci, s in enumerate(block):
        rraw, rsc, rmer, _, _ = pair.ref.longtarget_windows(s.sequence, F1=0.02)
        mine = raw[raw["seq"] == ci]
        assert len(mine) == len(rraw), (M, ci, len(mine), len(rraw))
        assert np.array_equal(mine["n"], rraw[:, 0]) and np.array_equal(mine["k"], rraw[:, 1]) and np.array_equal(mine["length"], rraw[:, 2]), (M, ci)
        assert np.array_equal(mine["score"], rsc), (M, ci)
        mm = merged[merged["seq"] == ci]
        assert np.array_equal(mm["n"], rmer[:, 0]) and np.array_equal(mm["length"], rmer[:, 1]), (M, ci, mm[:4], rmer[:4])
        nraw += len(rraw); nmer += len(rmer)
    assert nraw >= 10 and nmer >= 5                      # the planted homologs were found
    print("M=%d: %d diagonals, %d windows identical" % (M, nraw, nmer))


def test_longtarget_needs_max_length(make_pair):
    dna = easel.Alphabet.dna()
    pair = make_pair(synth.random_hmm(dna, 50, np.random.default_rng(1)))
    block = easel.DigitalSequenceBlock(dna, [easel.DigitalSequence(dna, name=b"c", sequence=np.zeros(100, np.uint8))])
    with pytest.raises(Exception):
        plan7.long_target_windows(pair.om, block)
