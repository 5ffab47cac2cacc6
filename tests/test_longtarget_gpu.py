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


def _scan_info(om, F1):
    import ctypes
    from pyhmmer_b200 import _lib
    thr, fast = ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.lib.b2h_longtarget_scan_info(om._device(_lib.context()), float(F1), ctypes.byref(thr), ctypes.byref(fast)), "b2h_longtarget_scan_info")
    return thr.value, bool(fast.value)


@pytest.mark.parametrize("cells", ["fast", "full"])              # the two-instruction cell (default where it is exact) / the byte arithmetic in full
@pytest.mark.parametrize("M", [9, 60, 121, 333, 600, 1100])      # SSV tile families G = 8 / 16 / 32, leftover words 0..3
def test_ssv_longtarget_windows(make_pair, monkeypatch, M, cells):
    if cells == "full":
        monkeypatch.setenv("B2H_LT_FULL_CELLS", "1")
    dna = easel.Alphabet.dna()
    rng = np.random.default_rng(7000 + M)
    h = synth.random_hmm(dna, M, rng, name="lt%d" % M)
    h.max_length = 2 * M + 50
    h._evparam[:] = np.array([-8.0 - np.log2(M) * 0.3, 0.70, -9.0, 0.70, -4.0, 0.70], np.float32)
    pair = make_pair(h)
    block = _chunks(dna, pair.hmm, rng, [40000, 7, 26214, 1, 15000, 3000, 262144 // 4], nplant=6)
    raw, merged = plan7.long_target_windows(pair.om, block, F1=0.02)
    assert _scan_info(pair.om, 0.02)[1] == (cells == "fast")    # (these models and F1 qualify for the short cell)
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


@pytest.mark.parametrize("F1", [0.5, 7e-9, 3e-9])
def test_ssv_longtarget_threshold_extremes(make_pair, F1):
    """A low threshold (diagonals every few hundred rows: resets, stretches scanned again), the highest one the short cell
    takes (256 - bias) and one beyond it (the cell arithmetic falls back to the saturating form by itself)."""
    dna = easel.Alphabet.dna()
    rng = np.random.default_rng(81)
    h = synth.random_hmm(dna, 150, rng, name="ltx")
    h.max_length = 350
    h._evparam[:] = np.array([-9.5, 0.70, -9.0, 0.70, -4.0, 0.70], np.float32)
    pair = make_pair(h)
    block = _chunks(dna, pair.hmm, rng, [30000, 20000], nplant=5)
    thr, fast = _scan_info(pair.om, F1)
    assert (thr, fast) == {0.5: (171, True), 7e-9: (250, True), 3e-9: (253, False)}[F1]
    raw, merged = plan7.long_target_windows(pair.om, block, F1=F1)
    n = 0
    for ci, s in enumerate(block):
        rraw, rsc, rmer, _, _ = pair.ref.longtarget_windows(s.sequence, F1=F1)
        mine = raw[raw["seq"] == ci]
        assert len(mine) == len(rraw), (F1, thr, fast, ci, len(mine), len(rraw))
        assert np.array_equal(mine["n"], rraw[:, 0]) and np.array_equal(mine["k"], rraw[:, 1]) and np.array_equal(mine["length"], rraw[:, 2])
        assert np.array_equal(mine["score"], rsc)
        mm = merged[merged["seq"] == ci]
        assert np.array_equal(mm["n"], rmer[:, 0]) and np.array_equal(mm["length"], rmer[:, 1])
        n += len(rraw)
    print("F1=%g: threshold %d, short cell %s, %d diagonals identical" % (F1, thr, fast, n))
    assert n >= 5


def test_longtarget_needs_max_length(make_pair):
    dna = easel.Alphabet.dna()
    pair = make_pair(synth.random_hmm(dna, 50, np.random.default_rng(1)))
    block = easel.DigitalSequenceBlock(dna, [easel.DigitalSequence(dna, name=b"c", sequence=np.zeros(100, np.uint8))])
    with pytest.raises(Exception):
        plan7.long_target_windows(pair.om, block)
