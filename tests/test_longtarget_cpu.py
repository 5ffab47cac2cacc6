"""CPU-only checks of the long-target (nhmmer) window stages.

1. The oracle is pinned: `ref_longtarget_stages` (the reference's static postSSV / postViterbi functions restated with its
   public calls) leaves the same pos_past_* counters as the real p7_Pipeline_LongTarget, and the scalar port of
   p7_ViterbiFilter_longtarget (oracle/hmmer_oracle.c) records the same landmarks in the same order as the SSE original.
2. The product's HOST logic (`pyhmmer_b200.longtarget.stages`: gates, B1/B2/B3 bias scaling in the reference's precision,
   counters; `b2h_longtarget_vit_finish`: landmark order, extend / merge, 80 kb cut; `b2h_longtarget_vit_threshold`) gives
   the reference's intermediates when every DP score is supplied by the reference's functions -- no device involved.
"""
import ctypes
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pyhmmer_b200 import _lib, longtarget
from oracle import port
import lt_common
from conftest import ModelPair


@pytest.mark.parametrize("M,mu_shift", [(60, 0.0), (60, -3.0), (333, -2.0)])
def test_restated_stages_match_the_real_pipeline(M, mu_shift):
    pair, rng = lt_common.dna_model(ModelPair, M, mu_shift=mu_shift)
    block = lt_common.dna_chunks(pair, rng, [120000, 30000], nplant=8)
    for s in block:
        st = pair.ref.longtarget_stages(s.sequence)
        cnt, hits = pair.ref.longtarget_pipeline(s.sequence)
        assert np.array_equal(st["counters"], cnt[:4]), (st["counters"], cnt)
        assert st["counters"][0] > 0 and cnt[4] >= 1 and len(st["vithit"]) > 0


@pytest.mark.parametrize("M", [5, 9, 17, 60, 333, 1100])
def test_port_of_viterbi_longtarget(M):
    pair, rng = lt_common.dna_model(ModelPair, M)
    pt = port.Port(pair.om)
    total = 0
    for L in (1, 7, 300, 3000):
        seq = lt_common.dna_chunks(pair, rng, [L], nplant=2)[0].sequence
        for filtersc in (-3.0, -8.0, -12.0):
            cfg = min(L, pair.hmm.max_length)
            a = pair.ref.vit_longtarget(seq, cfg, filtersc)
            b = pt.vit_longtarget(seq, cfg, filtersc)
            assert np.array_equal(a, b), (M, L, filtersc, len(a), len(b))
            total += len(a)
            # the threshold the host library derives is the one the port derives
            thr, xwm = ctypes.c_int32(), ctypes.c_int32()
            o = ctypes.c_void_p()
            assert _lib.lib.b2h_profile_create_host(ctypes.byref(pair.om._desc), ctypes.byref(o)) == 0
            assert _lib.lib.b2h_longtarget_vit_threshold(o, cfg, filtersc, 3e-3, ctypes.byref(thr), ctypes.byref(xwm)) == 0
            _lib.lib.b2h_profile_destroy(o)
            assert (thr.value, xwm.value) == pt.vit_longtarget_threshold(cfg, filtersc, 3e-3)
    assert total > 0


@pytest.mark.parametrize("M,mu_shift,bias_filter", [(60, -3.0, True), (121, -2.0, True), (333, -2.0, False)])
def test_host_logic_with_reference_scores(M, mu_shift, bias_filter):
    pair, rng = lt_common.dna_model(ModelPair, M, mu_shift=mu_shift)
    block = lt_common.dna_chunks(pair, rng, [60000, 0, 9, 25000], nplant=6)
    kw = dict(F1=0.02, F2=3e-3, F3=3e-5, bias_filter=bias_filter, B1=100, B2=240, B3=1000)
    got = longtarget.stages(pair.om, block, backend=lt_common.OracleBackend(pair, block), **kw)
    tot = lt_common.compare_with_reference(pair, block, got, exact_scores=True, **kw)
    assert tot["msvwin"] >= 5 and tot["vitmark"] >= 5 and tot["vitwin"] >= 3 and tot["passed"] >= 1, tot


def test_long_windows_are_cut_at_80kb():
    """b2h_longtarget_vit_finish against the reference's loop (p7_pipeline.c:1389-1406) on a window long enough to be cut:
    a repeat-rich region whose landmarks merge into one window above 80 kb."""
    M = 40
    pair, rng = lt_common.dna_model(ModelPair, M, mu_shift=-3.0)
    dom = lt_common.synth.emit_sequence(pair.hmm, rng)
    reps = 200000 // len(dom)
    seq = np.concatenate([dom] * reps).astype(np.uint8)
    dna = pair.hmm.alphabet
    block = lt_common.easel.DigitalSequenceBlock(dna, [lt_common.easel.DigitalSequence(dna, name=b"rep", sequence=seq)])
    kw = dict(F1=0.02, F2=3e-3, F3=3e-5, bias_filter=True, B1=100, B2=240, B3=1000)
    got = longtarget.stages(pair.om, block, backend=lt_common.OracleBackend(pair, block), **kw)
    tot = lt_common.compare_with_reference(pair, block, got, exact_scores=True, **kw)
    assert got["vitwin"]["length"].max() == 80000 and tot["vitwin"] >= 3, (tot, got["vitwin"]["length"])
