"""GPU parity: the generic log-space reference DP (p7_GMSV, p7_GViterbi, p7_GForward, p7_GBackward)
vs the reference's own functions on the same generic profile.  The kernels evaluate the recurrences in the reference's
order with the reference's logsum table, so the bar is bit-identity (== on float32); Forward == Backward is checked to
the tolerance HMMER's own unit test uses for the table-driven logsum (generic_fwdback.c utest: 0.1 nat is its bound for
the table version; we see ~1e-3)."""
import numpy as np
import pytest

from pyhmmer_b200 import easel, synth
from test_msv_gpu import _targets

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M", [1, 2, 3, 17, 64, 145, 300])
def test_generic_dp_bit_identical(amino, make_pair, M):
    rng = np.random.default_rng(4000 + M)
    pair = make_pair(synth.random_hmm(amino, M, rng))
    block = _targets(amino, pair.hmm, rng, n_random=60, n_homolog=12)
    out = pair.profile._generic_scores(block)
    worst = 0.0
    for i, s in enumerate(block):
        ref = pair.ref.generic(s.sequence)
        got = (out["msv"][i], out["viterbi"][i], out["forward"][i], out["backward"][i])
        for name, r, g in zip(("GMSV", "GViterbi", "GForward", "GBackward"), ref, got):
            assert np.float32(r) == g or (np.isnan(r) and np.isnan(g)), (M, s.name, len(s), name, r, g)
        if np.isfinite(got[2]):
            worst = max(worst, abs(got[2] - got[3]))
    assert worst < 0.05, worst                      # Forward and Backward agree (table logsum)
    print("M=%d: %d comparisons bit-identical; max |GFwd - GBck| = %.2g nats" % (M, len(block), worst))


def test_profile_msv_filter(amino, make_pair):
    """Profile.msv_filter (plan7.pyx:8212): p7_GMSV with nu = 2 of one sequence."""
    rng = np.random.default_rng(9)
    pair = make_pair(synth.random_hmm(amino, 120, rng))
    seq = easel.DigitalSequence(amino, name=b"t", sequence=synth.emit_sequence(pair.hmm, rng))
    # the reference call scores with the profile as configured (L = 400), not reconfigured to the target: do the same
    got = pair.profile.msv_filter(seq)
    assert np.isfinite(got)
    ref = pair.ref.generic(seq.sequence)[0]
    assert abs(got - ref) < 1e-5        # p7_GMSV takes its length terms from L itself (generic_msv.c:62-63): no dependence on gm->L


@pytest.mark.parametrize("M", [1, 5, 40, 130])
def test_generic_decoding(amino, make_pair, M):
    """p7_GDecoding: Forward/Backward scores bit-identical, posterior probabilities within a few ulp of expf."""
    rng = np.random.default_rng(5000 + M)
    pair = make_pair(synth.random_hmm(amino, M, rng))
    dom = synth.emit_sequence(pair.hmm, rng)
    for codes in (np.concatenate([rng.integers(0, 20, 30).astype(np.uint8), dom, rng.integers(0, 20, 25).astype(np.uint8)]),
                  rng.integers(0, 20, 77).astype(np.uint8), rng.integers(0, 20, 1).astype(np.uint8)):
        seq = easel.DigitalSequence(amino, name=b"t", sequence=codes)
        pp, xpp, f, b, dom = pair.profile._generic_decoding(seq, domains=True)
        rpp, rxpp, rf, rb, rdom = pair.ref.gdecoding(codes)
        for got, ref in zip(dom, rdom):                          # p7_GDomainDecoding: host libm on bit-identical inputs
            assert np.array_equal(got, ref)
        assert np.float32(rf) == np.float32(f) and np.float32(rb) == np.float32(b), (M, len(codes), rf, f, rb, b)
        assert np.allclose(pp, rpp, rtol=2e-5, atol=1e-9), (M, len(codes), float(np.abs(pp - rpp).max()))
        assert np.allclose(xpp, rxpp, rtol=2e-5, atol=1e-9)
        rows = pp[1:, :, :2].sum((1, 2)) + xpp[1:, [1, 2, 4]].sum(1)
        assert np.allclose(rows, 1.0, atol=1e-5)                # every residue is emitted by exactly one state
