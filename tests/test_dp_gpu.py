"""GPU parity: ViterbiFilter (bit-exact), Forward/Backward parsers (<= 1e-4 nats), null1 and bias filter
vs the reference's p7_ViterbiFilter / p7_ForwardParser / p7_BackwardParser / p7_bg_NullOne / p7_bg_FilterScore."""
import numpy as np
import pytest

from pyhmmer_b200 import _lib, easel, plan7, synth
from test_msv_gpu import _targets, _run

pytestmark = pytest.mark.gpu

FB_TOL = 1e-4      # nats; north_star's bound, and HMMER's own Fwd==Bck contract (fwdback.c:925-927)


# one model length (at least) per register size class of b2h_dpreg.cu, the class boundaries, and the shared-memory kernels beyond 1536
@pytest.mark.parametrize("M", [1, 2, 7, 8, 9, 31, 32, 33, 64, 65, 80, 100, 150, 180, 200, 210, 257, 270, 300, 340, 384, 420, 500, 513, 550, 640,
                               641, 700, 768, 850, 900, 1024, 1025, 1200, 1300, 1536, 1600])
def test_viterbi_bit_exact(ctx, amino, make_pair, M):
    rng = np.random.default_rng(2000 + M)
    pair = make_pair(synth.random_hmm(amino, M, rng))
    block = _targets(amino, pair.hmm, rng, n_random=120, n_homolog=24)
    sc, st = _run(_lib.lib.b2h_viterbi_filter, ctx, pair.om, block)
    n_inf = 0
    for i, s in enumerate(block):
        rsc, rst = pair.ref.vit(s.sequence)
        assert rst == st[i], (M, s.name, rst, st[i])
        assert rsc == sc[i], (M, s.name, len(s), rsc, sc[i])
        n_inf += np.isinf(rsc)
    print("M=%d: %d comparisons, %d overflowed" % (M, len(block), n_inf))


@pytest.mark.parametrize("M", [1, 2, 9, 33, 64, 80, 100, 150, 180, 200, 210, 257, 270, 300, 340, 420, 500, 513, 550, 640, 641, 700, 768, 850, 900,
                               1024, 1025, 1200, 1300, 1536, 1600])
def test_forward_backward_parsers(ctx, amino, make_pair, M):
    rng = np.random.default_rng(3000 + M)
    pair = make_pair(synth.random_hmm(amino, M, rng))
    block = _targets(amino, pair.hmm, rng, n_random=100, n_homolog=24)
    fsc, fst = _run(_lib.lib.b2h_forward_parser, ctx, pair.om, block)
    bsc, bst = _run(_lib.lib.b2h_backward_parser, ctx, pair.om, block)
    worst_f = worst_b = 0.0
    for i, s in enumerate(block):
        rf, rb, rst = pair.ref.fwdbck(s.sequence)
        assert rst == 0 and fst[i] == 0 and (bst[i] & 0xff) == 0, (M, s.name, rst, fst[i], bst[i])
        worst_f = max(worst_f, abs(rf - fsc[i]))
        worst_b = max(worst_b, abs(rb - bsc[i]))
        assert abs(rf - fsc[i]) <= FB_TOL + 2e-7 * abs(rf), (M, s.name, len(s), rf, fsc[i])
        assert abs(rb - bsc[i]) <= FB_TOL + 2e-7 * abs(rb), (M, s.name, len(s), rb, bsc[i])
    print("M=%d: max |dFwd| = %.3g, max |dBck| = %.3g nats" % (M, worst_f, worst_b))


def test_null_and_bias_scores(ctx, amino, make_pair):
    rng = np.random.default_rng(11)
    for M in (30, 200, 640):
        pair = make_pair(synth.random_hmm(amino, M, rng))
        block = _targets(amino, pair.hmm, rng, n_random=150, n_homolog=10)
        db = plan7.SequenceDatabase(ctx, block)
        n1 = np.empty(len(block), np.float32)
        fs = np.empty(len(block), np.float32)
        _lib.check(_lib.lib.b2h_null_scores(ctx.handle, pair.om._device(ctx), db.handle, _lib.ptr(n1), _lib.ptr(fs)), "null", ctx.handle)
        nbad = 0
        for i, s in enumerate(block):
            assert pair.ref.null1(s.sequence) == n1[i], (s.name, len(s))
            rb = pair.ref.bias(s.sequence)
            # device log() of a double may differ from glibc's in the last ulp; after rounding to float32 that
            # is almost always invisible: allow 2 ulp of float32 and count the exact matches
            assert abs(rb - fs[i]) <= 4 * np.spacing(np.float32(abs(rb))), (s.name, len(s), rb, fs[i])
            nbad += (rb != fs[i])
        print("M=%d: %d/%d bias-filter scores not bit-identical" % (M, nbad, len(block)))
