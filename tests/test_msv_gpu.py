"""GPU parity: SSV / MSV filters vs the reference's p7_SSVFilter / p7_MSVFilter, bit-exact.

Cases follow SURVEY 8(d): M in {1,2,7,8,9,15,16,17,63,64,65,100,200,...}, L in {1,2,15,16,17,...},
random targets (SSV decisive), planted homologs (overflow -> +inf, J state -> full MSV).
"""
import numpy as np
import pytest

from pyhmmer_b200 import _lib, easel, plan7, synth

pytestmark = pytest.mark.gpu


def _targets(abc, hmm, rng, n_random=200, n_homolog=30):
    block = synth.random_sequences(abc, n_random, rng, mean_len=200, sd_len=150, lo=1, hi=900)
    for L in (1, 2, 3, 4, 5, 15, 16, 17, 31, 32, 33, 127, 128, 129):
        block.append(easel.DigitalSequence(abc, name=b"L%d" % L, sequence=rng.integers(0, abc.K, L).astype(np.uint8)))
    # degenerate / special residue codes too (B J Z O U X and '*' '~' '-' never occur in real dsq but X etc. do)
    block.append(easel.DigitalSequence(abc, name=b"degenerate", sequence=rng.integers(abc.K + 1, abc.Kp - 2, 77).astype(np.uint8)))
    block.append(easel.DigitalSequence(abc, name=b"allX", sequence=np.full(50, abc.Kp - 3, np.uint8)))
    for i in range(n_homolog):                                  # homologs embedded in random flanks; some twice (J state)
        dom = synth.emit_sequence(hmm, rng)
        parts = [rng.integers(0, abc.K, rng.integers(0, 80)).astype(np.uint8), dom]
        if i % 3 == 0:
            parts += [rng.integers(0, abc.K, rng.integers(5, 60)).astype(np.uint8), synth.emit_sequence(hmm, rng)]
        if i % 5 == 0:
            parts = [dom[: max(1, len(dom) // 3)]]             # weak partial hit
        parts.append(rng.integers(0, abc.K, rng.integers(0, 80)).astype(np.uint8))
        s = np.concatenate(parts)
        if len(s):
            block.append(easel.DigitalSequence(abc, name=b"hom%d" % i, sequence=s))
    return block


def _run(fn, ctx, om, block):
    db = plan7.SequenceDatabase(ctx, block)
    sc = np.empty(len(block), np.float32)
    st = np.empty(len(block), np.int32)
    _lib.check(fn(ctx.handle, om._device(ctx), db.handle, _lib.ptr(sc), _lib.ptr(st)), "filter", ctx.handle)
    return sc, st


# every SSV register tile family is hit: G=8 (M <= 511; NR = ceil((M+1)/16), leftover words 0..3), G=16 (M <= 1023), G=32 beyond
@pytest.mark.parametrize("M", [1, 2, 7, 8, 9, 15, 16, 17, 31, 33, 47, 48, 63, 64, 65, 100, 127, 128, 143, 200, 255, 256, 300, 319, 400,
                               463, 511, 512, 600, 700, 1023, 1024])
def test_msv_and_ssv_bit_exact(ctx, amino, make_pair, M):
    rng = np.random.default_rng(1000 + M)
    pair = make_pair(synth.random_hmm(amino, M, rng))
    block = _targets(amino, pair.hmm, rng)
    msv_sc, msv_st = _run(_lib.lib.b2h_msv_filter, ctx, pair.om, block)
    ssv_sc, ssv_st = _run(_lib.lib.b2h_ssv_filter, ctx, pair.om, block)
    n_inf = n_redo = 0
    for i, s in enumerate(block):
        rsc, rst = pair.ref.msv(s.sequence)
        assert rst == msv_st[i], (M, s.name, rst, msv_st[i])
        assert rsc == msv_sc[i], (M, s.name, len(s), rsc, msv_sc[i])          # == on float32; inf == inf
        qsc, qst = pair.ref.ssv(s.sequence)
        assert qst == ssv_st[i], (M, s.name, qst, ssv_st[i])
        if qst == 0 or qst == 16:
            assert qsc == ssv_sc[i], (M, s.name, qsc, ssv_sc[i])
        n_inf += np.isinf(rsc)
        n_redo += (qst == 19)
    if M >= 64:
        assert n_inf > 0 and n_redo > 0          # the test really exercised overflow and the MSV fallback


def test_large_models(ctx, amino, make_pair):
    rng = np.random.default_rng(5)
    for M in (1100, 1500, 2300):
        pair = make_pair(synth.random_hmm(amino, M, rng))
        block = _targets(amino, pair.hmm, rng, n_random=40, n_homolog=6)
        sc, st = _run(_lib.lib.b2h_msv_filter, ctx, pair.om, block)
        for i, s in enumerate(block):
            rsc, rst = pair.ref.msv(s.sequence)
            assert (rst, rsc) == (st[i], sc[i]), (M, s.name, rsc, sc[i], rst, st[i])


def test_empty_database(ctx, amino, make_pair):
    pair = make_pair(synth.random_hmm(amino, 50, np.random.default_rng(0)))
    sc, st = _run(_lib.lib.b2h_msv_filter, ctx, pair.om, easel.DigitalSequenceBlock(amino))
    assert sc.size == 0
