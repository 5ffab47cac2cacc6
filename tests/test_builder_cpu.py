"""`pyhmmer_b200.builder.Builder` (single query sequences: phmmer / nhmmer sequence queries) against the reference's
p7_SingleBuilder: the HMM file we write is, line for line, the file the reference writes (DATE aside) -- emissions from the
score matrix's conditional probabilities, gap transitions, composition, consensus, E-value parameters calibrated on the same
random sequences (Easel's fast generator), MAXL for nucleotide models.  The calibration filters come from the reference here
(no device in this suite); the product runs them on the GPU."""
import io
import os
import sys
import tempfile

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pyhmmer_b200 import builder, easel, plan7
from oracle import refshim


def _reference_scorer(state):
    def scorer(om, seqs, which):
        with tempfile.NamedTemporaryFile(suffix=".hmm") as tmp:      # the reference's filters on OUR model
            state["hmm"].write(tmp)
            tmp.flush()
            rm = refshim.RefModel(tmp.name, 0, 400)
        f = {"msv": rm.msv, "vit": rm.vit, "fwd": rm.fwd}[which]
        return np.array([f(s)[0] for s in seqs], np.float32), np.array([rm.null1(s) for s in seqs], np.float32)
    return scorer


def _build(b, seq, bg):
    state = {}
    b._scorer = _reference_scorer(state)
    calibrate = b.calibrate

    def hooked(hmm, background):
        state["hmm"] = hmm
        if hmm._evparam[0] == plan7.P7_EVPARAM_UNSET:      # the file the scorer writes needs some statistics to be readable
            hmm._evparam[:] = np.array([-8, .7, -9, .7, -4, .7], np.float32)
        return calibrate(hmm, background)
    b.calibrate = hooked
    return b.build(seq, bg)


@pytest.mark.parametrize("alphabet,L,seed,kw", [("amino", 80, 42, {}), ("amino", 7, 42, {}), ("amino", 233, 7, dict(popen=0.05, pextend=0.5)),
                                                ("dna", 120, 42, {}), ("dna", 61, 3, {})])
def test_single_sequence_models_match_the_reference_builder(alphabet, L, seed, kw, tmp_path):
    abc = getattr(easel.Alphabet, alphabet)()
    bg = plan7.Background(abc)
    rng = np.random.default_rng(L + seed)
    codes = rng.integers(0, abc.K, L).astype(np.uint8)
    if L > 50:
        codes[5] = abc.Kp - 3                              # a fully degenerate residue in the query (X / N)
    path = str(tmp_path / "ref.hmm")
    refshim.single_builder({"amino": 3, "dna": 2}[alphabet], codes, "query1", path, matrix="BLOSUM62" if alphabet == "amino" else "DNA1",
                           popen=kw.get("popen", 0.02 if alphabet == "amino" else 0.03125),
                           pextend=kw.get("pextend", 0.4 if alphabet == "amino" else 0.75), seed=seed)
    b = builder.Builder(abc, seed=seed, **kw)
    hmm, profile, om = _build(b, easel.DigitalSequence(abc, name="query1", sequence=codes), bg)
    buf = io.BytesIO()
    hmm.write(buf)
    mine = [l for l in buf.getvalue().decode().splitlines() if not l.startswith("DATE")]
    ref = [l for l in open(path).read().splitlines() if not l.startswith("DATE")]
    # the statistics to 2e-4: the reference filters of this test score OUR model through its ASCII file (five decimals per
    # parameter), the reference builder scores the unrounded one -- Forward scores, hence tau, move in the fourth decimal
    stats = lambda ls: [[float(v) for v in l.split()[3:]] for l in ls if l.startswith("STATS")]
    rest = lambda ls: [l for l in ls if not l.startswith("STATS")]
    assert rest(mine) == rest(ref), [(a, r) for a, r in zip(rest(mine), rest(ref)) if a != r][:3]
    assert np.allclose(stats(mine), stats(ref), rtol=0, atol=2e-4) and len(stats(mine)) == 3
    assert (profile.M, om.M) == (L, L) and hmm.nseq == 1 and (alphabet == "amino" or hmm.max_length > 0)


def test_generators_follow_easel():
    assert np.array_equal(builder.Randomness(42).random(2000), refshim.mt_stream(42, 2000))
    r = builder.FastRandomness(42)
    a = r.random(5)
    r.reinit()
    assert [r.random() for _ in range(5)] == list(a) and all(0.0 <= v < 1.0 for v in a)
    with pytest.raises(ValueError):
        builder.Builder(easel.Alphabet.amino(), score_matrix="PAM30")
