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


def test_sequence_queries_through_the_long_target_pipeline(monkeypatch, tmp_path):
    """`LongTargetsPipeline.search_seq` / `hmmer.nhmmer` with a DigitalSequence query: the model comes from the Builder, the
    search is the nhmmer path -- compared with the reference's nhmmer loop run on the same model (host logic; the filters of
    the calibration and of the search come from the reference here)."""
    import lt_common
    from conftest import ModelPair
    from pyhmmer_b200 import hmmer
    dna = easel.Alphabet.dna()
    bg = plan7.Background(dna)
    rng = np.random.default_rng(77)
    q = rng.integers(0, 4, 150).astype(np.uint8)
    query = easel.DigitalSequence(dna, name="nquery", sequence=q)
    state = {}

    class RoundTripBuilder(builder.Builder):              # both sides search the model as its ASCII file holds it
        def build(self, sequence, background):
            inner = builder.Builder(dna, seed=self.seed, window_length=self.window_length, window_beta=self.window_beta)
            hmm, _, _ = _build(inner, sequence, background)
            state["pair"] = ModelPair(hmm)
            h = state["pair"].hmm
            prof = plan7.Profile(h.M, dna).configure(h, background, 200)
            return h, prof, prof.to_optimized()

    targets = []
    for i, L in enumerate((50000, 20000)):
        codes = rng.integers(0, 4, L).astype(np.uint8)
        for _ in range(6):                                # diverged copies of the query, both strands
            copy_ = q.copy()
            m = rng.random(len(copy_)) < 0.15
            copy_[m] = rng.integers(0, 4, int(m.sum()))
            if rng.random() < 0.5:
                copy_ = (3 - copy_[::-1]).astype(np.uint8)
            pos = int(rng.integers(0, L - len(copy_)))
            codes[pos:pos + len(copy_)] = copy_
        targets.append(easel.DigitalSequence(dna, name="t%d" % i, sequence=codes))
    block = easel.DigitalSequenceBlock(dna, targets)
    monkeypatch.setattr(plan7._lib, "context", lambda device=None: None)
    pli = plan7.LongTargetsPipeline(dna, block_length=20000)
    pli._backend_factory = lambda om, blk: lt_common.OracleBackend(state["pair"], blk)
    th = pli.search_seq(query, block, builder=RoundTripBuilder(dna))
    pair = state["pair"]
    rhits, rstats = pair.ref.nhmmer([s.sequence for s in block], block_length=20000, evalue_window=pair.ref.max_length(),
                                    names=[s.name for s in block])
    assert th.query is query and len(th) == len(rhits) >= 6
    for h, r in zip(th, rhits):
        d = h.domains[0]
        assert (h.name, d.alignment.target_from, d.alignment.target_to) == (block[r.seqidx].name, r.iali, r.jali)
        assert abs(h.score - r.score) < 2e-3 and h.reported == bool(r.flags & 2)
    assert any(d.alignment.target_from > d.alignment.target_to for h in th for d in h.domains)      # a hit on the reverse strand
    with pytest.raises(ValueError):
        pli.search_seq(query, block, builder=builder.Builder(dna, window_length=500))
    with pytest.raises(TypeError):
        next(hmmer.phmmer([pair.hmm], block))


def test_search_seq_glue_without_a_device(monkeypatch):
    """`Pipeline.search_seq`: the query goes through the Builder and its OptimizedProfile reaches the search call; the hits
    keep the SEQUENCE as their query (the device search itself is replaced by an empty result here)."""
    abc = easel.Alphabet.amino()
    bg = plan7.Background(abc)
    rng = np.random.default_rng(9)
    query = easel.DigitalSequence(abc, name="sq", sequence=rng.integers(0, 20, 60).astype(np.uint8))
    block = easel.DigitalSequenceBlock(abc, [easel.DigitalSequence(abc, name="t", sequence=rng.integers(0, 20, 90).astype(np.uint8))])
    seen = {}

    def fake_run(self, oms, blk):
        seen["oms"] = oms
        return [], [], b"", np.zeros((len(oms), 4), np.int64)

    monkeypatch.setattr(plan7._lib, "context", lambda device=None: None)
    monkeypatch.setattr(plan7.Pipeline, "_run", fake_run)
    pli = plan7.Pipeline(abc, background=bg)
    b = builder.Builder(abc)
    state = {}
    b._scorer = _reference_scorer(state)
    calibrate = b.calibrate

    def hooked(hmm, background):
        state["hmm"] = hmm
        hmm._evparam[:] = np.array([-8, .7, -9, .7, -4, .7], np.float32)
        return calibrate(hmm, background)
    b.calibrate = hooked
    th = pli.search_seq(query, block, builder=b)
    assert th.query is query and len(th) == 0 and th.searched_sequences == 1
    om = seen["oms"][0]
    assert isinstance(om, plan7.OptimizedProfile) and om.M == 60 and om.name == "sq" and om._evparam[0] != plan7.P7_EVPARAM_UNSET
    with pytest.raises(plan7.AlphabetMismatch):
        pli.search_seq(easel.DigitalSequence(easel.Alphabet.dna(), name="d", sequence=np.zeros(5, np.uint8)), block)
