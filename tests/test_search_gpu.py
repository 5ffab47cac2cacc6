"""GPU parity of the whole search path: Pipeline.search_hmm / hmmsearch / hmmscan vs (a) the reference's own
search loop run through oracle/_ref on the same inputs, (b) the committed golden results produced by the
reference Python package (tests/golden/hmmsearch.json) and the HMMER CLI tables the reference ships."""
import gzip
import json
import os
import tempfile

import numpy as np
import pytest

from pyhmmer_b200 import _lib, easel, plan7, synth, hmmer
from oracle import refshim

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _proteome(abc):
    with easel.SequenceFile(os.path.join(GOLD, "data", "proteome.faa.gz"), digital=True, alphabet=abc) as f:
        return f.read_block()


def _hmms(name):
    with gzip.open(os.path.join(GOLD, "data", name + ".hmm.gz")) as f:
        return list(plan7.HMMFile(f))


def _ref_model(name, index=0):
    tmp = tempfile.NamedTemporaryFile(suffix=".hmm")
    with gzip.open(os.path.join(GOLD, "data", name + ".hmm.gz")) as f:
        tmp.write(f.read())
    tmp.flush()
    m = refshim.RefModel(tmp.name, index, 400)
    m._tmp = tmp
    return m


def _compare_with_ref(th_raw, ref_out, tol=2e-3):
    hits, doms, text, counters = th_raw
    rh, rd, rtext, rc = ref_out
    assert list(counters) == list(rc), (counters, rc)
    assert [h.seq for h in hits] == [h.seq for h in rh]
    for a, b in zip(hits, rh):
        for f in ("score", "pre_score", "sum_score", "nexpected"):
            assert abs(getattr(a, f) - getattr(b, f)) <= tol, (a.seq, f, getattr(a, f), getattr(b, f))
        for f in ("nregions", "nclustered", "noverlaps", "nenvelopes", "ndom", "best_domain"):
            assert getattr(a, f) == getattr(b, f), (a.seq, f, getattr(a, f), getattr(b, f))
        for d in range(a.ndom):
            x, y = doms[a.dom_offset + d], rd[b.dom_offset + d]
            for f in ("ienv", "jenv", "iali", "jali", "hmmfrom", "hmmto", "sqfrom", "sqto", "N"):
                assert getattr(x, f) == getattr(y, f), (a.seq, d, f, getattr(x, f), getattr(y, f))
            for f in ("envsc", "domcorrection", "dombias", "oasc", "bitscore"):
                assert abs(getattr(x, f) - getattr(y, f)) <= tol, (a.seq, d, f, getattr(x, f), getattr(y, f))
            n = x.N
            assert text[x.text_offset:x.text_offset + 4 * (n + 1)] == rtext[y.text_offset:y.text_offset + 4 * (n + 1)]


@pytest.mark.parametrize("name", ["PF02826", "Thioesterase", "KR", "LuxC"])
def test_search_matches_reference_loop(amino, name):
    seqs = _proteome(amino)
    hmm = _hmms(name)[0]
    pli = plan7.Pipeline(amino)
    om = pli._optimized(hmm, len(seqs[0]))
    raw = pli._run([om], seqs)
    ref = _ref_model(name)
    out = ref.search([s.sequence for s in seqs])
    _compare_with_ref((raw[0], raw[1], raw[2], raw[3][0]), out)


def test_golden_hmmsearch_json(amino):
    """Every fixture HMM vs the proteome: the hit list pyhmmer itself produced (hmmsearch.json)."""
    gold = json.load(open(os.path.join(GOLD, "hmmsearch.json")))
    seqs = _proteome(amino)
    queries = [h for n in ("PF02826", "Thioesterase", "KR", "LuxC", "RREFam") for h in _hmms(n)]
    results = list(hmmer.hmmsearch(queries, seqs))
    assert len(results) == len(queries) == len(gold)
    for q, th in zip(queries, results):
        g = gold[q.name]
        assert [th.n_past_msv, th.n_past_bias, th.n_past_vit, th.n_past_fwd] == g["counters"], q.name
        assert len(th) == g["n_hits"], q.name
        assert th.Z == g["Z"] and th.domZ == g["domZ"]
        for h, gh in zip(th, g["hits"]):
            assert h.name == gh["name"]
            assert abs(h.score - gh["score"]) < 2e-3 and abs(h.bias - gh["bias"]) < 2e-3
            assert abs(h.evalue - gh["evalue"]) <= 1e-3 * gh["evalue"] + 1e-300
            assert h.reported == gh["reported"] and h.included == gh["included"]
            assert len(h.domains) == len(gh["domains"])
            for d, gd in zip(h.domains, gh["domains"]):
                assert (d.env_from, d.env_to) == (gd["env_from"], gd["env_to"])
                assert abs(d.score - gd["score"]) < 2e-3 and abs(d.bias - gd["bias"]) < 2e-3
                assert abs(d.c_evalue - gd["c_evalue"]) <= 1e-3 * gd["c_evalue"]
                assert abs(d.i_evalue - gd["i_evalue"]) <= 1e-3 * gd["i_evalue"]
                assert (d.reported, d.included) == (gd["reported"], gd["included"])
                a = d.alignment
                assert (a.hmm_from, a.hmm_to, a.target_from, a.target_to) == (gd["hmm_from"], gd["hmm_to"], gd["target_from"], gd["target_to"])
                assert a.hmm_sequence == gd["hmm_sequence"] and a.target_sequence == gd["target_sequence"]
                assert a.identity_sequence == gd["identity_sequence"] and a.posterior_probabilities == gd["posterior_probabilities"]


def test_pf02826_table(amino):
    """The reference's own golden table (HMMER CLI output, tests/data/tables/PF02826.tbl): 22 hits, names/scores/E-values."""
    rows = [l.split() for l in open(os.path.join(GOLD, "data", "PF02826.tbl")) if not l.startswith("#")]
    th = plan7.Pipeline(amino).search_hmm(_hmms("PF02826")[0], _proteome(amino))
    assert len(th) == len(rows) == 22
    for h, r in zip(th, rows):
        assert h.name == r[0]
        assert abs(h.score - float(r[5])) <= 0.1 and abs(h.bias - float(r[6])) <= 0.1
        assert "%9.2g" % h.evalue == "%9.2g" % float(r[4])


def test_synthetic_with_planted_homologs(amino):
    """Calibrated synthetic profiles vs random targets with planted domains: all stages exercised."""
    rng = np.random.default_rng(77)
    hmms = [synth.random_hmm(amino, M, rng, name="syn%d" % i) for i, M in enumerate((45, 130, 210, 330))]
    synth.calibrate(hmms)
    seqs = synth.random_sequences(amino, 3000, rng)
    for i in range(60):
        h = hmms[i % len(hmms)]
        dom = synth.emit_sequence(h, rng)
        s = seqs[int(rng.integers(0, len(seqs)))]
        cut = int(rng.integers(0, len(s)))
        parts = [s.sequence[:cut], dom] + ([s.sequence[cut:cut + 30], synth.emit_sequence(h, rng)] if i % 4 == 0 else []) + [s.sequence[cut:]]
        s.sequence = np.concatenate(parts)[:1500]
    seqs._cache = {}
    pli = plan7.Pipeline(amino)
    with tempfile.TemporaryDirectory() as td:
        nhit = 0
        for i, h in enumerate(hmms):
            path = os.path.join(td, "m%d.hmm" % i)
            with open(path, "wb") as f:
                h.write(f)
            with plan7.HMMFile(path) as f:
                h2 = f.read()
            om = pli._optimized(h2, len(seqs[0]))
            raw = pli._run([om], seqs)
            ref = refshim.RefModel(path, 0, 400)
            out = ref.search([s.sequence for s in seqs])
            _compare_with_ref((raw[0], raw[1], raw[2], raw[3][0]), out)
            nhit += len(out[0])
        assert nhit >= 40


def test_hmmscan_matches_search(amino):
    """hmmscan of one sequence against a profile block == the same comparisons via hmmsearch (Z = #models)."""
    seqs = _proteome(amino)
    hmms = _hmms("RREFam") + _hmms("PF02826")
    target = max(seqs, key=len)
    th = list(hmmer.hmmscan([target], hmms))[0]
    assert th.Z == len(hmms)
    names = {h.name for h in th}
    for q in hmms:
        r = plan7.Pipeline(amino, Z=len(hmms)).search_hmm(q, easel.DigitalSequenceBlock(amino, [target]))
        assert (len(r) > 0) == (q.name in names)
        if len(r):
            assert abs(r[0].score - [h for h in th if h.name == q.name][0].score) < 1e-4
