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
            ta, tb = text[x.text_offset:x.text_offset + 4 * (n + 1)], rtext[y.text_offset:y.text_offset + 4 * (n + 1)]
            # the four rows (model, match line, target, posterior line) are n+1 bytes each.  The first three must be
            # identical; a posterior digit is a float compared with .05/.15/... (p7_alidisplay.c: p7_alidisplay_EncodePostProb),
            # so a value within float tolerance of a class boundary may land in the neighbouring class.
            assert ta[:3 * (n + 1)] == tb[:3 * (n + 1)], (a.seq, d)
            pa, pb = ta[3 * (n + 1):], tb[3 * (n + 1):]
            off = [i for i in range(len(pa)) if pa[i] != pb[i]]
            cls = b"0123456789*"
            assert len(off) <= max(1, n // 100) and all(abs(cls.index(pa[i]) - cls.index(pb[i])) == 1 for i in off), (a.seq, d, off[:8])


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


def test_search_wave_by_wave_equals_one_call(amino, monkeypatch):
    """b2h_search_begin / _next / _end (results handed out wave by wave while the following waves are searched) against
    the blocking b2h_search: the same records, every profile final in exactly one wave, waves in order; `Pipeline._search_many`
    (which assembles `TopHits` per wave) against `_run` + `_assemble`; an early exit leaves the engine usable."""
    rng = np.random.default_rng(78)
    Ms = [int(m) for m in rng.integers(25, 420, 24)] + [560, 700]            # two multi-warp (W = 2) models as well
    hmms = [synth.random_hmm(amino, M, rng, name="wv%d" % i) for i, M in enumerate(Ms)]
    synth.calibrate(hmms)
    seqs = synth.random_sequences(amino, 2500, rng)
    for i in range(120):
        s = seqs[int(rng.integers(0, len(seqs)))]
        cut = int(rng.integers(0, len(s)))
        s.sequence = np.concatenate([s.sequence[:cut], synth.emit_sequence(hmms[i % len(hmms)], rng), s.sequence[cut:]])[:1500]
    seqs._cache = {}
    pli = plan7.Pipeline(amino)
    oms = [pli._optimized(h, len(seqs[0])) for h in hmms]
    for nwaves in (1, 3, 5):
        monkeypatch.setenv("B2H_WAVES", str(nwaves))
        hits, doms, text, counters = pli._run(oms, seqs)
        n, gen = pli._run_waves(oms, seqs)
        assert n == nwaves
        seen, got, waves = [], [], 0
        for profs, wh, wd, wt, wc in gen:
            waves += 1
            assert sorted({h.profile for h in wh}) <= sorted(profs) and not set(profs) & set(seen)
            assert [(h.profile, h.seq) for h in wh] == sorted((h.profile, h.seq) for h in wh)
            assert all(oms[a].M >= oms[b].M for a in seen for b in profs)        # longest models first
            assert np.array_equal(wc[profs], counters[profs]) and not wc[[p for p in range(len(oms)) if p not in profs]].any()
            seen += profs
            for h in wh:
                d0, r0 = wd[h.dom_offset], None
                got.append((h.profile, h.seq, round(float(h.score), 3), h.ndom, d0.ienv, d0.jenv, wt[d0.text_offset:d0.text_offset + d0.N + 1]))
        assert waves == nwaves and sorted(seen) == list(range(len(oms)))
        want = [(h.profile, h.seq, round(float(h.score), 3), h.ndom, doms[h.dom_offset].ienv, doms[h.dom_offset].jenv,
                 text[doms[h.dom_offset].text_offset:doms[h.dom_offset].text_offset + doms[h.dom_offset].N + 1]) for h in hits]
        assert sorted(got) == want and len(want) >= 100
        a = pli._search_many(hmms, seqs)
        b = pli._assemble(hmms, oms, seqs, hits, doms, text, counters)
        sig = lambda ths: [[(h.name, round(h.score, 3), h.reported, h.included, len(h.domains)) for h in th] + [(th.Z, th.n_past_msv, th.n_past_fwd)] for th in ths]
        assert sig(a) == sig(b)
    # leaving the generator early drains the job; the next search is unaffected
    n, gen = pli._run_waves(oms, seqs)
    next(gen)
    gen.close()
    again = pli._run(oms, seqs)
    assert [(h.profile, h.seq) for h in again[0]] == [(h.profile, h.seq) for h in hits]


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


def test_full_size_counters_match_reference(amino):
    """BASELINE configs[1] at full size (100 profiles x 50 000 sequences, bench.py's seeded generator): the four pipeline
    pass counters and the number of comparisons scored to completion are identical to the reference's p7_Pipeline
    (oracle/_ref, all host threads) -- i.e. every filter decision of 5 million comparisons agrees -- and a sample of the
    queries is compared hit by hit."""
    import bench
    import psutil
    abc, hmms, seqs = bench.build_inputs(0, 1)
    assert bench.apply_stats(hmms)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "q.hmm")
        bench.write_hmm_file(hmms, path)
        with plan7.HMMFile(path) as f:
            hmms2 = list(f)
        pli = plan7.Pipeline(abc)
        oms = [pli._optimized(h, len(seqs[0])) for h in hmms2]
        hits, doms, text, counters = pli._run(oms, seqs)
        models = [refshim.RefModel(path, i, 400) for i in range(len(hmms))]
        codes = [s.sequence for s in seqs]
        nh, ctr = refshim.search_mt(models, codes, psutil.cpu_count(logical=True) or 8)
        assert counters.sum(0).tolist() == ctr
        assert len(hits) == nh
        for qi in (0, 17, 63, 99):                                  # every field of every hit for a few queries
            raw = pli._run([oms[qi]], seqs)
            assert len(raw[0]) == sum(1 for h in hits if h.profile == qi)
            _compare_with_ref((raw[0], raw[1], raw[2], raw[3][0]), models[qi].search(codes))


def test_hmmscan_over_pressed_database(amino):
    """hmmscan against profiles read straight from the reference's hmmpress'ed fixtures (HMMPressedFile) gives the same
    hits as against the ASCII models converted on our side."""
    seqs = _proteome(amino)
    queries = sorted(seqs, key=len)[-3:]
    names = ("PF02826", "Thioesterase")
    pressed = []
    for n in names:
        with plan7.HMMPressedFile(os.path.join(GOLD, "data", "pressed", n + ".hmm")) as pf:
            pressed += list(pf)
    ascii_models = [h for n in names for h in _hmms(n)]
    for a, b in zip(hmmer.hmmscan(queries, pressed), hmmer.hmmscan(queries, ascii_models)):
        assert [(h.name, round(h.score, 4), len(h.domains)) for h in a] == [(h.name, round(h.score, 4), len(h.domains)) for h in b]
    # and the search orientation: PF02826 finds its 22 golden hits from the pressed profile too
    th = plan7.Pipeline(amino).search_hmm(pressed[0], seqs)
    assert len(th) == 22
