"""GPU parity of hmmscan and of the reference's committed result tables.

* `hmmer.hmmscan` / `Pipeline.scan_seq` against the reference's own scan loop (`refshim.scan` = Pipeline._scan_loop over
  oracle/_ref, plan7.pyx:6625-6677): hit list, every score, coordinates, alignments and the four pass counters PER QUERY.
* The reference's golden tables, used the way its own tests use them (src/pyhmmer/tests/test_hmmer.py): RREFam.scan.tbl
  (:823-904), RREFam.tbl / RREFam.domtbl (:161-199, c-/i-Evalues compared as %9.2g strings), PF02826.domtbl.
"""
import gzip
import itertools
import os
import tempfile

import numpy as np
import pytest

from pyhmmer_b200 import easel, plan7, synth, hmmer
from oracle import refshim

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _proteome(abc):
    with easel.SequenceFile(os.path.join(GOLD, "data", "proteome.faa.gz"), digital=True, alphabet=abc) as f:
        return f.read_block()


def _hmms(name):
    with gzip.open(os.path.join(GOLD, "data", name + ".hmm.gz")) as f:
        return list(plan7.HMMFile(f))


def _ref_models(name):
    tmp = tempfile.NamedTemporaryFile(suffix=".hmm")
    with gzip.open(os.path.join(GOLD, "data", name + ".hmm.gz")) as f:
        tmp.write(f.read())
    tmp.flush()
    out = []
    while True:
        try:
            out.append(refshim.RefModel(tmp.name, len(out), 400))
        except ValueError:
            break
    return out


def _table(name):
    rows = []
    for line in open(os.path.join(GOLD, "data", name)):
        if not line.startswith("#"):
            rows.append(line.split())
    return rows


def _txt(v):
    return v.decode() if isinstance(v, bytes) else v


def _compare_scan(ours, ref, tol=2e-3):
    """ours: (hits of ONE query sorted by profile, doms, text, counters[4]); ref: refshim.scan output."""
    hits, doms, text, counters = ours
    rh, rd, rtext, rc = ref
    assert [int(v) for v in counters] == list(rc), (list(counters), rc)
    assert [h.profile for h in hits] == [h.seq for h in rh]
    for a, b in zip(hits, rh):
        for f in ("score", "pre_score", "sum_score", "nexpected"):
            assert abs(getattr(a, f) - getattr(b, f)) <= tol, (a.profile, f, getattr(a, f), getattr(b, f))
        assert abs(a.lnP - b.lnP) <= 2e-3 * max(1.0, abs(b.lnP))
        for f in ("nregions", "nclustered", "noverlaps", "nenvelopes", "ndom", "best_domain"):
            assert getattr(a, f) == getattr(b, f), (a.profile, f, getattr(a, f), getattr(b, f))
        for d in range(a.ndom):
            x, y = doms[a.dom_offset + d], rd[b.dom_offset + d]
            for f in ("ienv", "jenv", "iali", "jali", "hmmfrom", "hmmto", "sqfrom", "sqto", "N"):
                assert getattr(x, f) == getattr(y, f), (a.profile, d, f, getattr(x, f), getattr(y, f))
            for f in ("envsc", "domcorrection", "dombias", "oasc", "bitscore"):
                assert abs(getattr(x, f) - getattr(y, f)) <= tol, (a.profile, d, f, getattr(x, f), getattr(y, f))
            n = x.N
            ta, tb = text[x.text_offset:x.text_offset + 4 * (n + 1)], rtext[y.text_offset:y.text_offset + 4 * (n + 1)]
            assert ta[:3 * (n + 1)] == tb[:3 * (n + 1)], (a.profile, d)
            pa, pb = ta[3 * (n + 1):], tb[3 * (n + 1):]
            off = [i for i in range(len(pa)) if pa[i] != pb[i]]
            cls = b"0123456789*"
            assert len(off) <= max(1, n // 100) and all(abs(cls.index(pa[i]) - cls.index(pb[i])) == 1 for i in off), (a.profile, d, off[:8])


def test_scan_matches_reference_scan_loop(amino):
    """Several query sequences at once against the fixture models (RREFam + PF02826 + Thioesterase + KR + LuxC): every
    comparison p7_Pipeline scores to completion, field by field, and the pass counters of EVERY query (the batched scan
    keeps them per sequence) equal the reference's scan loop run once per query."""
    seqs = _proteome(amino)
    names = ("RREFam", "PF02826", "Thioesterase", "KR", "LuxC")
    hmms = [h for n in names for h in _hmms(n)]
    refs = [m for n in names for m in _ref_models(n)]
    assert len(hmms) == len(refs) == 14
    byname = {_txt(s.name): s for s in seqs}
    picked = [byname[n] for n in ("938293.PRJEB85.HG003691_78", "938293.PRJEB85.HG003686_714", "938293.PRJEB85.HG003686_827",
                                  "938293.PRJEB85.HG003684_52")]
    picked += sorted(seqs, key=len)[-3:] + [seqs[0], seqs[1000]]
    pli = plan7.Pipeline(amino)
    oms = pli._optimized_many(hmms, 100)
    block = easel.DigitalSequenceBlock(amino, picked)
    hits, doms, text, counters = pli._run(oms, block, seq_counters=True)
    assert counters.shape == (len(picked), 4)
    total = 0
    for si, q in enumerate(picked):
        mine = sorted((h for h in hits if h.seq == si), key=lambda h: h.profile)
        ref = refshim.scan(refs, q.sequence)
        _compare_scan((mine, doms, text, counters[si]), ref)
        total += len(mine)
    assert total >= 8
    # the public entry point reports the same per-query counters and Z = number of models
    for si, th in enumerate(hmmer.hmmscan(picked, hmms)):
        assert (th.n_past_msv, th.n_past_bias, th.n_past_vit, th.n_past_fwd) == tuple(int(v) for v in counters[si])
        assert th.Z == len(hmms) and th.searched_models == len(hmms) and th.searched_residues == len(picked[si])
    # one query at a time (Pipeline.scan_seq) gives the very same records
    one = plan7.Pipeline(amino).scan_seq(picked[0], plan7.OptimizedProfileBlock(amino, oms))
    assert (one.n_past_msv, one.n_past_bias, one.n_past_vit, one.n_past_fwd) == tuple(int(v) for v in counters[0])
    assert sorted(_txt(h.name) for h in one) == sorted(_txt(hmms[h.profile].name) for h in hits if h.seq == 0 and
                                                       np.exp(h.lnP) * (h.profile + 1) <= 10.0)


def test_rrefam_scan_table(amino):
    """hmmscan of the whole proteome (2 100 queries) against RREFam: exactly the rows of the reference's golden
    RREFam.scan.tbl (HMMER CLI output; test_hmmer.py:823-904 compares name, score, bias, E-value) and no reported hit for
    any other query."""
    expected = {}
    for q, rows in itertools.groupby(_table("RREFam.scan.tbl"), key=lambda r: r[2]):
        expected[q] = list(rows)
    assert len(expected) == 7 and sum(len(v) for v in expected.values()) == 10
    seqs = _proteome(amino)
    hmms = _hmms("RREFam")
    seen = 0
    for th in hmmer.hmmscan(seqs, hmms):
        rows = expected.get(_txt(th.query.name))
        rep = [h for h in th if h.reported]
        if rows is None:
            assert len(rep) == 0, th.query.name
            continue
        seen += 1
        assert len(rep) == len(rows), th.query.name
        for h, r in zip(rep, rows):
            assert _txt(h.name) == r[0] and _txt(h.accession) == r[1]
            assert abs(h.score - float(r[5])) <= 0.1 and abs(h.bias - float(r[6])) <= 0.1
            assert "%9.2g" % h.evalue == "%9.2g" % float(r[4]), (h.name, h.evalue, r[4])
            d = h.best_domain
            assert abs(d.score - float(r[8])) <= 0.1 and "%9.2g" % d.i_evalue == "%9.2g" % float(r[7])
            assert "%.1f" % h._rec.nexpected == r[10] and [h._rec.nregions, h._rec.nclustered, h._rec.noverlaps, h._rec.nenvelopes,
                                                           h._rec.ndom] == [int(v) for v in r[11:16]]
            assert [len(h.domains.reported), len(h.domains.included)] == [int(r[16]), int(r[17])]
    assert seen == len(expected)


def test_rrefam_search_tables(amino):
    """hmmsearch of the 10 RREFam models against the proteome vs RREFam.tbl and RREFam.domtbl, as the reference's own
    test_rrefam does (test_hmmer.py:161-199): hits in table order; domain c-/i-Evalues equal as %9.2g strings."""
    seqs = _proteome(amino)
    all_hits = list(hmmer.hmmsearch(_hmms("RREFam"), seqs))
    hits = [h for th in all_hits for h in th]
    rows = _table("RREFam.tbl")
    assert len(hits) == len(rows)
    for h, r in zip(hits, rows):
        assert _txt(h.name) == r[0]
        assert (h.accession is None) if r[1] == "-" else (_txt(h.accession) == r[1])
        assert abs(h.score - float(r[5])) <= 0.1 and abs(h.bias - float(r[6])) <= 0.1 and abs(h.evalue - float(r[4])) <= 0.1
        assert "%9.2g" % h.evalue == "%9.2g" % float(r[4])
    doms = [d for th in all_hits for h in th for d in h.domains]
    drows = _table("RREFam.domtbl")
    assert len(doms) == len(drows)
    for d, r in zip(doms, drows):
        assert d.hit.hits.Z == len(seqs)
        assert _txt(d.hit.name) == r[0]
        assert round(abs(d.score - float(r[13])), 1) == 0 and round(abs(d.bias - float(r[14])), 1) == 0
        assert "%9.2g" % d.c_evalue == "%9.2g" % float(r[11]), (d.hit.name, d.c_evalue, r[11])
        assert "%9.2g" % d.i_evalue == "%9.2g" % float(r[12]), (d.hit.name, d.i_evalue, r[12])
        assert [d.alignment.hmm_from, d.alignment.hmm_to, d.alignment.target_from, d.alignment.target_to, d.env_from, d.env_to] == \
               [int(v) for v in r[15:21]]
        assert "%4.2f" % (d._rec.oasc / (1.0 + abs(float(d.env_to - d.env_from)))) == r[21]


def test_pf02826_domain_table(amino):
    """PF02826 vs the proteome: the reported domains are the rows of the reference's PF02826.domtbl, coordinate by coordinate."""
    th = plan7.Pipeline(amino).search_hmm(_hmms("PF02826")[0], _proteome(amino))
    doms = [d for h in th if h.reported for d in h.domains if d.reported]
    drows = _table("PF02826.domtbl")
    assert len(doms) == len(drows)
    for d, r in zip(doms, drows):
        assert _txt(d.hit.name) == r[0] and int(r[2]) == d.hit.length and int(r[5]) == 178
        assert abs(d.hit.score - float(r[7])) <= 0.1 and abs(d.score - float(r[13])) <= 0.1 and abs(d.bias - float(r[14])) <= 0.1
        assert "%9.2g" % d.c_evalue == "%9.2g" % float(r[11]) and "%9.2g" % d.i_evalue == "%9.2g" % float(r[12])
        assert [d.alignment.hmm_from, d.alignment.hmm_to, d.alignment.target_from, d.alignment.target_to, d.env_from, d.env_to] == \
               [int(v) for v in r[15:21]]


def test_written_tables_equal_golden_files(amino):
    """`TopHits.write` reproduces the data rows of the committed HMMER tables byte for byte where the printed precision
    allows (everything but the last digit of a score that sits on a rounding boundary): PF02826.tbl / .domtbl."""
    import io
    th = plan7.Pipeline(amino).search_hmm(_hmms("PF02826")[0], _proteome(amino))
    for fmt, name in (("targets", "PF02826.tbl"), ("domains", "PF02826.domtbl")):
        buf = io.BytesIO()
        th.write(buf, format=fmt, header=False)
        ours = [l.split() for l in buf.getvalue().decode().splitlines() if not l.startswith("#")]
        gold = _table(name)
        assert len(ours) == len(gold)
        ndiff = 0
        for a, b in zip(ours, gold):
            assert len(a) == len(b) and a[0] == b[0]
            for x, y in zip(a, b):
                if x != y:
                    assert abs(float(x) - float(y)) <= 0.1000001 * max(1.0, abs(float(y)) * 0.1), (a[0], x, y)
                    ndiff += 1
        assert ndiff <= 4


def test_empty_sequences_are_skipped(amino):
    """p7_Pipeline returns at once for a target of length 0 (p7_pipeline.c:713): it passes no filter, it is not a hit, but
    it counts in Z (p7_pli_NewSeq).  Search and scan orientation."""
    seqs = _proteome(amino)
    hmm = _hmms("PF02826")[0]
    sub = easel.DigitalSequenceBlock(amino, list(seqs[:400]))
    base = plan7.Pipeline(amino).search_hmm(hmm, sub)
    empty = lambda i: easel.DigitalSequence(amino, name=b"empty%d" % i, sequence=np.zeros(0, np.uint8))
    mixed = easel.DigitalSequenceBlock(amino, [empty(0)] + list(seqs[:200]) + [empty(1), empty(2)] + list(seqs[200:400]) + [empty(3)])
    th = plan7.Pipeline(amino, Z=float(len(sub))).search_hmm(hmm, mixed)
    ref = plan7.Pipeline(amino, Z=float(len(sub))).search_hmm(hmm, sub)
    assert (th.n_past_msv, th.n_past_bias, th.n_past_vit, th.n_past_fwd) == (base.n_past_msv, base.n_past_bias, base.n_past_vit, base.n_past_fwd)
    assert [(h.name, h.score, len(h.domains)) for h in th] == [(h.name, h.score, len(h.domains)) for h in ref]
    assert plan7.Pipeline(amino).search_hmm(hmm, mixed).Z == len(mixed)
    got = list(hmmer.hmmscan([empty(0), seqs[5]], _hmms("RREFam")))
    assert len(got[0]) == 0 and (got[0].n_past_msv, got[0].n_past_fwd) == (0, 0)


def test_reduced_c3_slice_counters(amino):
    """A stated slice of BASELINE configs[2] (20 000 Pfam-A-sized profiles x 100 000 proteins): 400 profiles with the
    Pfam-like length distribution of the full set (median ~150, up to 2 300 nodes, so every kernel family runs -- register
    tiles of all three group widths, the multi-warp and the shared-memory DP classes) x 20 000 proteins with planted
    homologs.  The four pass counters of all 8 million comparisons and the number of comparisons scored to completion equal
    the reference pipeline's (oracle/_ref, all host threads)."""
    import psutil
    import bench_inputs
    models = bench_inputs.make_models(400, seed=31, median_M=150.0, sigma=0.75, lo=20, hi=2300)
    seqs_arr = bench_inputs.make_sequences(20000, seed=32)
    bench_inputs.plant(seqs_arr, models, 200, seed=33)
    hmms = bench_inputs.to_hmms(models, amino)
    block = bench_inputs.to_block(seqs_arr, amino)
    assert max(h.M for h in hmms) > 1536 and min(h.M for h in hmms) <= 40
    pli = plan7.Pipeline(amino)
    oms = pli._optimized_many(hmms, 350)
    hits, doms, text, counters = pli._run(oms, block)
    refs = bench_inputs.to_ref_models(models, nthreads=psutil.cpu_count(logical=True) or 8)
    nh, ctr = refshim.search_mt(refs, seqs_arr, psutil.cpu_count(logical=True) or 8)
    assert counters.sum(0).tolist() == ctr
    assert len(hits) == nh and nh >= 150


_ORIENT_WORKER = r'''
import json, sys
import numpy as np
sys.path.insert(0, %r)
from pyhmmer_b200 import easel, plan7, synth
abc = easel.Alphabet.amino()
rng = np.random.default_rng(2024)
Ms = (1, 7, 60, 254, 255, 256, 300, 511, 512, 700, 1023, 1024, 1500, 2300)
hmms = [synth.random_hmm(abc, M, rng, name="m%%d" %% M) for M in Ms]
lens = (0, 1, 2, 37, 350, 1027, 5000, 12000, 4096)
seqs = []
for i, L in enumerate(lens):
    s = rng.integers(0, 20, L).astype(np.uint8)
    for j in range(L // 900):                                # homologs of various models all along the long ones
        d = synth.emit_sequence(hmms[(i + 3 * j) %% len(hmms)], rng)
        pos = int(rng.integers(0, max(1, L - len(d))))
        s[pos:pos + len(d)] = d[:L - pos]
    seqs.append(easel.DigitalSequence(abc, name="s%%d" %% i, sequence=s))
block = easel.DigitalSequenceBlock(abc, seqs)
pli = plan7.Pipeline(abc)
oms = [pli._optimized(h, 100) for h in hmms]
hits, doms, text, counters = pli._run(oms, block, seq_counters=True)
hits2, doms2, text2, pcounters = pli._run(oms, block)
print(json.dumps({"seq_counters": counters.tolist(), "profile_counters": pcounters.tolist(),
                  "hits": [(h.profile, h.seq, round(float(h.score), 4), h.ndom) for h in hits]}))
'''


def test_scan_orientation_equals_search_orientation():
    """The chunked SSV pass of the scan orientation (few, long sequences: every sequence cut into overlapping chunks whose
    maxima are folded per sequence) decides exactly what the search-orientation kernel decides: same pass counters per
    sequence and per profile, same hits -- models at every overlap-class boundary, sequences from 0 to 12 000 residues."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for env in ({}, {"B2H_SCAN_MAXSEQ": "0"}):
        r = subprocess.run([sys.executable, "-c", _ORIENT_WORKER % (root,)], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert outs[0] == outs[1]
    assert len(outs[0]["hits"]) >= 10 and sum(c[0] for c in outs[0]["seq_counters"]) > len(outs[0]["hits"])
