"""CPU: pin the oracle.  (1) oracle/_ref (the reference's own C code) reproduces the committed golden results that
the reference Python package produced (tests/golden/*.json) and the HMMER CLI tables the reference ships;
(2) the scalar C restatement oracle/hmmer_oracle.c agrees with oracle/_ref bit-for-bit on the integer filters and
to 1e-4 nats on Forward."""
import gzip
import json
import os
import tempfile

import numpy as np
import pytest

from pyhmmer_b200 import easel, plan7, synth
from oracle import refshim, port

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
pytestmark = pytest.mark.skipif(not refshim.available(), reason="oracle/_ref not built (make -C oracle)")


@pytest.fixture(scope="module")
def proteome(amino):
    with easel.SequenceFile(os.path.join(GOLD, "data", "proteome.faa.gz"), digital=True, alphabet=amino) as f:
        return f.read_block()


def _ref(name, index=0):
    tmp = tempfile.NamedTemporaryFile(suffix=".hmm")
    with gzip.open(os.path.join(GOLD, "data", name + ".hmm.gz")) as f:
        tmp.write(f.read())
    tmp.flush()
    m = refshim.RefModel(tmp.name, index, 400)
    m._tmp = tmp
    return m


def test_ref_reproduces_golden_msv_scores(proteome):
    gold = json.load(open(os.path.join(GOLD, "filters.json")))
    for name, rec in gold.items():
        ref = _ref(rec["file"])
        for s, g in zip(proteome[:300], rec["msv"]):
            sc, st = ref.msv(s.sequence)
            if g == "inf":
                assert np.isinf(sc)
            else:
                assert sc == np.float32(g), (name, s.name, sc, g)


@pytest.mark.parametrize("name,index", [("PF02826", 0), ("KR", 0), ("Thioesterase", 0), ("RREFam", 1)])
def test_ref_reproduces_golden_hmmsearch(proteome, name, index):
    gold = json.load(open(os.path.join(GOLD, "hmmsearch.json")))
    ref = _ref(name, index)
    g = gold[ref.name.decode()]
    hits, doms, text, counters = ref.search([s.sequence for s in proteome])
    assert counters == g["counters"]
    # p7_Pipeline admits with the running Z; opened wide here, so compare the subset the golden run reported
    by_name = {proteome[h.seq].name: h for h in hits}
    assert len(g["hits"]) <= len(hits)
    for gh in g["hits"]:
        h = by_name[gh["name"]]
        assert abs(h.score - gh["score"]) < 1e-4 and len(gh["domains"]) == h.ndom
        for d, gd in zip(doms[h.dom_offset:h.dom_offset + h.ndom], gh["domains"]):
            assert (d.ienv, d.jenv, d.hmmfrom, d.hmmto, d.sqfrom, d.sqto) == (gd["env_from"], gd["env_to"], gd["hmm_from"], gd["hmm_to"], gd["target_from"], gd["target_to"])
            assert abs(d.bitscore - gd["score"]) < 1e-4


def test_ref_matches_cli_table(proteome):
    rows = [l.split() for l in open(os.path.join(GOLD, "data", "PF02826.tbl")) if not l.startswith("#")]
    ref = _ref("PF02826")
    hits, doms, text, counters = ref.search([s.sequence for s in proteome])
    got = {proteome[h.seq].name: h.score for h in hits}
    assert len(rows) == 22
    for r in rows:
        assert abs(got[r[0]] - float(r[5])) <= 0.1


@pytest.mark.parametrize("M", [1, 2, 9, 17, 64, 130, 300])
def test_port_matches_ref(amino, make_pair, M):
    rng = np.random.default_rng(500 + M)
    pair = make_pair(synth.random_hmm(amino, M, rng))
    po = port.Port(pair.om)
    seqs = [rng.integers(0, amino.K, int(L)).astype(np.uint8) for L in (1, 2, 3, 15, 16, 17, 40, 100, 250, 400)]
    seqs += [np.concatenate([rng.integers(0, 20, 30).astype(np.uint8), synth.emit_sequence(pair.hmm, rng), rng.integers(0, 20, 25).astype(np.uint8)]) for _ in range(8)]
    seqs += [np.concatenate([synth.emit_sequence(pair.hmm, rng), rng.integers(0, 20, 40).astype(np.uint8), synth.emit_sequence(pair.hmm, rng)]) for _ in range(4)]
    seqs.append(rng.integers(amino.K + 1, amino.Kp - 2, 60).astype(np.uint8))
    for c in seqs:
        if len(c) == 0:
            continue
        assert po.msv(c) == pair.ref.msv(c)
        a, b = po.ssv(c), pair.ref.ssv(c)
        assert a[1] == b[1] and (a[1] == 19 or a[0] == b[0])
        assert po.vit(c) == pair.ref.vit(c)
        fa, fb = po.fwd(c), pair.ref.fwd(c)
        assert fa[1] == fb[1] == 0 and abs(fa[0] - fb[0]) <= 1e-4 + 2e-7 * abs(fb[0])
        assert po.null1(len(c)) == pair.ref.null1(c)


@pytest.mark.parametrize("M", [9, 120, 700])
def test_port_longtarget_scan_matches_ref(make_pair, M):
    """oracle_ssv_longtarget (the scalar restatement of p7_SSVFilter_longtarget) against the reference itself: every
    diagonal of a chunk with planted homologs -- start, model end, length, score -- identical, in order."""
    from pyhmmer_b200 import easel
    dna = easel.Alphabet.dna()
    rng = np.random.default_rng(M)
    h = synth.random_hmm(dna, M, rng, name="lt")
    h.max_length = 3 * M
    h._evparam[:] = np.array([-8.0 - np.log2(M) * 0.3, 0.70, -9.0, 0.70, -4.0, 0.70], np.float32)
    pair = make_pair(h)
    chunk = rng.integers(0, 4, 60000).astype(np.uint8)
    for _ in range(12):
        dom = synth.emit_sequence(pair.hmm, rng)
        pos = int(rng.integers(0, len(chunk) - len(dom)))
        chunk[pos:pos + len(dom)] = dom
    rraw, rsc, _, _, _ = pair.ref.longtarget_windows(chunk)
    w, sc = port.Port(pair.om).ssv_longtarget(chunk)
    assert len(rraw) >= 10
    assert np.array_equal(w, rraw) and np.array_equal(sc, rsc)
