"""Alignment queries and jackhmmer on the host side, no device needed:

* `Builder.build_msa` (pyhmmer_b200/msabuild.py) against the reference's p7_Builder on seeded random alignments -- relative
  weights bit for bit, the HMM file line for line (DATE aside), statistics to 3e-4 (the calibration filters of this suite come
  from the reference scoring OUR model through its five-decimal ASCII file);
* the jackhmmer loop (`IterativeSearch`: build, search, rank, align with the query first, rebuild) against golden iterations
  recorded from the reference's `Pipeline.iterate_seq` / `iterate_hmm` (tests/golden/make_jackhmmer_golden.py).  The searches of
  this test are done by the reference's p7_Pipeline (oracle/_ref); the GPU suite runs the same loop on the engine.
"""
import gzip
import io
import json
import os
import sys
import tempfile

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pyhmmer_b200 import _lib, builder, easel, msabuild, plan7
from oracle import refshim
from test_builder_cpu import _reference_scorer

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not refshim.available(), reason="oracle/_ref not built")


def hook_reference_calibration(b):
    """The builder's calibration filters from the reference (the product runs them on the GPU)."""
    state = {}
    b._scorer = _reference_scorer(state)
    calibrate = b.calibrate

    def hooked(hmm, background):
        state["hmm"] = hmm
        if hmm._evparam[0] == plan7.P7_EVPARAM_UNSET:      # the file the scorer writes needs some statistics to be readable
            hmm._evparam[:] = np.array([-8, .7, -9, .7, -4, .7], np.float32)
        return calibrate(hmm, background)
    b.calibrate = hooked
    return b


def random_alignment(abc, nseq, alen, seed):
    """Rows over a random consensus: substitutions, gappy (insert) columns, deletion runs, degenerate residues, fragments."""
    rng = np.random.default_rng(seed)
    K, Kp = abc.K, abc.Kp
    cons = rng.integers(0, K, alen)
    inscol = rng.random(alen) < 0.2
    rows = np.empty((nseq, alen), np.uint8)
    for i in range(nseq):
        r = np.where(rng.random(alen) < 0.7, cons, rng.integers(0, K, alen))
        gap = np.where(inscol, rng.random(alen) < 0.85, rng.random(alen) < 0.08)
        runs = np.zeros(alen, bool)
        for j in np.nonzero(rng.random(alen) < 0.03)[0]:
            runs[j:j + rng.integers(1, 6)] = True
        r = np.where(gap | runs, K, r)
        if rng.random() < 0.3:
            r[rng.integers(0, alen)] = Kp - 3              # X / N
        if K == 20 and rng.random() < 0.2:
            r[rng.integers(0, alen)] = 21                  # B = {D, N}
        if rng.random() < 0.2:                             # a fragment
            a = int(rng.integers(0, alen // 2))
            b = a + int(rng.integers(3, alen // 3))
            r[:a] = K
            r[b:] = K
        rows[i] = r
    return rows


@needs_ref
@pytest.mark.parametrize("abcname,nseq,alen,seed,kw", [
    ("amino", 12, 60, 1, {}), ("amino", 40, 150, 2, {}), ("dna", 25, 120, 3, {}), ("amino", 30, 100, 4, dict(architecture="hand")),
    ("amino", 30, 100, 5, dict(effective_number="none")), ("amino", 30, 100, 6, dict(effective_number=7.5)),
    ("amino", 30, 100, 7, dict(prior_scheme="laplace")), ("amino", 1, 50, 8, {}), ("amino", 200, 300, 9, {}),
    ("amino", 25, 80, 10, dict(symfrac=0.8, fragthresh=0.3, seed=7)), ("dna", 60, 200, 11, dict(architecture="hand"))])
def test_alignment_models_match_the_reference_builder(abcname, nseq, alen, seed, kw, tmp_path):
    abc = getattr(easel.Alphabet, abcname)()
    bg = plan7.Background(abc)
    rows = random_alignment(abc, nseq, alen, seed)
    names = ["s%d" % i for i in range(nseq)]
    hand = kw.get("architecture") == "hand"
    rf = "".join("x" if v else "." for v in np.random.default_rng(seed + 1).random(alen) < 0.7) if hand else None
    effn = kw.get("effective_number", "entropy")
    path = str(tmp_path / "ref.hmm")
    wref = refshim.msa_builder({"amino": 3, "dna": 2}[abcname], rows, names, "fam%d" % seed, path, rf=rf, architecture="hand" if hand else "fast",
                               symfrac=kw.get("symfrac", 0.5), fragthresh=kw.get("fragthresh", 0.5),
                               effn=-1.0 if effn == "entropy" else (0.0 if effn == "none" else float(effn)),
                               laplace=kw.get("prior_scheme") == "laplace", seed=kw.get("seed", 42))
    b = hook_reference_calibration(builder.Builder(abc, **kw))
    msa = easel.DigitalMSA(abc, name="fam%d" % seed, names=names, rows=list(rows), reference=rf)
    original = msa.copy()
    hmm, profile, om = b.build_msa(msa, bg)
    assert np.array_equal(wref, msa.sequence_weights)                       # esl_msaweight_PB_adv, bit for bit
    buf = io.BytesIO()
    hmm.write(buf)
    mine = [l for l in buf.getvalue().decode().splitlines() if not l.startswith("DATE")]
    ref = [l for l in open(path).read().splitlines() if not l.startswith("DATE")]
    stats = lambda ls: [[float(v) for v in l.split()[3:]] for l in ls if l.startswith("STATS")]
    rest = lambda ls: [l for l in ls if not l.startswith("STATS")]
    assert rest(mine) == rest(ref), [(a, r) for a, r in zip(rest(mine), rest(ref)) if a != r][:3]
    assert np.allclose(stats(mine), stats(ref), rtol=0, atol=3e-4) and len(stats(mine)) == 3
    assert profile.M == om.M == hmm.M and hmm.nseq == nseq and hmm.checksum == original.checksum
    # p7_Builder rewrites its alignment (weights, RF = the columns it chose); a copy keeps the original
    assert msa.reference.count("x") == hmm.M and original.reference == rf and original.sequence_weights is None
    assert (abcname == "amino") == (hmm.max_length <= 0)


def test_alignment_builder_errors_and_containers():
    abc = easel.Alphabet.amino()
    bg = plan7.Background(abc)
    txt = easel.TextMSA(name="x", names=["a", "b"], sequences=["AC-DE", "ACGDE"])
    msa = txt.digitize(abc)
    assert isinstance(msa, easel.DigitalMSA) and len(msa) == 5 and msa.ax.tolist() == [[0, 1, 20, 2, 3], [0, 1, 5, 2, 3]]
    assert [s.name for s in msa.sequences] == ["a", "b"] and msa.textize().alignment == ["AC-DE", "ACGDE"]
    with pytest.raises(ValueError):
        easel.DigitalMSA(abc, names=["a", "a"], rows=[[0, 1], [0, 1]])
    with pytest.raises(ValueError):
        easel.DigitalMSA(abc, names=["a", "b"], rows=[[0, 1], [0]])
    b = builder.Builder(abc)
    unnamed = easel.DigitalMSA(abc, names=["a", "b"], rows=[[0, 1, 2], [0, 1, 2]])
    b.calibrate = lambda hmm, background: hmm
    with pytest.raises(ValueError, match="Unable to name the HMM"):
        b.build_msa(unnamed, bg)
    allgap = easel.DigitalMSA(abc, name="g", names=["a", "b"], rows=[[20, 20], [20, 20]])
    with pytest.raises(ValueError, match="no consensus columns"):
        b.build_msa(allgap, bg)
    with pytest.raises(ValueError, match="no reference annotation"):
        builder.Builder(abc, architecture="hand").build_msa(msa.copy(), bg)
    with pytest.raises(plan7.AlphabetMismatch):
        b.build_msa(easel.DigitalMSA(easel.Alphabet.dna(), name="d", names=["a"], rows=[[0, 1]]), bg)
    inner = easel.DigitalMSA(abc, name="m", names=["a", "b"], rows=[[0, 28, 2], [0, 1, 2]])
    with pytest.raises(ValueError, match="missing data"):
        b.build_msa(inner, bg)
    for bad in (dict(architecture="slow"), dict(weighting="x"), dict(effective_number="many"), dict(prior_scheme="x")):
        with pytest.raises(ValueError):
            builder.Builder(abc, **bad)
    # Lanczos log-gamma against the standard library, and a two-component mixture's posterior mean against direct evaluation
    import math
    x = np.array([0.003, 0.5, 1.0, 7.25, 120.0])
    assert np.allclose(msabuild.log_gamma(x), [math.lgamma(v) for v in x], rtol=0, atol=1e-9)
    q, alpha = np.array([0.3, 0.7]), np.array([[1.0, 2.0, 0.5], [0.2, 0.2, 4.0]])
    c = np.array([[3.0, 0.0, 1.0]])
    lp = [math.log(q[k]) + sum(math.lgamma(c[0, a] + alpha[k, a]) - math.lgamma(alpha[k, a]) for a in range(3))
          + math.lgamma(alpha[k].sum()) - math.lgamma(c.sum() + alpha[k].sum()) for k in range(2)]
    w = np.exp(np.array(lp) - max(lp))
    w /= w.sum()
    want = sum(w[k] * (c[0] + alpha[k]) / (c.sum() + alpha[k].sum()) for k in range(2))
    assert np.allclose(msabuild.mp_parameters((q, alpha), c)[0], want, rtol=0, atol=1e-9)


class ReferenceSearchPipeline(plan7.Pipeline):
    """`Pipeline` whose comparisons are scored by the reference's p7_Pipeline (oracle/_ref): lets the host side of jackhmmer
    -- thresholds, ranking, alignment, model building -- run without a device.  Test infrastructure."""

    def __init__(self, alphabet, **kw):
        opts = dict(bias_filter=True, null2=True, seed=42, Z=None, domZ=None, F1=0.02, F2=1e-3, F3=1e-5, E=10.0, T=None, domE=10.0, domT=None,
                    incE=0.01, incT=None, incdomE=0.01, incdomT=None, bit_cutoffs=None, host_threads=1)
        opts.update(kw)
        self.alphabet, self.background = alphabet, plan7.Background(alphabet)
        for k, v in opts.items():
            setattr(self, k, v)
        self.clear()

    def search_hmm(self, query, sequences):
        hmm = query
        with tempfile.NamedTemporaryFile(suffix=".hmm") as tmp:
            hmm.write(tmp)
            tmp.flush()
            ref = refshim.RefModel(tmp.name, 0, 400)
        rh, rd, rtext, rc = ref.search([s.sequence for s in sequences])
        hits = (_lib.HitRec * len(rh))()
        doms = (_lib.DomainRec * len(rd))()
        for a, r in zip(hits, rh):
            a.profile = 0
            for fld in ("seq", "score", "pre_score", "sum_score", "nexpected", "lnP", "pre_lnP", "sum_lnP", "nregions", "nclustered", "noverlaps",
                        "nenvelopes", "ndom", "best_domain", "dom_offset"):
                setattr(a, fld, getattr(r, fld))
        for a, r in zip(doms, rd):
            for fld in ("ienv", "jenv", "iali", "jali", "envsc", "domcorrection", "dombias", "oasc", "bitscore", "lnP", "hmmfrom", "hmmto", "sqfrom",
                        "sqto", "N", "text_offset"):
                setattr(a, fld, getattr(r, fld))
        om = plan7.Profile(hmm.M, self.alphabet).configure(hmm, self.background, 400).to_optimized()
        order = sorted(range(len(hits)), key=lambda i: hits[i].seq)
        return self._assemble([hmm], [om], sequences, [hits[i] for i in order], list(doms), rtext, np.array([rc], np.int64).reshape(1, 4))[0]


def load_pksi(abc):
    with easel.SequenceFile(os.path.join(GOLD, "data", "PKSI.faa.gz"), digital=True, alphabet=abc) as f:
        return f.read_block()


def check_iteration(got, want, score_tol=None):
    """One `IterationResult` against a golden step of the reference."""
    tol = score_tol or (lambda v: max(0.05, 2e-5 * abs(v)))
    assert (got.iteration, got.converged, got.hmm.M) == (want["iteration"], want["converged"], want["M"])
    assert got.hmm.nseq == want["nseq"]
    if want["nseq"] > 1 and want["nseq_effective"] is not None and got.hmm.nseq_effective is not None:
        assert abs(got.hmm.nseq_effective - want["nseq_effective"]) < 0.011      # bisection tolerance of p7_EntropyWeight
    assert np.allclose(np.asarray(got.hmm.match_emissions)[1:4], want["match_head"], rtol=0, atol=2e-6)
    assert [h.name for h in got.hits] == [h["name"] for h in want["hits"]]
    for h, w in zip(got.hits, want["hits"]):
        assert (h.included, h.reported, h.new, h.dropped, len(h.domains)) == (w["included"], w["reported"], w["new"], w["dropped"], len(w["domains"]))
        assert abs(h.score - w["score"]) < tol(w["score"]) and abs(h.bias - w["bias"]) < tol(w["score"])
        for d, e in zip(h.domains, w["domains"]):
            assert [d.env_from, d.env_to] == e["env"] and [d.alignment.target_from, d.alignment.target_to] == e["ali"]
            assert [d.alignment.hmm_from, d.alignment.hmm_to] == e["hmm"] and d.included == e["included"]
    m = got.msa.textize()
    txt = lambda v: v.decode() if isinstance(v, bytes) else v
    assert txt(got.msa.name) == want["msa"]["name"] and [txt(n) for n in m.names] == want["msa"]["names"]
    assert m.reference == want["msa"]["rf"]
    # (an alignment that has already been through the builder carries its fragment marks: p7_Builder rewrites its input, and
    # with checkpoints=True the reference hands out the very objects it rebuilds from)
    norm = lambda r: r.upper().replace(".", "-").replace("~", "-")
    assert [norm(r) for r in m.alignment] == [norm(r) for r in want["msa"]["rows"]]


@needs_ref
@pytest.mark.parametrize("label", ["seq:-1", "hmm:KR"])
def test_jackhmmer_loop_matches_the_reference_iterations(label):
    abc = easel.Alphabet.amino()
    seqs = load_pksi(abc)
    with gzip.open(os.path.join(GOLD, "jackhmmer.json.gz")) as f:
        golden = {r["query"]: r["steps"] for r in json.load(f)["runs"]}
    pli = ReferenceSearchPipeline(abc, incE=1e-3, incdomE=1e-3)
    b = hook_reference_calibration(builder.Builder(abc, seed=pli.seed, architecture="hand"))
    if label.startswith("hmm"):
        with tempfile.NamedTemporaryFile(suffix=".hmm") as tmp:
            with gzip.open(os.path.join(GOLD, "data", "KR.hmm.gz")) as f:
                tmp.write(f.read())
            tmp.flush()
            with plan7.HMMFile(tmp.name) as f:
                query = f.read()
        it = pli.iterate_hmm(query, seqs, b)
    else:
        query = seqs[int(label.split(":")[1])]
        it = pli.iterate_seq(query, seqs, b)
    steps = golden[label]
    n = 0
    for want in steps:
        got = next(it)
        check_iteration(got, want)
        n += 1
    assert n >= 2 and steps[-1]["converged"]
    with pytest.raises(StopIteration):
        next(it)
    with pytest.raises(ValueError, match="hand"):
        pli.iterate_seq(seqs[0], seqs, builder.Builder(abc))
