"""jackhmmer and alignment queries on the engine: `hmmer.jackhmmer` / `Pipeline.iterate_seq` /
`iterate_hmm` / `search_msa` of the Python mirror -- models built on the host (`Builder.build` / `build_msa`, calibrated with
the GPU filters), every search on the GPU -- against the golden iterations recorded from the reference's own
`Pipeline.iterate_seq` / `iterate_hmm` (tests/golden/make_jackhmmer_golden.py): hits, flags, scores, domain coordinates, the
alignment of every round (query first), the model built from it, convergence."""
import gzip
import json
import os
import sys
import tempfile

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pyhmmer_b200 import builder, easel, hmmer, plan7
from test_msabuild_cpu import GOLD, check_iteration, load_pksi

pytestmark = pytest.mark.gpu


def _golden():
    with gzip.open(os.path.join(GOLD, "jackhmmer.json.gz")) as f:
        return {r["query"]: r["steps"] for r in json.load(f)["runs"]}


def _kr():
    with tempfile.NamedTemporaryFile(suffix=".hmm") as tmp:
        with gzip.open(os.path.join(GOLD, "data", "KR.hmm.gz")) as f:
            tmp.write(f.read())
        tmp.flush()
        with plan7.HMMFile(tmp.name) as f:
            return f.read()


def test_jackhmmer_matches_the_reference_iterations():
    abc = easel.Alphabet.amino()
    seqs = load_pksi(abc)
    golden = _golden()
    queries = [seqs[-1], seqs[0], _kr()]
    seen = []
    results = list(hmmer.jackhmmer(queries, seqs, checkpoints=True, callback=lambda q, n: seen.append(n)))
    assert seen == [1, 2, 3] and len(results) == 3
    for label, steps in zip(("seq:-1", "seq:0", "hmm:KR"), results):
        want = golden[label]
        assert len(steps) == len(want) and steps[-1].converged
        for got, w in zip(steps, want):
            check_iteration(got, w)
    # without checkpoints: the last iteration only; max_iterations bounds the loop
    last = next(hmmer.jackhmmer(seqs[-1], seqs))
    assert last.iteration == len(golden["seq:-1"]) and last.converged
    first = next(hmmer.jackhmmer(seqs[-1], seqs, max_iterations=1))
    assert first.iteration == 1 and not first.converged
    assert next(hmmer.jackhmmer([], seqs), None) is None


def test_alignment_queries_on_the_gpu():
    """`Pipeline.search_msa` / `hmmer.phmmer` with a `DigitalMSA`: the model `Builder.build_msa` makes of a jackhmmer round's
    alignment finds what that round's own next iteration found (same builder settings => same model => same hits)."""
    abc = easel.Alphabet.amino()
    seqs = load_pksi(abc)
    pli = plan7.Pipeline(abc, incE=1e-3, incdomE=1e-3)
    it = pli.iterate_seq(seqs[-1], seqs)
    r1 = next(it)
    r2 = next(it)
    hand = builder.Builder(abc, seed=pli.seed, architecture="hand")
    hits = plan7.Pipeline(abc, incE=1e-3, incdomE=1e-3).search_msa(r1.msa.copy(), seqs, hand)
    assert hits.query is not None and [h.name for h in hits] == [h.name for h in r2.hits]
    assert np.allclose([h.score for h in hits], [h.score for h in r2.hits], rtol=0, atol=1e-3)
    via = next(hmmer.phmmer(r1.msa.copy(), seqs, builder=hand, incE=1e-3, incdomE=1e-3))
    assert [h.name for h in via] == [h.name for h in hits]
    # the default (fast) architecture chooses its own consensus columns
    fast = plan7.Pipeline(abc).search_msa(r1.msa.copy(), seqs)
    assert len(fast) >= 4 and fast[0].score > 1000
