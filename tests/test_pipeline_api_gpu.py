"""The reference's own Pipeline tests (src/pyhmmer/tests/test_plan7/test_pipeline.py), restated against this package for
the cases that do not need `Builder` (model construction is outside the search path): alphabet checks, unsupported query
types, the Z bookkeeping, bit cutoffs, unnamed models, scan orientation."""
import gzip
import os

import numpy as np
import pytest

from pyhmmer_b200 import easel, plan7, hmmer
from pyhmmer_b200.easel import Alphabet, AlphabetMismatch, TextSequence, DigitalSequenceBlock

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")


@pytest.fixture(scope="module")
def references(amino):
    with easel.SequenceFile(os.path.join(GOLD, "proteome.faa.gz"), digital=True, alphabet=amino) as f:
        return f.read_block()


def _hmm(name):
    with gzip.open(os.path.join(GOLD, name + ".hmm.gz")) as f:
        return plan7.HMMFile(f).read()


def test_search_hmm_alphabet_mismatch(amino, references):
    """test_pipeline.py:35 -- pipeline alphabet vs database alphabet, and vs query alphabet."""
    pipeline = plan7.Pipeline(Alphabet.dna())
    hmm = _hmm("PF02826")
    with pytest.raises(AlphabetMismatch):
        pipeline.search_hmm(hmm, references)
    dna_block = DigitalSequenceBlock(Alphabet.dna(), [TextSequence(sequence="ATGC", name=b"d").digitize(Alphabet.dna())])
    with pytest.raises(AlphabetMismatch):
        pipeline.search_hmm(hmm, dna_block)


def test_search_hmm_unsupported(amino, references):
    """test_pipeline.py:101 -- a query that is no HMM / Profile / OptimizedProfile is a TypeError."""
    with pytest.raises(TypeError):
        plan7.Pipeline(amino).search_hmm(object(), references)
    with pytest.raises(TypeError):
        plan7.Pipeline(amino).search_hmm(_hmm("PF02826"), [s for s in references])      # not a DigitalSequenceBlock


def test_search_hmm_unnamed(amino, references):
    """test_pipeline.py:90 -- a model without accession goes through; the query is reported back on the TopHits."""
    hmm = _hmm("Thioesterase")
    hmm.name, hmm.accession = "test", None
    hits = plan7.Pipeline(amino).search_hmm(hmm, references)
    assert hits.query.name == "test" and hits.query.accession is None


def test_Z(amino, references):
    """test_pipeline.py:151 -- Z = number of targets unless given; clear() keeps a given Z."""
    hmm = _hmm("PF02826")
    pipeline = plan7.Pipeline(amino)
    assert pipeline.Z is None
    hits = pipeline.search_hmm(hmm, references[:100])
    assert hits.Z == 100 and pipeline.Z is None
    pipeline.clear()
    assert pipeline.Z is None
    pipeline = plan7.Pipeline(amino, Z=25)
    hits = pipeline.search_hmm(hmm, references[:100])
    assert pipeline.Z == 25 and hits.Z == 25
    pipeline.clear()
    assert pipeline.Z == 25


def test_bit_cutoffs(amino, references):
    """test_pipeline.py:176 -- missing cutoffs raise (a ValueError); thresholds then decide what is reported."""
    hmm = _hmm("Thioesterase")
    hmm.cutoffs.trusted = None
    pipeline = plan7.Pipeline(amino, bit_cutoffs="trusted")
    with pytest.raises(ValueError):
        pipeline.search_hmm(hmm, references)
    best = max(h.score for h in plan7.Pipeline(amino).search_hmm(hmm, references))
    hmm.cutoffs.trusted = (best - 1.0, 0.0)
    hits = pipeline.search_hmm(hmm, references)
    assert len(hits) >= 1 and all(h.score >= best - 1.0 for h in hits)
    hmm.cutoffs.trusted = (best + 50.0, best + 50.0)
    assert len(pipeline.search_hmm(hmm, references)) == 0


def test_scan_seq_alphabet_mismatch(amino, references):
    """test_pipeline.py:209 -- scan orientation: query sequence alphabet vs pipeline alphabet."""
    pipeline = plan7.Pipeline(Alphabet.dna())
    with pytest.raises(AlphabetMismatch):
        pipeline.scan_seq(references[0], [_hmm("PF02826")])


def test_scan_seq_block(amino, references):
    """test_pipeline.py:222 -- scan one sequence against a block of profiles: hits name the models, Z = number of models."""
    hmms = [_hmm("PF02826"), _hmm("Thioesterase")]
    pli = plan7.Pipeline(amino)
    target = max(references, key=len)
    hits = pli.scan_seq(target, plan7.OptimizedProfileBlock(amino, [pli._optimized(h, 400) for h in hmms]))
    assert hits.Z == 2
    assert {h.name for h in hits} <= {h.name for h in hmms}
    # empty profile list and empty database: no hits, no crash
    assert len(plan7.Pipeline(amino).search_hmm(hmms[0], DigitalSequenceBlock(amino))) == 0
    assert list(hmmer.hmmsearch([], references)) == []
