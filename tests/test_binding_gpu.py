"""The engine bound INTO the reference: `pyhmmer_cuda.CudaPipeline(pyhmmer.plan7.Pipeline)` (pyhmmer_b200/binding/), a Cython
extension compiled against the unmodified reference installed under baseline/_ref.  pyhmmer's own objects go in
(`HMM`, `DigitalSequenceBlock`, `OptimizedProfileBlock`), pyhmmer's own `TopHits` come out -- so the reference's OWN test
classes can be run against it, and its results compared object by object with the reference's CPU pipeline."""
import importlib
import os
import sys
import unittest

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


@pytest.fixture(scope="module")
def bound():
    if not os.path.isdir(os.path.join(REF, "pyhmmer")):
        pytest.skip("the reference is not installed under baseline/_ref (pip install --target baseline/_ref, DESIGN.md)")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import pyhmmer
    from pyhmmer_b200.binding import build as bbuild
    bbuild.build()                                        # no-op when the module is up to date
    from pyhmmer_b200.binding import pyhmmer_cuda
    return pyhmmer, pyhmmer_cuda


def _data(pyhmmer, *parts):
    return os.path.join(os.path.dirname(pyhmmer.__file__), "tests", "data", *parts)


def _run_cases(cases):
    suite = unittest.TestSuite()
    for c in cases:
        suite.addTests(unittest.defaultTestLoader.loadTestsFromTestCase(c))
    res = unittest.TestResult()
    suite.run(res)
    return res


def test_reference_pipeline_test_classes_run_on_cuda(bound):
    """The reference's TestSearchPipeline / TestScanPipeline (src/pyhmmer/tests/test_plan7/test_pipeline.py) with `Pipeline`
    replaced by `CudaPipeline` in the test module: every case passes, and the block-based ones launched GPU kernels."""
    pyhmmer, pyhmmer_cuda = bound
    mod = importlib.import_module("pyhmmer.tests.test_plan7.test_pipeline")
    saved = mod.Pipeline
    launches0 = pyhmmer_cuda.engine().launch_count
    mod.Pipeline = pyhmmer_cuda.CudaPipeline
    try:
        res = _run_cases([mod.TestSearchPipeline, mod.TestScanPipeline])
    finally:
        mod.Pipeline = saved
    assert res.testsRun >= 15
    assert not res.failures and not res.errors, [str(f[0]) + "\n" + f[1] for f in res.failures + res.errors]
    assert pyhmmer_cuda.engine().launch_count > launches0 + 50


def test_reference_hmmer_test_classes_run_on_cuda(bound):
    """The reference's TestHmmsearch* / TestHMMScan (src/pyhmmer/tests/test_hmmer.py: golden tables PF02826.tbl, RREFam.tbl,
    RREFam.domtbl, RREFam.scan.tbl) with the workers of pyhmmer.hmmsearch / hmmscan building CudaPipeline objects."""
    pyhmmer, pyhmmer_cuda = bound
    mod = importlib.import_module("pyhmmer.tests.test_hmmer")
    undo = pyhmmer_cuda.install()
    launches0 = pyhmmer_cuda.engine().launch_count
    try:
        cases = [getattr(mod, n) for n in ("TestHmmsearch", "TestHmmsearchSingle", "TestHmmsearchReverse", "TestHMMScan") if hasattr(mod, n)]
        res = _run_cases(cases)
    finally:
        undo()
    assert len(cases) >= 3 and res.testsRun >= 15
    assert not res.failures and not res.errors, [str(f[0]) + "\n" + f[1] for f in res.failures + res.errors]
    assert pyhmmer_cuda.engine().launch_count > launches0 + 100


def test_reference_phmmer_jackhmmer_test_classes_run_on_cuda(bound):
    """the reference's TestPhmmer / TestJackhmmer (src/pyhmmer/tests/test_hmmer.py:417-629; golden table
    A0A089QRB9.domtbl, the jackhmmer CLI's 3 iterations / 5 hits / 17 aligned sequences on PKSI) with the workers of
    pyhmmer.phmmer / jackhmmer building CudaPipeline objects: models come from pyhmmer's own Builder on the CPU, every search
    iteration of `IterativeSearch` (plan7.pyx:4332-4389) runs on the GPU."""
    pyhmmer, pyhmmer_cuda = bound
    mod = importlib.import_module("pyhmmer.tests.test_hmmer")
    undo = pyhmmer_cuda.install()
    launches0 = pyhmmer_cuda.engine().launch_count
    try:
        res = _run_cases([mod.TestPhmmer, mod.TestJackhmmer])
    finally:
        undo()
    assert res.testsRun >= 8
    assert not res.failures and not res.errors, [str(f[0]) + "\n" + f[1] for f in res.failures + res.errors]
    assert pyhmmer_cuda.engine().launch_count > launches0 + 100


def test_iterative_search_on_cuda_equals_reference(bound):
    """jackhmmer iteration by iteration: `CudaPipeline.iterate_seq` / `iterate_hmm` against `Pipeline.iterate_seq` /
    `iterate_hmm` -- the same hits, inclusion flags and scores, the same alignment handed to the builder, the same model
    length and convergence at every iteration; `search_seq` / `search_msa` likewise.  The GPU launch counter proves that the
    cpdef calls inside the reference (`self.search_hmm[...]`, `IterativeSearch._search_hmm`) reach the override."""
    pyhmmer, pyhmmer_cuda = bound
    abc = pyhmmer.easel.Alphabet.amino()
    with pyhmmer.easel.SequenceFile(_data(pyhmmer, "seqs", "PKSI.faa"), digital=True, alphabet=abc) as f:
        seqs = f.read_block()
    eng = pyhmmer_cuda.engine()

    def same(got, ref):
        assert len(got) == len(ref) and len(got.included) == len(ref.included) and len(got.reported) == len(ref.reported)
        for a, b in zip(got, ref):
            assert a.name == b.name and (a.included, a.reported, len(a.domains)) == (b.included, b.reported, len(b.domains))
            # 2e-3 bits, or 1e-5 of the score for the kilobit self-hits of these multi-kilobase proteins (a 4 675-bit score
            # is 3 240 nats: float32 carries it to 2.4e-4 nats, and both implementations sum ~2 000 rescaling logs into it)
            tol = max(2e-3, 1e-5 * abs(b.score))
            assert abs(a.score - b.score) < tol and abs(a.bias - b.bias) < tol
            for d, e in zip(a.domains, b.domains):
                assert (d.env_from, d.env_to, d.alignment.target_from, d.alignment.target_to, d.included) == \
                       (e.env_from, e.env_to, e.alignment.target_from, e.alignment.target_to, e.included)

    for query in (seqs[-1], seqs[0]):
        n0 = eng.launch_count
        same(pyhmmer_cuda.CudaPipeline(abc).search_seq(query, seqs), pyhmmer.plan7.Pipeline(abc).search_seq(query, seqs))
        assert eng.launch_count > n0 + 5
        it_ref = pyhmmer.plan7.Pipeline(abc).iterate_seq(query, seqs)
        it_got = pyhmmer_cuda.CudaPipeline(abc).iterate_seq(query, seqs)
        last = None
        for k in range(5):
            n0 = eng.launch_count
            r, g = next(it_ref), next(it_got)
            assert eng.launch_count > n0 + 5, "iteration %d did not run on the GPU" % (k + 1)
            same(g.hits, r.hits)
            assert (g.iteration, g.converged, g.hmm.M, len(g.msa.sequences), g.msa.name) == (r.iteration, r.converged, r.hmm.M, len(r.msa.sequences), r.msa.name)
            assert [bytes(x) for x in g.msa.alignment] == [bytes(x) for x in r.msa.alignment] if hasattr(g.msa, "alignment") else True
            last = g
            if r.converged:
                break
        assert last is not None and last.iteration >= 2
        # alignment query (phmmer / hmmsearch with an MSA): the alignment of the last iteration as the query
        n0 = eng.launch_count
        # (p7_Builder rewrites the alignment it is given -- weights, fragment marks, RF line: a copy each)
        same(pyhmmer_cuda.CudaPipeline(abc).search_msa(last.msa.copy(), seqs), pyhmmer.plan7.Pipeline(abc).search_msa(last.msa.copy(), seqs))
        assert eng.launch_count > n0 + 5
    with pyhmmer.plan7.HMMFile(_data(pyhmmer, "hmms", "txt", "KR.hmm")) as f:
        hmm = f.read()
    it_ref = pyhmmer.plan7.Pipeline(abc).iterate_hmm(hmm, seqs)
    it_got = pyhmmer_cuda.CudaPipeline(abc).iterate_hmm(hmm, seqs)
    for k in range(3):
        r, g = next(it_ref), next(it_got)
        same(g.hits, r.hits)
        assert (g.converged, g.hmm.M, len(g.msa.sequences)) == (r.converged, r.hmm.M, len(r.msa.sequences))
        if r.converged:
            break


def test_reference_nhmmer_test_classes_run_on_cuda(bound):
    """The long-target pipeline through the binding: the reference's TestNhmmer (src/pyhmmer/tests/test_hmmer.py:631-797;
    goldens bmyD1/2/3.tbl from the nhmmer CLI, RF00001 on the reverse strand, window_length) with pyhmmer.nhmmer's workers
    building `CudaLongTargetsPipeline`, and TestLongTargetsPipeline / TestIteratePipeline (tests/test_plan7/test_pipeline.py:
    257-400; the jackhmmer CLI's alignments of P12748 and KR, iteration by iteration) with the pipeline classes swapped."""
    pyhmmer, pyhmmer_cuda = bound
    mod = importlib.import_module("pyhmmer.tests.test_hmmer")
    undo = pyhmmer_cuda.install()
    launches0 = pyhmmer_cuda.engine().launch_count
    try:
        res = _run_cases([mod.TestNhmmer])
    finally:
        undo()
    assert res.testsRun >= 9
    assert not res.failures and not res.errors, [str(f[0]) + "\n" + f[1] for f in res.failures + res.errors]
    assert pyhmmer_cuda.engine().launch_count > launches0 + 50
    mod = importlib.import_module("pyhmmer.tests.test_plan7.test_pipeline")
    saved = mod.Pipeline, mod.LongTargetsPipeline
    mod.Pipeline, mod.LongTargetsPipeline = pyhmmer_cuda.CudaPipeline, pyhmmer_cuda.CudaLongTargetsPipeline
    try:
        res = _run_cases([mod.TestLongTargetsPipeline, mod.TestIteratePipeline])
    finally:
        mod.Pipeline, mod.LongTargetsPipeline = saved
    assert res.testsRun >= 5
    assert not res.failures and not res.errors, [str(f[0]) + "\n" + f[1] for f in res.failures + res.errors]


def test_cuda_long_targets_pipeline_equals_reference(bound):
    """Object-level comparison of `CudaLongTargetsPipeline` with `LongTargetsPipeline`: hit lists in order, flags, strands,
    coordinates, scores, E-values, the residue counters of the pipeline; sequence queries; a user-set Z; one strand only."""
    pyhmmer, pyhmmer_cuda = bound
    dna = pyhmmer.easel.Alphabet.dna()
    with pyhmmer.plan7.HMMFile(_data(pyhmmer, "hmms", "txt", "bmyD.hmm")) as f:
        bmyd = f.read()
    with pyhmmer.plan7.HMMFile(_data(pyhmmer, "hmms", "txt", "RF00001.hmm")) as f:
        rf = f.read()
    blocks = {}
    for name, abc in (("1390.SAMEA104415756.OFHT01000022", bmyd.alphabet), ("1390.SAMEA104415756.OFHT01000024", rf.alphabet)):
        with pyhmmer.easel.SequenceFile(_data(pyhmmer, "seqs", name + ".fna"), "fasta", digital=True, alphabet=abc) as f:
            blocks[name] = f.read_block()
    with pyhmmer.easel.SequenceFile(_data(pyhmmer, "seqs", "bmyD.fna"), "fasta", digital=True, alphabet=dna) as f:
        bmyd_seq = next(f)
    eng = pyhmmer_cuda.engine()

    def same(got, ref):
        assert type(got) is pyhmmer.plan7.TopHits and len(got) == len(ref)
        assert len(got.reported) == len(ref.reported) and len(got.included) == len(ref.included)
        assert (got.searched_sequences, got.searched_residues, got.searched_nodes) == (ref.searched_sequences, ref.searched_residues, ref.searched_nodes)
        for a, b in zip(got, ref):
            assert a.name == b.name and (a.reported, a.included, a.duplicate) == (b.reported, b.included, b.duplicate)
            assert abs(a.score - b.score) < 2e-3 and abs(a.evalue - b.evalue) <= 2e-3 * b.evalue + 1e-300
            d, e = a.best_domain, b.best_domain
            assert (d.strand, d.env_from, d.env_to, d.alignment.target_from, d.alignment.target_to, d.alignment.hmm_from, d.alignment.hmm_to) == \
                   (e.strand, e.env_from, e.env_to, e.alignment.target_from, e.alignment.target_to, e.alignment.hmm_from, e.alignment.hmm_to)
            assert d.alignment.target_sequence == e.alignment.target_sequence and d.alignment.hmm_sequence == e.alignment.hmm_sequence
            assert abs(d.bias - e.bias) < 2e-3 and d.alignment.target_length == e.alignment.target_length

    n = 0
    cases = [(bmyd, "1390.SAMEA104415756.OFHT01000022", {}), (rf, "1390.SAMEA104415756.OFHT01000024", {}),
             (rf, "1390.SAMEA104415756.OFHT01000024", {"window_length": 3878}), (bmyd, "1390.SAMEA104415756.OFHT01000022", {"Z": 50.0}),
             (bmyd, "1390.SAMEA104415756.OFHT01000022", {"strand": "watson"}), (rf, "1390.SAMEA104415756.OFHT01000024", {"strand": "crick", "E": 1e-3})]
    for hmm, key, kw in cases:
        n0 = eng.launch_count
        got = pyhmmer_cuda.CudaLongTargetsPipeline(hmm.alphabet, **kw).search_hmm(hmm, blocks[key])
        assert eng.launch_count > n0 + 5
        ref = pyhmmer.plan7.LongTargetsPipeline(hmm.alphabet, **kw).search_hmm(hmm, blocks[key])
        same(got, ref)
        n += len(got)
    assert n >= 8
    # a sequence query: pyhmmer's Builder on the host, then the search above
    with pyhmmer.easel.SequenceFile(_data(pyhmmer, "seqs", "BGC0001090.gbk"), digital=True, alphabet=dna) as f:
        bgc = f.read_block()
    n0 = eng.launch_count
    got = pyhmmer_cuda.CudaLongTargetsPipeline(dna).search_seq(bmyd_seq, bgc)
    assert eng.launch_count > n0 + 5
    same(got, pyhmmer.plan7.LongTargetsPipeline(dna).search_seq(bmyd_seq, bgc))
    assert len(got) == 1


def test_cuda_pipeline_equals_reference_pipeline(bound):
    """Object-level comparison: the same queries and targets through pyhmmer's Pipeline (CPU) and CudaPipeline (GPU) -- hit
    lists, flags, every score, domain coordinates and alignment rows, the accounting of the TopHits, pickling."""
    import pickle
    pyhmmer, pyhmmer_cuda = bound
    abc = pyhmmer.easel.Alphabet.amino()
    with pyhmmer.easel.SequenceFile(_data(pyhmmer, "seqs", "938293.PRJEB85.HG003687.faa"), digital=True, alphabet=abc) as f:
        seqs = f.read_block()
    hmms = []
    for name in ("PF02826", "Thioesterase", "RREFam"):
        with pyhmmer.plan7.HMMFile(_data(pyhmmer, "hmms", "txt", name + ".hmm")) as f:
            hmms += list(f)
    nhits = 0
    for kwargs in ({}, {"Z": 5000.0, "domE": 1e-3}, {"bias_filter": False}, {"null2": False}, {"T": 20.0, "domT": 15.0}):
        for hmm in hmms[:4] if kwargs else hmms:
            ref = pyhmmer.plan7.Pipeline(abc, **kwargs).search_hmm(hmm, seqs)
            got = pyhmmer_cuda.CudaPipeline(abc, **kwargs).search_hmm(hmm, seqs)
            assert type(got) is pyhmmer.plan7.TopHits
            assert len(got) == len(ref), (hmm.name, kwargs, len(got), len(ref))
            assert (got.Z, got.domZ, got.searched_sequences, got.searched_residues, got.searched_models, got.searched_nodes) == \
                   (ref.Z, ref.domZ, ref.searched_sequences, ref.searched_residues, ref.searched_models, ref.searched_nodes)
            assert len(got.reported) == len(ref.reported) and len(got.included) == len(ref.included)
            for a, b in zip(got, ref):
                assert a.name == b.name and a.accession == b.accession and a.description == b.description
                assert abs(a.score - b.score) < 2e-3 and abs(a.pre_score - b.pre_score) < 2e-3 and abs(a.bias - b.bias) < 2e-3
                assert abs(a.evalue - b.evalue) <= 2e-3 * b.evalue + 1e-300
                assert (a.reported, a.included, len(a.domains)) == (b.reported, b.included, len(b.domains))
                for d, e in zip(a.domains, b.domains):
                    assert (d.env_from, d.env_to, d.reported, d.included) == (e.env_from, e.env_to, e.reported, e.included)
                    assert abs(d.score - e.score) < 2e-3 and abs(d.bias - e.bias) < 2e-3 and abs(d.envelope_score - e.envelope_score) < 2e-3
                    x, y = d.alignment, e.alignment
                    assert (x.hmm_from, x.hmm_to, x.target_from, x.target_to, x.hmm_name, x.target_name, x.hmm_length, x.target_length) == \
                           (y.hmm_from, y.hmm_to, y.target_from, y.target_to, y.hmm_name, y.target_name, y.hmm_length, y.target_length)
                    assert x.hmm_sequence == y.hmm_sequence and x.target_sequence == y.target_sequence and x.identity_sequence == y.identity_sequence
                    assert sum(1 for p, q in zip(x.posterior_probabilities, y.posterior_probabilities) if p != q) <= max(1, len(x.hmm_sequence) // 100)
            nhits += len(got)
            clone = pickle.loads(pickle.dumps(got))                       # p7_hit_Serialize over the hits we filled
            assert [h.name for h in clone] == [h.name for h in got] and clone.Z == got.Z
    assert nhits >= 60
    # scan orientation: every proteome sequence with a golden hit, against the RREFam block
    with pyhmmer.plan7.HMMFile(_data(pyhmmer, "hmms", "txt", "RREFam.hmm")) as f:
        block = pyhmmer.plan7.OptimizedProfileBlock(abc, [h.to_profile(pyhmmer.plan7.Background(abc)).to_optimized() for h in f])
    picked = [s for s in seqs if s.name in (b"938293.PRJEB85.HG003691_78", b"938293.PRJEB85.HG003686_714", "938293.PRJEB85.HG003691_78", "938293.PRJEB85.HG003686_714")]
    picked += [seqs[0], seqs[7]]
    assert len(picked) == 4
    for q in picked:
        ref = pyhmmer.plan7.Pipeline(abc).scan_seq(q, block)
        got = pyhmmer_cuda.CudaPipeline(abc).scan_seq(q, block)
        assert [(h.name, h.reported, h.included, len(h.domains)) for h in got] == [(h.name, h.reported, h.included, len(h.domains)) for h in ref]
        assert all(abs(a.score - b.score) < 2e-3 for a, b in zip(got, ref)) and (got.Z, got.searched_models) == (ref.Z, ref.searched_models)


def test_one_call_wrappers(bound):
    """`pyhmmer_b200.binding.hmmsearch / hmmscan / nhmmer`: pyhmmer's functions with the CUDA pipelines installed for the duration of
    the call -- same hits as pyhmmer's own, kernels launched, and the workers' `pipeline_class` restored afterwards."""
    pyhmmer, pyhmmer_cuda = bound
    from pyhmmer_b200 import binding
    import pyhmmer.hmmer._hmmsearch as hs, pyhmmer.hmmer._nhmmer as nh
    abc = pyhmmer.easel.Alphabet.amino()
    with pyhmmer.easel.SequenceFile(_data(pyhmmer, "seqs", "938293.PRJEB85.HG003687.faa"), digital=True, alphabet=abc) as f:
        seqs = f.read_block()
    with pyhmmer.plan7.HMMFile(_data(pyhmmer, "hmms", "txt", "PF02826.hmm")) as f:
        hmm = f.read()
    eng = pyhmmer_cuda.engine()
    before = (hs._SEARCHWorker.__dict__.get("pipeline_class"), nh._NHMMERWorker.__dict__.get("pipeline_class"))
    n0 = eng.launch_count
    got = list(binding.hmmsearch(hmm, seqs, cpus=1))
    assert eng.launch_count > n0 + 5
    assert (hs._SEARCHWorker.__dict__.get("pipeline_class"), nh._NHMMERWorker.__dict__.get("pipeline_class")) == before
    ref = list(pyhmmer.hmmsearch(hmm, seqs, cpus=1))
    assert len(got) == len(ref) == 1 and len(got[0]) == len(ref[0]) > 0
    for a, b in zip(got[0], ref[0]):
        assert a.name == b.name and abs(a.score - b.score) < 2e-3 and (a.reported, a.included) == (b.reported, b.included)
    n0 = eng.launch_count
    scan = list(binding.hmmscan(seqs[:3], [hmm], cpus=1))
    assert eng.launch_count > n0 and len(scan) == 3
    ref = list(pyhmmer.hmmscan(seqs[:3], [hmm], cpus=1))
    assert [len(t) for t in scan] == [len(t) for t in ref]
    with pyhmmer.plan7.HMMFile(_data(pyhmmer, "hmms", "txt", "bmyD.hmm")) as f:
        bmyd = f.read()
    with pyhmmer.easel.SequenceFile(_data(pyhmmer, "seqs", "BGC0001090.gbk"), digital=True, alphabet=bmyd.alphabet) as f:
        bgc = f.read_block()
    n0 = eng.launch_count
    hits = next(binding.nhmmer(bmyd, bgc, cpus=1))
    assert eng.launch_count > n0 + 5 and len(hits.reported) == 2
    assert (hs._SEARCHWorker.__dict__.get("pipeline_class"), nh._NHMMERWorker.__dict__.get("pipeline_class")) == before
