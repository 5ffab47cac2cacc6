"""`hmmsearch` / `hmmscan` / `nhmmer` / `phmmer` / `jackhmmer` entry points with the reference's signatures
(src/pyhmmer/hmmer/_hmmsearch.py:294, _hmmscan.py:91, _nhmmer.py, _phmmer.py, _jackhmmer.py), running on the B200 engine.

The reference fans queries (or target slices) out over CPU threads; here one process drives one GPU and
every query batch is ONE fused cascade launch sequence over (batch x database).  With several GPUs
(one process per GPU under ``torch.distributed``) the *target database* is sharded by residues -- exactly the
reference's ``parallel="targets"`` strategy (_hmmsearch.py:115-289) -- and the per-rank hits are combined by a
single all-gather at the end (`pyhmmer_b200.parallel`).
"""
import collections
import os

from . import plan7
from .easel import DigitalSequenceBlock, DigitalSequence, SequenceFile
from .plan7 import Pipeline, LongTargetsPipeline, HMM, Profile, OptimizedProfile, OptimizedProfileBlock

__all__ = ["hmmsearch", "hmmscan", "nhmmer", "phmmer", "jackhmmer"]

_QUERY_BATCH = 256


def _as_block(sequences, alphabet):
    if isinstance(sequences, DigitalSequenceBlock):
        return sequences
    if isinstance(sequences, SequenceFile):
        if not sequences.digital:
            raise ValueError("target sequences file is not in digital mode")
        return sequences.read_block()
    return DigitalSequenceBlock(alphabet, sequences)


def hmmsearch(queries, sequences, *, cpus=0, callback=None, backend="threading", parallel="queries",
              batch_size=_QUERY_BATCH, **options):
    """Search HMM/profile queries against a sequence database; yields one `TopHits` per query, in order.

    ``cpus``, ``backend`` and ``parallel`` are accepted for signature compatibility; the work division is
    decided by the GPU engine (``cpus`` bounds the host threads used for domain definition).
    """
    if isinstance(queries, (HMM, Profile, OptimizedProfile)):
        queries = (queries,)
    it = iter(queries)
    first = next(it, None)
    if first is None:
        return
    alphabet = first.alphabet
    block = _as_block(sequences, alphabet)
    options.setdefault("host_threads", cpus or 0)
    pipeline = Pipeline(alphabet, **options)
    from . import parallel as par
    world = par.World.current()
    local = par.shard_block(block, world) if world.size > 1 else None
    batch = [first]
    total = 0

    def flush(batch):
        nonlocal total
        if world.size > 1:
            results = par.search_sharded(pipeline, batch, block, local, world)
        else:
            results = pipeline._search_many(batch, block)
        for q, th in zip(batch, results):
            total += 1
            if callback is not None:
                callback(q, total)
            yield th

    for q in it:
        batch.append(q)
        if len(batch) >= batch_size:
            yield from flush(batch)
            batch = []
    if batch:
        yield from flush(batch)


def hmmscan(queries, profiles, *, cpus=0, callback=None, backend="threading", background=None,
            batch_size=_QUERY_BATCH, **options):
    """Scan query sequences against a profile database; yields one `TopHits` per query sequence, in order."""
    if isinstance(queries, DigitalSequence):
        queries = (queries,)
    targets = profiles if isinstance(profiles, (list, OptimizedProfileBlock)) else list(profiles)
    if not targets:
        for q in queries:
            yield plan7.TopHits(q, "scan")
        return
    alphabet = targets[0].alphabet
    options.setdefault("host_threads", cpus or 0)
    pipeline = Pipeline(alphabet, background=background, **options)
    # convert HMM / Profile targets once (the reference does the same up front, _hmmscan.py:191-215)
    oms = OptimizedProfileBlock(alphabet, pipeline._optimized_many(list(targets), pipeline.L_HINT))
    from . import parallel as par
    world = par.World.current()              # several GPUs: the profile block is sharded, one all-gather per batch
    total = 0
    batch = []

    def flush(batch):
        nonlocal total
        for q, th in zip(batch, pipeline._scan_many(batch, oms, world=world)):
            total += 1
            if callback is not None:
                callback(q, total)
            yield th

    for q in queries:
        batch.append(q)
        if len(batch) >= batch_size:
            yield from flush(batch)
            batch = []
    if batch:
        yield from flush(batch)


def nhmmer(queries, sequences, *, cpus=0, callback=None, backend="threading", builder=None, **options):
    """Search nucleotide HMM / profile queries against long nucleotide targets; yields one `TopHits` per query, in order
    (``pyhmmer.hmmer.nhmmer``, src/pyhmmer/hmmer/_nhmmer.py).  `DigitalSequence` queries become single-sequence models,
    `DigitalMSA` queries go through ``Builder.build_msa`` (`pyhmmer_b200.builder.Builder`)."""
    if isinstance(queries, (HMM, Profile, OptimizedProfile)):
        queries = (queries,)
    it = iter(queries)
    first = next(it, None)
    if first is None:
        return
    if not isinstance(first, (HMM, Profile, OptimizedProfile, DigitalSequence)):
        raise TypeError("nhmmer queries must be HMM, Profile, OptimizedProfile or DigitalSequence (building models from "
                        "alignments is not part of pyhmmer_b200), found %s" % type(first).__name__)
    alphabet = first.alphabet
    block = _as_block(sequences, alphabet)
    options.setdefault("host_threads", cpus or 0)
    pipeline = LongTargetsPipeline(alphabet, **options)
    index = 0
    import itertools
    for query in itertools.chain((first,), it):
        if isinstance(query, DigitalSequence):
            hits = pipeline.search_seq(query, block, builder)
        else:
            hits = pipeline.search_hmm(query, block)
        if callback is not None:
            callback(query, index + 1)
        index += 1
        yield hits


def phmmer(queries, sequences, *, cpus=0, callback=None, backend="threading", builder=None, **options):
    """Search protein query SEQUENCES against a sequence database; yields one `TopHits` per query, in order
    (``pyhmmer.hmmer.phmmer``, src/pyhmmer/hmmer/_phmmer.py): every query becomes a single-sequence model
    (`pyhmmer_b200.builder.Builder`: BLOSUM62, gap open 0.02 / extend 0.4, calibrated) and is searched like an HMM.
    `DigitalMSA` queries go through ``Builder.build_msa``."""
    from .builder import Builder
    from .easel import DigitalMSA
    if isinstance(queries, (DigitalSequence, DigitalMSA)):
        queries = (queries,)
    it = iter(queries)
    first = next(it, None)
    if first is None:
        return
    if not isinstance(first, (DigitalSequence, DigitalMSA)):
        raise TypeError("phmmer queries must be DigitalSequence or DigitalMSA, found %s" % type(first).__name__)
    alphabet = first.alphabet
    block = _as_block(sequences, alphabet)
    options.setdefault("host_threads", cpus or 0)
    pipeline = Pipeline(alphabet, **options)
    builder = Builder(alphabet, seed=pipeline.seed) if builder is None else builder
    import itertools
    for index, query in enumerate(itertools.chain((first,), it)):
        hits = pipeline.search_msa(query, block, builder) if isinstance(query, DigitalMSA) else pipeline.search_seq(query, block, builder)
        if callback is not None:
            callback(query, index + 1)
        yield hits


def jackhmmer(queries, sequences, *, cpus=0, callback=None, backend="threading", builder=None, max_iterations=5, select_hits=None,
              checkpoints=False, **options):
    """Iterative search of query sequences or HMMs against a sequence database (``pyhmmer.hmmer.jackhmmer``,
    src/pyhmmer/hmmer/_jackhmmer.py:37-127): for every query, up to ``max_iterations`` rounds of `IterativeSearch` (build a
    model on the host, search on the GPU, align the included hits, rebuild); yields the last `IterationResult` per query,
    or the list of all of them with ``checkpoints``."""
    from .builder import Builder
    import itertools
    if isinstance(queries, (DigitalSequence, HMM)):
        queries = (queries,)
    it = iter(queries)
    first = next(it, None)
    if first is None:
        return
    alphabet = first.alphabet
    block = _as_block(sequences, alphabet)
    options.setdefault("host_threads", cpus or 0)
    options.setdefault("incE", 0.001)                     # jackhmmer's inclusion thresholds (_jackhmmer.py: incE = incdomE = 1e-3)
    options.setdefault("incdomE", 0.001)
    pipeline = Pipeline(alphabet, **options)
    builder = Builder(alphabet, seed=pipeline.seed, architecture="hand") if builder is None else builder
    for index, query in enumerate(itertools.chain((first,), it)):
        if isinstance(query, DigitalSequence):
            steps = pipeline.iterate_seq(query, block, builder, select_hits)
        elif isinstance(query, HMM):
            steps = pipeline.iterate_hmm(query, block, builder, select_hits)
        else:
            raise TypeError("Unsupported query type for `jackhmmer`: %s" % type(query).__name__)
        done = []
        last = None
        for last in itertools.islice(steps, max_iterations):
            if checkpoints:
                done.append(last)
            if last.converged:
                break
        if callback is not None:
            callback(query, index + 1)
        pipeline.clear()
        yield done if checkpoints else last
