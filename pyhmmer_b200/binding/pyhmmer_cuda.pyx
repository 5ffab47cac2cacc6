# cython: language_level=3
"""pyhmmer_cuda -- the B200 engine bound INTO pyhmmer: `CudaPipeline(pyhmmer.plan7.Pipeline)`.

A Cython extension compiled against an installed pyhmmer (its .pxd files, headers and shared libraries, as
PyHMMERConfig.cmake advertises them: src/cmake/PyHMMERConfig.cmake.in) and against include/b2h.h.  `search_hmm` and
`scan_seq` keep pyhmmer's signatures and semantics and return genuine `pyhmmer.plan7.TopHits` (a real P7_TOPHITS filled
through p7_tophits_CreateNextHit), but the loops over targets -- `Pipeline._search_loop` / `_scan_loop`
(src/pyhmmer/plan7.pyx:6394-6453, 6625-6677) -- are one `b2h_search` call each (b2h_pyhmmer_glue.c).

    import pyhmmer, pyhmmer_cuda
    hits = pyhmmer_cuda.CudaPipeline(alphabet).search_hmm(hmm, sequences)       # pyhmmer objects in, pyhmmer TopHits out
    pyhmmer_cuda.install()          # pyhmmer.hmmsearch / hmmscan workers build CudaPipeline objects (pipeline_class hook)
"""
from libc.stdint cimport uint8_t, int32_t, int64_t, uint32_t, uint64_t
from libc.stdlib cimport malloc, free

cimport libeasel
from libeasel.sq cimport ESL_SQ
from libhmmer.p7_bg cimport P7_BG
from libhmmer.p7_pipeline cimport P7_PIPELINE, p7_pipemodes_e
from libhmmer.p7_tophits cimport P7_TOPHITS
from libhmmer.impl_sse.p7_oprofile cimport P7_OPROFILE, P7_OM_BLOCK

from pyhmmer.easel cimport Alphabet, DigitalSequence, DigitalSequenceBlock
from pyhmmer.plan7 cimport Pipeline, LongTargetsPipeline, TopHits, HMM, Profile, OptimizedProfile, OptimizedProfileBlock, Background

from pyhmmer.errors import AlphabetMismatch, UnexpectedError, MissingCutoffs


cdef extern from "b2h.h" nogil:
    ctypedef struct b2h_ctx
    ctypedef struct b2h_seqdb
    ctypedef struct b2h_profile
    int  b2h_ctx_create(int device, b2h_ctx **out)
    void b2h_ctx_destroy(b2h_ctx *ctx)
    const char *b2h_ctx_last_error(const b2h_ctx *ctx)
    unsigned long long b2h_ctx_launch_count(const b2h_ctx *ctx)
    void b2h_seqdb_destroy(b2h_seqdb *db)
    void b2h_profile_destroy(b2h_profile *p)

cdef extern from "b2h_pyhmmer_glue.h" nogil:
    int b2h_glue_upload_oprofile(b2h_ctx *ctx, const P7_OPROFILE *om, const P7_BG *bg, b2h_profile **out)
    int b2h_glue_seqdb(b2h_ctx *ctx, ESL_SQ *const *sq, size_t n, b2h_seqdb **out)
    int b2h_glue_search_loop(b2h_ctx *ctx, const b2h_seqdb *db, P7_PIPELINE *pli, P7_OPROFILE *om, P7_BG *bg,
                             ESL_SQ *const *sq, size_t n_targets, P7_TOPHITS *th, unsigned seed, int host_threads)
    int b2h_glue_scan_loop(b2h_ctx *ctx, const b2h_profile *const *profs, P7_PIPELINE *pli, const ESL_SQ *sq, P7_BG *bg,
                           P7_OPROFILE *const *om, size_t n_targets, P7_TOPHITS *th, unsigned seed, int host_threads)
    int b2h_glue_longtarget_fill(P7_PIPELINE *pli, P7_TOPHITS *th, const void *hits, size_t nh, const void *doms, const char *text,
                                 const unsigned char *dup, ESL_SQ *const *sq, P7_OPROFILE *om, P7_BG *bg,
                                 int64_t nseqs, int64_t nres, const int64_t *pos_past)
    void b2h_glue_seqdb_destroy(b2h_seqdb *db)
    void b2h_glue_profile_destroy(b2h_profile *p)

cdef enum:
    eslOK = 0
    eslEINVAL = 11
    eslERANGE = 16
    B2H_ECUDA = 100

DEF HMMER_TARGET_LIMIT = 100000


cdef class _Engine:
    """One CUDA context per process, the target databases and profile blocks it keeps resident."""
    cdef b2h_ctx* ctx
    cdef dict _dbs          # id(block) -> (_SeqDB, length, first ESL_SQ*)
    cdef dict _blocks       # id(profile block) -> (_ProfBlock, length)

    def __cinit__(self):
        self.ctx = NULL
        self._dbs = {}
        self._blocks = {}

    def __init__(self, int device=0):
        cdef int status = b2h_ctx_create(device, &self.ctx)
        if status != eslOK:
            raise RuntimeError("b2h_ctx_create failed with status %d: no sm_100 CUDA device? (there is no CPU fallback)" % status)

    def __dealloc__(self):
        self._dbs = None
        self._blocks = None
        if self.ctx != NULL:
            b2h_ctx_destroy(self.ctx)
            self.ctx = NULL

    cdef str last_error(self):
        cdef const char* e = b2h_ctx_last_error(self.ctx)
        return e.decode("utf-8", "replace") if e != NULL else ""

    @property
    def launch_count(self):
        """Kernels launched for this process's searches: the engine's own context plus the host layer's, through which the
        long-target searches run (CudaLongTargetsPipeline)."""
        import sys
        n = int(b2h_ctx_launch_count(self.ctx))
        m = sys.modules.get("pyhmmer_b200._lib")
        if m is not None:
            n += sum(c.launch_count for c in list(m._contexts.values()) if c.handle)
        return n


cdef class _SeqDB:
    cdef b2h_seqdb* db
    cdef _Engine engine
    def __cinit__(self):
        self.db = NULL
    def __dealloc__(self):
        if self.db != NULL:
            b2h_glue_seqdb_destroy(self.db)


cdef class _ProfBlock:
    cdef b2h_profile** profs
    cdef size_t n
    cdef _Engine engine
    def __cinit__(self):
        self.profs = NULL
        self.n = 0
    def __dealloc__(self):
        cdef size_t i
        if self.profs != NULL:
            for i in range(self.n):
                if self.profs[i] != NULL:
                    b2h_glue_profile_destroy(self.profs[i])
            free(self.profs)


cdef _Engine _ENGINE = None

def engine(int device=0):
    """The process-wide engine (created on first use)."""
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = _Engine(device)
    return _ENGINE


cdef void _raise(_Engine eng, int status, str fn) except *:
    if status == eslERANGE:
        raise OverflowError("numerical overflow in the optimized vector implementation")
    if status == B2H_ECUDA:
        raise RuntimeError("%s: CUDA failure: %s" % (fn, eng.last_error()))
    raise UnexpectedError(status, fn)


cdef class _CudaPipelineBase(Pipeline):
    """`pyhmmer.plan7.Pipeline` whose comparison loops run on the GPU (the extension type; use `CudaPipeline`)."""

    cdef _Engine _engine
    cdef int     _host_threads

    def __init__(self, Alphabet alphabet, Background background=None, *, int device=0, int host_threads=0, **kwargs):
        super().__init__(alphabet, background, **kwargs)
        self._engine = engine(device)
        self._host_threads = host_threads

    cdef _SeqDB _database(self, DigitalSequenceBlock sequences):
        """The block, resident on the device.  The engine keeps the two most recent blocks (pyhmmer's blocks cannot be weakly
        referenced, so they are held strongly); a block is recognised by identity plus a checksum over its ESL_SQ pointers
        and lengths, i.e. appending, removing or replacing sequences re-uploads it."""
        cdef _SeqDB sdb
        cdef int status
        cdef size_t i, n = sequences._length
        cdef uint64_t sig = 1469598103934665603ULL
        cdef ESL_SQ *const * refs = <ESL_SQ *const *> sequences._refs
        cdef b2h_ctx* ctx = self._engine.ctx
        for i in range(n):
            sig = (sig ^ <uint64_t> <size_t> refs[i]) * 1099511628211ULL
            sig = (sig ^ <uint64_t> refs[i].n) * 1099511628211ULL
        key = id(sequences)
        hit = self._engine._dbs.get(key)
        if hit is not None and hit[1] is sequences and hit[2] == n and hit[3] == sig:
            return hit[0]
        sdb = _SeqDB()
        sdb.engine = self._engine
        with nogil:
            status = b2h_glue_seqdb(ctx, refs, n, &sdb.db)
        if status != eslOK:
            _raise(self._engine, status, "b2h_seqdb_create")
        self._engine._dbs.pop(key, None)
        while len(self._engine._dbs) >= 2:
            self._engine._dbs.pop(next(iter(self._engine._dbs)))
        self._engine._dbs[key] = (sdb, sequences, n, sig)
        return sdb

    def search_hmm(self, query, sequences):
        """`Pipeline.search_hmm` (plan7.pyx:6155-6258) for targets in a `DigitalSequenceBlock`."""
        cdef size_t       L
        cdef int          status
        cdef P7_OPROFILE* om
        cdef TopHits      hits
        cdef _SeqDB       sdb
        cdef DigitalSequenceBlock block
        cdef OptimizedProfile opt
        cdef Profile      gm
        cdef b2h_ctx*     ctx = self._engine.ctx
        cdef unsigned     seed = self._seed
        cdef int          nthreads = self._host_threads
        if not isinstance(sequences, DigitalSequenceBlock):
            return super().search_hmm(query, sequences)                  # SequenceFile targets: the reference's own loop
        if not isinstance(query, (HMM, Profile, OptimizedProfile)):
            raise TypeError("Expected HMM, Profile or OptimizedProfile, found %s" % type(query).__name__)
        block = sequences
        hits = TopHits(query)
        if not self.alphabet._eq(query.alphabet):
            raise AlphabetMismatch(self.alphabet, query.alphabet)
        if not self.alphabet._eq(block.alphabet):
            raise AlphabetMismatch(self.alphabet, block.alphabet)
        L = self.L_HINT if block._length == 0 else block._refs[0].L
        if block._length > 0 and len(block.largest()) > HMMER_TARGET_LIMIT:
            raise ValueError(f"sequence length over comparison pipeline limit ({HMMER_TARGET_LIMIT})")
        # the optimized profile of the query, as Pipeline._get_om_from_query builds it (plan7.pyx:5979-6013):
        # HMM -> Profile.configure(hmm, background, L) -> OptimizedProfile
        if isinstance(query, OptimizedProfile):
            opt = <OptimizedProfile> query
        elif isinstance(query, Profile):
            opt = (<Profile> query).to_optimized()
        else:
            gm = Profile((<HMM> query).M, self.alphabet)
            gm.configure(<HMM> query, self.background, L)
            opt = gm.to_optimized()
        om = opt._om
        sdb = self._database(block)
        cdef b2h_seqdb* db = sdb.db
        cdef P7_BG* bg = self.background._bg
        cdef ESL_SQ *const * refs = <ESL_SQ *const *> block._refs
        cdef size_t n = block._length
        with nogil:
            self._pli.mode = p7_pipemodes_e.p7_SEARCH_SEQS
            self._pli.nseqs = 0
            status = b2h_glue_search_loop(ctx, db, self._pli, om, bg, refs, n, hits._th, seed, nthreads)
        if status == eslEINVAL:
            raise MissingCutoffs(query.name, self.bit_cutoffs)
        elif status != eslOK:
            _raise(self._engine, status, "b2h_search")
        with nogil:
            hits._sort_by_key()
            hits._threshold(self)
        hits._query = query
        hits._empty = False
        return hits

    cdef _ProfBlock _profiles(self, OptimizedProfileBlock targets):
        """The profile block, resident on the device (same caching rule as `_database`)."""
        cdef _ProfBlock pb
        cdef size_t i
        cdef int status = eslOK
        cdef size_t n = <size_t> targets._block.count
        cdef P7_OPROFILE** oms = targets._block.list
        cdef b2h_ctx* ctx = self._engine.ctx
        cdef P7_BG* bg = self.background._bg
        cdef uint64_t sig = 1469598103934665603ULL
        for i in range(n):
            sig = (sig ^ <uint64_t> <size_t> oms[i]) * 1099511628211ULL
        key = id(targets)
        hit = self._engine._blocks.get(key)
        if hit is not None and hit[1] is targets and hit[2] == n and hit[3] == sig:
            return hit[0]
        pb = _ProfBlock()
        pb.engine = self._engine
        pb.n = n
        pb.profs = <b2h_profile**> malloc(sizeof(b2h_profile*) * max(1, n))
        if pb.profs == NULL:
            raise MemoryError()
        for i in range(n):
            pb.profs[i] = NULL
        with nogil:
            for i in range(n):
                status = b2h_glue_upload_oprofile(ctx, oms[i], bg, &pb.profs[i])
                if status != eslOK:
                    break
        if status != eslOK:
            _raise(self._engine, status, "b2h_profile_upload")
        self._engine._blocks.pop(key, None)
        while len(self._engine._blocks) >= 2:
            self._engine._blocks.pop(next(iter(self._engine._blocks)))
        self._engine._blocks[key] = (pb, targets, n, sig)
        return pb

    def scan_seq(self, DigitalSequence query, targets):
        """`Pipeline.scan_seq` (plan7.pyx:6534-6620) for targets pre-fetched into an `OptimizedProfileBlock`."""
        cdef int          status
        cdef TopHits      hits
        cdef _ProfBlock   pb
        cdef OptimizedProfileBlock block
        cdef b2h_ctx*     ctx = self._engine.ctx
        cdef unsigned     seed = self._seed
        cdef int          nthreads = self._host_threads
        if not isinstance(targets, OptimizedProfileBlock):
            return super().scan_seq(query, targets)                      # HMMPressedFile targets: the reference's own loop
        block = targets
        hits = TopHits(query)
        if not self.alphabet._eq(query.alphabet):
            raise AlphabetMismatch(self.alphabet, query.alphabet)
        if not self.alphabet._eq(block.alphabet):
            raise AlphabetMismatch(self.alphabet, block.alphabet)
        if len(query) > HMMER_TARGET_LIMIT:
            raise ValueError(f"sequence length over comparison pipeline limit ({HMMER_TARGET_LIMIT})")
        pb = self._profiles(block)
        cdef const b2h_profile *const * profs = <const b2h_profile *const *> pb.profs
        cdef P7_OPROFILE *const * oms = <P7_OPROFILE *const *> block._block.list
        cdef size_t n = <size_t> block._block.count
        cdef const ESL_SQ* sq = query._sq
        cdef P7_BG* bg = self.background._bg
        with nogil:
            self._pli.mode = p7_pipemodes_e.p7_SCAN_MODELS
            self._pli.nmodels = 0
            status = b2h_glue_scan_loop(ctx, profs, self._pli, sq, bg, oms, n, hits._th, seed, nthreads)
        if status == eslEINVAL:
            raise MissingCutoffs(b"?", self.bit_cutoffs)
        elif status != eslOK:
            _raise(self._engine, status, "b2h_search")
        with nogil:
            hits._sort_by_key()
            hits._threshold(self)
        hits._query = query
        hits._empty = False
        return hits


class CudaPipeline(_CudaPipelineBase):
    """`pyhmmer.plan7.Pipeline` whose comparison loops run on the GPU.  Same constructor, same results.

    A Python-level class on purpose: `Pipeline.search_seq`, `search_msa` and `IterativeSearch._search_hmm` call the cpdef
    `search_hmm` through the C vtable, and Cython only looks for an overriding method there when the object's type is a
    heap type (the dispatch check of cpdef methods is skipped for static extension types).  With this class phmmer's and
    jackhmmer's searches -- every iteration -- reach the GPU `search_hmm` above."""
    __slots__ = ()


# The long-target (nhmmer) pipeline.  Its window loop -- `LongTargetsPipeline._search_loop_longtargets` (plan7.pyx:7541-7663):
# p7_Pipeline_LongTarget window by window, both strands -- is replaced by the engine's long-target search (stage by stage over
# ALL windows: pyhmmer_b200/longtarget.py over the C ABI); the records it returns become real P7_HITs (b2h_glue_longtarget_fill)
# and the reference's own tail -- sort by key, p7_tophits_Threshold, output tallies -- finishes the job.
_MIRROR_BLOCKS = {}                 # id(pyhmmer block) -> (pyhmmer block, mirror block): the device copy stays with the mirror block


def _mirror_block(m_easel, m_abc, DigitalSequenceBlock sequences):
    hit = _MIRROR_BLOCKS.get(id(sequences))
    if hit is not None and hit[0] is sequences and len(hit[1]) == len(sequences) and \
            all(len(a) == len(b) for a, b in zip(hit[1], sequences)):
        return hit[1]
    import numpy as np
    block = m_easel.DigitalSequenceBlock(m_abc, [
        m_easel.DigitalSequence(m_abc, name=q.name, description=q.description or "", accession=q.accession or "",
                                sequence=np.frombuffer(q.sequence, dtype=np.uint8).copy()) for q in sequences])
    while len(_MIRROR_BLOCKS) >= 2:
        _MIRROR_BLOCKS.pop(next(iter(_MIRROR_BLOCKS)))
    _MIRROR_BLOCKS[id(sequences)] = (sequences, block)
    return block


class CudaLongTargetsPipeline(LongTargetsPipeline):
    """`pyhmmer.plan7.LongTargetsPipeline` whose search runs on the GPU.  Same constructor (+ ``host_threads``), same results:
    `search_hmm` with an `HMM` query and a `DigitalSequenceBlock`; `search_seq` / `search_msa` build their model with pyhmmer's
    `Builder` and arrive here.  Other argument types go to the reference's own loop."""

    def __init__(self, alphabet, background=None, *, host_threads=0, **kwargs):
        super().__init__(alphabet, background, **kwargs)
        self._b2h_host_threads = int(host_threads)
        engine(0)                                            # fails loudly without a device

    def search_hmm(self, query, sequences):
        import ctypes, io
        import numpy as np
        from pyhmmer_b200 import easel as m_easel, plan7 as m_plan7, _lib as m_lib
        cdef LongTargetsPipeline me = <LongTargetsPipeline> self
        cdef DigitalSequenceBlock block
        cdef TopHits hits
        cdef HMM hmm
        cdef Profile gm
        cdef OptimizedProfile opt
        cdef int status
        cdef size_t nh, j
        cdef size_t a_hits = 0, a_doms = 0, a_dup = 0, a_pos = 0
        cdef const char* c_text = NULL
        cdef int64_t nseqs, nres, span
        if not isinstance(sequences, DigitalSequenceBlock) or not isinstance(query, HMM):
            return super().search_hmm(query, sequences)      # SequenceFile targets, Profile / OptimizedProfile queries: the reference's loop
        block, hmm = sequences, query
        if not me.alphabet._eq(hmm.alphabet):
            raise AlphabetMismatch(me.alphabet, hmm.alphabet)
        if not me.alphabet._eq(block.alphabet):
            raise AlphabetMismatch(me.alphabet, block.alphabet)
        # the same query and targets as objects of the engine's host layer (binary HMM format: every float32 as it is)
        m_abc = m_easel.Alphabet.rna() if me.alphabet.is_rna() else m_easel.Alphabet.dna()
        buf = io.BytesIO()
        hmm.write(buf, binary=True)
        buf.seek(0)
        with m_plan7.HMMFile(buf) as f:
            m_hmm = f.read()
        m_block = _mirror_block(m_easel, m_abc, block)
        # (read from the struct: enum p7_strands_e is TOPONLY=0, BOTTOMONLY=1, BOTH=2, hmmer.h:90 -- the `strand` getter of
        # plan7.pyx:7077 indexes (None, "watson", "crick") with it and answers "crick" for a search of both strands)
        strand = {0: "watson", 1: "crick"}.get(<int> me._pli.strands, None)
        opts = dict(F1=me._pli.F1, F2=me._pli.F2, F3=me._pli.F3, B1=me._pli.B1, B2=me._pli.B2, B3=me._pli.B3,
                    bias_filter=bool(me._pli.do_biasfilter), null2=bool(me._pli.do_null2), seed=self.seed, Z=self.Z, domZ=self.domZ,
                    bit_cutoffs=self.bit_cutoffs, strand=strand, block_length=int(me._pli.block_length),
                    window_length=self.window_length, window_beta=self.window_beta, host_threads=self._b2h_host_threads)
        m_pli = m_plan7.LongTargetsPipeline(m_abc, **opts)
        try:
            m_om, m_cut, (m_hits, m_doms, m_text, m_dup, stats) = m_pli._search_records(m_hmm, m_block)
        except m_plan7.MissingCutoffs:
            raise MissingCutoffs(query.name, self.bit_cutoffs)
        # the reference's optimized profile of the query (thresholds, names and model length of the hits come from it)
        L = self.L_HINT if len(block) == 0 else block._refs[0].L
        gm = Profile(hmm.M, me.alphabet)
        gm.configure(hmm, me.background, L)
        opt = gm.to_optimized()
        opt._om.max_length = int(m_om._desc.max_length)
        hits = TopHits(query)
        nh = len(m_hits)
        keep = []                                            # (the ctypes buffers must outlive the call)
        if nh:
            harr = (m_lib.HitRec * nh)(*m_hits); keep.append(harr)
            a_hits = <size_t> ctypes.addressof(harr)
            darr = m_doms.raw if getattr(m_doms, "raw", None) is not None and len(m_doms.raw) == len(m_doms) else (m_lib.DomainRec * len(m_doms))(*m_doms)
            keep.append(darr)
            a_doms = <size_t> ctypes.addressof(darr)
            dupa = np.ascontiguousarray(np.array(m_dup, dtype=np.uint8)); keep.append(dupa)
            a_dup = <size_t> dupa.ctypes.data
            text_b = bytes(m_text); keep.append(text_b)
            c_text = text_b
        pos = np.array([stats["pos_past_msv"], stats["pos_past_bias"], stats["pos_past_vit"], stats["pos_past_fwd"]], dtype=np.int64)
        a_pos = <size_t> pos.ctypes.data
        nseqs, nres = stats["nseqs"], stats["nres"]
        with nogil:
            me._pli.mode = p7_pipemodes_e.p7_SEARCH_SEQS
            me._pli.nseqs = 0
            status = b2h_glue_longtarget_fill(me._pli, hits._th, <const void*> a_hits, nh, <const void*> a_doms, c_text,
                                              <const unsigned char*> a_dup, <ESL_SQ *const *> block._refs, opt._om, me.background._bg,
                                              nseqs, nres, <const int64_t*> a_pos)
        if status == eslEINVAL:
            raise MissingCutoffs(query.name, self.bit_cutoffs)
        elif status != eslOK:
            raise UnexpectedError(status, "b2h_glue_longtarget_fill")
        with nogil:
            hits._sort_by_key()
            hits._threshold(me)
            hits._pli.n_output = hits._pli.pos_output = 0
            for j in range(hits._th.N):
                if (hits._th.hit[j].flags & 1) or (hits._th.hit[j].flags & 2):          # p7_IS_INCLUDED | p7_IS_REPORTED
                    hits._pli.n_output += 1
                    span = hits._th.hit[j].dcl[0].jali - hits._th.hit[j].dcl[0].iali
                    hits._pli.pos_output += 1 + (span if span >= 0 else -span)
        hits._query = query
        hits._empty = False
        return hits


def install():
    """Make `pyhmmer.hmmsearch` / `hmmscan` / `phmmer` / `jackhmmer` / `nhmmer` build `CudaPipeline` (`CudaLongTargetsPipeline`) objects in their workers (the
    `pipeline_class` hook of pyhmmer.hmmer._base._BaseWorker).  phmmer's `search_seq` / `search_msa` and jackhmmer's
    `IterativeSearch` (plan7.pyx:4273-4389) build their models with pyhmmer's own `Builder` and then call
    `pipeline.search_hmm`, i.e. every search iteration runs on the GPU.  Returns a function that undoes it."""
    import pyhmmer.hmmer._hmmsearch as hs, pyhmmer.hmmer._hmmscan as sc, pyhmmer.hmmer._phmmer as ph, pyhmmer.hmmer._jackhmmer as jk
    import pyhmmer.hmmer._nhmmer as nh
    saved = [(cls, cls.__dict__.get("pipeline_class")) for cls in (hs._SEARCHWorker, sc._SCANWorker, ph._PHMMERWorker, jk._JACKHMMERWorker, nh._NHMMERWorker)]
    for cls, _ in saved:
        cls.pipeline_class = CudaLongTargetsPipeline if cls is nh._NHMMERWorker else CudaPipeline

    def uninstall():
        for cls, old in saved:
            if old is None:
                try:
                    del cls.pipeline_class
                except AttributeError:
                    pass
            else:
                cls.pipeline_class = old
    return uninstall
