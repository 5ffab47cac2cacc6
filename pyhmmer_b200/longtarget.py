"""The window-level stages of the long-target (nhmmer) pipeline, between the SSV scan and domain definition.

``p7_Pipeline_LongTarget`` (vendor/hmmer/src/p7_pipeline.c:1496-1706) scans a target chunk with the SSV filter, turns the
diagonals it finds into windows, and hands every window to ``p7_pli_postSSV_LongTarget`` (:1331-1436: bias gate, Viterbi
scan for landmarks, second round of windows) and ``p7_pli_postViterbi_LongTarget`` (:1065-1112: Forward gate).  Here the
windows of ALL chunks go through each stage together: the DP runs on the GPU through the C ABI (one launch per stage over
a database whose "sequences" are the windows), and this module does what the reference does between the DP calls -- the
P-value gates, the --B1/--B2/--B3 scaling of the bias correction, the bookkeeping of the ``pos_past_*`` counters -- in the
reference's own precision (which operands are float, which double, is part of the result).

`stages()` returns every intermediate, so the parity tests can compare stage by stage with the reference
(``oracle/ref_shim.c: ref_longtarget_stages``).  The compute sits behind a small backend interface; the only backend
shipped is the CUDA one (there is no CPU path in this package) -- the CPU-only tests drive the same host logic with the
reference's functions as the backend.

`search()` is the whole nhmmer loop for one profile (LongTargetsPipeline.search_hmm, src/pyhmmer/plan7.pyx:7258-7412): the
targets cut into windows with context, both strands, `stages()`, then the hit stage -- Forward / Backward parser specials of
the surviving windows from the GPU, the long-target branch of domain definition and the hit arithmetic of
``p7_pli_postViterbi_LongTarget`` on the host threads (``b2h_longtarget_hits``) -- and what ``p7_tophits`` does to nhmmer
hits: E-values over residues / window length, the position sort, duplicate removal.  `plan7.LongTargetsPipeline` wraps it.
"""
import ctypes
import math
import time

import numpy as np

from . import _lib
from ._lib import lib, check, ptr

LOG2 = 0.69314718055994529          # eslCONST_LOG2


def gumbel_surv(x, mu, lam):
    """esl_gumbel_surv (vendor/easel/esl_gumbel.c:129), double precision, libm's exp."""
    y = lam * (x - mu)
    ey = -math.exp(-y)
    return -ey if abs(ey) < 5e-9 else 1.0 - math.exp(ey)


def exp_surv(x, mu, lam):
    """esl_exp_surv (vendor/easel/esl_exponential.c:128)."""
    return 1.0 if x < mu else math.exp(-lam * (x - mu))


_null1_cache = {}


def null1(L):
    """p7_bg_SetLength + p7_bg_NullOne for a target of length L (p7_bg.c:189, 357), through the host library."""
    v = _null1_cache.get(L)
    if v is None:
        lp = _lib.LenParams()
        check(lib.b2h_length_params(int(L), 1.0, ctypes.byref(lp)), "b2h_length_params")
        v = _null1_cache[L] = np.float32(lp.null1)
    return v


def window_residues(block, seq, start, length):
    """Concatenated residues and offsets of the windows (seq[i], 1-based start[i], length[i]) of a sequence block: one memcpy
    per window on the host library's threads (b2h_pack_windows over the block's packed residues)."""
    res, off = block._packed()
    res = np.ascontiguousarray(res, np.uint8)
    length = np.ascontiguousarray(length, np.int64)
    n = len(length)
    woff = np.zeros(n + 1, np.int64)
    np.cumsum(length, out=woff[1:])
    src0 = np.ascontiguousarray(np.asarray(off, np.int64)[np.asarray(seq, np.int64)] + np.asarray(start, np.int64) - 1)
    if n and (src0.min() < 0 or (src0 + length).max() > len(res) or length.min() < 0):
        raise IndexError("window outside its sequence block")
    out = np.empty(int(woff[-1]), np.uint8)
    zero32 = np.zeros(n, np.int32)
    ptrs = (ctypes.c_void_p * 1)(res.ctypes.data)
    check(lib.b2h_pack_windows(ptrs, n, ptr(zero32), ptr(src0), ptr(length), ptr(zero32), None, 0, ptr(out), ptr(woff), 0), "b2h_pack_windows")
    return out, woff


class CudaBackend:
    """The DP of every stage on the GPU, one profile against the windows of all chunks (C ABI of include/b2h.h)."""

    def __init__(self, om, block):
        from . import plan7
        self.om, self.block = om, block
        self.ctx = _lib.context()
        self.prof = om._device(self.ctx)
        self.db = plan7.SequenceDatabase.of(self.ctx, block)
        self.max_length = int(om._desc.max_length)

    def ssv_windows(self, F1):
        from . import plan7
        return plan7.long_target_windows(self.om, self.block, F1)[1]

    class _WindowDB:
        def __init__(self, ctx, res, off, n):
            out = ctypes.c_void_p()
            check(lib.b2h_seqdb_create_packed(ctx.handle, ptr(res), ptr(off), n, ctypes.byref(out)), "b2h_seqdb_create_packed", ctx.handle)
            self.handle, self.n, self.length = out, n, np.diff(off).astype(np.int32)

        def __del__(self):
            try:
                lib.b2h_seqdb_destroy(self.handle)
            except Exception:
                pass

    def window_db(self, seq, start, length):
        res, off = window_residues(self.block, seq, start, length)
        return self._WindowDB(self.ctx, res, off, len(length))

    def _scores(self, fn, name, wdb):
        sc, st = np.empty(wdb.n, np.float32), np.empty(wdb.n, np.int32)
        check(fn(self.ctx.handle, self.prof, wdb.handle, ptr(sc), ptr(st)), name, self.ctx.handle)
        return sc

    def null_bias(self, wdb):
        n1, fs = np.empty(wdb.n, np.float32), np.empty(wdb.n, np.float32)
        check(lib.b2h_null_scores(self.ctx.handle, self.prof, wdb.handle, ptr(n1), ptr(fs)), "b2h_null_scores", self.ctx.handle)
        return n1, fs

    def msv(self, wdb):
        return self._scores(lib.b2h_msv_filter, "b2h_msv_filter", wdb)

    def forward(self, wdb):
        return self._scores(lib.b2h_forward_parser, "b2h_forward_parser", wdb)

    def hits(self, wdb, window_start, seq_start, complement, target, prm):
        """Forward / Backward parsers on the device for the windows behind the Forward gate, domain definition on the host
        threads (b2h_longtarget_hits): (hits, domains, text); hit.profile = window, hit.seq = target."""
        out = ctypes.c_void_p()
        a = [np.ascontiguousarray(window_start, np.int64), np.ascontiguousarray(seq_start, np.int64),
             np.ascontiguousarray(complement, np.int32), np.ascontiguousarray(target, np.int32)]
        check(lib.b2h_longtarget_hits(self.ctx.handle, self.prof, wdb.handle, ptr(a[0]), ptr(a[1]), ptr(a[2]), ptr(a[3]),
                                      ctypes.byref(prm), ctypes.byref(out)), "b2h_longtarget_hits", self.ctx.handle)
        try:
            return _lib.read_results(out)
        finally:
            lib.b2h_results_destroy(out)

    def viterbi_windows(self, wdb, filtersc, active, F2):
        marks, out = ctypes.c_void_p(), ctypes.c_void_p()
        nm, no = ctypes.c_size_t(), ctypes.c_size_t()
        act = np.ascontiguousarray(active, np.uint8)
        fsc = np.ascontiguousarray(filtersc, np.float32)
        check(lib.b2h_longtarget_viterbi_windows(self.ctx.handle, self.prof, wdb.handle, ptr(fsc), ptr(act), float(F2),
                                                 ctypes.byref(marks), ctypes.byref(nm), ctypes.byref(out), ctypes.byref(no)),
              "b2h_longtarget_viterbi_windows", self.ctx.handle)
        return _take_windows(marks, nm.value), _take_windows(out, no.value)


def _take_windows(p, n):
    dt = np.dtype(_lib.WindowRec)
    try:
        return np.frombuffer(ctypes.string_at(p, n * dt.itemsize), dtype=dt).copy() if n else np.zeros(0, dt)
    finally:
        lib.b2h_free(p)


class _Clock:
    """Wall-clock seconds per stage into a caller's dict (every backend call returns host results, i.e. is synchronous)."""

    def __init__(self, sink):
        self.sink, self.t = sink, time.perf_counter()

    def lap(self, name):
        now = time.perf_counter()
        if self.sink is not None:
            self.sink[name] = self.sink.get(name, 0.0) + (now - self.t)
        self.t = now


def stages(om, chunks, F1=0.02, F2=3e-3, F3=3e-5, bias_filter=True, B1=100, B2=240, B3=1000, backend=None, timings=None):
    """SSV windows -> MSV / bias gates -> Viterbi landmarks and windows -> Forward gate, for every chunk of ``chunks``.

    Returns a dict of numpy arrays:
      ``msvwin``  record array (seq = chunk, n, length) of the merged SSV windows, ``msvsc`` [n,3] = null1, FilterScore, MSV
      score per window, ``msvflag`` bit 0 = passed the MSV gate, bit 1 = passed the bias gate;
      ``vitmark`` record array (seq = index into msvwin, n = row, k) of the Viterbi landmarks, ``vitwin`` (seq = index into
      msvwin, n = start inside that window, length) the windows behind p7_pli_ExtendAndMergeWindows(.., 0.5) and the 80 kb cut,
      ``vitsc`` [v,3] = null1, FilterScore, Forward score, ``vitpass`` = passed the Forward gate;
      ``counters`` [nchunks,4] = pos_past_msv, pos_past_bias, pos_past_vit, pos_past_fwd per chunk (P7_PIPELINE, hmmer.h).
    """
    clock = _Clock(timings)
    be = backend if backend is not None else CudaBackend(om, chunks)
    clock.lap("upload")
    f32, f64 = np.float32, np.float64
    ev = [float(v) for v in om._evparam]                 # MMU MLAMBDA VMU VLAMBDA FTAU FLAMBDA
    max_length = int(om._desc.max_length)
    nchunks = len(chunks)
    counters = np.zeros((nchunks, 4), np.int64)
    wdt = np.dtype(_lib.WindowRec)
    out = dict(msvwin=np.zeros(0, wdt), msvsc=np.zeros((0, 3), f32), msvflag=np.zeros(0, np.int32), vitmark=np.zeros(0, wdt),
               vitwin=np.zeros(0, wdt), vitsc=np.zeros((0, 3), f32), vitpass=np.zeros(0, np.int32), counters=counters)

    mw = be.ssv_windows(F1)
    clock.lap("ssv_scan")
    n = len(mw)
    out["msvwin"] = mw
    if n == 0:
        return out
    wlen = mw["length"].astype(np.int64)
    wdb = be.window_db(mw["seq"], mw["n"], wlen)
    nul, bias = be.null_bias(wdb)                        # p7_bg_SetLength(window) + p7_bg_NullOne / p7_bg_FilterScore
    usc = be.msv(wdb)                                    # p7_oprofile_ReconfigMSVLength(window) + p7_MSVFilter
    clock.lap("window_msv_bias")
    out["msvsc"] = np.stack([nul, bias, usc], axis=1)
    # p7_Pipeline_LongTarget's gate (:1637): float difference, double division
    x = (usc - nul).astype(f32).astype(f64) / LOG2
    pass_msv = np.array([not (gumbel_surv(v, ev[0], ev[1]) > F1) for v in x], bool)
    flen = wlen.astype(f32)
    if bias_filter:
        # p7_pli_postSSV_LongTarget (:1359-1368): everything in float
        b = (bias - nul).astype(f32)
        filtersc = (nul + (b * (np.minimum(wlen, B1).astype(f32) / flen)).astype(f32)).astype(f32)
        seq_score = ((usc - filtersc).astype(f32).astype(f64) / LOG2).astype(f32)
        pass_bias = pass_msv & np.array([not (gumbel_surv(float(v), ev[0], ev[1]) > F1) for v in seq_score], bool)
    else:
        b = np.zeros(n, f32)
        pass_bias = pass_msv.copy()
    out["msvflag"] = pass_msv.astype(np.int32) + 2 * pass_bias.astype(np.int32)
    np.add.at(counters[:, 0], mw["seq"][pass_msv], wlen[pass_msv])
    np.add.at(counters[:, 1], mw["seq"][pass_bias], wlen[pass_bias])
    # null1 of the possibly shorter length model, B2-scaled bias: the ternary makes this one double (:1376-1382)
    nul_loc = np.array([null1(int(min(L, max_length))) for L in wlen], f32)
    ratio2 = np.minimum(wlen, B2).astype(f32) / flen
    filtersc2 = (nul_loc.astype(f64) + b.astype(f64) * ratio2.astype(f64)).astype(f32)
    clock.lap("host_gates")
    marks, vw = be.viterbi_windows(wdb, filtersc2, pass_bias, F2)
    clock.lap("viterbi_scan")
    out["vitmark"], out["vitwin"] = marks, vw
    nv = len(vw)
    if nv == 0:
        return out
    vlen = vw["length"].astype(np.int64)
    vchunk = mw["seq"][vw["seq"]]
    vdb = be.window_db(vchunk, mw["n"][vw["seq"]] + vw["n"] - 1, vlen)
    vnul, vbias = be.null_bias(vdb)
    fwd = be.forward(vdb)                                # p7_oprofile_ReconfigRestLength(window) + p7_ForwardParser
    clock.lap("window_forward")
    out["vitsc"] = np.stack([vnul, vbias, fwd], axis=1)
    vb = (vbias - vnul).astype(f32) if bias_filter else np.zeros(nv, f32)
    ratio3 = np.minimum(vlen, B3).astype(f32) / vlen.astype(f32)
    filtersc3 = (vnul.astype(f64) + vb.astype(f64) * ratio3.astype(f64)).astype(f32)
    seq_score = ((fwd - filtersc3).astype(f32).astype(f64) / LOG2).astype(f32)
    passed = np.array([not (exp_surv(float(v), ev[4], ev[5]) > F3) for v in seq_score], bool)
    out["vitpass"] = passed.astype(np.int32)
    # pos_past_vit / pos_past_fwd: lengths minus the overlap with the preceding window of the same SSV window (:1409-1430)
    vend = vw["n"].astype(np.int64) + vlen
    overlap = 0
    for i in range(nv):
        first = i == 0 or vw["seq"][i] != vw["seq"][i - 1]
        last = i == nv - 1 or vw["seq"][i + 1] != vw["seq"][i]
        if first:
            overlap = 0
        c = int(vchunk[i])
        counters[c, 2] += vlen[i] - (0 if first else max(0, int(vend[i - 1] - vw["n"][i])))
        if passed[i]:
            counters[c, 3] += vlen[i] - overlap
            overlap = 0 if last else max(0, int(vend[i] - vw["n"][i + 1]))
        else:
            overlap = 0
    clock.lap("host_gates")
    return out


# =====================================================================================================
# The whole search: windows over the targets, both strands, hits, E-values, duplicate removal
# (LongTargetsPipeline.search_hmm and _search_loop_longtargets, src/pyhmmer/plan7.pyx:7258-7412, 7541-7663)
# =====================================================================================================
_DNA_COMPLEMENT = "TGCA-YRKMSWDVBHN*~"      # of ACGT-RYMKSWHBVDN*~ (esl_alphabet.c: set_complementarity)


def complement_table(alphabet):
    """The digital complement of every residue code (ESL_ALPHABET.complement, esl_alphabet.c:create_dna / create_rna)."""
    if alphabet.K != 4:
        raise ValueError("reverse complement needs a nucleotide alphabet")
    return np.array([alphabet.symbols.index(c if c in alphabet.symbols else {"T": "U"}.get(c, c)) for c in
                     (_DNA_COMPLEMENT if "T" in alphabet.symbols else _DNA_COMPLEMENT.replace("T", "U"))], np.uint8)


def reverse_complement(alphabet, codes):
    """esl_sq_ReverseComplement on digital residues (nucleotide alphabets only)."""
    return np.ascontiguousarray(complement_table(alphabet)[codes[::-1]])


def pack_windows(alphabet, sequences, wins, host_threads=0):
    """The windows of `target_windows` as one packed residue buffer + offsets, cut (and reverse-complemented) by the host
    library's threads (b2h_pack_windows)."""
    n = len(wins)
    w = np.array(wins, dtype=np.int64).reshape(n, 6) if n else np.zeros((0, 6), np.int64)
    off = np.zeros(n + 1, np.int64)
    np.cumsum(w[:, 2], out=off[1:])
    res = np.empty(int(off[-1]), np.uint8)
    used = sorted(set(int(t) for t in w[:, 0]))
    codes = {t: np.ascontiguousarray(sequences[t].sequence, dtype=np.uint8) for t in used}
    ptrs = (ctypes.c_void_p * max(1, len(sequences)))()
    for t, c in codes.items():
        ptrs[t] = c.ctypes.data
    # a reverse-strand window is the same stretch of the target, reversed and complemented
    target, offset, length, comp = (np.ascontiguousarray(w[:, 0], np.int32), np.ascontiguousarray(w[:, 1], np.int64),
                                    np.ascontiguousarray(w[:, 2], np.int64), np.ascontiguousarray(w[:, 3], np.int32))
    table = complement_table(alphabet) if comp.any() else None
    check(lib.b2h_pack_windows(ptrs, n, ptr(target), ptr(offset), ptr(length), ptr(comp), None if table is None else ptr(table),
                               0 if table is None else len(table), ptr(res), ptr(off), int(host_threads)), "b2h_pack_windows")
    return res, off


class _Chunk:
    def __init__(self, codes):
        self.sequence = codes

    def __len__(self):
        return len(self.sequence)


class _ChunkBlock:
    """The windows of a search as one block: residue views into the targets (and their reverse complements), packed once
    for the upload -- the part of `easel.DigitalSequenceBlock` the backends use, without one object and one copy per window."""

    def __init__(self, alphabet, pieces):
        self.alphabet, self._pieces, self._cache = alphabet, pieces, {}

    @classmethod
    def from_packed(cls, alphabet, res, off):
        """A block over an already packed buffer: the pieces are views into it."""
        self = cls(alphabet, [res[off[i]:off[i + 1]] for i in range(len(off) - 1)])
        self._cache["packed"] = (res, off)
        return self

    def __len__(self):
        return len(self._pieces)

    def __getitem__(self, i):
        return _Chunk(self._pieces[i])

    def __iter__(self):
        return (_Chunk(p) for p in self._pieces)

    def _packed(self):
        hit = self._cache.get("packed")
        if hit is None:
            off = np.zeros(len(self._pieces) + 1, np.int64)
            np.cumsum([len(p) for p in self._pieces], out=off[1:])
            res = np.concatenate(self._pieces) if self._pieces else np.zeros(0, np.uint8)
            hit = self._cache["packed"] = (np.ascontiguousarray(res, dtype=np.uint8), off)
        return hit


def target_windows(lengths, W, C, strand=None):
    """The windows _search_loop_longtargets cuts the targets into: (target, offset i, n, complement) in its order --
    windows of W residues that keep C residues of the previous one as context; each on the strands asked for."""
    if C <= 0 or W <= C:
        raise ValueError("block length must be a strictly positive integer greater than the model's max_length")
    out = []
    for t, n in enumerate(lengths):
        i = 0
        while i < n:
            c = 0 if i == 0 else min(C, n - i)
            w = min(W, n - i - c)
            if strand != "crick":
                out.append((t, i, c + w, 0, w, c))
            if strand != "watson":
                out.append((t, i, c + w, 1, w, c))
            i += W - C
    return out


def search(om, sequences, F1=0.02, F2=3e-3, F3=3e-5, bias_filter=True, null2=True, B1=100, B2=240, B3=1000,
           block_length=0x40000, strand=None, seed=42, host_threads=0, backend_factory=None, evalue_window=None, world=None,
           timings=None, evalue_residues=None):
    """nhmmer for one profile: every target of ``sequences`` in windows, on both strands (or one), through `stages`
    and the hit stage; then p7_tophits_ComputeNhmmerEvalues, the seqidx / position sort, p7_tophits_RemoveDuplicates
    (p7_tophits.c:796, 426, 823).  Returns (hits, doms, text, duplicate flags, stats) with the hits in target order;
    stats = dict(nres, nseqs, pos_past_msv, pos_past_bias, pos_past_vit, pos_past_fwd).  hit.seq = target index.

    With several GPUs (one process each, ``torch.distributed``) the windows are dealt to the ranks in contiguous runs
    balanced by residues, every rank takes its windows through all stages, and ONE all-gather of the hit records
    (`parallel.all_gather_bytes`) gives every rank the same hits before the E-value / duplicate pass."""
    from . import parallel
    max_length = int(om._desc.max_length)
    lens = [len(s) for s in sequences]
    wins = target_windows(lens, int(block_length), max_length, strand)
    nres = sum(w for (_, _, _, _, w, _) in wins)         # pli->nres: W per window and strand (plan7.pyx:7606-7640)
    stats = dict(nres=int(nres), nseqs=len(lens), pos_past_msv=0, pos_past_bias=0, pos_past_vit=0, pos_past_fwd=0)
    if world is None:
        world = parallel.World.current()
    mine = wins
    if world.size > 1:
        bounds = parallel.shard_bounds([w[2] for w in wins], world.size)
        mine = wins[bounds[world.rank]:bounds[world.rank + 1]]
    hits, doms, text, counters = _search_windows(om, sequences, mine, F1, F2, F3, bias_filter, null2, B1, B2, B3, seed, host_threads,
                                                 backend_factory, timings)
    if world.size > 1:
        parts = [parallel.unpack_records(buf) for buf in parallel.all_gather_bytes(parallel.pack_records(hits, doms, text, counters, 0), world)]
        hits, doms, tbuf, counters = _lib.RecList(), _lib.RecList(), bytearray(), np.zeros(4, np.int64)
        for h, d, t, c in parts:                            # rank order = window order
            for r in h:
                r.dom_offset += len(doms)
            for r in d:
                r.text_offset += len(tbuf)
            hits.extend(h); doms.extend(d); tbuf.extend(t)
            counters = counters + c
        text = bytes(tbuf)
    for k, col in (("pos_past_msv", 0), ("pos_past_bias", 1), ("pos_past_vit", 2), ("pos_past_fwd", 3)):
        stats[k] = int(counters[col])
    if not hits:
        return _lib.RecList(), _lib.RecList(), b"", [], stats
    # p7_tophits_ComputeNhmmerEvalues: the search space is residues / window length
    # (a user-set Z replaces the residues searched by 1e6 Z per strand, plan7.pyx:7389-7396)
    add = math.log(float(np.float32(evalue_residues if evalue_residues is not None else stats["nres"])) / float(np.float32(evalue_window or max_length)))
    for h in hits:
        h.lnP += add
        doms[h.dom_offset].lnP = h.lnP
    # hit_sorter_by_seqidx_aliposition, then p7_tophits_RemoveDuplicates
    def poskey(h):
        d = doms[h.dom_offset]
        s, e = (d.iali, d.jali) if d.iali < d.jali else (d.jali, d.iali)
        return (h.seq, 0 if d.iali < d.jali else 1, s, -e)
    order = sorted(range(len(hits)), key=lambda q: poskey(hits[q]))
    dup = [False] * len(hits)
    j = 0
    for a in range(1, len(order)):
        hi, hj, hp = hits[order[a]], hits[order[j]], hits[order[a - 1]]
        di, dj = doms[hi.dom_offset], doms[hj.dom_offset]
        s_j, e_j = dj.iali, dj.jali
        dir_j = 1 if s_j < e_j else -1
        if dir_j == -1:
            s_j, e_j = e_j, s_j
        s_i, e_i = di.iali, di.jali
        dir_i = 1 if s_i < e_i else -1
        if dir_i == -1:
            s_i, e_i = e_i, s_i
        len_i, len_j = e_i - s_i + 1, e_j - s_j + 1
        ialilen = min(e_i, e_j) - max(s_i, s_j) + 1
        ihmmlen = min(di.hmmto, dj.hmmto) - max(di.hmmfrom, dj.hmmfrom) + 1
        if (hi.seq == hp.seq and dir_i == dir_j and ihmmlen > 0 and
                (s_j - 3 <= s_i <= s_j + 3 or e_j - 3 <= e_i <= e_j + 3 or ialilen >= len_i * 0.95 or ialilen >= len_j * 0.95)):
            remove = j if hi.lnP < hj.lnP else a
            dup[order[remove]] = True
            j = a if remove == j else j
        else:
            j = a
    hits_sorted = _lib.RecList(hits[q] for q in order)
    hits_sorted.raw = getattr(hits, "raw", None)
    return hits_sorted, doms, text, [dup[q] for q in order], stats


def _search_windows(om, sequences, wins, F1, F2, F3, bias_filter, null2, B1, B2, B3, seed, host_threads, backend_factory, timings=None):
    """All stages for a list of target windows (`target_windows` tuples): (hits, domains, text, pos_past_* [4])."""
    abc = om.alphabet
    none = (_lib.RecList(), _lib.RecList(), b"", np.zeros(4, np.int64))
    if not wins:
        return none
    clock = _Clock(timings)
    block = _ChunkBlock.from_packed(abc, *pack_windows(abc, sequences, wins, host_threads))
    clock.lap("cut_windows")
    be = (backend_factory or CudaBackend)(om, block)
    clock.lap("upload")
    st = stages(om, block, F1=F1, F2=F2, F3=F3, bias_filter=bias_filter, B1=B1, B2=B2, B3=B3, backend=be, timings=timings)
    clock.t = time.perf_counter()
    counters = st["counters"].sum(axis=0).astype(np.int64)
    sel = np.flatnonzero(st["vitpass"])
    if len(sel) == 0:
        return none[0], none[1], b"", counters
    mw, vw = st["msvwin"], st["vitwin"]
    chunk = mw["seq"][vw["seq"][sel]]                                   # chunk of every surviving window
    wstart = (mw["n"][vw["seq"][sel]] + vw["n"][sel] - 1).astype(np.int64)  # its first residue in the chunk
    wlen = vw["length"][sel].astype(np.int64)
    # sq->start of the chunk: its first target coordinate, or -- after esl_sq_ReverseComplement -- its last
    seq_start = np.array([(wins[c][1] + wins[c][2]) if wins[c][3] else (wins[c][1] + 1) for c in chunk], np.int64)
    comp = np.array([wins[c][3] for c in chunk], np.int32)
    target = np.array([wins[c][0] for c in chunk], np.int32)
    prm = _lib.SearchParams(F1, F2, F3, int(bias_filter), int(null2), seed, int(host_threads))
    hits, doms, text = be.hits(be.window_db(chunk, wstart, wlen), wstart, seq_start, comp, target, prm)
    clock.lap("hits")
    return hits, doms, text, counters
