"""Shared pieces of the long-target (nhmmer) stage tests: the synthetic DNA workload, a backend that computes every DP
of `pyhmmer_b200.longtarget.stages` with the REFERENCE's functions (oracle/_ref) so that the host logic can be checked
without a GPU, and the stage-by-stage comparison with `ref_longtarget_stages`."""
import ctypes

import numpy as np

from pyhmmer_b200 import _lib, easel, longtarget, synth


def dna_model(make_pair, M, seed=0, mu_shift=0.0, mask=False):
    dna = easel.Alphabet.dna()
    rng = np.random.default_rng(7000 + M + seed)
    h = synth.random_hmm(dna, M, rng, name="lt%d" % M)
    if mask:                                             # a model mask (MM line): nodes M/4 .. M/2 masked
        h.model_mask = "".join("m" if M // 4 <= k < M // 2 else "." for k in range(1, M + 1))
        h.match_emissions[M // 4:M // 2] = 0.25          # ... which emit the background, as hmmbuild leaves masked columns
    h.max_length = 2 * M + 50
    mu = -8.0 - np.log2(M) * 0.3 + mu_shift
    h._evparam[:] = np.array([mu, 0.70, mu - 1.0, 0.70, -4.0 + mu_shift / 2, 0.70], np.float32)
    return make_pair(h), rng


def dna_chunks(pair, rng, sizes, nplant):
    dna = pair.hmm.alphabet
    out = []
    for ci, L in enumerate(sizes):
        seq = rng.integers(0, dna.K, L).astype(np.uint8)
        for _ in range(nplant):
            dom = synth.emit_sequence(pair.hmm, rng)
            if len(dom) < L:
                pos = int(rng.integers(0, L - len(dom)))
                seq[pos:pos + len(dom)] = dom
        out.append(easel.DigitalSequence(dna, name=b"chunk%d" % ci, sequence=seq))
    return easel.DigitalSequenceBlock(dna, out)


class OracleBackend:
    """`longtarget.stages` backend over the reference library: every score comes from HMMER's own functions, the windows'
    extension / merging / cutting from the product's HOST code (b2h_longtarget_vit_finish) -- no device involved."""

    class _WDB:
        def __init__(self, res, off):
            self.res, self.off, self.n = res, off, len(off) - 1
            self.length = np.diff(off).astype(np.int32)

        def seq(self, i):
            return self.res[self.off[i]:self.off[i + 1]]

    def __init__(self, pair, block):
        self.pair, self.block, self.ref = pair, block, pair.ref
        self.hprof = ctypes.c_void_p()
        assert _lib.lib.b2h_profile_create_host(ctypes.byref(pair.om._desc), ctypes.byref(self.hprof)) == 0

    def __del__(self):
        _lib.lib.b2h_profile_destroy(self.hprof)

    def ssv_windows(self, F1):
        rows = []
        for ci, s in enumerate(self.block):
            if len(s) == 0:
                continue
            _, _, mer, _, _ = self.ref.longtarget_windows(s.sequence, F1=F1)
            rows += [(ci, 0, int(n), int(l), 0.0) for n, l in mer]
        return np.array(rows, dtype=np.dtype(_lib.WindowRec)) if rows else np.zeros(0, np.dtype(_lib.WindowRec))

    def window_db(self, seq, start, length):
        return self._WDB(*longtarget.window_residues(self.block, seq, start, length))

    def null_bias(self, wdb):
        return (np.array([self.ref.null1(wdb.seq(i)) for i in range(wdb.n)], np.float32),
                np.array([self.ref.bias(wdb.seq(i)) for i in range(wdb.n)], np.float32))

    def msv(self, wdb):
        return np.array([self.ref.msv(wdb.seq(i))[0] for i in range(wdb.n)], np.float32)

    def forward(self, wdb):
        return np.array([self.ref.fwd(wdb.seq(i))[0] for i in range(wdb.n)], np.float32)

    def viterbi_windows(self, wdb, filtersc, active, F2):
        maxl = int(self.pair.om._desc.max_length)
        rows = []
        for i in range(wdb.n):
            if active[i]:
                hit = self.ref.vit_longtarget(wdb.seq(i), min(int(wdb.length[i]), maxl), float(filtersc[i]), F2)
                rows += [(i, int(k), int(r), 1, 0.0) for r, k in hit]
        marks = np.array(rows, dtype=np.dtype(_lib.WindowRec)) if rows else np.zeros(0, np.dtype(_lib.WindowRec))
        rng = np.random.default_rng(len(rows))
        shuffled = marks[rng.permutation(len(marks))]               # the device reports landmarks in no particular order
        out, no = ctypes.c_void_p(), ctypes.c_size_t()
        assert _lib.lib.b2h_longtarget_vit_finish(self.hprof, _lib.ptr(shuffled), len(shuffled), _lib.ptr(wdb.length), wdb.n,
                                                  ctypes.byref(out), ctypes.byref(no)) == 0
        assert np.array_equal(shuffled, marks)                      # ... and the host code restores the reference's order
        return shuffled, longtarget._take_windows(out, no.value)


def _oracle_hits(self, wdb, window_start, seq_start, complement, target, prm):
    """The hit stage with the reference's parser specials: p7_ForwardParser / p7_BackwardParser per window (oracle/_ref),
    then the product's host code (b2h_longtarget_domains)."""
    n = wdb.n
    wins = (_lib.LtWindow * max(n, 1))()
    keep = []
    for i in range(n):
        codes = np.ascontiguousarray(wdb.seq(i))
        _, _, st, fx, bx = self.ref.fwdbck(codes, want_x=True)
        assert st == 0
        keep.append((codes, fx, bx))
        w = wins[i]
        w.dsq, w.L, w.fwd_xmx, w.bck_xmx = codes.ctypes.data, len(codes), fx.ctypes.data, bx.ctypes.data
        w.window_start, w.seq_start, w.complement, w.seq = int(window_start[i]), int(seq_start[i]), int(complement[i]), int(target[i])
    _lib.lib.b2h_profile_set_annotation(self.hprof, (self.pair.hmm.consensus or "x" * self.pair.hmm.M).encode(), None, None,
                                        self.pair.hmm.alphabet.symbols.encode())
    if self.pair.hmm.model_mask:
        _lib.lib.b2h_profile_set_model_mask(self.hprof, self.pair.hmm.model_mask.encode())
    out = ctypes.c_void_p()
    _lib.check(_lib.lib.b2h_longtarget_domains(self.hprof, wins, n, ctypes.byref(prm), ctypes.byref(out)), "b2h_longtarget_domains")
    try:
        return _lib.read_results(out)
    finally:
        _lib.lib.b2h_results_destroy(out)


OracleBackend.hits = _oracle_hits


def compare_nhmmer(pair, seqs, got, **kw):
    """`longtarget.search` against ref_nhmmer (pyhmmer's LongTargetsPipeline loop over the reference's functions): same hits
    with the same coordinates, scores to 2e-3 bits, same duplicates, same residue / position counters."""
    hits, doms, text, dup, stats = got
    rhits, rstats = pair.ref.nhmmer(seqs, **kw)
    assert [stats[k] for k in ("nres", "nseqs", "pos_past_msv", "pos_past_bias", "pos_past_vit", "pos_past_fwd")] == rstats, (stats, rstats)
    assert len(hits) == len(rhits), (len(hits), len(rhits))
    # the reference's final order: hit_sorter_by_sortkey (sortkey = -lnP, then name, strand, start)
    mine = sorted(range(len(hits)), key=lambda q: (hits[q].lnP, "seq%d" % hits[q].seq,
                                                    0 if doms[hits[q].dom_offset].iali < doms[hits[q].dom_offset].jali else 1,
                                                    doms[hits[q].dom_offset].iali))
    ndup = 0
    for q, r in zip(mine, rhits):
        h, d = hits[q], doms[hits[q].dom_offset]
        assert (h.seq, d.ienv, d.jenv, d.iali, d.jali, d.hmmfrom, d.hmmto) == (r.seqidx, r.ienv, r.jenv, r.iali, r.jali, r.hmmfrom, r.hmmto), (q, r.seqidx, r.iali, r.jali)
        assert abs(h.score - r.score) < 2e-3 and abs(d.dombias - r.bias) < 2e-3 and abs(h.pre_score - r.pre_score) < 2e-3
        assert abs(h.lnP - r.lnP) < 2e-3 and abs(d.envsc - r.envsc) < 2e-3 and abs(d.oasc - r.oasc) < 2e-3
        assert dup[q] == bool(r.flags & 16), (q, dup[q], r.flags)
        ndup += dup[q]
    return len(hits), ndup


def compare_with_reference(pair, block, got, exact_scores, fwd_ulps=2.0, **kw):
    """Stage by stage against ref_longtarget_stages, chunk by chunk.  Returns totals for the caller's sanity checks."""
    tot = dict(msvwin=0, vitmark=0, vitwin=0, passed=0)
    mw, vm, vw = got["msvwin"], got["vitmark"], got["vitwin"]
    for ci, s in enumerate(block):
        if len(s) == 0:
            assert not np.any(mw["seq"] == ci)
            continue
        ref = pair.ref.longtarget_stages(s.sequence, **kw)
        sel = np.flatnonzero(mw["seq"] == ci)
        assert len(sel) == len(ref["msvwin"]), (ci, len(sel), len(ref["msvwin"]))
        assert np.array_equal(mw["n"][sel], ref["msvwin"][:, 0]) and np.array_equal(mw["length"][sel], ref["msvwin"][:, 1]), ci
        sc = got["msvsc"][sel]
        assert np.array_equal(sc[:, 0], ref["msvsc"][:, 0]) and np.array_equal(sc[:, 2], ref["msvsc"][:, 2]), ci     # null1, MSV: exact
        if exact_scores:
            assert np.array_equal(sc[:, 1], ref["msvsc"][:, 1]), ci
        else:
            assert np.all(np.abs(sc[:, 1] - ref["msvsc"][:, 1]) <= 4 * np.spacing(np.abs(ref["msvsc"][:, 1]))), ci
        assert np.array_equal(got["msvflag"][sel], ref["msvflag"]), (ci, got["msvflag"][sel], ref["msvflag"])
        base = sel[0] if len(sel) else 0
        m = vm[(vm["seq"] >= base) & (vm["seq"] < base + len(sel))] if len(sel) else vm[:0]
        assert len(m) == len(ref["vithit"]), (ci, len(m), len(ref["vithit"]))
        assert np.array_equal(m["seq"] - base, ref["vithit"][:, 0]) and np.array_equal(m["n"], ref["vithit"][:, 1]) \
            and np.array_equal(m["k"], ref["vithit"][:, 2]), ci
        vsel = np.flatnonzero((vw["seq"] >= base) & (vw["seq"] < base + len(sel))) if len(sel) else np.zeros(0, np.int64)
        assert len(vsel) == len(ref["vitwin"]), (ci, len(vsel), len(ref["vitwin"]))
        assert np.array_equal(vw["seq"][vsel] - base, ref["vitwin"][:, 0]) and np.array_equal(vw["n"][vsel], ref["vitwin"][:, 1]) \
            and np.array_equal(vw["length"][vsel], ref["vitwin"][:, 2]), ci
        if len(vsel):
            vs = got["vitsc"][vsel]
            assert np.array_equal(vs[:, 0], ref["vitsc"][:, 0]), ci
            if not kw.get("bias_filter", True):
                vs = vs.copy(); vs[:, 1] = ref["vitsc"][:, 1]        # the reference does not run its bias filter here at all
            if exact_scores:
                assert np.array_equal(vs[:, 1:], ref["vitsc"][:, 1:]), ci
            else:
                assert np.all(np.abs(vs[:, 1] - ref["vitsc"][:, 1]) <= 4 * np.spacing(np.abs(ref["vitsc"][:, 1]))), ci
                # Forward: 1e-4 nats ABSOLUTE wherever float32 can express that; a score so large that its float32 spacing
                # exceeds 5e-5 nats (> 512 nats: windows of tens of kilobases) is compared in units of that spacing -- both
                # implementations add a rounded log(scale) to a float at every rescaling row (fwdback.c:429), so each carries a
                # few spacings of its own rounding noise
                dev = np.abs(vs[:, 2] - ref["vitsc"][:, 2])
                tol = np.maximum(1e-4, fwd_ulps * np.spacing(np.abs(ref["vitsc"][:, 2]).astype(np.float32)))
                assert np.all(dev <= tol), (ci, float(dev.max()), float(ref["vitsc"][:, 2][np.argmax(dev)]), float((dev / tol).max()))
            assert np.array_equal(got["vitpass"][vsel], ref["vitpass"]), ci
        assert np.array_equal(got["counters"][ci], ref["counters"]), (ci, got["counters"][ci], ref["counters"])
        tot["msvwin"] += len(sel); tot["vitmark"] += len(m); tot["vitwin"] += len(vsel); tot["passed"] += int(ref["vitpass"].sum())
    return tot
