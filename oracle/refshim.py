"""ctypes wrapper over oracle/_ref/libhmmer_ref.so (the UNMODIFIED reference C library + ref_shim.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by pyhmmer_b200/.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libhmmer_ref.so")

_lib = None


class RefHit(ctypes.Structure):
    _fields_ = [("seq", ctypes.c_int), ("score", ctypes.c_float), ("pre_score", ctypes.c_float), ("sum_score", ctypes.c_float),
                ("nexpected", ctypes.c_float), ("lnP", ctypes.c_double), ("pre_lnP", ctypes.c_double), ("sum_lnP", ctypes.c_double),
                ("nregions", ctypes.c_int), ("nclustered", ctypes.c_int), ("noverlaps", ctypes.c_int), ("nenvelopes", ctypes.c_int),
                ("ndom", ctypes.c_int), ("best_domain", ctypes.c_int), ("dom_offset", ctypes.c_long)]


class RefDom(ctypes.Structure):
    _fields_ = [("ienv", ctypes.c_int), ("jenv", ctypes.c_int), ("iali", ctypes.c_int), ("jali", ctypes.c_int),
                ("envsc", ctypes.c_float), ("domcorrection", ctypes.c_float), ("dombias", ctypes.c_float), ("oasc", ctypes.c_float),
                ("bitscore", ctypes.c_float), ("lnP", ctypes.c_double),
                ("hmmfrom", ctypes.c_int), ("hmmto", ctypes.c_int), ("sqfrom", ctypes.c_int), ("sqto", ctypes.c_int), ("N", ctypes.c_int),
                ("text_offset", ctypes.c_long)]


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libhmmer_ref.so missing: run `make -C oracle` where /root/reference exists")
        L = ctypes.CDLL(LIB)
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        L.refm_read.restype = vp
        L.refm_read.argtypes = [ctypes.c_char_p, ci, ci]
        L.refm_free.argtypes = [vp]
        for f in ("refm_M", "refm_K", "refm_Kp", "refm_abc_type", "refm_max_length", "refm_Q16", "refm_Q8", "refm_Q4"):
            getattr(L, f).restype = ci
            getattr(L, f).argtypes = [vp]
        for f in ("refm_name", "refm_acc", "refm_desc"):
            getattr(L, f).restype = ctypes.c_char_p
            getattr(L, f).argtypes = [vp]
        for f in ("refm_evparam", "refm_cutoff", "refm_compo", "refm_bg_f", "refm_om_ints", "refm_om_floats"):
            getattr(L, f).argtypes = [vp, vp]
        L.refm_hmm_params.argtypes = [vp, vp, vp, vp]
        L.refm_gm_params.argtypes = [vp, vp, vp, vp]
        L.refm_om_tables.argtypes = [vp] * 7
        L.refm_set_length.argtypes = [vp, ci]
        L.refm_set_multihit.argtypes = [vp, ci, ci]
        for f in ("ref_ssv", "ref_msv", "ref_vit", "ref_fwd"):
            getattr(L, f).restype = ci
            getattr(L, f).argtypes = [vp, vp, ci, ctypes.POINTER(cf)]
        L.ref_fwdbck.restype = ci
        L.ref_fwdbck.argtypes = [vp, vp, ci, ctypes.POINTER(cf), ctypes.POINTER(cf), vp, vp]
        L.ref_null1.restype = cf
        L.ref_null1.argtypes = [vp, vp, ci]
        L.ref_bias.restype = cf
        L.ref_bias.argtypes = [vp, vp, ci]
        L.ref_generic.argtypes = [vp, vp, ci] + [ctypes.POINTER(cf)] * 4
        L.ref_longtarget_windows.argtypes = [vp, vp, ci, ctypes.c_double, ci, vp, vp, vp, vp, vp, vp, vp]
        L.ref_longtarget_stages.argtypes = [vp, vp, ci, ctypes.c_double, ctypes.c_double, ctypes.c_double, ci, ci, ci, ci, ci] + [vp] * 11
        L.ref_nhmmer.restype = ctypes.c_long
        L.ref_nhmmer.argtypes = [vp, ci, vp, vp, ctypes.c_long, ci, ctypes.c_double, ctypes.c_double, ctypes.c_double, ci, ci,
                                 ctypes.c_double, ctypes.c_double, ctypes.c_long, ctypes.c_long, vp, vp, vp, ctypes.c_char_p]
        L.ref_search_tables.restype = ctypes.c_long
        L.ref_search_tables.argtypes = [vp, vp, vp, ci, vp, vp, vp, ctypes.c_char_p]
        L.ref_max_length.argtypes = [vp, ctypes.c_double]
        L.ref_vit_longtarget.argtypes = [vp, vp, ci, ci, cf, ctypes.c_double, ci, vp]
        L.ref_longtarget_pipeline.argtypes = [vp, vp, ci, ctypes.c_double, ctypes.c_double, ctypes.c_double, ci, ci, ctypes.c_long, ci, vp, ci, vp, vp, ctypes.c_long]
        L.ref_gdecoding.argtypes = [vp, vp, ci, vp, vp, ctypes.POINTER(cf), ctypes.POINTER(cf), vp]
        L.ref_gumbel_surv.restype = ctypes.c_double
        L.ref_gumbel_surv.argtypes = [ctypes.c_double] * 3
        L.ref_exp_surv.restype = ctypes.c_double
        L.ref_exp_surv.argtypes = [ctypes.c_double] * 3
        L.ref_nxcells.restype = ci
        L.ref_search.restype = vp
        L.ref_search.argtypes = [vp, vp, vp, ci, ctypes.c_double, ctypes.c_double, ctypes.c_double, ci, ci, ctypes.c_uint]
        L.ref_result_free.argtypes = [vp]
        for f in ("ref_result_nhits", "ref_result_ndoms"):
            getattr(L, f).restype = ctypes.c_long
            getattr(L, f).argtypes = [vp]
        L.ref_result_hits.restype = ctypes.POINTER(RefHit); L.ref_result_hits.argtypes = [vp]
        L.ref_result_doms.restype = ctypes.POINTER(RefDom); L.ref_result_doms.argtypes = [vp]
        L.ref_result_text.restype = vp; L.ref_result_text.argtypes = [vp]
        L.ref_result_counters.restype = ctypes.POINTER(ctypes.c_long); L.ref_result_counters.argtypes = [vp]
        L.ref_search_mt.restype = ctypes.c_long
        L.ref_search_mt.argtypes = [vp, ci, vp, vp, ci, ci, ctypes.c_double, ctypes.c_double, ctypes.c_double, ci, ci, vp]
        L.refm_from_arrays.restype = vp
        L.refm_from_arrays.argtypes = [ci, ci, vp, vp, vp, vp, vp, ci, ctypes.c_char_p, ctypes.c_char_p, ci]
        L.refm_from_arrays_many.restype = ci
        L.refm_from_arrays_many.argtypes = [ci, ci, vp, vp, vp, vp, vp, vp, ctypes.c_char_p, ctypes.c_char_p, vp, ci, ci, vp]
        L.ref_scan.restype = vp
        L.ref_scan.argtypes = [vp, ci, vp, ctypes.c_long, ctypes.c_double, ctypes.c_double, ctypes.c_double, ci, ci, ctypes.c_uint]
        L.ref_scan_mt.restype = ctypes.c_long
        L.ref_scan_mt.argtypes = [vp, ci, vp, ctypes.c_long, ci, ctypes.c_double, ctypes.c_double, ctypes.c_double, ci, ci, vp]
        L.ref_nhmmer_mt.restype = ctypes.c_long
        L.ref_nhmmer_mt.argtypes = [vp, ci, vp, vp, ctypes.c_long, ci, ci, ctypes.c_double, ctypes.c_double, ctypes.c_double, ci, ci, ctypes.c_long, vp]
        _lib = L
    return _lib


def dsq_of(codes):
    """Residue codes -> Easel digital sequence with sentinels (esl_sq.h:100-102)."""
    codes = np.asarray(codes, dtype=np.uint8)
    d = np.full(codes.size + 2, 255, dtype=np.uint8)
    d[1:-1] = codes
    return d


class RefLtHit(ctypes.Structure):
    _fields_ = [(n, ctypes.c_long) for n in ("seqidx", "ienv", "jenv", "iali", "jali", "hmmfrom", "hmmto")] + \
               [(n, ctypes.c_float) for n in ("score", "bias", "pre_score", "envsc", "oasc")] + \
               [("lnP", ctypes.c_double), ("flags", ctypes.c_int), ("dom_reported", ctypes.c_int), ("dom_included", ctypes.c_int), ("pad", ctypes.c_int)]


def press(hmmpath, outbase):
    """hmmpress's .h3f / .h3p for an HMM file, written by the reference's p7_oprofile_Write (ref_press)."""
    L = lib()
    L.ref_press.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    n = L.ref_press(os.fsencode(hmmpath), os.fsencode(outbase))
    if n < 0:
        raise ValueError("cannot press %s" % hmmpath)
    return n


def scan_tables(models, codes, qname, qacc, qdesc, prefix):
    """hmmscan of one sequence against RefModel objects with default thresholds; the reference's tabular writers leave
    <prefix>.tbl / .domtbl / .pfam (ref_scan_tables).  Returns the hit count."""
    L = lib()
    L.ref_scan_tables.restype = ctypes.c_long
    L.ref_scan_tables.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_long, ctypes.c_char_p, ctypes.c_char_p,
                                  ctypes.c_char_p, ctypes.c_char_p]
    d = dsq_of(codes)
    hs = (ctypes.c_void_p * len(models))(*[m.h for m in models])
    enc = lambda v: None if v is None else (v if isinstance(v, bytes) else v.encode())
    return L.ref_scan_tables(hs, len(models), d.ctypes.data, d.size - 2, enc(qname), enc(qacc) or b"", enc(qdesc) or b"", os.fsencode(prefix))


def single_builder(alphabet_type, codes, name, path, matrix="BLOSUM62", popen=0.02, pextend=0.4, seed=42):
    """Builder.build of one query sequence by the reference (p7_SingleBuilder), HMM written to <path> (ref_single_builder)."""
    L = lib()
    L.ref_single_builder.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_double,
                                     ctypes.c_double, ctypes.c_uint, ctypes.c_char_p]
    d = dsq_of(codes)
    st = L.ref_single_builder(alphabet_type, d.ctypes.data, d.size - 2, name if isinstance(name, bytes) else name.encode(),
                              matrix.encode(), popen, pextend, seed, os.fsencode(path))
    if st != 0:
        raise ValueError("p7_SingleBuilder failed with status %d" % st)


def msa_builder(alphabet_type, rows, names, msaname, path, rf=None, architecture="fast", symfrac=0.5, fragthresh=0.5, effn=-1.0,
                laplace=False, seed=42):
    """Builder.build_msa of a digital alignment by the reference (p7_Builder), HMM written to <path>; returns the relative
    weights the builder left in the alignment (ref_msa_builder)."""
    L = lib()
    rows = np.ascontiguousarray(rows, np.uint8)
    nseq, alen = rows.shape
    L.ref_msa_builder.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.c_char_p,
                                  ctypes.c_char_p, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_uint,
                                  ctypes.c_char_p, ctypes.c_void_p]
    enc = lambda v: v if isinstance(v, bytes) else v.encode()
    arr = (ctypes.c_char_p * nseq)(*[enc(n) for n in names])
    wgt = np.zeros(nseq, np.float64)
    st = L.ref_msa_builder(alphabet_type, rows.ctypes.data, nseq, alen, arr, enc(msaname), enc(rf) if rf is not None else None,
                           1 if architecture == "hand" else 0, symfrac, fragthresh, effn, int(laplace), seed, os.fsencode(path), wgt.ctypes.data)
    if st != 0:
        raise ValueError("p7_Builder failed with status %d" % st)
    return wgt


def mt_stream(seed, n):
    """The first n esl_random() values of Easel's Mersenne Twister seeded <seed>."""
    L = lib()
    L.ref_mt_stream.argtypes = [ctypes.c_uint, ctypes.c_int, ctypes.c_void_p]
    out = np.zeros(n, np.float64)
    L.ref_mt_stream(seed, n, out.ctypes.data)
    return out


class RefModel:
    """One HMM of a file, configured as Pipeline.search_hmm configures an HMM query."""

    def __init__(self, path, index=0, L=400, _handle=None):
        self.L = lib()
        self.h = _handle if _handle is not None else self.L.refm_read(os.fsencode(path), index, L)
        if not self.h:
            raise ValueError("cannot read HMM %d of %s" % (index, path))
        self.M, self.K, self.Kp = self.L.refm_M(self.h), self.L.refm_K(self.h), self.L.refm_Kp(self.h)

    @classmethod
    def from_arrays(cls, abc_type, t, mat, ins, evparam, name, compo=None, consensus=None, max_length=0, L=400):
        """A model from probability arrays (refm_from_arrays): t [(M+1), 7], mat / ins [(M+1), K] float32; abc_type 2 = DNA,
        3 = amino (esl_alphabet.h)."""
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
        t, mat, ins, ev, compo = f32(t), f32(mat), f32(ins), f32(evparam), f32(compo)
        M = t.shape[0] - 1
        enc = lambda v: None if v is None else (v if isinstance(v, bytes) else v.encode())
        h = lib().refm_from_arrays(abc_type, M, t.ctypes.data, mat.ctypes.data, ins.ctypes.data, None if compo is None else compo.ctypes.data,
                                   ev.ctypes.data, int(max_length), enc(name), enc(consensus), L)
        return cls(None, _handle=h)

    def __del__(self):
        try:
            if self.h:
                self.L.refm_free(self.h)
        except Exception:
            pass

    def _vec(self, fn, n, dtype=np.float32):
        out = np.zeros(n, dtype=dtype)
        getattr(self.L, fn)(self.h, out.ctypes.data)
        return out

    evparam = property(lambda s: s._vec("refm_evparam", 6))
    cutoff = property(lambda s: s._vec("refm_cutoff", 6))
    compo = property(lambda s: s._vec("refm_compo", 20))
    bg_f = property(lambda s: s._vec("refm_bg_f", s.K))
    name = property(lambda s: s.L.refm_name(s.h))
    max_length = property(lambda s: s.L.refm_max_length(s.h))

    def hmm_params(self):
        M, K = self.M, self.K
        t = np.zeros((M + 1, 7), np.float32); mat = np.zeros((M + 1, K), np.float32); ins = np.zeros((M + 1, K), np.float32)
        self.L.refm_hmm_params(self.h, t.ctypes.data, mat.ctypes.data, ins.ctypes.data)
        return t, mat, ins

    def gm_params(self):
        M, Kp = self.M, self.Kp
        tsc = np.zeros((M + 1, 8), np.float32); rsc = np.zeros((Kp, M + 1, 2), np.float32); xsc = np.zeros((4, 2), np.float32)
        self.L.refm_gm_params(self.h, tsc.ctypes.data, rsc.ctypes.data, xsc.ctypes.data)
        return tsc, rsc, xsc

    def om_scalars(self):
        i = self._vec("refm_om_ints", 17, np.int32)
        f = self._vec("refm_om_floats", 12)
        return dict(tbm_b=i[0], tec_b=i[1], tjb_b=i[2], base_b=i[3], bias_b=i[4], base_w=i[5], ddbound_w=i[6],
                    xw=i[7:15].reshape(4, 2), L=i[15], mode=i[16],
                    scale_b=f[0], scale_w=f[1], nj=f[3], xf=f[4:12].reshape(4, 2))

    def om_tables_striped(self):
        M, Kp = self.M, self.Kp
        Q16, Q8, Q4 = self.L.refm_Q16(self.h), self.L.refm_Q8(self.h), self.L.refm_Q4(self.h)
        rbv = np.zeros((Kp, Q16, 16), np.uint8)
        rwv = np.zeros((Kp, Q8, 8), np.int16); twv = np.zeros((8 * Q8, 8), np.int16)
        rfv = np.zeros((Kp, Q4, 4), np.float32); tfv = np.zeros((8 * Q4, 4), np.float32)
        self.L.refm_om_tables(self.h, rbv.ctypes.data, None, rwv.ctypes.data, twv.ctypes.data, rfv.ctypes.data, tfv.ctypes.data)
        return rbv, rwv, twv, rfv, tfv

    def om_tables(self):
        """De-striped (node-major) tables: k = q + z*Q + 1 (p7_oprofile.c:800,856,949)."""
        M, Kp = self.M, self.Kp
        rbv, rwv, twv, rfv, tfv = self.om_tables_striped()

        def destripe(v):            # (Q, W) -> [k-1] for k = q + z*Q + 1
            return v.T.reshape(-1)[:M]

        msv = np.stack([destripe(rbv[x]) for x in range(Kp)])
        vr = np.stack([destripe(rwv[x]) for x in range(Kp)])
        fr = np.stack([destripe(rfv[x]) for x in range(Kp)])
        Q8, Q4 = rwv.shape[1], rfv.shape[1]
        vt = np.stack([destripe(twv[t:7 * Q8:7]) for t in range(7)] + [destripe(twv[7 * Q8:])])
        ft = np.stack([destripe(tfv[t:7 * Q4:7]) for t in range(7)] + [destripe(tfv[7 * Q4:])])
        return msv, vr, vt, fr, ft

    def _score(self, fn, codes):
        d = dsq_of(codes)
        sc = ctypes.c_float()
        st = getattr(self.L, fn)(self.h, d.ctypes.data, d.size - 2, ctypes.byref(sc))
        return sc.value, st

    def ssv(self, codes): return self._score("ref_ssv", codes)
    def msv(self, codes): return self._score("ref_msv", codes)
    def vit(self, codes): return self._score("ref_vit", codes)
    def fwd(self, codes): return self._score("ref_fwd", codes)

    def fwdbck(self, codes, want_x=False):
        d = dsq_of(codes); n = d.size - 2
        f, b = ctypes.c_float(), ctypes.c_float()
        nx = self.L.ref_nxcells()
        fx = np.zeros((n + 1, nx), np.float32); bx = np.zeros((n + 1, nx), np.float32)
        st = self.L.ref_fwdbck(self.h, d.ctypes.data, n, ctypes.byref(f), ctypes.byref(b), fx.ctypes.data, bx.ctypes.data)
        return (f.value, b.value, st, fx, bx) if want_x else (f.value, b.value, st)

    def generic(self, codes):
        """(p7_GMSV, p7_GViterbi, p7_GForward, p7_GBackward) of the generic profile reconfigured to the target's length."""
        d = dsq_of(codes)
        v = [ctypes.c_float() for _ in range(4)]
        self.L.ref_generic(self.h, d.ctypes.data, d.size - 2, *[ctypes.byref(x) for x in v])
        return tuple(x.value for x in v)

    def gdecoding(self, codes):
        """p7_GDecoding of the generic Forward/Backward matrices: (pp[(L+1),(M+1),3], xpp[(L+1),5], fwdsc, bcksc)."""
        d = dsq_of(codes); n = d.size - 2
        pp = np.zeros((n + 1, self.M + 1, 3), np.float32); xpp = np.zeros((n + 1, 5), np.float32)
        f, b = ctypes.c_float(), ctypes.c_float()
        dom = np.zeros((3, n + 1), np.float32)
        self.L.ref_gdecoding(self.h, d.ctypes.data, n, pp.ctypes.data, xpp.ctypes.data, ctypes.byref(f), ctypes.byref(b), dom.ctypes.data)
        return pp, xpp, f.value, b.value, dom

    def longtarget_windows(self, codes, F1=0.02, cap=100000):
        """First stage of nhmmer on one chunk: (raw SSV diagonals [n,3] = start n, model end k, length; their scores;
        merged windows [m,2] = start, length; prefix and suffix length tables)."""
        d = dsq_of(codes); n = d.size - 2
        nr, nm = ctypes.c_int(), ctypes.c_int()
        raw = np.zeros((cap, 3), np.int64); rsc = np.zeros(cap, np.float32); mer = np.zeros((cap, 2), np.int64)
        pre = np.zeros(self.M + 1, np.float32); suf = np.zeros(self.M + 1, np.float32)
        self.L.ref_longtarget_windows(self.h, d.ctypes.data, n, F1, cap, ctypes.byref(nr), raw.ctypes.data, rsc.ctypes.data,
                                      ctypes.byref(nm), mer.ctypes.data, pre.ctypes.data, suf.ctypes.data)
        assert nr.value <= cap
        return raw[:nr.value].copy(), rsc[:nr.value].copy(), mer[:nm.value].copy(), pre, suf

    def longtarget_stages(self, codes, F1=0.02, F2=3e-3, F3=3e-5, bias_filter=True, B1=100, B2=240, B3=1000, cap=200000):
        """The window-level stages of nhmmer on one chunk, restated with the reference's public calls (ref_longtarget_stages):
        a dict of msvwin [n,2], msvsc [n,3] (null1, FilterScore, MSV), msvflag [n], vithit [h,3] (msv window, i, k),
        vitwin [v,3] (msv window, n, length), vitsc [v,3] (null1, FilterScore, Forward), vitpass [v], counters [4]."""
        d = dsq_of(codes); n = d.size - 2
        nm, nh, nv = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        msvwin = np.zeros((cap, 2), np.int64); msvsc = np.zeros((cap, 3), np.float32); msvflag = np.zeros(cap, np.int32)
        vithit = np.zeros((cap, 3), np.int64)
        vitwin = np.zeros((cap, 3), np.int64); vitsc = np.zeros((cap, 3), np.float32); vitpass = np.zeros(cap, np.int32)
        counters = np.zeros(4, np.int64)
        a = lambda x: x.ctypes.data
        self.L.ref_longtarget_stages(self.h, a(d), n, F1, F2, F3, int(bias_filter), B1, B2, B3, cap,
                                     ctypes.addressof(nm), a(msvwin), a(msvsc), a(msvflag), ctypes.addressof(nh), a(vithit),
                                     ctypes.addressof(nv), a(vitwin), a(vitsc), a(vitpass), a(counters))
        assert max(nm.value, nh.value, nv.value) <= cap
        return dict(msvwin=msvwin[:nm.value].copy(), msvsc=msvsc[:nm.value].copy(), msvflag=msvflag[:nm.value].copy(),
                    vithit=vithit[:nh.value].copy(), vitwin=vitwin[:nv.value].copy(), vitsc=vitsc[:nv.value].copy(),
                    vitpass=vitpass[:nv.value].copy(), counters=counters)

    def nhmmer(self, seqs, block_length=0x40000, strand=None, F1=0.02, F2=3e-3, F3=3e-5, bias_filter=True, null2=True,
               E=10.0, incE=0.01, evalue_window=0, cap=100000, names=None, table_prefix=None):
        """nhmmer as pyhmmer's LongTargetsPipeline.search_hmm runs it (ref_nhmmer): (hits in final order as RefLtHit records,
        stats [6] = nres, nseqs, pos_past_msv, pos_past_bias, pos_past_vit, pos_past_fwd).  flags: 1 included, 2 reported,
        16 duplicate (p7_hitflags_e)."""
        dsqs = [dsq_of(c) for c in seqs]
        n = len(dsqs)
        ptrs = (ctypes.c_void_p * max(n, 1))(*[d.ctypes.data for d in dsqs])
        lens = (ctypes.c_long * max(n, 1))(*[d.size - 2 for d in dsqs])
        out = (RefLtHit * cap)()
        stats = (ctypes.c_long * 6)()
        nh = self.L.ref_nhmmer(self.h, n, ptrs, lens, block_length, {None: 0, "watson": 1, "crick": 2}[strand], F1, F2, F3,
                               int(bias_filter), int(null2), E, incE, int(evalue_window), cap, out, stats,
                               None if names is None else (ctypes.c_char_p * max(n, 1))(*[v if isinstance(v, bytes) else v.encode() for v in names]),
                               None if table_prefix is None else os.fsencode(table_prefix))
        assert 0 <= nh <= cap, nh
        return [out[i] for i in range(nh)], list(stats)

    def search_tables(self, seqs, names, accs, descs, prefix):
        """The reference search with default thresholds, its hits written by the reference's own tabular writers to
        <prefix>.tbl / .domtbl / .pfam (p7_tophits_TabularTargets / TabularDomains / TabularXfam).  Returns the hit count."""
        dsqs = [dsq_of(c) for c in seqs]
        n = len(dsqs)
        ptrs = (ctypes.c_void_p * max(n, 1))(*[d.ctypes.data for d in dsqs])
        lens = np.array([d.size - 2 for d in dsqs], dtype=np.int64)
        enc = lambda v: None if v is None else (v if isinstance(v, bytes) else v.encode())
        arr = lambda vals: (ctypes.c_char_p * max(n, 1))(*[enc(v) for v in vals])
        return self.L.ref_search_tables(self.h, ptrs, lens.ctypes.data, n, arr(names), arr(accs), arr(descs), os.fsencode(prefix))

    def search_msa(self, seqs, names, accs, descs, path, all_consensus_cols=False, trim=False):
        """The reference search with default thresholds, then TopHits.to_msa = p7_tophits_Alignment of the included domains,
        written to <path> in Pfam Stockholm format.  Returns the number of aligned sequences (0 = nothing included)."""
        dsqs = [dsq_of(c) for c in seqs]
        n = len(dsqs)
        ptrs = (ctypes.c_void_p * max(n, 1))(*[d.ctypes.data for d in dsqs])
        lens = np.array([d.size - 2 for d in dsqs], dtype=np.int64)
        enc = lambda v: None if v is None else (v if isinstance(v, bytes) else v.encode())
        arr = lambda vals: (ctypes.c_char_p * max(n, 1))(*[enc(v) for v in vals])
        self.L.ref_search_msa.restype = ctypes.c_long
        self.L.ref_search_msa.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p]
        return self.L.ref_search_msa(self.h, ptrs, lens.ctypes.data, n, arr(names), arr(accs), arr(descs),
                                     int(all_consensus_cols) | 2 * int(trim), os.fsencode(path))

    def max_length(self, beta=1e-7):
        """p7_Builder_MaxLength of the model's HMM."""
        return self.L.ref_max_length(self.h, beta)

    def vit_longtarget(self, codes, cfg_len, filtersc, F2=3e-3, cap=100000):
        """p7_ViterbiFilter_longtarget on one window: landmarks [n,2] = i, k in the reference's order."""
        d = dsq_of(codes); n = d.size - 2
        hit = np.zeros((cap, 2), np.int64)
        nh = self.L.ref_vit_longtarget(self.h, d.ctypes.data, n, int(cfg_len), float(filtersc), F2, cap, hit.ctypes.data)
        assert nh <= cap
        return hit[:nh].copy()

    def longtarget_pipeline(self, codes, F1=0.02, F2=3e-3, F3=3e-5, bias_filter=True, null2=True, start=1, complement=False, cap=20000,
                            want_text=False):
        """p7_Pipeline_LongTarget itself on one chunk: (counters [5] = pos_past_msv, pos_past_bias, pos_past_vit, pos_past_fwd,
        number of hits; hits [n,12] = ienv, jenv, iali, jali, score, bias, pre_score, lnP, envsc, oasc, hmmfrom, hmmto).
        <start> = sq->start; for the complement strand pass the reverse-complemented codes and the chunk's last coordinate."""
        d = dsq_of(codes); n = d.size - 2
        counters = np.zeros(5, np.int64); hits = np.zeros((cap, 12), np.float64)
        tcap = 1 << 22
        tbuf = ctypes.create_string_buffer(tcap) if want_text else None
        st = self.L.ref_longtarget_pipeline(self.h, d.ctypes.data, n, F1, F2, F3, int(bias_filter), int(null2), int(start), int(complement),
                                            counters.ctypes.data, cap, hits.ctypes.data, tbuf, tcap)
        if want_text:
            assert st == 0, st
            lines = tbuf.raw.split(b"\0")
            nh = min(cap, int(counters[4]))
            return counters, hits[:nh].copy(), [tuple(lines[4 * i:4 * i + 4]) for i in range(nh)]
        assert st == 0, st
        return counters, hits[:min(cap, int(counters[4]))].copy()

    def null1(self, codes):
        d = dsq_of(codes); return self.L.ref_null1(self.h, d.ctypes.data, d.size - 2)

    def bias(self, codes):
        d = dsq_of(codes); return self.L.ref_bias(self.h, d.ctypes.data, d.size - 2)


    def search(self, seqs, F1=0.02, F2=1e-3, F3=1e-5, bias_filter=True, null2=True, seed=42):
        """The reference search loop over <seqs> (list of residue-code arrays); returns (hits, doms, text, counters)."""
        dsqs = [dsq_of(c) for c in seqs]
        n = len(dsqs)
        ptrs = (ctypes.c_void_p * max(n, 1))(*[d.ctypes.data for d in dsqs])
        lens = np.array([d.size - 2 for d in dsqs], dtype=np.int64)
        r = self.L.ref_search(self.h, ptrs, lens.ctypes.data, n, F1, F2, F3, int(bias_filter), int(null2), seed)
        try:
            nh, nd = self.L.ref_result_nhits(r), self.L.ref_result_ndoms(r)
            hp, dp = self.L.ref_result_hits(r), self.L.ref_result_doms(r)
            hits = [RefHit.from_buffer_copy(hp[i]) for i in range(nh)]
            doms = [RefDom.from_buffer_copy(dp[i]) for i in range(nd)]
            ntext = sum(4 * (d.N + 1) for d in doms)
            text = ctypes.string_at(self.L.ref_result_text(r), ntext) if ntext else b""
            cp = self.L.ref_result_counters(r)
            counters = [cp[i] for i in range(4)]
        finally:
            self.L.ref_result_free(r)
        return hits, doms, text, counters


def search_mt(models, seqs, nthreads, F1=0.02, F2=1e-3, F3=1e-5, bias_filter=True, null2=True):
    """Timing run: every model against <seqs> with <nthreads> threads; returns (nhits, counters)."""
    L = lib()
    dsqs = [dsq_of(c) for c in seqs]
    n = len(dsqs)
    ptrs = (ctypes.c_void_p * max(n, 1))(*[d.ctypes.data for d in dsqs])
    lens = np.array([d.size - 2 for d in dsqs], dtype=np.int64)
    mp = (ctypes.c_void_p * len(models))(*[m.h for m in models])
    ctr = (ctypes.c_long * 4)()
    nh = L.ref_search_mt(mp, len(models), ptrs, lens.ctypes.data, n, nthreads, F1, F2, F3, int(bias_filter), int(null2), ctr)
    return nh, list(ctr)


def models_from_arrays(abc_type, Ms, t, mat, ins, evparam, names, compo=None, consensus=None, max_length=None, L=400, nthreads=1):
    """n models from concatenated arrays (refm_from_arrays_many, converted on <nthreads> threads): Ms [n]; t [sum(M+1), 7],
    mat / ins [sum(M+1), K], evparam [n, 6], compo [n, K] (None: p7_hmm_SetComposition) float32; names [n]; consensus = the
    concatenated consensus strings (sum(M) characters) or None; max_length [n] or None.  Returns RefModel objects."""
    Lb = lib()
    Ms = np.ascontiguousarray(Ms, dtype=np.int32)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    t, mat, ins, ev = f32(t), f32(mat), f32(ins), f32(evparam)
    compo = None if compo is None else f32(compo)
    maxl = None if max_length is None else np.ascontiguousarray(max_length, dtype=np.int32)
    n = len(Ms)
    out = (ctypes.c_void_p * max(n, 1))()
    enc = lambda v: v if isinstance(v, bytes) else v.encode()
    Lb.refm_from_arrays_many(abc_type, n, Ms.ctypes.data, t.ctypes.data, mat.ctypes.data, ins.ctypes.data,
                             None if compo is None else compo.ctypes.data, ev.ctypes.data,
                             None if consensus is None else enc(consensus), b"".join(enc(v) + b"\0" for v in names),
                             None if maxl is None else maxl.ctypes.data, L, max(1, int(nthreads)), out)
    return [RefModel(None, _handle=out[i]) for i in range(n)]


def _result(L, r):
    try:
        nh, nd = L.ref_result_nhits(r), L.ref_result_ndoms(r)
        hp, dp = L.ref_result_hits(r), L.ref_result_doms(r)
        hits = [RefHit.from_buffer_copy(hp[i]) for i in range(nh)]
        doms = [RefDom.from_buffer_copy(dp[i]) for i in range(nd)]
        ntext = sum(4 * (d.N + 1) for d in doms)
        text = ctypes.string_at(L.ref_result_text(r), ntext) if ntext else b""
        cp = L.ref_result_counters(r)
        counters = [cp[i] for i in range(4)]
    finally:
        L.ref_result_free(r)
    return hits, doms, text, counters


def scan(models, codes, F1=0.02, F2=1e-3, F3=1e-5, bias_filter=True, null2=True, seed=42):
    """hmmscan of one sequence against RefModel objects as Pipeline.scan_seq runs it (ref_scan), thresholds wide open:
    (hits, doms, text, counters); hit.seq = index of the MODEL."""
    L = lib()
    d = dsq_of(codes)
    hs = (ctypes.c_void_p * len(models))(*[m.h for m in models])
    return _result(L, L.ref_scan(hs, len(models), d.ctypes.data, d.size - 2, F1, F2, F3, int(bias_filter), int(null2), seed))


def scan_mt(models, codes, nthreads, F1=0.02, F2=1e-3, F3=1e-5, bias_filter=True, null2=True):
    """Timing run of hmmscan, the models spread over <nthreads> threads (ref_scan_mt): (number of reported hits, counters)."""
    L = lib()
    d = dsq_of(codes)
    hs = (ctypes.c_void_p * len(models))(*[m.h for m in models])
    ctr = (ctypes.c_long * 4)()
    nh = L.ref_scan_mt(hs, len(models), d.ctypes.data, d.size - 2, nthreads, F1, F2, F3, int(bias_filter), int(null2), ctr)
    return nh, list(ctr)


def nhmmer_mt(model, seqs, nthreads, block_length=0x40000, strand=None, F1=0.02, F2=3e-3, F3=3e-5, bias_filter=True, null2=True,
              evalue_window=0):
    """Timing run of nhmmer, the windows spread over <nthreads> threads (ref_nhmmer_mt): (hits after duplicate removal, stats [6])."""
    L = lib()
    dsqs = [dsq_of(c) for c in seqs]
    n = len(dsqs)
    ptrs = (ctypes.c_void_p * max(n, 1))(*[d.ctypes.data for d in dsqs])
    lens = (ctypes.c_long * max(n, 1))(*[d.size - 2 for d in dsqs])
    stats = (ctypes.c_long * 6)()
    nh = L.ref_nhmmer_mt(model.h, n, ptrs, lens, block_length, {None: 0, "watson": 1, "crick": 2}[strand], nthreads, F1, F2, F3,
                         int(bias_filter), int(null2), int(evalue_window), stats)
    assert nh >= 0, nh
    return nh, list(stats)
