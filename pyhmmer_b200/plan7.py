"""The search-path surface of ``pyhmmer.plan7`` on top of the B200 engine.

Model objects (`HMM`, `HMMFile`, `Background`, `Profile`, `OptimizedProfile`) hold plain numpy
arrays; every number that feeds the integer filters is produced by the host C++ in
``libb2h.so`` (``csrc/b2h_host.cpp``) so that it is bit-identical to what the reference's
``p7_ProfileConfig`` / ``p7_oprofile_Convert`` produce.  All dynamic programming happens in
CUDA kernels behind the C ABI of ``include/b2h.h``; nothing here scores on the CPU.

Reference: src/pyhmmer/plan7.pyx (HMM 2236-3321, HMMFile 3323-3800, Background 427-560,
Profile 7767-8310, OptimizedProfile 4392-5070, Pipeline 5423-6906, TopHits 8312-9278).
"""
import ctypes
import math
import os

import numpy as np

from . import _lib
from ._lib import lib, ptr, check, OProfileDesc
from .easel import Alphabet, DigitalSequence, DigitalSequenceBlock, AlphabetMismatch

__all__ = ["HMM", "HMMFile", "Background", "Profile", "OptimizedProfile", "EvalueParameters", "Cutoffs"]

P7_EVPARAM_UNSET = -99999.0
P7_CUTOFF_UNSET = -99999.0
P7_COMPO_UNSET = -1.0

# amino-acid background frequencies (Swiss-Prot 50.8; the data table of p7_AminoFrequencies, hmmer.c:161)
_AMINO_FREQ = np.array([0.0787945, 0.0151600, 0.0535222, 0.0668298, 0.0397062, 0.0695071, 0.0229198,
                        0.0590092, 0.0594422, 0.0963728, 0.0237718, 0.0414386, 0.0482904, 0.0395639,
                        0.0540978, 0.0683364, 0.0540687, 0.0673417, 0.0114135, 0.0304133], dtype=np.float32)


class Background:
    """The null model (``P7_BG``, vendor/hmmer/src/p7_bg.c:54-100)."""

    def __init__(self, alphabet, uniform=False):
        self.alphabet = alphabet
        self.uniform = uniform
        if alphabet.is_amino() and not uniform:
            self.residue_frequencies = _AMINO_FREQ.copy()
        else:
            self.residue_frequencies = np.full(alphabet.K, np.float32(1.0) / np.float32(alphabet.K), dtype=np.float32)
        self.L = 350
        self.omega = 1.0 / 256.0

    def copy(self):
        b = Background(self.alphabet, self.uniform)
        b.residue_frequencies = self.residue_frequencies.copy()
        b.L = self.L
        return b


class EvalueParameters:
    """``hmm.evparam`` accessor (plan7.pyx:1689-1848)."""
    _names = ("m_mu", "m_lambda", "v_mu", "v_lambda", "f_tau", "f_lambda")

    def __init__(self, vec):
        self._v = vec

    def as_vector(self):
        return self._v.copy()

    def __getattr__(self, name):
        if name in EvalueParameters._names:
            v = float(self._v[EvalueParameters._names.index(name)])
            return None if v == P7_EVPARAM_UNSET else v
        raise AttributeError(name)


class Cutoffs:
    """``hmm.cutoff`` accessor (plan7.pyx:1198-1439)."""

    def __init__(self, vec):
        self._v = vec

    def as_vector(self):
        return self._v.copy()

    def _pair(self, i):
        a, b = float(self._v[i]), float(self._v[i + 1])
        return None if a == P7_CUTOFF_UNSET or b == P7_CUTOFF_UNSET else (a, b)

    gathering = property(lambda self: self._pair(0))
    trusted = property(lambda self: self._pair(2))
    noise = property(lambda self: self._pair(4))

    def gathering_available(self):
        return self._pair(0) is not None

    def trusted_available(self):
        return self._pair(2) is not None

    def noise_available(self):
        return self._pair(4) is not None


class HMM:
    """A core profile HMM in probability space (``P7_HMM``).

    ``transition_probabilities`` is (M+1, 7) in the order MM MI MD IM II DM DD (hmmer.h:129),
    ``match_emissions`` / ``insert_emissions`` are (M+1, K); row 0 is the begin node.
    """

    def __init__(self, alphabet, M, name=b""):
        self.alphabet = alphabet
        self.M = int(M)
        self.name = bytes(name)
        self.accession = None
        self.description = None
        K = alphabet.K
        self.transition_probabilities = np.zeros((M + 1, 7), dtype=np.float32)
        self.match_emissions = np.zeros((M + 1, K), dtype=np.float32)
        self.insert_emissions = np.zeros((M + 1, K), dtype=np.float32)
        self._evparam = np.full(6, P7_EVPARAM_UNSET, dtype=np.float32)
        self._cutoff = np.full(6, P7_CUTOFF_UNSET, dtype=np.float32)
        self._compo = np.full(20, P7_COMPO_UNSET, dtype=np.float32)
        self.max_length = -1
        self.consensus = None
        self.consensus_structure = None
        self.reference = None
        self.model_mask = None
        self.map = None
        self.nseq = None
        self.nseq_effective = None
        self.checksum = None
        self.creation_time = None
        self.command_line = None

    @property
    def evalue_parameters(self):
        return EvalueParameters(self._evparam)

    @property
    def cutoffs(self):
        return Cutoffs(self._cutoff)

    @property
    def composition(self):
        return None if self._compo[0] == P7_COMPO_UNSET else self._compo[: self.alphabet.K].copy()

    def __repr__(self):
        return "<HMM name=%r M=%d alphabet=%r>" % (self.name, self.M, self.alphabet)

    def set_composition(self):
        """``p7_hmm_SetComposition`` (p7_hmm.c:621): occupancy-weighted mean emission."""
        t = self.transition_probabilities.astype(np.float64)
        M, K = self.M, self.alphabet.K
        mocc = np.zeros(M + 1)
        mocc[1] = t[0, 1] + t[0, 0]
        for k in range(2, M + 1):
            mocc[k] = mocc[k - 1] * (t[k - 1, 0] + t[k - 1, 1]) + (1.0 - mocc[k - 1]) * t[k - 1, 5]
        iocc = np.zeros(M + 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            iocc[0] = t[0, 1] / t[0, 3]
            iocc[1:] = mocc[1:] * t[1:, 1] / t[1:, 3]
        iocc = np.nan_to_num(iocc, nan=0.0, posinf=0.0)
        compo = (self.insert_emissions.astype(np.float64) * iocc[:, None]).sum(0) + \
                (self.match_emissions.astype(np.float64)[1:] * mocc[1:, None]).sum(0)
        compo /= compo.sum()
        self._compo[:K] = compo.astype(np.float32)

    def write(self, fh, binary=False):
        """Write the model in HMMER3/f ASCII format (``p7_hmmfile_WriteASCII``, p7_hmmfile.c:560-700)."""
        if binary:
            raise NotImplementedError("binary HMM output is outside the search path")
        abc = self.alphabet
        K = abc.K

        def w(s):
            fh.write(s.encode() if isinstance(s, str) else s)

        def prob(p):
            return "      *" if p == 0.0 else " %8.5f" % (-math.log(p))

        w("HMMER3/f [3.4 | Aug 2023]\n")
        w("NAME  %s\n" % self.name.decode())
        if self.accession:
            w("ACC   %s\n" % self.accession.decode())
        if self.description:
            w("DESC  %s\n" % self.description.decode())
        w("LENG  %d\n" % self.M)
        if self.max_length > 0:
            w("MAXL  %d\n" % self.max_length)
        w("ALPH  %s\n" % abc.type)
        w("RF    %s\n" % ("yes" if self.reference else "no"))
        w("MM    %s\n" % ("yes" if self.model_mask else "no"))
        w("CONS  %s\n" % ("yes" if self.consensus else "no"))
        w("CS    %s\n" % ("yes" if self.consensus_structure else "no"))
        w("MAP   %s\n" % ("yes" if self.map is not None else "no"))
        if self.nseq is not None:
            w("NSEQ  %d\n" % self.nseq)
        if self.nseq_effective is not None:
            w("EFFN  %f\n" % self.nseq_effective)
        for tag, i in (("GA", 0), ("TC", 2), ("NC", 4)):
            if self._cutoff[i] != P7_CUTOFF_UNSET:
                w("%s    %.2f %.2f\n" % (tag, self._cutoff[i], self._cutoff[i + 1]))
        if self._evparam[0] != P7_EVPARAM_UNSET:
            w("STATS LOCAL MSV      %8.4f %8.5f\n" % (self._evparam[0], self._evparam[1]))
            w("STATS LOCAL VITERBI  %8.4f %8.5f\n" % (self._evparam[2], self._evparam[3]))
            w("STATS LOCAL FORWARD  %8.4f %8.5f\n" % (self._evparam[4], self._evparam[5]))
        w("HMM     " + "".join("     %c   " % c for c in abc.symbols[:K]) + "\n")
        w("        %8s %8s %8s %8s %8s %8s %8s\n" % ("m->m", "m->i", "m->d", "i->m", "i->i", "d->m", "d->d"))
        if self._compo[0] != P7_COMPO_UNSET:
            w("  COMPO  " + " ".join(prob(p).strip().rjust(8) for p in self._compo[:K]) + "\n")
        for k in range(0, self.M + 1):
            if k > 0:
                w(" %6d  " % k + " ".join(prob(p).strip().rjust(8) for p in self.match_emissions[k]))
                w(" %6s" % (str(self.map[k]) if self.map is not None else "-"))
                w(" %c" % (self.consensus[k - 1] if self.consensus else "-"))
                w(" %c" % (self.reference[k - 1] if self.reference else "-"))
                w(" %c" % (self.model_mask[k - 1] if self.model_mask else "-"))
                w(" %c\n" % (self.consensus_structure[k - 1] if self.consensus_structure else "-"))
            w("         " + " ".join(prob(p).strip().rjust(8) for p in self.insert_emissions[k]) + "\n")
            w("         " + " ".join(prob(p).strip().rjust(8) for p in self.transition_probabilities[k]) + "\n")
        w("//\n")


def _decode_probs(tokens):
    """ASCII '-log p' fields -> float32 probabilities with the reference's expf (p7_hmmfile.c:1486)."""
    vals = np.array([math.inf if t == "*" else float(t) for t in tokens], dtype=np.float64)
    out = np.empty(vals.size, dtype=np.float32)
    check(lib.b2h_hmm_decode_probs(ptr(vals), ptr(out), vals.size), "b2h_hmm_decode_probs")
    return out


class HMMFile:
    """Iterate over the HMMs of a HMMER3 ASCII file (``pyhmmer.plan7.HMMFile``; read_asc30hmm, p7_hmmfile.c:1245)."""

    def __init__(self, file, db=True):
        if isinstance(file, (str, os.PathLike)):
            self.name = os.fspath(file)
            self._fh = open(file, "rb")
            self._own = True
        else:
            self.name = None
            self._fh = file
            self._own = False
        self._alphabet = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        if self._own:
            self._fh.close()

    def __iter__(self):
        return self

    def __next__(self):
        hmm = self.read()
        if hmm is None:
            raise StopIteration
        return hmm

    def _line(self):
        while True:
            line = self._fh.readline()
            if not line:
                return None
            if isinstance(line, bytes):
                line = line.decode("ascii", "replace")
            if line.strip() and not line.lstrip().startswith("#"):
                return line.rstrip("\n")

    def read(self):
        line = self._line()
        if line is None:
            return None
        if not line.startswith("HMMER3/"):
            raise ValueError("not a HMMER3 ASCII profile file (found %r)" % line[:20])
        fmt = line[7:8]
        hdr = {}
        ev = np.full(6, P7_EVPARAM_UNSET, dtype=np.float32)
        cut = np.full(6, P7_CUTOFF_UNSET, dtype=np.float32)
        stats = 0
        abc = None
        while True:
            line = self._line()
            if line is None:
                raise ValueError("premature end of HMM file in header")
            tag, _, rest = line.strip().partition(" ")
            rest = rest.strip()
            if tag == "HMM":
                break
            if tag == "ALPH":
                abc = {"amino": Alphabet.amino, "dna": Alphabet.dna, "rna": Alphabet.rna}[rest.lower()]()
            elif tag == "STATS":
                f = rest.split()
                if f[0] != "LOCAL":
                    raise ValueError("failed to parse STATS line")
                i = {"MSV": 0, "VITERBI": 2, "FORWARD": 4}[f[1].upper()]
                ev[i], ev[i + 1] = np.float32(float(f[2])), np.float32(float(f[3]))
                stats |= 1 << (i // 2)
            elif tag in ("GA", "TC", "NC"):
                f = rest.split()
                i = {"GA": 0, "TC": 2, "NC": 4}[tag]
                cut[i] = np.float32(float(f[0]))
                cut[i + 1] = cut[i] if (abc is not None and abc.is_nucleotide()) else np.float32(float(f[1]))
            else:
                hdr[tag] = rest
        if stats not in (0, 7):
            raise ValueError("missing one or more STATS parameter lines")
        if abc is None:
            raise ValueError("no ALPH found for HMM")
        if self._alphabet is not None and abc != self._alphabet:
            raise AlphabetMismatch(self._alphabet, abc)
        self._alphabet = abc
        M = int(hdr.get("LENG", "0"))
        if M <= 0 or "NAME" not in hdr:
            raise ValueError("no NAME / LENG found for HMM")
        K = abc.K
        self._line()                                   # the "m->m m->i ..." column header
        hmm = HMM(abc, M, hdr["NAME"].split()[0].encode())
        if "ACC" in hdr:
            hmm.accession = hdr["ACC"].split()[0].encode()
        if "DESC" in hdr:
            hmm.description = hdr["DESC"].encode()
        if "MAXL" in hdr:
            hmm.max_length = int(hdr["MAXL"])
        if "NSEQ" in hdr:
            hmm.nseq = int(hdr["NSEQ"])
        if "EFFN" in hdr:
            hmm.nseq_effective = float(hdr["EFFN"])
        if "CKSUM" in hdr:
            hmm.checksum = int(hdr["CKSUM"])
        if "COM" in hdr:
            hmm.command_line = hdr["COM"]
        flags = {k: hdr.get(k, "no").lower() == "yes" for k in ("RF", "MM", "CONS", "CS", "MAP")}
        hmm._evparam, hmm._cutoff = ev, cut

        toks = self._line().split()
        hmm._compo[:] = 0.0                             # p7_hmm_CreateBody zeroes compo; COMPO is optional
        if toks[0] == "COMPO":
            hmm._compo[:K] = _decode_probs(toks[1:1 + K])
            toks = self._line().split()
        rows_mat, rows_ins, rows_t = [None] * (M + 1), [None] * (M + 1), [None] * (M + 1)
        rows_ins[0] = toks[:K]
        rows_t[0] = self._line().split()[:7]
        anno = {"MAP": [], "CONS": [], "RF": [], "MM": [], "CS": []}
        for k in range(1, M + 1):
            f = self._line().split()
            if int(f[0]) != k:
                raise ValueError("expected match line to start with %d; saw %s" % (k, f[0]))
            rows_mat[k] = f[1:1 + K]
            extra = f[1 + K:]
            names = ["MAP"] + (["CONS"] if fmt >= "e" else []) + ["RF"] + (["MM"] if fmt >= "f" else []) + ["CS"]
            for nme, val in zip(names, extra):
                anno[nme].append(val)
            rows_ins[k] = self._line().split()[:K]
            rows_t[k] = self._line().split()[:7]
        end = self._line()
        if end is None or end.strip() != "//":
            raise ValueError("expected closing //")
        flat = _decode_probs([t for k in range(1, M + 1) for t in rows_mat[k]])
        hmm.match_emissions[1:] = flat.reshape(M, K)
        hmm.match_emissions[0, 0] = 1.0               # p7_hmm_CreateBody convention for the unused node 0
        hmm.insert_emissions[:] = _decode_probs([t for k in range(M + 1) for t in rows_ins[k]]).reshape(M + 1, K)
        hmm.transition_probabilities[:] = _decode_probs([t for k in range(M + 1) for t in rows_t[k]]).reshape(M + 1, 7)
        if flags["CONS"]:
            hmm.consensus = "".join(anno["CONS"])
        if flags["RF"]:
            hmm.reference = "".join(anno["RF"])
        if flags["MM"]:
            hmm.model_mask = "".join(anno["MM"])
        if flags["CS"]:
            hmm.consensus_structure = "".join(anno["CS"])
        if flags["MAP"]:
            hmm.map = np.array([0] + [int(v) for v in anno["MAP"]], dtype=np.int64)
        return hmm


class Profile:
    """A search profile in log-odds space (``P7_PROFILE``; p7_ProfileConfig, modelconfig.c:48)."""

    def __init__(self, M, alphabet):
        self.alphabet = alphabet
        self.M = int(M)
        self.L = 0
        self.multihit = True
        self.local = True
        self._configured = False

    def configure(self, hmm, background, L=400, multihit=True, local=True):
        if hmm.alphabet != self.alphabet:
            raise AlphabetMismatch(self.alphabet, hmm.alphabet)
        if background.alphabet != self.alphabet:
            raise AlphabetMismatch(self.alphabet, background.alphabet)
        if not local:
            raise NotImplementedError("glocal profiles are outside the search path (p7_Pipeline uses p7_LOCAL)")
        abc = self.alphabet
        M, K, Kp = hmm.M, abc.K, abc.Kp
        self.M = M
        self.tsc = np.empty((M, 8), dtype=np.float32)
        self.msc = np.empty((Kp, M + 1), dtype=np.float32)
        self.xsc = np.empty((4, 2), dtype=np.float32)
        t = np.ascontiguousarray(hmm.transition_probabilities, dtype=np.float32)
        mat = np.ascontiguousarray(hmm.match_emissions, dtype=np.float32)
        bgf = np.ascontiguousarray(background.residue_frequencies, dtype=np.float32)
        check(lib.b2h_profile_config(M, K, Kp, ptr(abc.degen), ptr(t), ptr(mat), ptr(bgf), int(L), int(bool(multihit)),
                                     ptr(self.tsc), ptr(self.msc), ptr(self.xsc)), "b2h_profile_config")
        self.L = int(L)
        self.multihit = bool(multihit)
        self.local = True
        self.name, self.accession, self.description = hmm.name, hmm.accession, hmm.description
        self.consensus, self.consensus_structure = hmm.consensus, hmm.consensus_structure
        self.reference, self.model_mask = hmm.reference, hmm.model_mask
        self._evparam, self._cutoff, self._compo = hmm._evparam.copy(), hmm._cutoff.copy(), hmm._compo.copy()
        self.max_length = hmm.max_length
        self._bgf = bgf
        self._configured = True
        return self

    @property
    def evalue_parameters(self):
        return EvalueParameters(self._evparam)

    @property
    def cutoffs(self):
        return Cutoffs(self._cutoff)

    def to_optimized(self):
        om = OptimizedProfile(self.M, self.alphabet)
        om.convert(self)
        return om


class OptimizedProfile:
    """The device-ready form of a profile (``P7_OPROFILE`` re-imagined node-major).

    ``convert`` runs the reference's three quantisations (p7_oprofile_Convert, p7_oprofile.c:1014)
    in the host C++ and keeps node-major tables; ``_device(ctx)`` uploads them once per context.
    """

    def __init__(self, M, alphabet):
        self.alphabet = alphabet
        self.M = int(M)
        self._desc = None
        self._dev = {}

    def convert(self, profile):
        if not profile._configured:
            raise ValueError("profile is not configured")
        abc = self.alphabet
        if profile.alphabet != abc:
            raise AlphabetMismatch(abc, profile.alphabet)
        M, K, Kp = profile.M, abc.K, abc.Kp
        self.M = M
        self.msv_cost = np.empty((Kp, M), dtype=np.uint8)
        self.vit_rsc = np.empty((Kp, M), dtype=np.int16)
        self.vit_tsc = np.empty((8, M), dtype=np.int16)
        self.fwd_rsc = np.empty((Kp, M), dtype=np.float32)
        self.fwd_tsc = np.empty((8, M), dtype=np.float32)
        d = OProfileDesc()
        check(lib.b2h_oprofile_convert(M, K, Kp, profile.L, int(profile.multihit),
                                       ptr(profile.tsc), ptr(profile.msc), ptr(profile.xsc),
                                       ptr(self.msv_cost), ptr(self.vit_rsc), ptr(self.vit_tsc),
                                       ptr(self.fwd_rsc), ptr(self.fwd_tsc), ctypes.byref(d)), "b2h_oprofile_convert")
        d.msv_cost, d.vit_rsc, d.vit_tsc = ptr(self.msv_cost), ptr(self.vit_rsc), ptr(self.vit_tsc)
        d.fwd_rsc, d.fwd_tsc = ptr(self.fwd_rsc), ptr(self.fwd_tsc)
        d.max_length = int(profile.max_length)
        for i in range(6):
            d.evparam[i] = float(profile._evparam[i])
            d.cutoff[i] = float(profile._cutoff[i])
        for i in range(20):
            d.compo[i] = float(profile._compo[i])
            d.bgf[i] = float(profile._bgf[i]) if i < K else 0.0
        d.degen = ptr(abc.degen)
        self._desc = d
        self._dev = {}
        self.name, self.accession, self.description = profile.name, profile.accession, profile.description
        self.consensus = profile.consensus
        self._evparam, self._cutoff, self._compo = profile._evparam, profile._cutoff, profile._compo
        self.L = profile.L
        self.multihit = profile.multihit
        return self

    # scalar views, named as on the reference object (plan7.pyx:4560-4860)
    tbm = property(lambda self: self._desc.tbm_b)
    tec = property(lambda self: self._desc.tec_b)
    tjb = property(lambda self: self._desc.tjb_b)
    base = property(lambda self: self._desc.base_b)
    bias = property(lambda self: self._desc.bias_b)
    scale_b = property(lambda self: self._desc.scale_b)
    base_w = property(lambda self: self._desc.base_w)
    scale_w = property(lambda self: self._desc.scale_w)
    ddbound_w = property(lambda self: self._desc.ddbound_w)

    @property
    def evalue_parameters(self):
        return EvalueParameters(self._evparam)

    @property
    def cutoffs(self):
        return Cutoffs(self._cutoff)

    def _device(self, ctx):
        h = self._dev.get(ctx)
        if h is None:
            out = ctypes.c_void_p()
            check(lib.b2h_profile_upload(ctx.handle, ctypes.byref(self._desc), ctypes.byref(out)),
                  "b2h_profile_upload", ctx.handle)
            h = self._dev[ctx] = _DeviceHandle(out, lib.b2h_profile_destroy)
        return h.handle

    def _filter_one(self, fn, seq):
        ctx = _lib.context()
        db = SequenceDatabase(ctx, DigitalSequenceBlock(self.alphabet, [seq]))
        sc = np.empty(1, np.float32)
        st = np.empty(1, np.int32)
        check(fn(ctx.handle, self._device(ctx), db.handle, ptr(sc), ptr(st)), fn.__name__, ctx.handle)
        return float(sc[0]), int(st[0])

    def msv_filter(self, seq):
        """``OptimizedProfile.msv_filter`` (plan7.pyx:4969): MSV score in nats, or ``inf`` on overflow."""
        return self._filter_one(lib.b2h_msv_filter, seq)[0]

    def ssv_filter(self, seq):
        """``OptimizedProfile.ssv_filter`` (plan7.pyx:5022): SSV score in nats, ``None`` if SSV cannot decide."""
        sc, st = self._filter_one(lib.b2h_ssv_filter, seq)
        return None if st == _lib.B2H_ENORESULT else sc


class _DeviceHandle:
    def __init__(self, handle, destroy):
        self.handle, self._destroy = handle, destroy

    def __del__(self):
        try:
            if self.handle:
                self._destroy(self.handle)
        except Exception:
            pass


class SequenceDatabase:
    """A `DigitalSequenceBlock` resident in HBM (``b2h_seqdb``)."""

    def __init__(self, ctx, block):
        res, off = block._packed()
        out = ctypes.c_void_p()
        check(lib.b2h_seqdb_create_packed(ctx.handle, ptr(res), ptr(off), len(block), ctypes.byref(out)),
              "b2h_seqdb_create_packed", ctx.handle)
        self._h = _DeviceHandle(out, lib.b2h_seqdb_destroy)
        self.ctx = ctx
        self.n = len(block)

    @property
    def handle(self):
        return self._h.handle

    @classmethod
    def of(cls, ctx, block):
        hit = block._cache.get(("db", ctx))
        if hit is None:
            hit = block._cache[("db", ctx)] = cls(ctx, block)
        return hit
