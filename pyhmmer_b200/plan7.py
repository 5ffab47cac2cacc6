"""The search-path surface of ``pyhmmer.plan7`` on top of the B200 engine.

Model objects (`HMM`, `HMMFile`, `Background`, `Profile`, `OptimizedProfile`) hold plain numpy
arrays; every number that feeds the integer filters is produced by the host C++ in
``libb2h.so`` (``csrc/b2h_host.cpp``) so that it is bit-identical to what the reference's
``p7_ProfileConfig`` / ``p7_oprofile_Convert`` produce.  All dynamic programming happens in
CUDA kernels behind the C ABI of ``include/b2h.h``; nothing here scores on the CPU.

Reference: src/pyhmmer/plan7.pyx (HMM 2236-3321, HMMFile 3323-3800, Background 427-560,
Profile 7767-8310, OptimizedProfile 4392-5070, Pipeline 5423-6906, TopHits 8312-9278).
"""
import collections
import copy
import ctypes
import math
import time
import os

import numpy as np

from . import _lib
from ._lib import lib, ptr, check, OProfileDesc
from .easel import Alphabet, DigitalSequence, DigitalSequenceBlock, AlphabetMismatch

__all__ = ["HMM", "HMMFile", "Background", "Profile", "OptimizedProfile", "OptimizedProfileBlock", "EvalueParameters", "Cutoffs",
           "Pipeline", "TopHits", "Hit", "Domain", "Domains", "Alignment", "MissingCutoffs", "SequenceDatabase"]

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

    def _set(self, i, value):                              # (plan7.pyx:1290-1439: a pair of floats, or None to unset)
        if value is None:
            self._v[i] = self._v[i + 1] = P7_CUTOFF_UNSET
        else:
            a, b = value
            self._v[i], self._v[i + 1] = float(a), float(b)

    gathering = property(lambda self: self._pair(0), lambda self, v: self._set(0, v))
    trusted = property(lambda self: self._pair(2), lambda self, v: self._set(2, v))
    noise = property(lambda self: self._pair(4), lambda self, v: self._set(4, v))

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

    def __init__(self, alphabet, M, name=""):
        self.alphabet = alphabet
        self.M = int(M)
        self.name = name.decode() if isinstance(name, (bytes, bytearray)) else str(name)
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

    def to_profile(self, background=None, L=400, multihit=True, local=True):
        """``HMM.to_profile`` (plan7.pyx:3329): a `Profile` configured for target length ``L``."""
        return Profile(self.M, self.alphabet).configure(self, background if background is not None else Background(self.alphabet),
                                                        L, multihit, local)

    def match_occupancy(self, inserts=False):
        """``HMM.match_occupancy`` = p7_hmm_CalculateOccupancy (p7_hmm.c:1338): the probability that a path visits each
        match state (index 0 unused), in the reference's float arithmetic; with ``inserts`` also the expected number of
        residues each insert state emits."""
        f32, f64 = np.float32, np.float64
        t, M = self.transition_probabilities, self.M
        mocc, iocc = np.zeros(M + 1, f32), np.zeros(M + 1, f32)
        mocc[1] = f32(t[0, 1] + t[0, 0])
        for k in range(2, M + 1):
            mocc[k] = f32(f64(f32(mocc[k - 1] * f32(t[k - 1, 0] + t[k - 1, 1]))) + (1.0 - f64(mocc[k - 1])) * f64(t[k - 1, 5]))
        if not inserts:
            return mocc
        iocc[0] = f32(t[0, 1] / t[0, 3])
        for k in range(1, M + 1):
            iocc[k] = f32(f32(mocc[k] * t[k, 1]) / t[k, 3])
        return mocc, iocc

    def set_composition(self):
        """``HMM.set_composition`` = p7_hmm_SetComposition (p7_hmm.c:1391): the occupancy-weighted mean emission distribution."""
        f32 = np.float32
        K = self.alphabet.K
        mocc, iocc = self.match_occupancy(inserts=True)
        compo = np.zeros(K, f32)
        compo += self.insert_emissions[0] * iocc[0]
        for k in range(1, self.M + 1):
            compo += self.match_emissions[k] * mocc[k]
            compo += self.insert_emissions[k] * iocc[k]
        s = c = f32(0.0)
        for v in compo:                                   # esl_vec_FNorm over a compensated sum
            y = f32(v - c)
            tt = f32(s + y)
            c = f32(f32(tt - s) - y)
            s = tt
        self._compo[:] = 0.0
        self._compo[:K] = compo / s

    def set_consensus(self, sequence=None):
        """``HMM.set_consensus`` = p7_hmm_SetConsensus (p7_hmm.c:1449): the most probable residue per node (or the residues of
        ``sequence``), upper case where its emission probability reaches 0.5 (0.9 for nucleotides)."""
        K, sym = self.alphabet.K, self.alphabet.symbols
        thresh = 0.9 if K == 4 else 0.5
        codes = np.asarray(sequence.sequence) if sequence is not None else self.match_emissions[1:, :K].argmax(axis=1)
        if len(codes) != self.M:
            raise ValueError("sequence length differs from the model length")
        # (a degenerate residue of the query has no emission probability of its own: lower case)
        self.consensus = "".join(sym[x].upper() if x < K and self.match_emissions[k + 1, x] >= thresh else sym[x].lower()
                                 for k, x in enumerate(codes))

    def mean_match_relative_entropy(self, background):
        """``HMM.mean_match_relative_entropy`` = p7_MeanMatchRelativeEntropy (modelstats.c:95), in bits."""
        K = self.alphabet.K
        bgf = np.asarray(background.residue_frequencies, np.float32)
        KL = 0.0
        for k in range(1, self.M + 1):                    # esl_vec_FRelEntropy: a float sum of p log2(p/q) per node
            kl = np.float32(0.0)
            for p, q in zip(self.match_emissions[k], bgf[:K]):
                if p > 0:
                    kl = np.float32(np.float64(kl) + np.float64(p) * math.log2(float(np.float32(p / q))))
            KL += float(kl)
        return KL / float(self.M)

    def compute_max_length(self, beta=1e-7):
        """``p7_Builder_MaxLength`` (p7_builder.c:651): the window length beyond which the model emits less than ``beta``
        of its probability mass -- what hmmbuild stores as MAXL and nhmmer recomputes for ``--w_beta``."""
        t = np.ascontiguousarray(self.transition_probabilities, dtype=np.float32)
        out = ctypes.c_int32()
        check(lib.b2h_hmm_max_length(self.M, ptr(t), float(beta), ctypes.byref(out)), "b2h_hmm_max_length")
        return int(out.value)

    def write(self, fh, binary=False):
        """Write the model in HMMER3/f ASCII format (``p7_hmmfile_WriteASCII``, p7_hmmfile.c:560-700)."""
        if binary:
            return self._write_binary(fh)
        abc = self.alphabet
        K = abc.K

        def w(s):
            fh.write(s.encode() if isinstance(s, str) else s)

        def prob(p):                                     # printprob (p7_hmmfile.c:2091): single-precision logf
            return "      *" if p == 0.0 else (" %8.5f" % 0.0 if p == 1.0 else " %8.5f" % (-_logf(float(p))))

        w("HMMER3/f [3.4 | Aug 2023]\n")
        w("NAME  %s\n" % self.name)
        if self.accession:
            w("ACC   %s\n" % self.accession)
        if self.description:
            w("DESC  %s\n" % self.description)
        w("LENG  %d\n" % self.M)
        if self.max_length > 0:
            w("MAXL  %d\n" % self.max_length)
        w("ALPH  %s\n" % abc.type)
        w("RF    %s\n" % ("yes" if self.reference else "no"))
        w("MM    %s\n" % ("yes" if self.model_mask else "no"))
        w("CONS  %s\n" % ("yes" if self.consensus else "no"))
        w("CS    %s\n" % ("yes" if self.consensus_structure else "no"))
        w("MAP   %s\n" % ("yes" if self.map is not None else "no"))
        if self.creation_time:
            w("DATE  %s\n" % self.creation_time)
        if self.command_line:
            for n, cmd in enumerate(self.command_line.split("\n")):
                w("COM   [%d] %s\n" % (n + 1, cmd))
        if self.nseq is not None and self.nseq > 0:
            w("NSEQ  %d\n" % self.nseq)
        if self.nseq_effective is not None and self.nseq_effective >= 0:
            w("EFFN  %f\n" % self.nseq_effective)
        if self.checksum is not None:
            w("CKSUM %d\n" % self.checksum)
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


def _libm_logf():
    m = ctypes.CDLL("libm.so.6")
    m.logf.restype, m.logf.argtypes = ctypes.c_float, [ctypes.c_float]
    return m.logf


_logf = _libm_logf()

_BIN_MAGIC = {0xe8ededb7: "b", 0xe8ededb8: "c", 0xe8ededb9: "d", 0xe8ededb0: "e", 0xe8ededba: "f"}     # p7_hmmfile.c:47-52
_ABC_TYPE = {"amino": 3, "dna": 2, "rna": 1}                                                           # esl_alphabet.h
_H = dict(DESC=1 << 1, RF=1 << 2, CS=1 << 3, STATS=1 << 7, MAP=1 << 8, ACC=1 << 9, GA=1 << 10, TC=1 << 11, NC=1 << 12, CA=1 << 13,
          COMPO=1 << 14, CHKSUM=1 << 15, CONS=1 << 16, MMASK=1 << 17)                                  # hmmer.h:109-126


def _hmm_write_binary(self, fh):
    """``p7_hmmfile_WriteBinary`` (p7_hmmfile.c:714-815), format 3/f."""
    import struct
    abc, M, K = self.alphabet, self.M, self.alphabet.K
    flags = 0
    for cond, bit in ((self.description, "DESC"), (self.reference, "RF"), (self.consensus_structure, "CS"),
                      (self._evparam[0] != P7_EVPARAM_UNSET, "STATS"), (self.map is not None, "MAP"), (self.accession, "ACC"),
                      (self._cutoff[0] != P7_CUTOFF_UNSET, "GA"), (self._cutoff[2] != P7_CUTOFF_UNSET, "TC"),
                      (self._cutoff[4] != P7_CUTOFF_UNSET, "NC"), (self._compo[0] != P7_COMPO_UNSET, "COMPO"),
                      (self.checksum is not None, "CHKSUM"), (self.consensus, "CONS"), (self.model_mask, "MMASK")):
        if cond:
            flags |= _H[bit]
    w = fh.write
    bstr = lambda v: struct.pack("<i", 0) if v is None else struct.pack("<i", len(v.encode()) + 1) + v.encode() + b"\0"
    line = lambda v: b" " + v.encode() + b"\0"                               # annotation lines are 1..M with a leading blank
    w(struct.pack("<Iiii", 0xe8ededba, flags, M, _ABC_TYPE[abc.type.lower()]))
    w(np.ascontiguousarray(self.match_emissions[1:], np.float32).tobytes())
    w(np.ascontiguousarray(self.insert_emissions, np.float32).tobytes())
    w(np.ascontiguousarray(self.transition_probabilities, np.float32).tobytes())
    w(bstr(self.name))
    if self.accession:
        w(bstr(self.accession))
    if self.description:
        w(bstr(self.description))
    for v in (self.reference, self.model_mask, self.consensus, self.consensus_structure):
        if v:
            w(line(v))
    w(bstr(self.command_line))
    w(struct.pack("<if", int(self.nseq or 0), float(self.nseq_effective or 0.0)))
    w(struct.pack("<i", int(self.max_length)))
    w(bstr(self.creation_time))
    if self.map is not None:
        w(np.ascontiguousarray(self.map, np.int32).tobytes())
    w(struct.pack("<I", int(self.checksum or 0)))
    w(np.ascontiguousarray(self._evparam, np.float32).tobytes())
    w(np.ascontiguousarray(self._cutoff, np.float32).tobytes())
    if flags & _H["COMPO"]:
        w(np.ascontiguousarray(self._compo[:K], np.float32).tobytes())


HMM._write_binary = _hmm_write_binary


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
        self._binary = None                    # format letter of a binary file (read_bin30hmm), False for ASCII, None = not looked yet
        self._push = b""                       # bytes read while looking, handed back to the line reader

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

    def is_pressed(self):
        """Whether the auxiliary files of a pressed database (hmmpress) sit next to this file (plan7.pyx:4009)."""
        return self.name is not None and all(os.path.exists(self.name + e) for e in (".h3f", ".h3p"))

    def optimized_profiles(self):
        """An iterator over the optimized profiles of the pressed database (plan7.pyx:4030)."""
        if not self.is_pressed():
            raise ValueError("HMM file does not contain optimized profiles")
        return HMMPressedFile(self.name)

    def _line(self):
        while True:
            line = self._fh.readline()
            if self._push:
                line, self._push = self._push + (line if isinstance(line, bytes) else line.encode()), b""
            if not line:
                return None
            if isinstance(line, bytes):
                line = line.decode("ascii", "replace")
            if line.strip() and not line.lstrip().startswith("#"):
                return line.rstrip("\n")

    def _need(self, n, what):
        b = self._fh.read(n)
        if len(b) != n:
            raise ValueError("binary HMM file: failed to read %s" % what)
        return b

    def _read_binary(self, first):
        """One model of a HMMER3 binary file (read_bin30hmm, p7_hmmfile.c:1585-1700); formats 3/b .. 3/f."""
        import struct
        if not first:
            magic = self._fh.read(4)
            if not magic:
                return None
            if len(magic) != 4 or _BIN_MAGIC.get(struct.unpack("<I", magic)[0]) != self._binary:
                raise ValueError("bad magic number at start of HMM")
        fmt = self._binary
        flags, M, atype = struct.unpack("<iii", self._need(12, "flags, model size, alphabet"))
        name = {v: k for k, v in _ABC_TYPE.items()}.get(atype)
        if name is None or M < 1:
            raise ValueError("binary HMM file: unsupported alphabet type %d or model size %d" % (atype, M))
        abc = getattr(Alphabet, name)()
        if self._alphabet is not None and abc != self._alphabet:
            raise AlphabetMismatch(self._alphabet, abc)
        self._alphabet = abc
        K = abc.K
        arr = lambda dt, n, what: np.frombuffer(self._need(n * np.dtype(dt).itemsize, what), dtype=dt).copy()

        def bstr(what):
            n = struct.unpack("<i", self._need(4, what))[0]
            return self._need(n, what).rstrip(b"\0").decode("ascii", "replace") if n > 0 else None

        line = lambda what: self._need(M + 2, what)[1:M + 1].decode("ascii", "replace")
        mat = arr(np.float32, M * K, "match emissions").reshape(M, K)
        ins = arr(np.float32, (M + 1) * K, "insert emissions").reshape(M + 1, K)
        tr = arr(np.float32, (M + 1) * 7, "transitions").reshape(M + 1, 7)
        hmm = HMM(abc, M, bstr("name") or "")
        hmm.match_emissions[1:] = mat
        hmm.match_emissions[0, 0] = 1.0
        hmm.insert_emissions[:] = ins
        hmm.transition_probabilities[:] = tr
        if flags & _H["ACC"]:
            hmm.accession = bstr("accession")
        if flags & _H["DESC"]:
            hmm.description = bstr("description")
        if flags & _H["RF"]:
            hmm.reference = line("rf")
        if flags & _H["MMASK"]:
            hmm.model_mask = line("mm")
        if flags & _H["CONS"]:
            hmm.consensus = line("consensus")
        if flags & _H["CS"]:
            hmm.consensus_structure = line("cs")
        if flags & _H["CA"]:
            line("ca")
        hmm.command_line = bstr("comlog")
        hmm.nseq, hmm.nseq_effective = struct.unpack("<if", self._need(8, "nseq"))
        if fmt >= "c":
            hmm.max_length = struct.unpack("<i", self._need(4, "max_length"))[0]
        hmm.creation_time = bstr("ctime")
        if flags & _H["MAP"]:
            hmm.map = arr(np.int32, M + 1, "map").astype(np.int64)
        hmm.checksum = struct.unpack("<I", self._need(4, "checksum"))[0]
        hmm._evparam[:] = arr(np.float32, 6, "statistical parameters")
        hmm._cutoff[:] = arr(np.float32, 6, "score cutoffs")
        hmm._compo[:] = 0.0
        if flags & _H["COMPO"]:
            hmm._compo[:K] = arr(np.float32, K, "composition")
        if not flags & _H["CHKSUM"]:
            hmm.checksum = None
        return hmm

    def read(self):
        first = self._binary is None
        if first:
            import struct
            head = self._fh.read(4)
            if isinstance(head, str):
                head = head.encode()
            self._binary = _BIN_MAGIC.get(struct.unpack("<I", head)[0], False) if len(head) == 4 else False
            if not self._binary:
                self._push = head
        if self._binary:
            return self._read_binary(first)
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
            elif tag == "COM":                           # "[n] command": the number is skipped, lines are joined (p7_hmmfile.c:1371-1381)
                cmd = rest.partition(" ")[2].strip() if rest.startswith("[") else rest
                hdr["COM"] = cmd if "COM" not in hdr else hdr["COM"] + "\n" + cmd
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
        hmm = HMM(abc, M, hdr["NAME"].split()[0])
        if "ACC" in hdr:
            hmm.accession = hdr["ACC"].split()[0]
        if "DESC" in hdr:
            hmm.description = hdr["DESC"]
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
        if "DATE" in hdr:
            hmm.creation_time = hdr["DATE"]
        flags = {k: hdr.get(k, "no").lower() == "yes" for k in ("RF", "MM", "CONS", "CS", "MAP")}
        hmm._evparam, hmm._cutoff = ev, cut

        # The body is a regular table -- per node one match line (k, K scores, annotation columns), one insert line (K) and one
        # transition line (7) -- so it is tokenised in one go and sliced as an array instead of line by line.
        raw = []
        while True:
            line = self._fh.readline()
            if not line:
                raise ValueError("expected closing //")
            if isinstance(line, str):
                line = line.encode("ascii", "replace")
            if b"//" in line and line.lstrip()[:2] == b"//":
                break
            raw.append(line)                            # (comment and blank lines are skipped by the field scanner)
        names = ["MAP"] + (["CONS"] if fmt >= "e" else []) + ["RF"] + (["MM"] if fmt >= "f" else []) + ["CS"]
        body = b"".join(raw)
        hmm._compo[:] = 0.0                             # p7_hmm_CreateBody zeroes compo; COMPO is optional
        compo = np.zeros(K, np.float32)
        has_compo = ctypes.c_int32()
        amap = np.zeros(M + 1, np.int64)
        achr = np.zeros((len(names) - 1, M), dtype="S1")
        mat, ins, tr = hmm.match_emissions, hmm.insert_emissions, hmm.transition_probabilities
        st = lib.b2h_hmm_parse_body(body, len(body), M, K, len(names), ptr(compo), ctypes.byref(has_compo), ptr(mat), ptr(ins), ptr(tr),
                                    ptr(amap), ptr(achr))
        if st == _lib.B2H_ERANGE:
            raise ValueError("HMM body of %r: a match line does not start with its node number" % hdr["NAME"])
        if st != _lib.B2H_OK:
            raise ValueError("HMM body of %r: malformed node table (%d nodes expected)" % (hdr["NAME"], M))
        if has_compo.value:
            hmm._compo[:K] = compo
        hmm.match_emissions[0, 0] = 1.0               # p7_hmm_CreateBody convention for the unused node 0
        anno = {nme: achr[j - 1] for j, nme in enumerate(names) if j > 0}
        text = lambda col: col.tobytes().decode("ascii")
        if flags["CONS"]:
            hmm.consensus = text(anno["CONS"])
        if flags["RF"]:
            hmm.reference = text(anno["RF"])
        if flags["MM"]:
            hmm.model_mask = text(anno["MM"])
        if flags["CS"]:
            hmm.consensus_structure = text(anno["CS"])
        if flags["MAP"]:
            hmm.map = amap
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

    def _generic_scores(self, sequences, nu=2.0, which=("msv", "viterbi", "forward", "backward")):
        """p7_GMSV / p7_GViterbi / p7_GForward / p7_GBackward (generic_msv.c:56, generic_viterbi.c:64,
        generic_fwdback.c:48,164) of this profile against every sequence of a block, on the GPU; each target is scored
        with the profile reconfigured to its length (p7_ReconfigLength).  Returns a dict of float32 arrays (nats)."""
        if not self._configured:
            raise ValueError("profile is not configured")
        ctx = _lib.context()
        block = sequences if isinstance(sequences, DigitalSequenceBlock) else DigitalSequenceBlock(self.alphabet, sequences)
        db = SequenceDatabase(ctx, block)
        n = len(block)
        out = {k: np.empty(n, np.float32) for k in which}
        check(lib.b2h_generic_scores(ctx.handle, self.M, self.alphabet.K, self.alphabet.Kp, ptr(self.tsc), ptr(self.msc), ptr(self.xsc),
                                     1.0 if self.multihit else 0.0, db.handle, float(nu),
                                     ptr(out.get("msv")), ptr(out.get("viterbi")), ptr(out.get("forward")), ptr(out.get("backward"))),
              "b2h_generic_scores", ctx.handle)
        return out

    def _generic_decoding(self, seq, domains=False):
        """p7_GDecoding (generic_decoding.c:77) of this profile (reconfigured to the target's length) and one sequence:
        (pp[(L+1), (M+1), 3], xpp[(L+1), 5], forward score, backward score); with ``domains`` also the (btot, etot, mocc)
        arrays of p7_GDomainDecoding (generic_decoding.c:207)."""
        ctx = _lib.context()
        codes = np.ascontiguousarray(seq.sequence, dtype=np.uint8)
        L = len(codes)
        pp = np.zeros((L + 1, self.M + 1, 3), np.float32)
        xpp = np.zeros((L + 1, 5), np.float32)
        f, b = ctypes.c_float(), ctypes.c_float()
        dom = [np.zeros(L + 1, np.float32) for _ in range(3)] if domains else [None] * 3
        check(lib.b2h_generic_decoding(ctx.handle, self.M, self.alphabet.K, self.alphabet.Kp, ptr(self.tsc), ptr(self.msc), ptr(self.xsc),
                                       1.0 if self.multihit else 0.0, ptr(codes), L, ptr(pp), ptr(xpp), ctypes.byref(f), ctypes.byref(b),
                                       ptr(dom[0]), ptr(dom[1]), ptr(dom[2])),
              "b2h_generic_decoding", ctx.handle)
        return (pp, xpp, f.value, b.value) + ((tuple(dom),) if domains else ())

    def msv_filter(self, seq, nu=2.0):
        """``Profile.msv_filter`` (plan7.pyx:8212-8253): the generic MSV score (p7_GMSV) of one sequence, in nats."""
        if seq.alphabet != self.alphabet:
            raise AlphabetMismatch(self.alphabet, seq.alphabet)
        return float(self._generic_scores([seq], nu=nu, which=("msv",))["msv"][0])


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
        self.reference, self.consensus_structure = profile.reference, profile.consensus_structure
        self.model_mask = profile.model_mask
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

    # -- the reference's striped views (plan7.pyx:4626-4813; impl_sse.h:75-142), rebuilt from the node-major tables:
    #    vector q, lane z holds node k = q + z*Q + 1; cells beyond M hold the "minus infinity" of their score system --
    @staticmethod
    def _stripe(row, width, pad):
        M = row.shape[-1]
        Q = max(2, (M - 1) // width + 1)
        out = np.full(row.shape[:-1] + (width, Q), pad, dtype=row.dtype)
        out.reshape(row.shape[:-1] + (width * Q,))[..., :M] = row
        return np.ascontiguousarray(np.swapaxes(out, -1, -2))      # [..., Q, width]

    @staticmethod
    def _stripe_transitions(t, width, pad):
        v = OptimizedProfile._stripe(t, width, pad)                 # [8, Q, width]
        Q = v.shape[1]
        out = np.empty((8 * Q, width), dtype=t.dtype)
        out[:7 * Q] = np.swapaxes(v[:7], 0, 1).reshape(7 * Q, width)     # per q: BM MM IM DM MD MI II
        out[7 * Q:] = v[7]                                           # then all DD
        return out

    @property
    def rbv(self):
        """Match costs of the MSV filter, striped: uint8 [Kp, Q16*16] (``P7_OPROFILE.rbv``)."""
        v = self._stripe(self.msv_cost, 16, 255)
        return v.reshape(v.shape[0], -1)

    @property
    def rwv(self):
        """ViterbiFilter match scores, striped: int16 [Kp, Q8*8] (``P7_OPROFILE.rwv``; the reference does not expose it)."""
        v = self._stripe(self.vit_rsc, 8, -32768)
        return v.reshape(v.shape[0], -1)

    @property
    def twv(self):
        """ViterbiFilter transition scores, striped: int16 [8*Q8*8] (``P7_OPROFILE.twv``)."""
        return self._stripe_transitions(self.vit_tsc, 8, -32768).reshape(-1)

    @property
    def rfv(self):
        """Forward / Backward match odds, striped: float32 [Kp, Q4*4] (``P7_OPROFILE.rfv``)."""
        v = self._stripe(self.fwd_rsc, 4, 0.0)
        return v.reshape(v.shape[0], -1)

    @property
    def tfv(self):
        """Forward / Backward transition odds, striped: float32 [8*Q4*4] (``P7_OPROFILE.tfv``)."""
        return self._stripe_transitions(self.fwd_tsc, 4, 0.0).reshape(-1)

    @property
    def xf(self):
        """Special-state odds, float32 [4, 2]: rows E N J C, columns MOVE LOOP as impl_sse orders them (``P7_OPROFILE.xf``)."""
        return np.array([list(r) for r in self._desc.xf], np.float32)

    def copy(self):
        """``OptimizedProfile.copy`` (p7_oprofile_Copy): own tables and scalars, nothing resident yet."""
        new = OptimizedProfile(self.M, self.alphabet)
        new.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ("_dev", "_desc", "_batch")})
        for t in ("msv_cost", "vit_rsc", "vit_tsc", "fwd_rsc", "fwd_tsc"):
            setattr(new, t, getattr(self, t).copy())
        new._evparam, new._cutoff, new._compo = self._evparam.copy(), self._cutoff.copy(), self._compo.copy()
        d = OProfileDesc.from_buffer_copy(self._desc)
        d.msv_cost, d.vit_rsc, d.vit_tsc = ptr(new.msv_cost), ptr(new.vit_rsc), ptr(new.vit_tsc)
        d.fwd_rsc, d.fwd_tsc = ptr(new.fwd_rsc), ptr(new.fwd_tsc)
        new._keep = (getattr(self, "_keep", None), self)            # what bgf / degen point at
        new._desc, new._dev = d, {}
        return new

    def __eq__(self, other):
        """``p7_oprofile_Compare`` (p7_oprofile.c:1500): tables and scalars, floats to a relative 1e-3."""
        if not isinstance(other, OptimizedProfile):
            return NotImplemented
        if (self.M, self.alphabet) != (other.M, other.alphabet) or self._desc is None or other._desc is None:
            return False
        a, b = self._desc, other._desc
        if any(getattr(a, f) != getattr(b, f) for f in ("tbm_b", "tec_b", "tjb_b", "base_b", "bias_b", "base_w", "ddbound_w", "mode_multihit", "L")):
            return False
        close = lambda x, y: np.allclose(x, y, rtol=1e-3, atol=0)
        if not (close(a.scale_b, b.scale_b) and close(a.scale_w, b.scale_w) and [list(r) for r in a.xw] == [list(r) for r in b.xw]):
            return False
        return (np.array_equal(self.msv_cost, other.msv_cost) and np.array_equal(self.vit_rsc, other.vit_rsc)
                and np.array_equal(self.vit_tsc, other.vit_tsc) and close(self.fwd_rsc, other.fwd_rsc) and close(self.fwd_tsc, other.fwd_tsc)
                and close(self.xf, other.xf))

    __hash__ = object.__hash__

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
            enc = lambda t: t.encode("ascii") if t else None
            lib.b2h_profile_set_annotation(out, enc(self.consensus), enc(self.reference), enc(self.consensus_structure),
                                           self.alphabet.symbols.encode("ascii"))
            if getattr(self, "model_mask", None):
                lib.b2h_profile_set_model_mask(out, enc(self.model_mask))
            h = self._dev[ctx] = _DeviceHandle(out, lib.b2h_profile_destroy)
        return h.handle

    @staticmethod
    def _device_many(ctx, oms):
        """Device handles for a list of profiles; the ones not yet resident are uploaded in one call."""
        todo = [om for om in oms if om._dev.get(ctx) is None]
        if len(todo) > 1:
            descs = (ctypes.c_void_p * len(todo))(*[ctypes.addressof(om._desc) for om in todo])
            outs = (ctypes.c_void_p * len(todo))()
            check(lib.b2h_profile_upload_many(ctx.handle, descs, len(todo), outs), "b2h_profile_upload_many", ctx.handle)
            enc = lambda t: t.encode("ascii") if t else None
            for om, h in zip(todo, outs):
                lib.b2h_profile_set_annotation(h, enc(om.consensus), enc(om.reference), enc(om.consensus_structure),
                                               om.alphabet.symbols.encode("ascii"))
                if getattr(om, "model_mask", None):
                    lib.b2h_profile_set_model_mask(h, enc(om.model_mask))
                om._dev[ctx] = _DeviceHandle(ctypes.c_void_p(h), lib.b2h_profile_destroy)
        return [om._device(ctx) for om in oms]

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


class _PyPressedFile:
    """Pure-Python reader of a pressed HMM database: the format documentation in executable form and the cross-check of
    the C reader behind `HMMPressedFile` (tests/test_host_cpu.py).

    ``hmmpress`` writes every model's vectorised score tables to ``<db>.h3f`` (the MSV part, read by
    p7_oprofile_ReadMSV, impl_sse/io.c:231) and ``<db>.h3p`` (everything else, p7_oprofile_ReadRest, io.c:498) in the
    SSE build's striped layout.  Both are read in step here and de-striped to the node-major tables the device wants
    (``b2h_destripe_oprofile``); no P7_HMM / P7_PROFILE is built and ``p7_oprofile_Convert`` never runs -- the path
    hmmscan takes over a Pfam-sized database.  Format 3/f (HMMER 3.1b2 .. 3.4); byte order and ``off_t`` as written by
    the x86-64 reference build.
    """

    _FMAGIC, _PMAGIC = 0xb3e6e6f3, 0xb3e6f0f3            # v3f_fmagic / v3f_pmagic, impl_sse/io.c:46-47
    _EXTRA_SB = 17                                        # p7O_EXTRA_SB, impl_sse.h:28
    _ABC = {3: "amino", 2: "dna", 1: "rna"}               # eslAMINO / eslDNA / eslRNA (esl_alphabet.h)

    def __init__(self, file):
        base = os.fspath(file)
        for ext in (".h3f", ".h3p"):
            if not os.path.exists(base + ext):
                raise ValueError("%r is not a pressed HMM database (%s is missing)" % (base, base + ext))
        self.name = base
        self._f = open(base + ".h3f", "rb")
        self._p = open(base + ".h3p", "rb")
        self._alphabet = None
        self._bg = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        self._f.close()
        self._p.close()

    @property
    def closed(self):
        return self._f.closed

    def rewind(self):
        self._f.seek(0)
        self._p.seek(0)

    def __iter__(self):
        return self

    def __next__(self):
        om = self.read()
        if om is None:
            raise StopIteration
        return om

    @staticmethod
    def _get(fh, fmt):
        import struct
        n = struct.calcsize(fmt)
        b = fh.read(n)
        if len(b) != n:
            raise ValueError("truncated pressed HMM database")
        v = struct.unpack(fmt, b)
        return v[0] if len(v) == 1 else v

    @staticmethod
    def _arr(fh, dtype, count):
        a = np.frombuffer(fh.read(np.dtype(dtype).itemsize * count), dtype=dtype)
        if a.size != count:
            raise ValueError("truncated pressed HMM database")
        return a

    def read(self):
        f, p, get, arr = self._f, self._p, self._get, self._arr
        head = f.read(4)
        if not head:
            return None
        import struct
        if struct.unpack("<I", head)[0] != self._FMAGIC:
            raise ValueError("bad magic in %s.h3f: not a 3/f pressed database (hmmpress it again with HMMER >= 3.1)" % self.name)
        # ---- .h3f: the MSV part (p7_oprofile_ReadMSV) ----
        M, atype, n = get(f, "<iii")
        name = f.read(n + 1)[:n].decode("ascii")
        max_length = get(f, "<i")
        tbm_b, tec_b, tjb_b = get(f, "<BBB")
        scale_b = get(f, "<f")
        base_b, bias_b = get(f, "<BB")
        if atype not in self._ABC:
            raise ValueError("unsupported alphabet type %d in %s" % (atype, self.name))
        abc = getattr(Alphabet, self._ABC[atype])()
        if self._alphabet is None:
            self._alphabet, self._bg = abc, Background(abc)
        elif abc != self._alphabet:
            raise AlphabetMismatch(self._alphabet, abc)
        K, Kp = abc.K, abc.Kp
        Q16, Q8, Q4 = max(2, (M - 1) // 16 + 1), max(2, (M - 1) // 8 + 1), max(2, (M - 1) // 4 + 1)
        f.seek(Kp * (Q16 + self._EXTRA_SB) * 16, 1)       # sbv: the SSV copy of the same scores, not needed (our table is built from rbv)
        rbv = arr(f, np.uint8, Kp * Q16 * 16)
        evparam = arr(f, np.float32, 6)
        f.seek(3 * 8, 1)                                  # offs[p7_NOFFSETS] (off_t): disk offsets into the .h3m/.h3f/.h3p files
        compo = arr(f, np.float32, 20)
        if get(f, "<I") != self._FMAGIC:
            raise ValueError("bad sentinel in %s.h3f: file corrupted?" % self.name)
        # ---- .h3p: the rest (p7_oprofile_ReadRest) ----
        if get(p, "<I") != self._PMAGIC:
            raise ValueError("bad magic in %s.h3p" % self.name)
        M2, atype2, n = get(p, "<iii")
        name2 = p.read(n + 1)[:n].decode("ascii")
        if M2 != M or atype2 != atype or name2 != name:
            raise ValueError("%s.h3f and .h3p are out of step (%s / %s)" % (self.name, name, name2))
        n = get(p, "<i")
        acc = p.read(n + 1)[:n].decode("ascii") if n > 0 else None
        n = get(p, "<i")
        desc = p.read(n + 1)[:n].decode("ascii") if n > 0 else None
        ann = [p.read(M + 2) for _ in range(4)]           # rf, mm, cs, consensus: 1..M, NUL at 0 when absent
        twv = arr(p, np.int16, 8 * Q8 * 8)
        rwv = arr(p, np.int16, Kp * Q8 * 8)
        xw = arr(p, np.int16, 8)
        scale_w = get(p, "<f")
        base_w, ddbound_w = get(p, "<hh")
        p.seek(4, 1)                                      # ncj_roundoff
        tfv = arr(p, np.float32, 8 * Q4 * 4)
        rfv = arr(p, np.float32, Kp * Q4 * 4)
        xf = arr(p, np.float32, 8)
        cutoff = arr(p, np.float32, 6)
        nj = get(p, "<f")
        mode, L = get(p, "<ii")
        if get(p, "<I") != self._PMAGIC:
            raise ValueError("bad sentinel in %s.h3p: file corrupted?" % self.name)

        om = OptimizedProfile(M, abc)
        om.msv_cost = np.empty((Kp, M), dtype=np.uint8)
        om.vit_rsc = np.empty((Kp, M), dtype=np.int16)
        om.vit_tsc = np.empty((8, M), dtype=np.int16)
        om.fwd_rsc = np.empty((Kp, M), dtype=np.float32)
        om.fwd_tsc = np.empty((8, M), dtype=np.float32)
        check(lib.b2h_destripe_oprofile(M, Kp, ptr(np.ascontiguousarray(rbv)), ptr(np.ascontiguousarray(rwv)), ptr(np.ascontiguousarray(twv)),
                                        ptr(np.ascontiguousarray(rfv)), ptr(np.ascontiguousarray(tfv)),
                                        ptr(om.msv_cost), ptr(om.vit_rsc), ptr(om.vit_tsc), ptr(om.fwd_rsc), ptr(om.fwd_tsc)),
              "b2h_destripe_oprofile")
        d = OProfileDesc()
        d.M, d.K, d.Kp, d.L, d.max_length = M, K, Kp, int(L), int(max_length)
        d.mode_multihit = int(nj > 0.0)                    # p7_LOCAL (multihit) has nj = 1, p7_UNILOCAL nj = 0
        d.msv_cost, d.vit_rsc, d.vit_tsc = ptr(om.msv_cost), ptr(om.vit_rsc), ptr(om.vit_tsc)
        d.fwd_rsc, d.fwd_tsc = ptr(om.fwd_rsc), ptr(om.fwd_tsc)
        d.tbm_b, d.tec_b, d.tjb_b, d.base_b, d.bias_b, d.scale_b = tbm_b, tec_b, tjb_b, base_b, bias_b, scale_b
        for i in range(4):
            for j in range(2):
                d.xw[i][j] = int(xw[i * 2 + j])
                d.xf[i][j] = float(xf[i * 2 + j])
        d.base_w, d.ddbound_w, d.scale_w = int(base_w), int(ddbound_w), scale_w
        bgf = self._bg.residue_frequencies
        for i in range(6):
            d.evparam[i] = float(evparam[i])
            d.cutoff[i] = float(cutoff[i])
        for i in range(20):
            d.compo[i] = float(compo[i])
            d.bgf[i] = float(bgf[i]) if i < K else 0.0
        d.degen = ptr(abc.degen)
        om._desc, om._dev = d, {}
        text = lambda b: b[1:M + 1].decode("ascii") if b[1:2] != b"\0" else None      # absent annotation = NUL at position 1 (io.c:552-555)
        om.name, om.accession, om.description = name, acc, desc
        om.reference, om.model_mask, om.consensus_structure, om.consensus = (text(a) for a in ann)
        om._evparam, om._cutoff, om._compo = evparam.copy(), cutoff.copy(), compo.copy()
        om.L, om.multihit = int(L), bool(d.mode_multihit)
        return om


class _PressedBatch:
    """One batch of b2h_pressed_read: owns the table block the batch's OptimizedProfile arrays are views of."""

    def __init__(self, models, n, block, nbytes, text, ntext):
        self.models, self.n, self.block, self.nbytes = models, n, block, nbytes
        self.text = ctypes.string_at(text, ntext) if ntext else b""
        lib.b2h_free(text)
        self.buf = (ctypes.c_uint8 * max(nbytes, 1)).from_address(block.value)
        self.base = block.value

    def view(self, addr, dtype, rows, cols):
        return np.frombuffer(self.buf, dtype=dtype, count=rows * cols, offset=addr - self.base).reshape(rows, cols)

    def __del__(self):
        try:
            lib.b2h_free(self.models)
            lib.b2h_free(self.block)
        except Exception:
            pass


class _ConvertedBlock:
    """The descriptors and the table block of one b2h_hmm_convert_many call; the OptimizedProfile arrays are views of it."""

    def __init__(self, descs, n, block, nbytes):
        self.descs, self.n, self.block = descs, n, block
        self.buf = (ctypes.c_uint8 * max(nbytes, 1)).from_address(block.value)
        self.base = block.value

    view = _PressedBatch.view

    def __del__(self):
        try:
            lib.b2h_free(self.descs)
            lib.b2h_free(self.block)
        except Exception:
            pass


def _convert_hmms(hmms, background, L, multihit=True, threads=0):
    """`OptimizedProfile` of every HMM of a list, configured for target length ``L`` against ``background``: the
    per-query ``Profile.configure`` + ``to_optimized`` of Pipeline.search_hmm (plan7.pyx:5979-6013) run for the whole
    query block by the host library on its own threads (``b2h_hmm_convert_many``), tables in one block."""
    n = len(hmms)
    if n == 0:
        return []
    abc = background.alphabet
    K, Kp = abc.K, abc.Kp
    bgf = np.ascontiguousarray(background.residue_frequencies, dtype=np.float32)
    keep = []
    arr = (_lib.HMMDesc * n)()
    for i, hmm in enumerate(hmms):
        if hmm.alphabet != abc:
            raise AlphabetMismatch(abc, hmm.alphabet)
        t = np.ascontiguousarray(hmm.transition_probabilities, dtype=np.float32)
        mat = np.ascontiguousarray(hmm.match_emissions, dtype=np.float32)
        keep.append((t, mat))
        h = arr[i]
        h.M, h.max_length = hmm.M, int(hmm.max_length)
        h.t, h.mat = t.ctypes.data, mat.ctypes.data
        h.evparam[:] = hmm._evparam.tolist()
        h.cutoff[:] = hmm._cutoff.tolist()
        h.compo[:] = hmm._compo.tolist()
    descs, block, nb = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t()
    nthreads = int(threads) if threads else min(n, max(1, len(os.sched_getaffinity(0))))
    check(lib.b2h_hmm_convert_many(K, Kp, ptr(abc.degen), ptr(bgf), int(L), int(bool(multihit)), arr, n, nthreads,
                                   ctypes.byref(descs), ctypes.byref(block), ctypes.byref(nb)), "b2h_hmm_convert_many")
    blk = _ConvertedBlock(descs, n, block, nb.value)
    blk.keep = (bgf, abc.degen)                            # the descriptors point at these
    recs = (OProfileDesc * n).from_address(descs.value)
    out = []
    for i, hmm in enumerate(hmms):
        d = recs[i]
        M = hmm.M
        om = OptimizedProfile(M, abc)
        om._batch = blk                                    # keeps the table block alive
        om.msv_cost = blk.view(d.msv_cost, np.uint8, Kp, M)
        om.vit_rsc = blk.view(d.vit_rsc, np.int16, Kp, M)
        om.vit_tsc = blk.view(d.vit_tsc, np.int16, 8, M)
        om.fwd_rsc = blk.view(d.fwd_rsc, np.float32, Kp, M)
        om.fwd_tsc = blk.view(d.fwd_tsc, np.float32, 8, M)
        om._desc, om._dev = d, {}
        om.name, om.accession, om.description = hmm.name, hmm.accession, hmm.description
        om.consensus = hmm.consensus
        om.reference, om.consensus_structure = hmm.reference, hmm.consensus_structure
        om.model_mask = hmm.model_mask
        om._evparam, om._cutoff, om._compo = hmm._evparam.copy(), hmm._cutoff.copy(), hmm._compo.copy()
        om.L, om.multihit = int(L), bool(multihit)
        out.append(om)
    return out


class HMMPressedFile:
    """Iterate over the `OptimizedProfile` of a pressed HMM database (``pyhmmer.plan7.HMMPressedFile``, plan7.pyx:4051).

    ``hmmpress`` writes every model's vectorised score tables to ``<db>.h3f`` (the MSV part, read by
    p7_oprofile_ReadMSV, impl_sse/io.c:231) and ``<db>.h3p`` (everything else, p7_oprofile_ReadRest, io.c:498) in the
    SSE build's striped layout.  The C library reads them in batches and de-stripes them straight into node-major
    tables (``b2h_pressed_read``); no P7_HMM / P7_PROFILE is built and ``p7_oprofile_Convert`` never runs -- the path
    hmmscan takes over a Pfam-sized database.  Format 3/f (HMMER 3.1b2 .. 3.4).
    """

    BATCH = 512
    _ABC = {3: "amino", 2: "dna", 1: "rna"}               # eslAMINO / eslDNA / eslRNA (esl_alphabet.h)

    def __init__(self, file):
        self._h = None
        base = os.fspath(file)
        for ext in (".h3f", ".h3p"):
            if not os.path.exists(base + ext):
                raise ValueError("%r is not a pressed HMM database (%s is missing)" % (base, base + ext))
        self.name = base
        h = ctypes.c_void_p()
        check(lib.b2h_pressed_open(base.encode(), ctypes.byref(h)), "b2h_pressed_open")
        self._h = h
        self._queue = []
        self._alphabet = self._bg = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        if self._h is not None:
            lib.b2h_pressed_close(self._h)
            self._h = None

    def __del__(self):
        self.close()

    @property
    def closed(self):
        return self._h is None

    def rewind(self):
        check(lib.b2h_pressed_rewind(self._h), "b2h_pressed_rewind")
        self._queue = []

    def __iter__(self):
        return self

    def __next__(self):
        om = self.read()
        if om is None:
            raise StopIteration
        return om

    def _fill(self):
        models, block, text = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        n, nb, nt = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
        st = lib.b2h_pressed_read(self._h, self.BATCH, ctypes.byref(models), ctypes.byref(n), ctypes.byref(block), ctypes.byref(nb),
                                  ctypes.byref(text), ctypes.byref(nt))
        if st != _lib.B2H_OK:
            raise ValueError("%s: %s" % (self.name, (lib.b2h_pressed_last_error(self._h) or b"").decode()))
        if n.value == 0:
            return
        batch = _PressedBatch(models, n.value, block, nb.value, text, nt.value)
        recs = (_lib.PressedModel * n.value).from_address(models.value)
        txt = batch.text

        def s(off):
            return None if off < 0 else txt[off:txt.index(b"\0", off)].decode("ascii")

        out = []
        for i in range(n.value):
            r = recs[i]
            d = r.desc
            if self._alphabet is None:
                abc = getattr(Alphabet, self._ABC[r.alphabet_type])()
                self._alphabet, self._bg, self._atype = abc, Background(abc), r.alphabet_type
                self._bgf = (ctypes.c_float * 20)(*([float(v) for v in self._bg.residue_frequencies] + [0.0] * (20 - abc.K)))
                self._degen = ptr(abc.degen)
            elif r.alphabet_type != self._atype:
                raise AlphabetMismatch(self._alphabet, getattr(Alphabet, self._ABC[r.alphabet_type])())
            abc = self._alphabet
            M, Kp = d.M, d.Kp
            om = OptimizedProfile(M, abc)
            om._batch = batch                              # keeps the table block alive
            om.msv_cost = batch.view(d.msv_cost, np.uint8, Kp, M)
            om.vit_rsc = batch.view(d.vit_rsc, np.int16, Kp, M)
            om.vit_tsc = batch.view(d.vit_tsc, np.int16, 8, M)
            om.fwd_rsc = batch.view(d.fwd_rsc, np.float32, Kp, M)
            om.fwd_tsc = batch.view(d.fwd_tsc, np.float32, 8, M)
            d.bgf = self._bgf
            d.degen = self._degen
            om._desc, om._dev = d, {}
            om.name, om.accession, om.description = s(r.name), s(r.acc), s(r.descr)
            om.reference, om.model_mask, om.consensus_structure, om.consensus = s(r.rf), s(r.mm), s(r.cs), s(r.consensus)
            om._evparam = np.array(d.evparam[:], np.float32)
            om._cutoff = np.array(d.cutoff[:], np.float32)
            om._compo = np.array(d.compo[:], np.float32)
            om.L, om.multihit = int(d.L), bool(d.mode_multihit)
            out.append(om)
        out.reverse()
        self._queue = out

    def read(self):
        if self._h is None:
            raise ValueError("I/O operation on closed file")
        if not self._queue:
            self._fill()
        return self._queue.pop() if self._queue else None


def long_target_windows(om, chunks, F1=0.02):
    """First stage of the long-target (nhmmer) pipeline on the GPU: ``p7_SSVFilter_longtarget`` over every sequence of
    ``chunks`` (the pieces a long target was cut into), then ``p7_pli_ExtendAndMergeWindows`` (p7_pipeline.c:1535-1565).
    Returns ``(raw, merged)`` numpy record arrays with fields seq, k, n, length, score (`_lib.WindowRec`)."""
    ctx = _lib.context()
    block = chunks if hasattr(chunks, "_packed") else DigitalSequenceBlock(om.alphabet, chunks)
    if block.alphabet != om.alphabet:
        raise AlphabetMismatch(om.alphabet, block.alphabet)
    db = SequenceDatabase.of(ctx, block)
    raw, mer = ctypes.c_void_p(), ctypes.c_void_p()
    nr, nm = ctypes.c_size_t(), ctypes.c_size_t()
    check(lib.b2h_longtarget_windows(ctx.handle, om._device(ctx), db.handle, float(F1), ctypes.byref(raw), ctypes.byref(nr),
                                     ctypes.byref(mer), ctypes.byref(nm)), "b2h_longtarget_windows", ctx.handle)
    try:
        dt = np.dtype(_lib.WindowRec)
        out = tuple(np.frombuffer(ctypes.string_at(ptr_, n_.value * dt.itemsize), dtype=dt).copy() if n_.value else np.zeros(0, dt)
                    for ptr_, n_ in ((raw, nr), (mer, nm)))
    finally:
        lib.b2h_free(raw)
        lib.b2h_free(mer)
    return out


class OptimizedProfileBlock(list):
    """An ordered block of `OptimizedProfile` sharing one alphabet (``pyhmmer.plan7.OptimizedProfileBlock``).

    What a scan needs of the whole block -- the array of device handles, the node total -- is cached per block and dropped
    by every mutating method, so that a scan of one query does not walk 20 000 Python objects."""

    def __init__(self, alphabet, iterable=()):
        super().__init__()
        self.alphabet = alphabet
        self._cache = {}
        for om in iterable:
            self.append(om)

    def _check(self, om):
        if not isinstance(om, OptimizedProfile):
            raise TypeError("expected OptimizedProfile, found %s" % type(om).__name__)
        if om.alphabet != self.alphabet:
            raise AlphabetMismatch(self.alphabet, om.alphabet)

    def append(self, om):
        self._check(om)
        self._cache = {}
        super().append(om)

    def extend(self, oms):
        for om in oms:
            self.append(om)

    def insert(self, i, om):
        self._check(om)
        self._cache = {}
        super().insert(i, om)

    def __setitem__(self, i, v):
        for om in (v if isinstance(i, slice) else (v,)):
            self._check(om)
        self._cache = {}
        super().__setitem__(i, v)

    def __delitem__(self, i):
        self._cache = {}
        super().__delitem__(i)

    def pop(self, i=-1):
        self._cache = {}
        return super().pop(i)

    def remove(self, om):
        self._cache = {}
        super().remove(om)

    def clear(self):
        self._cache = {}
        super().clear()

    def sort(self, *, key=None, reverse=False):
        self._cache = {}
        super().sort(key=key, reverse=reverse)

    def reverse(self):
        self._cache = {}
        super().reverse()

    def __iadd__(self, oms):
        self.extend(oms)
        return self

    def __getitem__(self, i):
        if isinstance(i, slice):
            return OptimizedProfileBlock(self.alphabet, list.__getitem__(self, i))
        return list.__getitem__(self, i)

    def copy(self):
        return OptimizedProfileBlock(self.alphabet, self)

    @property
    def total_nodes(self):
        n = self._cache.get("nodes")
        if n is None:
            n = self._cache["nodes"] = sum(om.M for om in self)
        return n

    def _handles(self, ctx):
        """ctypes array of the profiles' device handles (uploading what is not resident yet)."""
        h = self._cache.get(("handles", ctx))
        if h is None:
            h = self._cache[("handles", ctx)] = (ctypes.c_void_p * len(self))(*OptimizedProfile._device_many(ctx, self))
        return h


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


# =====================================================================================================
# Results: TopHits / Hit / Domain / Alignment   (reference: plan7.pyx 8312-9278, 1850-2234, 1441-1687, 229-425)
# =====================================================================================================
class Alignment:
    """Alignment of one domain to the model (``P7_ALIDISPLAY``).  The text lines are decoded on first access."""

    def __init__(self, domain, rec, text):
        self.domain = domain
        self._rec, self._text, self._lines = rec, text, None
        self.hmm_from, self.hmm_to = rec.hmmfrom, rec.hmmto
        self.target_from, self.target_to = rec.sqfrom, rec.sqto

    def _line(self, i):
        f = self._lines
        if f is None:
            rec, text = self._rec, self._text
            n, off = rec.N, rec.text_offset
            f = self._lines = [text[off + j * (n + 1): off + j * (n + 1) + n].decode("ascii")
                               for j in range(4 + rec.has_rf + rec.has_cs)]
        return f[i]

    hmm_sequence = property(lambda self: self._line(0))
    identity_sequence = property(lambda self: self._line(1))
    target_sequence = property(lambda self: self._line(2))
    posterior_probabilities = property(lambda self: self._line(3))
    hmm_reference = property(lambda self: self._line(4) if self._rec.has_rf else None)
    hmm_consensus_structure = property(lambda self: self._line(4 + self._rec.has_rf) if self._rec.has_cs else None)

    def __str__(self):
        """The alignment block as HMMER prints it (``Alignment.__str__`` = p7_nontranslated_alidisplay_Print with no line
        width, p7_alidisplay.c:715-790): optional CS / RF lines, model, match, target and posterior-probability lines."""
        txt = lambda v: "" if v is None else (v.decode() if isinstance(v, bytes) else str(v))
        r = self._rec
        hits = self.domain.hit.hits
        scan = hits.mode == "scan"
        hmmname = txt(self.domain.hit.name if scan else getattr(hits.query, "name", ""))
        sqname = txt(getattr(hits.query, "name", "") if scan else self.domain.hit.name)
        namew = max(len(hmmname), len(sqname))
        coordw = max(len(str(v)) for v in (r.hmmfrom, r.hmmto, r.sqfrom, r.sqto))
        model, mline, aseq, pp = self._line(0), self._line(1), self._line(2), self._line(3)
        nk = sum(1 for c in model if c != ".")
        ni = sum(1 for c in aseq if c != "-")
        k2 = r.hmmfrom + nk - 1
        i2 = r.sqfrom + ni - 1 if r.sqfrom < r.sqto else r.sqfrom - ni + 1
        out = []
        if r.has_cs:
            out.append("  %*s %s CS\n" % (namew + coordw + 1, "", self.hmm_consensus_structure))
        if r.has_rf:
            out.append("  %*s %s RF\n" % (namew + coordw + 1, "", self.hmm_reference))
        out.append("  %*s %*d %s %-*d\n" % (namew, hmmname, coordw, r.hmmfrom, model, coordw, k2))
        out.append("  %*s %s\n" % (namew + coordw + 1, " ", mline))
        if ni > 0:
            out.append("  %*s %*d %s %-*d\n" % (namew, sqname, coordw, r.sqfrom, aseq, coordw, i2))
        else:
            out.append("  %*s %*s %s %*s\n" % (namew, sqname, coordw, "-", aseq, coordw, "-"))
        out.append("  %*s %s PP\n" % (namew + coordw + 1, "", pp))
        return "".join(out)

    hmm_name = property(lambda self: self.domain.hit.hits.query.name)
    hmm_accession = property(lambda self: self.domain.hit.hits.query.accession)
    hmm_length = property(lambda self: self.domain.hit.hits.query.M)
    target_name = property(lambda self: self.domain.hit.name)
    target_length = property(lambda self: self.domain.hit.length)

    def __len__(self):
        return self._rec.N


class Domain:
    """One domain of a hit (``P7_DOMAIN``)."""

    def __init__(self, hit, rec, text):
        self.hit = hit
        self._rec = rec
        self.env_from, self.env_to = rec.ienv, rec.jenv
        self.score = float(rec.bitscore)
        self.bias = float(rec.dombias) * _LN2_INV_F        # dcl.dombias is kept in nats; reported in bits (plan7.pyx:1502)
        self.correction = float(rec.domcorrection) * _LN2_INV_F
        self.envelope_score = float(rec.envsc) * _LN2_INV_F
        self.expected_accuracy = float(rec.oasc)
        self.lnP = float(rec.lnP)
        self.reported = False
        self.included = False
        self._text, self._alignment = text, None

    @property
    def alignment(self):
        a = self._alignment
        if a is None:
            a = self._alignment = Alignment(self, self._rec, self._text)
        return a

    @alignment.setter
    def alignment(self, value):
        self._alignment = value

    @property
    def pvalue(self):
        return math.exp(self.lnP)

    @property
    def c_evalue(self):
        """Conditional E-value (plan7.pyx:1557-1565): for long-target hits the search space is already inside lnP."""
        hits = self.hit.hits
        return math.exp(self.lnP) if hits.long_targets else math.exp(self.lnP) * hits.domZ

    @property
    def i_evalue(self):
        """Independent E-value (plan7.pyx:1566-1574)."""
        hits = self.hit.hits
        return math.exp(self.lnP) if hits.long_targets else math.exp(self.lnP) * hits.Z


_LN2_INV_F = 1.0 / 0.69314718055994529


class Domains:
    def __init__(self, hit):
        self.hit = hit

    def __len__(self):
        return len(self.hit._domains)

    def __getitem__(self, i):
        return self.hit._domains[i]

    def __iter__(self):
        return iter(self.hit._domains)

    @property
    def reported(self):
        return [d for d in self.hit._domains if d.reported]

    @property
    def included(self):
        return [d for d in self.hit._domains if d.included]


class Hit:
    """One target (search) or model (scan) that p7_Pipeline scored to completion (``P7_HIT``)."""

    def __init__(self, hits, rec, target, doms, text):
        self.hits = hits
        self._rec = rec
        self.name = target.name
        self.accession = getattr(target, "accession", None) or None
        self.description = getattr(target, "description", None) or None
        self.length = len(target) if not hasattr(target, "M") else target.M
        self.score = float(rec.score)
        self.pre_score = float(rec.pre_score)
        self.sum_score = float(rec.sum_score)
        self.bias = self.pre_score - self.score
        self.lnP = float(rec.lnP)
        self.sortkey = -self.lnP if hits._params["inc_by_E"] else self.score
        self.reported = False
        self.included = False
        self.dropped = False
        self.new = False
        self.duplicate = False
        self._index = rec.seq if hits.mode == "search" else rec.profile
        off = rec.dom_offset
        self._domains = [Domain(self, doms[off + d], text) for d in range(rec.ndom)]
        self.best_domain = self._domains[rec.best_domain]

    @property
    def domains(self):
        return Domains(self)

    @domains.setter
    def domains(self, value):
        pass

    @property
    def pvalue(self):
        return math.exp(self.lnP)

    @property
    def evalue(self):
        return math.exp(self.lnP) * (1.0 if self.hits.long_targets else self.hits.Z)

    def __repr__(self):
        return "<Hit name=%r score=%.1f evalue=%.2g>" % (self.name, self.score, self.evalue)


class Trace:
    """A state path (``pyhmmer.plan7.Trace``): parallel lists of state letters, node indices and sequence positions.  Only what
    `TopHits.to_msa` reads -- jackhmmer aligns its query with `Trace.from_sequence` (plan7.pyx:9288: B, M1..ML, E)."""

    def __init__(self, states=(), k=(), i=(), M=0, L=0):
        self.states, self.k, self.i, self.M, self.L = list(states), list(k), list(i), int(M), int(L)

    @classmethod
    def from_sequence(cls, sequence):
        n = len(sequence)
        return cls(["B"] + ["M"] * n + ["E"], [0] + list(range(1, n + 1)) + [0], [0] + list(range(1, n + 1)) + [0], n, n)

    def __len__(self):
        return len(self.states)


IterationResult = collections.namedtuple("IterationResult", ["hmm", "hits", "msa", "converged", "iteration"])


class IterativeSearch:
    """``pyhmmer.plan7.IterativeSearch`` (plan7.pyx:4273-4389): jackhmmer's loop.  Every step builds a model (from the query
    on the first step, from the previous step's alignment afterwards), searches the targets on the GPU, ranks the included
    hits against the previous step's (p7_tophits_CompareRanking) and aligns them -- query first -- into the next
    alignment; converged when no new sequence was included and the alignment did not grow."""

    def __init__(self, pipeline, builder, query, targets, select_hits=None):
        self.pipeline, self.builder, self.query, self.targets, self.select_hits = pipeline, builder, query, targets, select_hits
        self.background = pipeline.background
        self.converged = False
        self.ranking = {}
        self.msa = None
        self.iteration = 0

    def __iter__(self):
        return self

    def __next__(self):
        if self.converged:
            raise StopIteration
        is_hmm = isinstance(self.query, HMM)
        if self.iteration == 0:
            hmm = self.query if is_hmm else self.builder.build(self.query, self.background)[0]
            n_prev = 1
        else:
            hmm = self.builder.build_msa(self.msa, self.background)[0]
            n_prev = len(self.msa.names)
        extra_sequences = None if is_hmm else [self.query]
        extra_traces = None if is_hmm else [Trace.from_sequence(self.query)]
        hits = self._search_hmm(hmm)
        hits.sort(by="key")
        if self.select_hits is not None:
            self.select_hits(hits)
        n_new = hits.compare_ranking(self.ranking)
        self.msa = hits.to_msa(self.pipeline.alphabet, sequences=extra_sequences, traces=extra_traces, all_consensus_cols=True, digitize=True)
        txt = lambda v: v.decode() if isinstance(v, (bytes, bytearray)) else v
        self.msa.name = "%s-i%d" % (txt(self.query.name), self.iteration + 1)
        self.msa.description = txt(self.query.description) or None
        self.msa.accession = txt(self.query.accession) or None
        self.msa.author = "jackhmmer (pyHMMER)"
        if n_new == 0 and len(self.msa.names) <= n_prev:
            self.converged = True
        self.pipeline.clear()
        self.iteration += 1
        return IterationResult(hmm, hits, self.msa, self.converged, self.iteration)

    def _search_hmm(self, hmm):
        return self.pipeline.search_hmm(hmm, self.targets)


class TopHits:
    """A sorted, thresholded list of hits (``P7_TOPHITS`` + the P7_PIPELINE snapshot pyhmmer keeps)."""

    def __init__(self, query=None, mode="search"):
        self.query = query
        self.mode = mode
        self._hits = []
        self._params = dict(E=10.0, T=None, domE=10.0, domT=None, incE=0.01, incT=None, incdomE=0.01, incdomT=None,
                            inc_by_E=True, by_E=True, dom_by_E=True, incdom_by_E=True, bit_cutoffs=None, Z=None, domZ=None)
        self.Z = 0.0
        self.domZ = 0.0
        self.long_targets = False          # nhmmer hits: the search space is already inside lnP (p7_pipeline.c:411)
        self.searched_models = 0
        self.searched_nodes = 0
        self.searched_sequences = 0
        self.searched_residues = 0
        self.n_past_msv = self.n_past_bias = self.n_past_vit = self.n_past_fwd = 0

    # -- list protocol --
    def __len__(self):
        return len(self._hits)

    def __getitem__(self, i):
        return self._hits[i]

    def __iter__(self):
        return iter(self._hits)

    def __bool__(self):
        return bool(self._hits)

    @property
    def reported(self):
        return [h for h in self._hits if h.reported]

    @property
    def included(self):
        return [h for h in self._hits if h.included]

    @property
    def hits_reported(self):
        return sum(1 for h in self._hits if h.reported)

    @property
    def hits_included(self):
        return sum(1 for h in self._hits if h.included)

    # -- p7_pli_*Reportable / Includable (p7_pipeline.c:407-464) --
    def _target_reportable(self, score, lnP, Z):
        p = self._params
        if self.long_targets:
            Z = 1.0
        return (math.exp(lnP) * Z <= p["E"]) if p["by_E"] else (score >= p["T"])

    def _target_includable(self, score, lnP, Z):
        p = self._params
        if self.long_targets:
            Z = 1.0
        return (math.exp(lnP) * Z <= p["incE"]) if p["inc_by_E"] else (score >= p["incT"])

    def _domain_reportable(self, score, lnP):
        p = self._params
        return (math.exp(lnP) * self.domZ <= p["domE"]) if p["dom_by_E"] else (score >= p["domT"])

    def _domain_includable(self, score, lnP):
        p = self._params
        return (math.exp(lnP) * self.domZ <= p["incdomE"]) if p["incdom_by_E"] else (score >= p["incdomT"])

    def _sort_by_key(self):
        """p7_tophits_SortBySortkey (p7_tophits.c:393; comparator hit_sorter_by_sortkey)."""
        if self.long_targets:              # ties: name, then the positive strand first, then position (p7_tophits.c:304-325)
            self._hits.sort(key=lambda h: (-h.sortkey, h.name, 0 if h._domains[0]._rec.iali < h._domains[0]._rec.jali else 1,
                                           h._domains[0]._rec.iali))
            return
        self._hits.sort(key=lambda h: (-h.sortkey, h.name, h._domains[0]._rec.sqfrom))

    def _threshold(self):
        """p7_tophits_Threshold (p7_tophits.c:1057)."""
        if not self._params["bit_cutoffs"]:
            for h in self._hits:
                h.reported = h.included = False
                if not h.duplicate and self._target_reportable(h.score, h.lnP, self.Z):
                    h.reported = True
                    h.included = self._target_includable(h.score, h.lnP, self.Z)
        if self._params["domZ"] is None:
            self.domZ = float(sum(1 for h in self._hits if h.reported))
        if self.long_targets:              # "no domains in dna search": the one domain of a hit follows the hit (p7_tophits.c:1078)
            if not self._params["bit_cutoffs"]:
                for h in self._hits:
                    for d in h._domains:
                        d.reported, d.included = h.reported, h.included
            return
        if not self._params["bit_cutoffs"]:
            for h in self._hits:
                for d in h._domains:
                    d.reported = d.included = False
                if h.reported:
                    for d in h._domains:
                        d.reported = self._domain_reportable(d.score, d.lnP)
                        d.included = h.included and self._domain_includable(d.score, d.lnP)
        # workaround_bug_h74 (p7_tophits.c:756-778, always applied at the end of p7_tophits_Threshold): when envelopes of
        # a target overlapped, two domains may carry the same alignment; only the better-scoring one stays reported
        for h in self._hits:
            if h._rec.noverlaps:
                ds = h._domains
                for d1 in range(len(ds)):
                    for d2 in range(d1 + 1, len(ds)):
                        a, b = ds[d1]._rec, ds[d2]._rec
                        if a.iali == b.iali and a.jali == b.jali:
                            gone = ds[d2] if a.bitscore >= b.bitscore else ds[d1]
                            gone.reported = gone.included = False

    def sort(self, by="key"):
        """``TopHits.sort`` (plan7.pyx:8932): ``key`` = by sort key (p7_tophits_SortBySortkey), ``seqidx`` = by target index and
        alignment position (p7_tophits_SortBySeqidxAndAlipos, p7_tophits.c:335-360)."""
        if by == "key":
            self._sort_by_key()
        elif by == "seqidx":
            self._hits.sort(key=self._seqidx_key)
        else:
            raise ValueError("invalid value for `by`: %r (expected 'key' or 'seqidx')" % (by,))

    @staticmethod
    def _seqidx_key(h):
        d = h._domains[0]._rec
        s, e = (d.iali, d.jali) if d.iali < d.jali else (d.jali, d.iali)
        return (h._index, 0 if d.iali < d.jali else 1, s, -e)

    def is_sorted(self, by="key"):
        """``TopHits.is_sorted`` (plan7.pyx:8951)."""
        if by not in ("key", "seqidx"):
            raise ValueError("invalid value for `by`: %r (expected 'key' or 'seqidx')" % (by,))
        before = list(self._hits)
        probe = TopHits(self.query, self.mode)
        probe.long_targets, probe._hits = self.long_targets, list(self._hits)
        probe.sort(by)
        return all(a is b for a, b in zip(before, probe._hits))

    def copy(self):
        """``TopHits.copy`` (plan7.pyx:8897): an independent list over the same (immutable) hit records."""
        import copy as _copy
        new = TopHits(self.query, self.mode)
        new.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ("_hits", "_params")})
        new._params = dict(self._params)
        for h in self._hits:
            c = _copy.copy(h)
            c.hits = new
            c._domains = []
            for d in h._domains:
                dc = _copy.copy(d)
                dc.hit = c
                dc._alignment = None                         # rebuilt on demand for the copy
                c._domains.append(dc)
            c.best_domain = c._domains[h._domains.index(h.best_domain)]
            new._hits.append(c)
        return new

    # thresholds and search-space figures, named as on the reference object (plan7.pyx:8560-8760)
    E = property(lambda self: self._params["E"])
    T = property(lambda self: None if self._params["by_E"] else self._params["T"])
    domE = property(lambda self: self._params["domE"])
    domT = property(lambda self: None if self._params["dom_by_E"] else self._params["domT"])
    incE = property(lambda self: self._params["incE"])
    incT = property(lambda self: None if self._params["inc_by_E"] else self._params["incT"])
    incdomE = property(lambda self: self._params["incdomE"])
    incdomT = property(lambda self: None if self._params["incdom_by_E"] else self._params["incdomT"])
    bit_cutoffs = property(lambda self: self._params["bit_cutoffs"])

    def __getstate__(self):
        """Pickling: the hit / domain records as bytes plus what is needed to rebuild the objects (names, lengths, flags)."""
        hits = self._hits
        recs = b"".join(bytes(h._rec) for h in hits)
        doms, text, meta = [], [], []
        for h in hits:
            off = []
            for d in h._domains:
                r = _lib.DomainRec.from_buffer_copy(d._rec)
                a = d.alignment
                lines = a._text[d._rec.text_offset:d._rec.text_offset + (4 + int(d._rec.has_rf) + int(d._rec.has_cs)) * (d._rec.N + 1)]
                r.text_offset = sum(len(t) for t in text)
                text.append(lines)
                off.append((bytes(r), d.reported, d.included))
            doms.append(off)
            meta.append((h.name, h.accession, h.description, h.length, h.reported, h.included, h.duplicate, h.dropped, h.new, h.sortkey,
                         h._index, h._domains.index(h.best_domain)))
        state = {k: v for k, v in self.__dict__.items() if k != "_hits"}
        state["_pickled"] = (recs, doms, b"".join(text), meta)
        return state

    def __setstate__(self, state):
        recs, doms, text, meta = state.pop("_pickled")
        self.__dict__.update(state)
        self._hits = []
        hs = ctypes.sizeof(_lib.HitRec)

        class _T:                                          # what Hit() reads from a target
            def __init__(self, name, acc, desc, length):
                self.name, self.accession, self.description, self._n = name, acc, desc, length

            def __len__(self):
                return self._n

        for i, (m, dl) in enumerate(zip(meta, doms)):
            rec = _lib.HitRec.from_buffer_copy(recs[i * hs:(i + 1) * hs])
            drecs = [_lib.DomainRec.from_buffer_copy(b) for b, _, _ in dl]
            rec.dom_offset, rec.ndom = 0, len(drecs)
            rec.best_domain = m[11]
            h = Hit(self, rec, _T(m[0], m[1], m[2], m[3]), drecs, text)
            h.reported, h.included, h.duplicate, h.dropped, h.new, h.sortkey, h._index = m[4], m[5], m[6], m[7], m[8], m[9], m[10]
            for d, (_, rep, inc) in zip(h._domains, dl):
                d.reported, d.included = rep, inc
            self._hits.append(h)

    def compare_ranking(self, ranking):
        """``TopHits.compare_ranking`` = p7_tophits_CompareRanking (p7_tophits.c:925): flag the included hits that were not
        in ``ranking`` (a name -> rank mapping from the previous iteration) as new, the known ones that are no longer
        included as dropped; ``ranking`` becomes the list of this round's included hits.  Returns the number of new hits."""
        nnew = 0
        for h in self._hits:
            old = ranking.get(h.name, -1)
            if h.included:
                if old == -1:
                    h.new = True
                    nnew += 1
            elif old >= 0:
                h.dropped = True
        ranking.clear()
        for h in self._hits:
            if h.included and h.name not in ranking:
                ranking[h.name] = len(ranking)
        return nnew

    def to_msa(self, alphabet, sequences=None, traces=None, trim=False, digitize=False, all_consensus_cols=False):
        """A multiple alignment of all included domains (``TopHits.to_msa``, plan7.pyx:8960-9080 = p7_tophits_Alignment,
        p7_tophits.c:1251): every included domain becomes a row named ``target/from-to``, rebuilt from its alignment display
        as the reference does (p7_alidisplay_Backconvert), and the rows are laid out by p7_tracealign_Seqs (tracealign.c:
        map_new_msa, make_text_msa, annotate_rf, annotate_posterior_probability, rejustify_insertions_text) -- match columns
        in upper case, insertions in lower case split half left / half right, ``x`` in the RF line on consensus columns.
        ``sequences`` / ``traces``: extra rows placed first (jackhmmer's query with `Trace.from_sequence`), without posterior
        probabilities.  Returns an `easel.TextMSA`, or an `easel.DigitalMSA` with ``digitize`` (p7_DIGITIZE)."""
        from .easel import TextMSA
        sequences, traces = list(sequences or ()), list(traces or ())
        if len(sequences) != len(traces):
            raise ValueError("`sequences` and `traces` must have the same length")
        txt = lambda v: "" if v is None else (v.decode() if isinstance(v, bytes) else str(v))
        rows = []
        M = 0
        for h in self._hits:
            if not h.included:
                continue
            for d in h._domains:
                if d.included:
                    a = d.alignment
                    rows.append((h, d, a.hmm_sequence, a.target_sequence, a.posterior_probabilities))
                    M = M or int(a.hmm_length or 0) or int(getattr(self.query, "M", 0)) or a.hmm_to
        if traces:
            if M == 0:
                M = traces[0].M
            elif M != traces[0].M:
                raise ValueError("top hits and included trace(s) have different profile lengths")
        if not rows and not traces:
            raise ValueError("No included domains found")
        gap = "-_."
        # states per display column: k advances on every non-gap model character; I = insert after node k
        paths = []
        inscount = [0] * (M + 1)
        matuse = [bool(all_consensus_cols)] * (M + 1)
        matuse[0] = False
        for sq, tr in zip(sequences, traces):             # the extra rows: states of the given trace over the given sequence
            text = sq.sequence if isinstance(sq.sequence, str) else alphabet.decode(sq.sequence)
            path, insnum = [], {}
            for st, k, i in zip(tr.states, tr.k, tr.i):
                if st == "M":
                    path.append(("M", k, text[i - 1], None))
                    matuse[k] = True
                elif st == "D":
                    path.append(("D", k, None, None))
                elif st == "I" or (st in "NCJ" and i > 0):
                    kk = 0 if st == "N" else (M if st == "C" else k)
                    path.append(("I", kk, text[i - 1], None))
                    insnum[kk] = insnum.get(kk, 0) + 1
            for kk, v in insnum.items():
                inscount[kk] = max(inscount[kk], v)
            paths.append(path)
        for h, d, model, aseq, pp in rows:
            k = d._rec.hmmfrom - 1
            path, insnum = [], {}
            for z in range(len(model)):
                if model[z] not in gap:
                    k += 1
                    if aseq[z] not in gap:
                        path.append(("M", k, aseq[z], pp[z]))
                        matuse[k] = True
                    else:
                        path.append(("D", k, None, None))
                else:
                    path.append(("I", k, aseq[z], pp[z]))
                    insnum[k] = insnum.get(k, 0) + 1
            for kk, v in insnum.items():
                inscount[kk] = max(inscount[kk], v)
            paths.append(path)
        if trim:
            inscount[0] = inscount[M] = 0
        matmap = [0] * (M + 1)
        alen = inscount[0]
        for k in range(1, M + 1):
            if matuse[k]:
                matmap[k] = alen + 1
                alen += 1 + inscount[k]
            else:
                matmap[k] = alen
                alen += inscount[k]
        # p7_alidisplay_DecodePostProb / EncodePostProb (p7_alidisplay.c:386-410): float values, double arithmetic on them
        f32 = lambda v: float(np.float32(v))
        decode = lambda c: 1.0 if c == "*" else (0.0 if c == "." else (f32(0.01) if c == "0" else f32((ord(c) - 48) / 10.0)))
        encode = lambda p: "*" if f32(p) + 0.05 >= 1.0 else chr(int((f32(p) + 0.05) * 10.0) + 48)
        totp, npp = [0.0] * alen, [0] * alen
        arows, prows = [], []
        for path in paths:
            row, ppr = ["."] * alen, ["."] * alen
            for k in range(1, M + 1):
                if matuse[k]:
                    row[matmap[k] - 1] = "-"
            apos = 0
            for st, k, c, p in path:
                if st == "M":
                    row[matmap[k] - 1] = c.upper()
                    if p is not None:
                        ppr[matmap[k] - 1] = encode(decode(p))
                        totp[matmap[k] - 1] += decode(p)
                        npp[matmap[k] - 1] += 1
                    apos = matmap[k]
                elif st == "D":
                    if matuse[k]:
                        row[matmap[k] - 1] = "-"
                    apos = matmap[k]
                elif not trim or (k != 0 and k != M):
                    row[apos] = c.lower()
                    if p is not None:
                        ppr[apos] = encode(decode(p))
                    apos += 1
            # rejustify_insertions_text: the second half of every insertion longer than one goes to the right edge
            for k in range(0, M):
                if inscount[k] > 1:
                    lo, hi = matmap[k], matmap[k + 1] - (1 if matuse[k + 1] else 0)
                    nins = sum(1 for x in row[lo:hi] if x not in gap and x != "~")
                    nins = 0 if k == 0 else nins // 2
                    opos = npos = hi - 1
                    while opos >= lo + nins:
                        if row[opos] in gap:
                            opos -= 1
                        else:
                            row[npos], ppr[npos] = row[opos], ppr[opos]
                            npos -= 1
                            opos -= 1
                    while npos >= lo + nins:
                        row[npos], ppr[npos] = ".", "."
                        npos -= 1
            arows.append("".join(row))
            prows.append("".join(ppr))
        rf = ["."] * alen
        for k in range(1, M + 1):
            if matuse[k]:
                rf[matmap[k] - 1] = "x"
        ppcons = "".join(encode(totp[i] / npp[i]) if npp[i] else "." for i in range(alen))
        nx = len(sequences)
        names = [txt(q.name).encode() for q in sequences] + \
                [("%s/%d-%d" % (txt(h.name), d._rec.sqfrom, d._rec.sqto)).encode() for h, d, _, _, _ in rows]
        descs = [txt(q.description).encode() if q.description else None for q in sequences] + \
                [("[subseq from] %s" % (txt(h.description) if h.description else txt(h.name))).encode() for h, d, _, _, _ in rows]
        accs = [txt(q.accession).encode() if q.accession else None for q in sequences] + \
               [txt(h.accession).encode() if h.accession else None for h, d, _, _, _ in rows]
        prows = [None] * nx + prows[nx:]
        msa = TextMSA(names=names, sequences=arows, accessions=accs, descriptions=descs, reference="".join(rf),
                      posterior_probabilities=prows if rows else None, consensus_posterior_probabilities=ppcons if rows else None)
        return msa.digitize(alphabet) if digitize else msa

    def write(self, fh, format="targets", header=True):
        """Write the hits in tabular form to a file opened in binary mode (``TopHits.write``, plan7.pyx:9096-9168):
        ``targets`` = hmmsearch ``--tblout`` (p7_tophits_TabularTargets, p7_tophits.c:1402), ``domains`` = ``--domtblout``
        (p7_tophits_TabularDomains, :1500), ``pfam`` = ``--pfamtblout`` (p7_tophits_TabularXfam, :1590)."""
        if format not in ("targets", "domains", "pfam"):
            raise ValueError("invalid format %r (expected 'targets', 'domains' or 'pfam')" % (format,))
        txt = lambda v: "" if v is None else (v.decode() if isinstance(v, bytes) else str(v))
        q = self.query
        qname = txt(getattr(q, "name", None))
        qacc = getattr(q, "accession", None)
        qacc = None if qacc is None else txt(qacc)
        hits = self._hits
        names = [txt(h.name) for h in hits]
        accs = [None if h.accession is None else txt(h.accession) for h in hits]
        descs = [None if h.description is None else txt(h.description) for h in hits]
        tnamew = max([20] + [len(v) for v in names])
        taccw = max([10] + [len(v) for v in accs if v is not None])
        qnamew = max(20, len(qname))
        qaccw = max(10, len(qacc)) if qacc is not None else 10
        qacc_s = qacc if qacc else "-"
        scan = self.mode == "scan"
        lt = self.long_targets
        posw = 0
        if lt:
            posw = max([7] + [len(str(v)) for h in hits if h._domains[0]._rec.iali > 0 for v in (h._domains[0]._rec.iali, h._domains[0]._rec.jali)])
        LOG2R = 1.44269504088896341
        out = []
        w = out.append
        rep = [(h, n, a, d) for h, n, a, d in zip(hits, names, accs, descs) if h.reported]
        if format == "targets":
            if header:
                if lt:
                    w("#%-*s %-*s %-*s %-*s %s %s %*s %*s %*s %*s %*s %6s %9s %6s %5s  %s\n" % (
                        tnamew - 1, " target name", taccw, "accession", qnamew, "query name", qaccw, "accession", "hmmfrom", "hmm to",
                        posw, "alifrom", posw, "ali to", posw, "envfrom", posw, "env to", posw, "modlen" if scan else "sq len",
                        "strand", "  E-value", " score", " bias", "description of target"))
                    w("#%*s %*s %*s %*s %s %s %*s %*s %*s %*s %*s %6s %9s %6s %5s %s\n" % (
                        tnamew - 1, "-------------------", taccw, "----------", qnamew, "--------------------", qaccw, "----------", "-------",
                        "-------", posw, "-------", posw, "-------", posw, "-------", posw, "-------", posw, "-------", "------", "---------",
                        "------", "-----", "---------------------"))
                else:
                    w("#%*s %22s %22s %33s\n" % (tnamew + qnamew + taccw + qaccw + 2, "", "--- full sequence ----", "--- best 1 domain ----",
                                                 "--- domain number estimation ----"))
                    w("#%-*s %-*s %-*s %-*s %9s %6s %5s %9s %6s %5s %5s %3s %3s %3s %3s %3s %3s %3s %s\n" % (
                        tnamew - 1, " target name", taccw, "accession", qnamew, "query name", qaccw, "accession", "  E-value", " score", " bias",
                        "  E-value", " score", " bias", "exp", "reg", "clu", " ov", "env", "dom", "rep", "inc", "description of target"))
                    w("#%*s %*s %*s %*s %9s %6s %5s %9s %6s %5s %5s %3s %3s %3s %3s %3s %3s %3s %s\n" % (
                        tnamew - 1, "-------------------", taccw, "----------", qnamew, "--------------------", qaccw, "----------", "---------",
                        "------", "-----", "---------", "------", "-----", "---", "---", "---", "---", "---", "---", "---", "---",
                        "---------------------"))
            for h, name, acc, desc in rep:
                d = h.best_domain
                r, dr = h._rec, d._rec
                if lt:
                    w("%-*s %-*s %-*s %-*s %7d %7d %*d %*d %*d %*d %*d %6s %9.2g %6.1f %5.1f  %s\n" % (
                        tnamew, name, taccw, acc if acc else "-", qnamew, qname, qaccw, qacc_s, dr.hmmfrom, dr.hmmto,
                        posw, dr.iali, posw, dr.jali, posw, dr.ienv, posw, dr.jenv, posw, h.length,
                        "   +  " if dr.iali < dr.jali else "   -  ", math.exp(h.lnP), h.score, dr.dombias * LOG2R, desc if desc else "-"))
                else:
                    w("%-*s %-*s %-*s %-*s %9.2g %6.1f %5.1f %9.2g %6.1f %5.1f %5.1f %3d %3d %3d %3d %3d %3d %3d %s\n" % (
                        tnamew, name, taccw, acc if acc else "-", qnamew, qname, qaccw, qacc_s,
                        math.exp(h.lnP) * self.Z, h.score, h.pre_score - h.score, math.exp(d.lnP) * self.Z, dr.bitscore, dr.dombias * LOG2R,
                        r.nexpected, r.nregions, r.nclustered, r.noverlaps, r.nenvelopes, r.ndom,
                        sum(1 for x in h._domains if x.reported), sum(1 for x in h._domains if x.included), desc if desc else "-"))
        elif format == "domains":
            if header:
                w("#%*s %22s %40s %11s %11s %11s\n" % (tnamew + qnamew - 1 + 15 + taccw + qaccw, "", "--- full sequence ---",
                                                       "-------------- this domain -------------", "hmm coord", "ali coord", "env coord"))
                w("#%-*s %-*s %5s %-*s %-*s %5s %9s %6s %5s %3s %3s %9s %9s %6s %5s %5s %5s %5s %5s %5s %5s %4s %s\n" % (
                    tnamew - 1, " target name", taccw, "accession", "tlen", qnamew, "query name", qaccw, "accession", "qlen", "E-value", "score",
                    "bias", "#", "of", "c-Evalue", "i-Evalue", "score", "bias", "from", "to", "from", "to", "from", "to", "acc",
                    "description of target"))
                w("#%*s %*s %5s %*s %*s %5s %9s %6s %5s %3s %3s %9s %9s %6s %5s %5s %5s %5s %5s %5s %5s %4s %s\n" % (
                    tnamew - 1, "-------------------", taccw, "----------", "-----", qnamew, "--------------------", qaccw, "----------", "-----",
                    "---------", "------", "-----", "---", "---", "---------", "---------", "------", "-----", "-----", "-----", "-----", "-----",
                    "-----", "-----", "----", "---------------------"))
            qM = int(getattr(q, "M", 0)) if not scan else 0
            for h, name, acc, desc in rep:
                nrep = sum(1 for x in h._domains if x.reported)
                nd = 0
                for d in h._domains:
                    if not d.reported:
                        continue
                    nd += 1
                    dr = d._rec
                    # the display's M / L are model / sequence lengths; which one is the target depends on the mode
                    tlen, qlen = (h.length, qM) if not scan else (h.length, len(q))
                    acc_ = dr.oasc / (1.0 + abs(float(dr.jenv - dr.ienv)))
                    w("%-*s %-*s %5d %-*s %-*s %5d %9.2g %6.1f %5.1f %3d %3d %9.2g %9.2g %6.1f %5.1f %5d %5d %5d %5d %5d %5d %4.2f %s\n" % (
                        tnamew, name, taccw, acc if acc else "-", tlen, qnamew, qname, qaccw, qacc_s, qlen,
                        math.exp(h.lnP) * self.Z, h.score, h.pre_score - h.score, nd, nrep, math.exp(d.lnP) * self.domZ, math.exp(d.lnP) * self.Z,
                        dr.bitscore, dr.dombias * LOG2R, dr.hmmfrom, dr.hmmto, dr.sqfrom, dr.sqto, dr.ienv, dr.jenv, acc_, desc if desc else "-"))
        else:
            taccw = max([20] + [len(v) for v in accs if v is not None])
            if lt:
                w("# hit scores\n# ----------\n#\n")
                w("# %-*s %-*s %-*s %6s %9s %5s  %s  %s %6s %*s %*s %*s %*s %*s   %s\n" % (
                    tnamew - 1, "target name", taccw, "acc", qnamew, "query name", "bits", "  e-value", " bias", "hmm-st", "hmm-en", "strand",
                    posw, "ali-st", posw, "ali-en", posw, "env-st", posw, "env-en", posw, "modlen" if scan else "sq-len", "description of target"))
                w("# %-*s %-*s %-*s %6s %9s %5s %s %s %6s %*s %*s %*s %*s %*s   %s\n" % (
                    tnamew - 1, "-------------------", taccw, "-------------------", qnamew, "-------------------", "------", "---------", "-----",
                    "-------", "-------", "------", posw, "-------", posw, "-------", posw, "-------", posw, "-------", posw, "-------",
                    "---------------------"))
                for h, name, acc, desc in rep:
                    dr = h._domains[0]._rec
                    w("%-*s  %-*s %-*s %6.1f %9.2g %5.1f %7d %7d %s %*d %*d %*d %*d %*d   %s\n" % (
                        tnamew, name, taccw, (acc if acc else "-") if scan else qacc_s, qnamew, qname, h.score, math.exp(h.lnP), dr.dombias * LOG2R,
                        dr.hmmfrom, dr.hmmto, "   +  " if dr.iali < dr.jali else "   -  ", posw, dr.iali, posw, dr.jali, posw, dr.ienv,
                        posw, dr.jenv, posw, h.length, desc if desc else "-"))
            else:
                w("# Sequence scores\n# ---------------\n#\n")
                w("# %-*s %6s %9s %3s %5s %5s    %s\n" % (tnamew - 1, "name", " bits", "  E-value", "n", "exp", " bias", "description"))
                w("# %*s %6s %9s %3s %5s %5s    %s\n" % (tnamew - 1, "-------------------", "------", "---------", "---", "-----", "-----",
                                                         "---------------------"))
                for h, name, acc, desc in rep:
                    w("%-*s  %6.1f %9.2g %3d %5.1f %5.1f    %s\n" % (tnamew, name, h.score, math.exp(h.lnP) * self.Z, h._rec.ndom, h._rec.nexpected,
                                                                    h.pre_score - h.score, desc if desc else "-"))
                w("\n")
                # one pseudo-hit per reported domain, sorted like hits (hit_sorter_by_sortkey: key, name, start)
                pseudo = []
                for h, name, acc, desc in rep:
                    k = 0
                    for d in h._domains:
                        if d.reported:
                            k += 1
                            key = -d.lnP if self._params["inc_by_E"] else d._rec.bitscore
                            pseudo.append((-key, name, d._rec.iali, k, d, desc))
                pseudo.sort(key=lambda t: t[:3])
                w("# Domain scores\n# -------------\n#\n")
                w("# %-*s %6s %9s %5s %5s %6s %6s %6s %6s %6s %6s     %s\n" % (tnamew - 1, " name", "bits", "E-value", "hit", "bias", "env-st", "env-en",
                                                                            "ali-st", "ali-en", "hmm-st", "hmm-en", "description"))
                w("# %*s %6s %9s %5s %5s %6s %6s %6s %6s %6s %6s      %s\n" % (tnamew - 1, "-------------------", "------", "---------", "-----", "-----",
                                                                            "------", "------", "------", "------", "------", "------",
                                                                            "---------------------"))
                for _, name, _, k, d, desc in pseudo:
                    dr = d._rec
                    w("%-*s  %6.1f %9.2g %5d %5.1f %6d %6d %6d %6d %6d %6d     %s\n" % (
                        tnamew, name, dr.bitscore, math.exp(d.lnP) * self.Z, k, dr.dombias * LOG2R, dr.ienv, dr.jenv, dr.sqfrom, dr.sqto,
                        dr.hmmfrom, dr.hmmto, desc if desc else "-"))
        fh.write("".join(out).encode())

    def _check_threshold_parameters(self, other):
        """``TopHits._check_threshold_parameters`` (plan7.pyx:8832-8861)."""
        p, q = self._params, other._params
        if self.long_targets and not other.long_targets:
            raise ValueError("Trying to merge a `TopHits` from a long targets pipeline to a `TopHits` from a regular pipeline.")
        if (p["Z"] is None) != (q["Z"] is None):
            raise ValueError("Trying to merge `TopHits` with `Z` values obtained with different methods.")
        if p["Z"] is not None and self.Z != other.Z:
            raise ValueError("Trying to merge `TopHits` obtained from pipelines manually configured to different `Z` values.")
        if (p["domZ"] is None) != (q["domZ"] is None):
            raise ValueError("Trying to merge `TopHits` with `domZ` values obtained with different methods.")
        if p["domZ"] is not None and self.domZ != other.domZ:
            raise ValueError("Trying to merge `TopHits` obtained from pipelines manually configured to different `domZ` values.")
        for mode, what in (("by_E", "reporting"), ("dom_by_E", "domain reporting"), ("inc_by_E", "inclusion"), ("incdom_by_E", "domain inclusion")):
            if p[mode] != q[mode]:
                raise ValueError("Trying to merge `TopHits` obtained from pipelines with different %s threshold modes" % what)
        for mode, e, t, what in (("by_E", "E", "T", "reporting"), ("inc_by_E", "incE", "incT", "inclusion"),
                                 ("dom_by_E", "domE", "domT", "domain reporting"), ("incdom_by_E", "incdomE", "incdomT", "domain inclusion")):
            k = e if p[mode] else t
            if p[k] != q[k]:
                raise ValueError("Trying to merge `TopHits` obtained from pipelines with different %s thresholds." % what)

    def merge(self, *others):
        """``TopHits.merge`` (plan7.pyx:9172-9273): combine hits of target-sharded searches of one query.  The parts are
        copied (the inputs keep their own flags and E-values), their queries and threshold parameters must agree."""
        merged = self.copy()
        parts = [self]
        for other in others:
            mq, oq = merged.query, other.query
            ident = lambda q: (type(q).__name__ if not isinstance(q, (HMM, Profile, OptimizedProfile)) else "model",
                               getattr(q, "name", None), getattr(q, "M", None) or (len(q) if hasattr(q, "__len__") else None),
                               getattr(q, "accession", None))
            mismatch = mq is not oq and ident(mq) != ident(oq)
            if mismatch:
                raise ValueError("Trying to merge `TopHits` obtained from different queries")
            merged._check_threshold_parameters(other)
            oc = other.copy()
            for h in oc._hits:
                h.hits = merged
                merged._hits.append(h)
            parts.append(other)
        for a in ("searched_sequences", "searched_residues", "n_past_msv", "n_past_bias", "n_past_vit", "n_past_fwd"):
            setattr(merged, a, sum(getattr(part, a) for part in parts))
        merged.searched_models, merged.searched_nodes = self.searched_models, self.searched_nodes
        if self.mode == "scan":
            merged.searched_models = sum(p.searched_models for p in parts)
            merged.searched_nodes = sum(p.searched_nodes for p in parts)
            merged.searched_sequences, merged.searched_residues = self.searched_sequences, self.searched_residues
        if self._params["Z"] is None:       # Z_setby == NTARGETS: search spaces add up (p7_pipeline_Merge)
            merged.Z = float(sum(p.Z for p in parts))
        else:
            merged.Z = self.Z
        merged.domZ = self.domZ
        merged._sort_by_key()
        merged._threshold()
        return merged


# =====================================================================================================
# Pipeline   (reference: plan7.pyx 5423-6906; p7_pipeline.c)
# =====================================================================================================
class Pipeline:
    """The accelerated comparison pipeline, one query at a time -- B200 back-end of ``pyhmmer.plan7.Pipeline``.

    Keyword arguments, defaults and meaning are those of the reference (plan7.pyx:5471-5560).
    """
    M_HINT = 100
    L_HINT = 100

    def __init__(self, alphabet, background=None, *, bias_filter=True, null2=True, seed=42, Z=None, domZ=None,
                 F1=0.02, F2=1e-3, F3=1e-5, E=10.0, T=None, domE=10.0, domT=None, incE=0.01, incT=None,
                 incdomE=0.01, incdomT=None, bit_cutoffs=None, device=None, host_threads=0):
        self.alphabet = alphabet
        self.background = background if background is not None else Background(alphabet)
        self.bias_filter, self.null2, self.seed = bool(bias_filter), bool(null2), int(seed)
        self.Z, self.domZ = Z, domZ
        self.F1, self.F2, self.F3 = float(F1), float(F2), float(F3)
        self.E, self.T, self.domE, self.domT = E, T, domE, domT
        self.incE, self.incT, self.incdomE, self.incdomT = incE, incT, incdomE, incdomT
        if bit_cutoffs not in (None, "gathering", "trusted", "noise"):
            raise ValueError("invalid bit cutoffs: %r" % (bit_cutoffs,))
        self.bit_cutoffs = bit_cutoffs
        self._ctx = _lib.context(device)
        self.host_threads = host_threads
        self.clear()

    def clear(self):
        self._nseqs = self._nres = self._nmodels = self._nnodes = 0

    # -- query preparation (plan7.pyx:5979-6013) --
    def _optimized(self, query, L):
        if isinstance(query, OptimizedProfile):
            return query
        if isinstance(query, HMM):
            query = Profile(query.M, self.alphabet).configure(query, self.background, L)
        if isinstance(query, Profile):
            return query.to_optimized()
        raise TypeError("Expected HMM, Profile or OptimizedProfile, found %s" % type(query).__name__)

    def _optimized_many(self, queries, L):
        """``_optimized`` for a block of queries: the HMMs among them are configured and converted in one library call."""
        idx = [i for i, q in enumerate(queries) if isinstance(q, HMM)]
        if len(idx) < 4:
            return [self._optimized(q, L) for q in queries]
        conv = dict(zip(idx, _convert_hmms([queries[i] for i in idx], self.background, L, threads=self.host_threads)))
        return [conv[i] if i in conv else self._optimized(q, L) for i, q in enumerate(queries)]

    def _params_struct(self, seq_counters=False):
        return _lib.SearchParams(self.F1, self.F2, self.F3, int(self.bias_filter), int(self.null2), self.seed, int(self.host_threads),
                                 int(bool(seq_counters)), 0)

    def _cutoffs(self, om):
        """p7_pli_NewModelThresholds (p7_pipeline.c:535): model-specific bit thresholds."""
        if self.bit_cutoffs is None:
            return None
        i = {"gathering": 0, "trusted": 2, "noise": 4}[self.bit_cutoffs]
        if om._cutoff[i] == P7_CUTOFF_UNSET:
            raise MissingCutoffs(om.name, self.bit_cutoffs)
        return float(om._cutoff[i]), float(om._cutoff[i + 1])

    def _tophits(self, query, mode, cut):
        th = TopHits(query, mode)
        p = th._params
        p.update(E=self.E, domE=self.domE, incE=self.incE, incdomE=self.incdomE, Z=self.Z, domZ=self.domZ,
                 bit_cutoffs=self.bit_cutoffs)
        p["by_E"], p["T"] = (self.T is None), (0.0 if self.T is None else self.T)
        p["dom_by_E"], p["domT"] = (self.domT is None), (0.0 if self.domT is None else self.domT)
        p["inc_by_E"], p["incT"] = (self.incT is None), (0.0 if self.incT is None else self.incT)
        p["incdom_by_E"], p["incdomT"] = (self.incdomT is None), (0.0 if self.incdomT is None else self.incdomT)
        if cut is not None:                                   # --cut_ga/--cut_tc/--cut_nc (p7_pipeline.c:165-190, 540-560)
            p.update(by_E=False, dom_by_E=False, inc_by_E=False, incdom_by_E=False,
                     T=cut[0], incT=cut[0], domT=cut[1], incdomT=cut[1])
        return th

    def _run(self, oms, block, seq_counters=False):
        """One ``b2h_search`` of the profiles against the block.  Returns the hit / domain records, the alignment text and
        the pass counters: per profile [P][4], or per SEQUENCE [N][4] with ``seq_counters`` (scan mode)."""
        ctx = self._ctx
        t0 = time.perf_counter()
        if block._cache.get(("db", ctx)) is None and sum(1 for om in oms if om._dev.get(ctx) is None) > 8:
            # neither side is resident yet: pack + upload the database on a helper thread while this one builds and
            # uploads the profile tables (both are C calls that release the GIL and end in stream-ordered copies)
            import threading
            box = []
            th = threading.Thread(target=lambda: box.append(SequenceDatabase.of(ctx, block)))
            th.start()
            try:
                handles = (ctypes.c_void_p * len(oms))(*OptimizedProfile._device_many(ctx, oms))
            finally:
                th.join()
            db = SequenceDatabase.of(ctx, block)
        else:
            db = SequenceDatabase.of(ctx, block)
            handles = oms._handles(ctx) if isinstance(oms, OptimizedProfileBlock) else \
                (ctypes.c_void_p * len(oms))(*OptimizedProfile._device_many(ctx, oms))
        prm = self._params_struct(seq_counters)
        out = ctypes.c_void_p()
        t1 = time.perf_counter()
        st = lib.b2h_search(ctx.handle, handles, len(oms), db.handle, ctypes.byref(prm), ctypes.byref(out))
        t2 = time.perf_counter()
        if st == _lib.B2H_ERANGE:
            raise OverflowError("numerical overflow in the optimized vector implementation")
        check(st, "b2h_search", ctx.handle)
        try:
            hits, doms, text = _lib.read_results(out)
            if seq_counters:
                cp = lib.b2h_results_seq_counters(out)
                counters = np.ctypeslib.as_array(cp, shape=(len(block), 4)).copy()
            else:
                cp = lib.b2h_results_counters(out)
                counters = np.ctypeslib.as_array(cp, shape=(len(oms), 4)).copy()
        finally:
            lib.b2h_results_destroy(out)
        self._last_run_s = (t1 - t0, t2 - t1, time.perf_counter() - t2)    # (uploads / handles, b2h_search, reading the results)
        return hits, doms, text, counters

    def _run_waves(self, oms, block):
        """The same search with its results wave by wave (b2h_search_begin / _next / _end, include/b2h.h): returns
        ``(n_waves, generator)``; the generator yields ``(profile_indices, hits, doms, text, counters[P][4])``, every tuple
        final for the listed profiles.  The search starts at once and goes on with the following waves on the engine's own
        driver thread while the caller works on what it was handed (assembling `TopHits`, exchanging hit records)."""
        ctx = self._ctx
        t0 = time.perf_counter()
        if block._cache.get(("db", ctx)) is None and sum(1 for om in oms if om._dev.get(ctx) is None) > 8:
            import threading
            box = []
            th = threading.Thread(target=lambda: box.append(SequenceDatabase.of(ctx, block)))
            th.start()
            try:
                handles = (ctypes.c_void_p * len(oms))(*OptimizedProfile._device_many(ctx, oms))
            finally:
                th.join()
            db = SequenceDatabase.of(ctx, block)
        else:
            db = SequenceDatabase.of(ctx, block)
            handles = oms._handles(ctx) if isinstance(oms, OptimizedProfileBlock) else \
                (ctypes.c_void_p * len(oms))(*OptimizedProfile._device_many(ctx, oms))
        prm = self._params_struct(False)
        job, nw = ctypes.c_void_p(), ctypes.c_size_t()
        t1 = time.perf_counter()
        check(lib.b2h_search_begin(ctx.handle, handles, len(oms), db.handle, ctypes.byref(prm), ctypes.byref(job), ctypes.byref(nw)),
              "b2h_search_begin", ctx.handle)
        self._last_run_s = (t1 - t0, 0.0, 0.0)
        return int(nw.value), self._wave_results(job, len(oms), t1 - t0)

    def _wave_results(self, job, P, t_upload):
        ctx = self._ctx
        waited = read = 0.0
        try:
            while True:
                out = ctypes.c_void_p()
                tw = time.perf_counter()
                st = lib.b2h_search_next(job, ctypes.byref(out))            # (the GIL is released while this waits)
                waited += time.perf_counter() - tw
                if st == _lib.B2H_ERANGE:
                    raise OverflowError("numerical overflow in the optimized vector implementation")
                check(st, "b2h_search", ctx.handle)
                if not out.value:
                    break
                tr = time.perf_counter()
                try:
                    hits, doms, text = _lib.read_results(out)
                    counters = np.ctypeslib.as_array(lib.b2h_results_counters(out), shape=(P, 4)).copy()
                    n = ctypes.c_size_t()
                    pp = lib.b2h_results_profiles(out, ctypes.byref(n))
                    profs = [pp[i] for i in range(n.value)]
                finally:
                    lib.b2h_results_destroy(out)
                read += time.perf_counter() - tr
                self._last_run_s = (t_upload, waited, read)
                yield profs, hits, doms, text, counters
        finally:
            st = lib.b2h_search_end(job)
        if st == _lib.B2H_ERANGE:
            raise OverflowError("numerical overflow in the optimized vector implementation")
        check(st, "b2h_search", ctx.handle)

    def _admit(self, th, rec, target, doms, text, Z_running, cut):
        """Hit admission as p7_Pipeline does it at the moment the comparison finishes (p7_pipeline.c:838):
        with the *running* Z when the search space is being counted (p7_pipeline.c:580)."""
        p = th._params
        if p["by_E"]:
            ok = math.exp(rec.lnP) * Z_running <= p["E"]
        else:
            ok = rec.score >= p["T"]
        if not ok:
            return
        h = Hit(th, rec, target, doms, text)
        if cut is not None:                                   # flags set immediately under bit cutoffs (p7_pipeline.c:905-925)
            h.reported = h.score >= p["T"]
            h.included = h.reported and h.score >= p["incT"]
            for d in h._domains:
                d.reported = d.score >= p["domT"]
                d.included = d.reported and d.score >= p["incdomT"]
        th._hits.append(h)

    def search_hmm(self, query, sequences):
        """Search ``sequences`` (a `DigitalSequenceBlock`) with one query; returns `TopHits` (plan7.pyx:6121-6260)."""
        return self._search_many([query], sequences)[0]

    def search_seq(self, query, sequences, builder=None):
        """Search with a query SEQUENCE (phmmer; plan7.pyx:6330-6390): a single-sequence model is built with ``builder``
        (default: ``Builder(alphabet, seed=seed)``) and searched like any HMM; the hits remember the sequence as their query."""
        from .builder import Builder
        if query.alphabet != self.alphabet:
            raise AlphabetMismatch(self.alphabet, query.alphabet)
        builder = Builder(self.alphabet, seed=self.seed) if builder is None else builder
        hmm, profile, opt = builder.build(query, self.background)
        hits = self.search_hmm(opt, sequences)
        hits.query = query
        return hits

    def search_msa(self, query, sequences, builder=None):
        """Search with a query ALIGNMENT (plan7.pyx:6264-6328): a model is built with ``builder.build_msa`` (default:
        ``Builder(alphabet, seed=seed)``) and searched like any HMM; the hits remember the alignment as their query."""
        from .builder import Builder
        if query.alphabet != self.alphabet:
            raise AlphabetMismatch(self.alphabet, query.alphabet)
        builder = Builder(self.alphabet, seed=self.seed) if builder is None else builder
        hmm, profile, opt = builder.build_msa(query, self.background)
        hits = self.search_hmm(opt, sequences)
        hits.query = query
        return hits

    def iterate_hmm(self, query, sequences, builder=None, select_hits=None):
        """jackhmmer from an HMM query (plan7.pyx:6739-6804): an iterator of `IterationResult`."""
        from .builder import Builder
        if query.alphabet != self.alphabet:
            raise AlphabetMismatch(self.alphabet, query.alphabet)
        if sequences.alphabet != self.alphabet:
            raise AlphabetMismatch(self.alphabet, sequences.alphabet)
        if builder is None:
            builder = Builder(self.alphabet, seed=self.seed, architecture="hand")
        elif builder.architecture != "hand":
            raise ValueError("`iterate_seq` only supports a builder with 'hand' architecture")
        return IterativeSearch(self, builder, query, sequences, select_hits)

    def iterate_seq(self, query, sequences, builder=None, select_hits=None):
        """jackhmmer from a sequence query (plan7.pyx:6806-6906): an iterator of `IterationResult`."""
        return self.iterate_hmm(query, sequences, builder, select_hits)

    def _search_many(self, queries, sequences):
        if not isinstance(sequences, DigitalSequenceBlock):
            raise TypeError("expected DigitalSequenceBlock, found %s" % type(sequences).__name__)
        if sequences.alphabet != self.alphabet:
            raise AlphabetMismatch(self.alphabet, sequences.alphabet)
        for q in queries:
            if not isinstance(q, (HMM, Profile, OptimizedProfile)):
                raise TypeError("Expected HMM, Profile or OptimizedProfile, found %s" % type(q).__name__)
            if q.alphabet != self.alphabet:
                raise AlphabetMismatch(self.alphabet, q.alphabet)
        if sequences and len(sequences.largest()) > 100000:
            raise ValueError("sequence length over comparison pipeline limit (100000)")
        L = len(sequences[0]) if len(sequences) else self.L_HINT
        oms = self._optimized_many(queries, L)
        if not len(sequences):
            return self._assemble(queries, oms, sequences, [], [], b"", np.zeros((len(oms), 4), np.int64))
        if len(oms) < 16:                                      # a single wave anyway (plan_waves): one blocking call
            return self._assemble(queries, oms, sequences, *self._run(oms, sequences))
        # wave by wave: the `TopHits` of a wave's queries are built while the GPU searches the following waves
        results = [None] * len(oms)
        first = True
        for profs, hits, doms, text, counters in self._run_waves(oms, sequences)[1]:
            for qi, th in zip(profs, self._assemble(queries, oms, sequences, hits, doms, text, counters, only=profs, count_targets=first)):
                results[qi] = th
            first = False
        return results

    def _assemble(self, queries, oms, sequences, hits, doms, text, counters, only=None, count_targets=True):
        """Turn raw hit records (ordered by profile, then target) into thresholded `TopHits`, one per query -- or, with
        ``only``, one per listed query index, in that order (``count_targets``: add the targets to the pipeline's totals)."""
        which = range(len(oms)) if only is None else only
        cuts = {qi: self._cutoffs(oms[qi]) for qi in which}
        n = len(sequences)
        nres = sequences.total_residues
        results = []
        by_query = {qi: [] for qi in which}
        for rec in hits:
            by_query[rec.profile].append(rec)
        for qi in which:
            query, om = queries[qi], oms[qi]
            th = self._tophits(query, "search", cuts[qi])
            self._nmodels += 1
            self._nnodes += om.M
            for rec in by_query[qi]:                           # already in target order
                Z_running = self.Z if self.Z is not None else float(rec.seq + 1)
                self._admit(th, rec, sequences[rec.seq], doms, text, Z_running, cuts[qi])
            th.Z = float(self.Z) if self.Z is not None else float(n)
            th.domZ = float(self.domZ) if self.domZ is not None else 0.0
            th.searched_models, th.searched_nodes = 1, om.M
            th.searched_sequences, th.searched_residues = n, nres
            th.n_past_msv, th.n_past_bias, th.n_past_vit, th.n_past_fwd = (int(v) for v in counters[qi])
            th._sort_by_key()
            th._threshold()
            results.append(th)
        if count_targets:
            self._nseqs += n
            self._nres += nres
        return results

    def scan_seq(self, query, targets):
        """Scan one sequence against a list of profiles (``OptimizedProfileBlock``); returns `TopHits` (plan7.pyx:6534-6677)."""
        return self._scan_many([query], targets)[0]

    def _scan_many(self, queries, targets, world=None):
        """Query sequences against a block of profiles.  With several ranks (``world``, one process per GPU) the PROFILE block
        is sharded -- contiguous runs balanced by nodes -- every rank scans its profiles, and one all-gather of the hit
        records gives every rank the same `TopHits` (Z = all models)."""
        if any(q.alphabet != self.alphabet for q in queries):
            raise AlphabetMismatch(self.alphabet, [q.alphabet for q in queries if q.alphabet != self.alphabet][0])
        if isinstance(targets, OptimizedProfileBlock):      # the pre-fetched database of hmmscan: nothing to convert
            if targets.alphabet != self.alphabet:
                raise AlphabetMismatch(self.alphabet, targets.alphabet)
            oms = targets
        else:
            oms = OptimizedProfileBlock(self.alphabet, self._optimized_many(targets, self.L_HINT))
        cuts = [self._cutoffs(om) for om in oms] if self.bit_cutoffs is not None else None
        block = DigitalSequenceBlock(self.alphabet, queries)
        from . import parallel
        lo, local = 0, oms
        if world is not None and world.size > 1:
            key = ("shard", world.size, world.rank)
            if key not in oms._cache:                           # this rank's run of the block, kept with the block
                b = parallel.shard_bounds([om.M for om in oms], world.size)
                oms._cache[key] = (b[world.rank], oms[b[world.rank]:b[world.rank + 1]])
            lo, local = oms._cache[key]
        # the pass counters are kept per QUERY SEQUENCE: every scan_seq result reports its own (plan7.pyx:6534-6677)
        if local and len(block):
            hits, doms, text, counters = self._run(local, block, seq_counters=True)
        else:
            hits, doms, text, counters = [], [], b"", np.zeros((len(block), 4), np.int64)
        if world is not None and world.size > 1:
            mine = parallel.pack_records(hits, doms, text, counters, 0, profile_offset=lo)
            hits, doms, tbuf, cparts = [], [], bytearray(), []
            for h, d, t, c in (parallel.unpack_records(buf) for buf in parallel.all_gather_bytes(mine, world)):
                for r in h:
                    r.dom_offset += len(doms)
                for r in d:
                    r.text_offset += len(tbuf)
                hits.extend(h); doms.extend(d); tbuf.extend(t); cparts.append(c.reshape(-1, 4))
            text, counters = bytes(tbuf), np.sum(cparts, axis=0) if cparts else np.zeros((len(block), 4), np.int64)
        results = []
        for si, seq in enumerate(queries):
            th = self._tophits(seq, "scan", None)
            mine = sorted((r for r in hits if r.seq == si), key=lambda r: r.profile)
            for rec in mine:
                cut = cuts[rec.profile] if cuts is not None else None
                if cut is not None:
                    th._params.update(by_E=False, dom_by_E=False, inc_by_E=False, incdom_by_E=False,
                                      T=cut[0], incT=cut[0], domT=cut[1], incdomT=cut[1])
                Z_running = self.Z if self.Z is not None else float(rec.profile + 1)
                self._admit(th, rec, oms[rec.profile], doms, text, Z_running, cut)
            th.Z = float(self.Z) if self.Z is not None else float(len(oms))
            th.domZ = float(self.domZ) if self.domZ is not None else 0.0
            th.searched_models, th.searched_nodes = len(oms), oms.total_nodes
            th.searched_sequences, th.searched_residues = 1, len(seq)
            th.n_past_msv, th.n_past_bias, th.n_past_vit, th.n_past_fwd = (int(v) for v in counters[si])
            th._sort_by_key()
            th._threshold()
            results.append(th)
        return results


class MissingCutoffs(ValueError):
    """Same meaning as ``pyhmmer.errors.MissingCutoffs``."""

    def __init__(self, name, kind):
        super().__init__("model %r is missing %s bit cutoffs" % (name, kind))


class LongTargetsPipeline(Pipeline):
    """The pipeline for long (nucleotide) targets -- ``pyhmmer.plan7.LongTargetsPipeline`` (plan7.pyx:6917-7412), nhmmer.

    Targets of any length are cut into windows of ``block_length`` residues that keep ``max_length`` residues of context
    and are searched on both strands; hits are single domains in target coordinates (``env_from > env_to`` on the reverse
    strand), their E-values count the search space in windows of the model's ``max_length``.  The stages run in
    `pyhmmer_b200.longtarget` (windows of ALL targets and strands per kernel launch).

    The model must carry its window length (``MAXL`` in the HMM file, as hmmbuild writes it) or ``window_length`` must be
    given: ``p7_Builder_MaxLength`` is part of the builder, which is outside this package's path.
    """
    M_HINT = 100
    L_HINT = 100

    def __init__(self, alphabet, background=None, *, F1=0.02, F2=3e-3, F3=3e-5, strand=None, B1=100, B2=240, B3=1000,
                 block_length=0x40000, window_length=None, window_beta=None, **kwargs):
        if alphabet.K != 4:
            raise ValueError("Expected nucleotide alphabet, found %r" % (alphabet,))
        super().__init__(alphabet, background, F1=F1, F2=F2, F3=F3, **kwargs)
        if strand not in (None, "watson", "crick"):
            raise ValueError("invalid value for `strand`: %r" % (strand,))
        if window_length is not None and window_length < 4:
            raise ValueError("invalid window length: %r" % (window_length,))
        if window_beta is not None and not (0.0 < window_beta < 1.0):
            raise ValueError("invalid window beta: %r" % (window_beta,))
        self.strand = strand
        self.B1, self.B2, self.B3 = int(B1), int(B2), int(B3)
        self.block_length = int(block_length)
        self.window_length, self.window_beta = window_length, window_beta
        self._backend_factory = None

    def search_hmm(self, query, sequences):
        """nhmmer: one query against a `DigitalSequenceBlock` of long targets; returns `TopHits` (plan7.pyx:7258-7412)."""
        om, cut, res = self._search_records(query, sequences)
        return self._long_target_tophits(query, om, sequences, res, cut)

    def _search_records(self, query, sequences):
        """The search itself: ``(optimized profile, bit cutoffs or None, (hits, doms, text, duplicate flags, stats))`` -- the
        records of every hit with its final (search-space corrected) lnP, before any `TopHits` is built.  The Cython binding
        fills a real ``P7_TOPHITS`` from them (pyhmmer_cuda.CudaLongTargetsPipeline)."""
        from . import longtarget
        if not isinstance(sequences, DigitalSequenceBlock):
            raise TypeError("expected DigitalSequenceBlock, found %s" % type(sequences).__name__)
        if not isinstance(query, (HMM, Profile, OptimizedProfile)):
            raise TypeError("Expected HMM, Profile or OptimizedProfile, found %s" % type(query).__name__)
        if query.alphabet != self.alphabet:
            raise AlphabetMismatch(self.alphabet, query.alphabet)
        if sequences.alphabet != self.alphabet:
            raise AlphabetMismatch(self.alphabet, sequences.alphabet)
        L = len(sequences[0]) if len(sequences) else self.L_HINT
        om = self._optimized(query, L)
        max_length = int(om._desc.max_length)
        if self.window_length is not None and self.window_length > 0:
            max_length = int(self.window_length)
            if max_length != int(om._desc.max_length):      # the resident copy of the profile carries the window length
                if isinstance(query, OptimizedProfile):
                    om = copy.copy(om)
                    om._desc = _lib.OProfileDesc.from_buffer_copy(om._desc)
                om._desc.max_length = max_length
                om._dev = {}
        elif isinstance(query, HMM):
            # the windows keep the model's own MAXL as context; the E-values count windows of p7_Builder_MaxLength(beta)
            max_length = query.compute_max_length(self.window_beta if self.window_beta is not None else 1e-7)
        elif max_length <= 0:
            raise TypeError("Cannot use `Profile` or `OptimizedProfile` query without `max_length` set")
        if int(om._desc.max_length) <= 0:
            raise ValueError("the model carries no window length (MAXL); pass window_length=...")
        if self.block_length <= int(om._desc.max_length):
            raise ValueError("block length (%d) must be greater than the model's window length (%d)" % (self.block_length, int(om._desc.max_length)))
        cut = self._cutoffs(om)
        self._evalue_window = max_length
        residues = None
        if self.Z is not None:                               # Z counts megabases per strand (plan7.pyx:7389-7396)
            residues = int(1000000 * self.Z) * (2 if self.strand is None else 1)
        res = longtarget.search(om, sequences, evalue_window=max_length, evalue_residues=residues, F1=self.F1, F2=self.F2, F3=self.F3, bias_filter=self.bias_filter, null2=self.null2,
                                B1=self.B1, B2=self.B2, B3=self.B3, block_length=self.block_length, strand=self.strand,
                                seed=self.seed, host_threads=self.host_threads, backend_factory=self._backend_factory)
        return om, cut, res

    def search_seq(self, query, sequences, builder=None):
        """nhmmer with a query SEQUENCE (plan7.pyx:7420-7540): the model comes from a `Builder` that shares this pipeline's
        window length / window beta."""
        from .builder import Builder
        if query.alphabet != self.alphabet:
            raise AlphabetMismatch(self.alphabet, query.alphabet)
        if builder is None:
            builder = Builder(self.alphabet, seed=self.seed, window_length=self.window_length, window_beta=self.window_beta)
        elif builder.window_length != self.window_length:
            raise ValueError("builder and long targets pipeline have different window lengths")
        elif self.window_beta is not None and builder.window_beta != self.window_beta:
            raise ValueError("builder and long targets pipeline have different window beta")
        hmm, profile, opt = builder.build(query, self.background)
        hits = self.search_hmm(hmm, sequences)
        hits.query = query
        return hits

    def _long_target_tophits(self, query, om, sequences, res, cut):
        """Hits of `longtarget.search` -> thresholded `TopHits` (the tail of search_hmm, plan7.pyx:7389-7412)."""
        hits, doms, text, dup, stats = res
        th = self._tophits(query, "search", cut)
        th.long_targets = True
        th._params["inc_by_E"] = True                        # sortkey = -lnP whatever the thresholds (p7_tophits.c:804)
        for rec, is_dup in zip(hits, dup):
            h = Hit(th, rec, sequences[rec.seq], doms, text)
            h.sortkey = -h.lnP
            h.duplicate = bool(is_dup)
            if cut is not None and not is_dup:               # flags set by the pipeline under bit cutoffs (p7_pipeline.c:1230-1246)
                p = th._params
                h.reported = h.score >= p["T"]
                h.included = h.reported and h.score >= p["incT"]
                for d in h._domains:
                    d.reported = d.score >= p["domT"]
                    d.included = d.reported and d.score >= p["incdomT"]
            th._hits.append(h)
        th._params["inc_by_E"] = (self.incT is None) and cut is None
        th.Z = float(self.Z) if self.Z is not None else 0.0
        th.domZ = float(self.domZ) if self.domZ is not None else 0.0
        th.searched_models, th.searched_nodes = 1, om.M
        th.searched_sequences, th.searched_residues = stats["nseqs"], stats["nres"]
        th.pos_past_msv, th.pos_past_bias = stats["pos_past_msv"], stats["pos_past_bias"]
        th.pos_past_vit, th.pos_past_fwd = stats["pos_past_vit"], stats["pos_past_fwd"]
        th._sort_by_key()
        th._threshold()
        self._nmodels += 1
        self._nnodes += om.M
        self._nseqs += stats["nseqs"]
        self._nres += stats["nres"]
        return th


def __getattr__(name):                     # ``plan7.Builder``, as in the reference's module (defined in builder.py)
    if name == "Builder":
        from .builder import Builder
        return Builder
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
