"""Input containers of the search path, mirroring ``pyhmmer.easel``.

Only what `Pipeline.search_hmm` / `scan_seq` / `hmmsearch` / `hmmscan` read from their
arguments is provided: `Alphabet`, `TextSequence`, `DigitalSequence`, `DigitalSequenceBlock`
and a FASTA `SequenceFile`.  Names, argument meaning and error behaviour follow the reference
(src/pyhmmer/easel.pyx) so that code written against pyhmmer keeps working.
"""
import gzip
import io
import os

import numpy as np

__all__ = ["Alphabet", "TextSequence", "DigitalSequence", "DigitalSequenceBlock", "SequenceFile",
           "AlphabetMismatch"]


class AlphabetMismatch(ValueError):
    """Same meaning as ``pyhmmer.errors.AlphabetMismatch`` (src/pyhmmer/errors.pyx)."""

    def __init__(self, expected, actual):
        super().__init__("alphabets mismatch: expected %r, found %r" % (expected, actual))
        self.expected, self.actual = expected, actual


_ESL_RNA, _ESL_DNA, _ESL_AMINO = 1, 2, 3


class Alphabet:
    """A biological alphabet (``ESL_ALPHABET``; vendor/easel/esl_alphabet.c:172-300).

    Symbols are ordered as Easel orders them: K canonical residues, the gap, the degenerate
    residues, the any-residue, then ``*`` (non-residue) and ``~`` (missing data).
    """

    def __init__(self, type_, symbols, K, degeneracies, equivalents):
        self.type_code = type_
        self.symbols = symbols
        self.K = K
        self.Kp = len(symbols)
        degen = np.zeros((self.Kp, K), dtype=np.uint8)
        for x in range(K):
            degen[x, x] = 1
        degen[self.Kp - 3, :] = 1                       # the "any" symbol (X / N)
        for sym, members in degeneracies.items():
            for m in members:
                degen[symbols.index(sym), symbols.index(m)] = 1
        self.degen = degen
        # input map: byte -> digital code, 255 = illegal (esl_alphabet.c inmap)
        inmap = np.full(256, 255, dtype=np.uint8)
        for i, c in enumerate(symbols):
            inmap[ord(c)] = i
            inmap[ord(c.lower())] = i
        for src, dst in equivalents.items():
            inmap[ord(src)] = symbols.index(dst)
            inmap[ord(src.lower())] = symbols.index(dst)
        self._inmap = inmap
        self._outmap = np.frombuffer(symbols.encode(), dtype=np.uint8)

    @classmethod
    def amino(cls):
        return cls(_ESL_AMINO, "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~", 20,
                   {"B": "ND", "J": "IL", "Z": "QE", "U": "C", "O": "K"}, {"_": "-", ".": "-"})

    @classmethod
    def dna(cls):
        return cls(_ESL_DNA, "ACGT-RYMKSWHBVDN*~", 4,
                   {"R": "AG", "Y": "CT", "M": "AC", "K": "GT", "S": "CG", "W": "AT", "H": "ACT",
                    "B": "CGT", "V": "ACG", "D": "AGT"}, {"U": "T", "X": "N", "I": "A", "_": "-", ".": "-"})

    @classmethod
    def rna(cls):
        return cls(_ESL_RNA, "ACGU-RYMKSWHBVDN*~", 4,
                   {"R": "AG", "Y": "CU", "M": "AC", "K": "GU", "S": "CG", "W": "AU", "H": "ACU",
                    "B": "CGU", "V": "ACG", "D": "AGU"}, {"T": "U", "X": "N", "I": "A", "_": "-", ".": "-"})

    @property
    def type(self):
        return {_ESL_RNA: "RNA", _ESL_DNA: "DNA", _ESL_AMINO: "amino"}[self.type_code]

    def is_dna(self):
        return self.type_code == _ESL_DNA

    def is_rna(self):
        return self.type_code == _ESL_RNA

    def is_amino(self):
        return self.type_code == _ESL_AMINO

    def is_nucleotide(self):
        return self.type_code in (_ESL_DNA, _ESL_RNA)

    def __eq__(self, other):
        return isinstance(other, Alphabet) and self.type_code == other.type_code

    def __hash__(self):
        return hash(self.type_code)

    def __repr__(self):
        return "Alphabet.%s()" % self.type.lower()

    def encode(self, sequence):
        """Text -> digital codes (``esl_abc_Digitize``); raises ``ValueError`` on an illegal symbol."""
        if isinstance(sequence, str):
            sequence = sequence.encode("ascii")
        raw = np.frombuffer(bytes(sequence), dtype=np.uint8)
        codes = self._inmap[raw]
        if codes.size and codes.max() == 255:
            bad = chr(int(raw[int(np.argmax(codes == 255))]))
            raise ValueError("invalid symbol %r in sequence for alphabet %r" % (bad, self))
        return codes

    def decode(self, codes):
        return self._outmap[np.asarray(codes, dtype=np.uint8)].tobytes().decode("ascii")


def _text(v):
    """Names are `str` as in pyhmmer >= 0.11; `bytes` are accepted and decoded."""
    if v is None:
        return ""
    return v.decode("utf-8", "replace") if isinstance(v, (bytes, bytearray)) else str(v)


class _Sequence:
    def __init__(self, name="", description="", accession="", source=""):
        self.name = _text(name)
        self.description = _text(description)
        self.accession = _text(accession)
        self.source = _text(source)


class TextSequence(_Sequence):
    """A sequence in text mode (``pyhmmer.easel.TextSequence``)."""

    def __init__(self, name="", description="", accession="", sequence="", source=""):
        super().__init__(name, description, accession, source)
        self.sequence = sequence

    def __len__(self):
        return len(self.sequence)

    def digitize(self, alphabet):
        return DigitalSequence(alphabet, name=self.name, description=self.description,
                               accession=self.accession, sequence=alphabet.encode(self.sequence),
                               source=self.source)


class DigitalSequence(_Sequence):
    """A sequence in digital mode (``pyhmmer.easel.DigitalSequence``; ``ESL_SQ`` with ``dsq``).

    ``sequence`` holds the residue codes 0..Kp-1 *without* Easel's two sentinel bytes.
    """

    def __init__(self, alphabet, name="", description="", accession="", sequence=None, source=""):
        super().__init__(name, description, accession, source)
        self.alphabet = alphabet
        if sequence is None:
            seq = np.zeros(0, dtype=np.uint8)
        else:
            seq = np.ascontiguousarray(np.frombuffer(bytes(sequence), dtype=np.uint8)
                                       if isinstance(sequence, (bytes, bytearray, memoryview))
                                       else np.asarray(sequence, dtype=np.uint8))
        if seq.size and int(seq.max()) >= alphabet.Kp:
            raise ValueError("invalid alphabet character in digital sequence: %d" % int(seq.max()))
        self.sequence = seq

    def __len__(self):
        return int(self.sequence.size)

    def textize(self):
        return TextSequence(name=self.name, description=self.description, accession=self.accession,
                            sequence=self.alphabet.decode(self.sequence), source=self.source)

    def copy(self):
        return DigitalSequence(self.alphabet, self.name, self.description, self.accession,
                               self.sequence.copy(), self.source)


class DigitalSequenceBlock(list):
    """An ordered block of `DigitalSequence` sharing one alphabet (``pyhmmer.easel.DigitalSequenceBlock``).

    The block is the unit the search path uploads to the GPU: `_packed()` returns the
    concatenated residues and offsets that ``b2h_seqdb_create_packed`` consumes; the device
    copy is cached per context until the block is modified.
    """

    def __init__(self, alphabet, iterable=()):
        super().__init__()
        self.alphabet = alphabet
        self._cache = {}
        for s in iterable:
            self.append(s)

    def _check(self, seq):
        if not isinstance(seq, DigitalSequence):
            raise TypeError("expected DigitalSequence, found %s" % type(seq).__name__)
        if seq.alphabet != self.alphabet:
            raise AlphabetMismatch(self.alphabet, seq.alphabet)

    def append(self, seq):
        self._check(seq)
        self._cache = {}
        super().append(seq)

    def extend(self, seqs):
        for s in seqs:
            self.append(s)

    def __setitem__(self, i, v):
        if isinstance(i, slice):
            v = list(v)
            for s in v:
                self._check(s)
        else:
            self._check(v)
        self._cache = {}
        super().__setitem__(i, v)

    def __delitem__(self, i):
        self._cache = {}
        super().__delitem__(i)

    # every other mutating method of `list` drops the packed / device copies too (a search after a mutation must never
    # run against the previous database) and keeps the type / alphabet check
    def insert(self, i, seq):
        self._check(seq)
        self._cache = {}
        super().insert(i, seq)

    def pop(self, i=-1):
        self._cache = {}
        return super().pop(i)

    def remove(self, seq):
        self._cache = {}
        super().remove(seq)

    def clear(self):
        self._cache = {}
        super().clear()

    def sort(self, *, key=None, reverse=False):
        self._cache = {}
        super().sort(key=key, reverse=reverse)

    def reverse(self):
        self._cache = {}
        super().reverse()

    def __iadd__(self, seqs):
        self.extend(seqs)
        return self

    def __imul__(self, n):
        self._cache = {}
        return super().__imul__(n)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return DigitalSequenceBlock(self.alphabet, list.__getitem__(self, i))
        return list.__getitem__(self, i)

    def copy(self):
        return DigitalSequenceBlock(self.alphabet, self)

    def largest(self):
        if not self:
            raise ValueError("block is empty")
        return max(self, key=len)

    @property
    def total_residues(self):
        n = self._cache.get("nres")
        if n is None:
            n = self._cache["nres"] = sum(len(s) for s in self)
        return n

    def _packed(self):
        hit = self._cache.get("packed")
        if hit is None:
            lens = np.fromiter((len(s) for s in self), dtype=np.int64, count=len(self))
            off = np.zeros(len(self) + 1, dtype=np.int64)
            np.cumsum(lens, out=off[1:])
            res = np.concatenate([s.sequence for s in self]) if len(self) else np.zeros(0, np.uint8)
            hit = self._cache["packed"] = (np.ascontiguousarray(res, dtype=np.uint8), off)
        return hit


class SequenceFile:
    """Minimal FASTA reader with the ``pyhmmer.easel.SequenceFile`` surface used by the search path."""

    def __init__(self, file, format=None, digital=False, alphabet=None):
        if format not in (None, "fasta", "afa"):
            raise ValueError("only FASTA input is supported by pyhmmer_b200.easel.SequenceFile")
        self.name = file if isinstance(file, (str, os.PathLike)) else None
        if self.name is not None:
            fh = open(file, "rb")
            if fh.peek(2)[:2] == b"\x1f\x8b":
                fh = gzip.open(fh)
            self._fh = fh
        else:
            self._fh = file
        self.digital = digital
        self.alphabet = alphabet
        if digital and alphabet is None:
            self.alphabet = Alphabet.amino()
        self._pending = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        if self.name is not None:
            self._fh.close()

    def __iter__(self):
        return self

    def __next__(self):
        s = self.read()
        if s is None:
            raise StopIteration
        return s

    def read(self):
        header = self._pending
        self._pending = None
        chunks = []
        while True:
            line = self._fh.readline()
            if not line:
                break
            if isinstance(line, str):
                line = line.encode()
            if line.startswith(b">"):
                if header is None:
                    header = line
                    continue
                self._pending = line
                break
            if header is not None:
                chunks.append(line.strip())
        if header is None:
            return None
        parts = header[1:].strip().split(None, 1)
        name = parts[0] if parts else b""
        desc = parts[1] if len(parts) > 1 else b""
        text = b"".join(chunks).replace(b" ", b"").decode("ascii")
        seq = TextSequence(name=name, description=desc, sequence=text)
        return seq.digitize(self.alphabet) if self.digital else seq

    def read_block(self, sequences=None, residues=None):
        if not self.digital:
            raise ValueError("read_block requires a SequenceFile in digital mode")
        block = DigitalSequenceBlock(self.alphabet)
        n = r = 0
        while (sequences is None or n < sequences) and (residues is None or r < residues):
            s = self.read()
            if s is None:
                break
            block.append(s)
            n += 1
            r += len(s)
        return block


class TextMSA:
    """A multiple alignment in text mode: the part of ``pyhmmer.easel.TextMSA`` that `TopHits.to_msa` fills -- aligned
    rows with names, accessions and descriptions, the reference (RF) line and posterior-probability annotation."""

    def __init__(self, name=None, names=(), sequences=(), accessions=None, descriptions=None, reference=None,
                 posterior_probabilities=None, consensus_posterior_probabilities=None):
        self.name = name
        self.names = list(names)
        self.alignment = list(sequences)
        self.accessions = list(accessions) if accessions is not None else [None] * len(self.names)
        self.descriptions = list(descriptions) if descriptions is not None else [None] * len(self.names)
        self.reference = reference
        self.posterior_probabilities = list(posterior_probabilities) if posterior_probabilities is not None else None
        self.consensus_posterior_probabilities = consensus_posterior_probabilities
        self.description = self.accession = self.author = None
        self.model_mask = self.secondary_structure = None
        if len({len(r) for r in self.alignment}) > 1:
            raise ValueError("aligned sequences of different lengths")

    def digitize(self, alphabet):
        """``TextMSA.digitize``: the same alignment in digital mode (esl_msa_Digitize)."""
        msa = DigitalMSA(alphabet, name=self.name, names=self.names, rows=[alphabet.encode(r) for r in self.alignment],
                         accessions=self.accessions, descriptions=self.descriptions, reference=self.reference)
        msa.description, msa.accession, msa.author = self.description, self.accession, self.author
        msa.model_mask, msa.secondary_structure = self.model_mask, self.secondary_structure
        msa.posterior_probabilities = self.posterior_probabilities
        msa.consensus_posterior_probabilities = self.consensus_posterior_probabilities
        return msa

    def __len__(self):
        return len(self.alignment[0]) if self.alignment else 0

    @property
    def sequences(self):
        """The rows as `TextSequence` (aligned text)."""
        return [TextSequence(name=n, description=d or b"", accession=a or b"", sequence=s)
                for n, a, d, s in zip(self.names, self.accessions, self.descriptions, self.alignment)]

    def write(self, fh, format="stockholm"):
        """Write the alignment to a binary file handle: ``stockholm`` / ``pfam`` (one block) or ``afa`` (aligned FASTA)."""
        txt = lambda v: v.decode() if isinstance(v, bytes) else str(v)
        out = []
        if format in ("stockholm", "pfam"):
            names = [txt(n) for n in self.names]
            w = max([len(n) for n in names] + [0])
            gr = w + 9 if self.posterior_probabilities else 0
            margin = max(w + 1, gr, 13 if self.consensus_posterior_probabilities else (8 if self.reference else 0))
            out.append("# STOCKHOLM 1.0\n\n")
            for n, a in zip(names, self.accessions):
                if a:
                    out.append("#=GS %-*s AC %s\n" % (w, n, txt(a)))
            for n, d in zip(names, self.descriptions):
                if d:
                    out.append("#=GS %-*s DE %s\n" % (w, n, txt(d)))
            out.append("\n")
            for i, (n, row) in enumerate(zip(names, self.alignment)):
                out.append("%-*s%s\n" % (margin, n, row))
                if self.posterior_probabilities and self.posterior_probabilities[i]:
                    out.append("%-*s%s\n" % (margin, "#=GR %-*s PP" % (w, n), self.posterior_probabilities[i]))
            if self.consensus_posterior_probabilities:
                out.append("%-*s%s\n" % (margin, "#=GC PP_cons", self.consensus_posterior_probabilities))
            if self.reference:
                out.append("%-*s%s\n" % (margin, "#=GC RF", self.reference))
            out.append("//\n")
        elif format == "afa":
            for n, d, row in zip(self.names, self.descriptions, self.alignment):
                out.append(">%s%s\n" % (txt(n), (" " + txt(d)) if d else ""))
                out.extend(row[i:i + 60] + "\n" for i in range(0, len(row), 60))
        else:
            raise ValueError("invalid format %r (expected 'stockholm', 'pfam' or 'afa')" % (format,))
        fh.write("".join(out).encode())


class DigitalMSA:
    """A multiple alignment in digital mode (``pyhmmer.easel.DigitalMSA``): ``ax[nseq][alen]`` residue codes (gap = K, missing
    data = Kp - 1), row names, the per-column annotation lines a model builder reads (RF, model mask, consensus structure),
    sequence weights and cutoffs.  `Builder.build_msa` rewrites its argument the way p7_Builder does (weights, fragment
    marks, RF line): hand it a `copy()` to keep the original."""

    def __init__(self, alphabet, name=None, description=None, accession=None, sequences=None, author=None, *, names=None, rows=None,
                 accessions=None, descriptions=None, reference=None):
        self.alphabet = alphabet
        self.name, self.description, self.accession, self.author = name, description, accession, author
        if sequences is not None:                            # aligned DigitalSequence objects, as pyhmmer takes them
            names = [q.name for q in sequences]
            rows = [np.asarray(q.sequence, np.uint8) for q in sequences]
            accessions = [q.accession or None for q in sequences]
            descriptions = [q.description or None for q in sequences]
        rows = [np.asarray(r, np.uint8) for r in (rows or [])]
        if len({len(r) for r in rows}) > 1:
            raise ValueError("all sequences must have the same length")
        self.ax = np.array(rows, np.uint8).reshape(len(rows), len(rows[0]) if rows else 0)
        self.names = list(names or [])
        if len(self.names) != len(rows):
            raise ValueError("one name per aligned sequence is needed")
        if len(set(self.names)) != len(self.names):
            raise ValueError("duplicate name in alignment")
        self.accessions = list(accessions) if accessions is not None else [None] * len(rows)
        self.descriptions = list(descriptions) if descriptions is not None else [None] * len(rows)
        self.reference = reference
        self.model_mask = self.secondary_structure = None
        self.posterior_probabilities = self.consensus_posterior_probabilities = None
        self.sequence_weights = None
        self.cutoffs = {}

    def __len__(self):
        return int(self.ax.shape[1])

    @property
    def sequences(self):
        """The rows as `DigitalSequence` (aligned codes)."""
        return [DigitalSequence(self.alphabet, name=n, description=d or "", accession=a or "", sequence=r)
                for n, a, d, r in zip(self.names, self.accessions, self.descriptions, self.ax)]

    @property
    def alignment(self):
        return [r.copy() for r in self.ax]

    @property
    def checksum(self):
        from .msabuild import msa_checksum
        return msa_checksum(self.ax)

    def copy(self):
        m = DigitalMSA(self.alphabet, self.name, self.description, self.accession, None, self.author, names=self.names,
                       rows=list(self.ax), accessions=self.accessions, descriptions=self.descriptions, reference=self.reference)
        m.model_mask, m.secondary_structure = self.model_mask, self.secondary_structure
        m.posterior_probabilities, m.consensus_posterior_probabilities = self.posterior_probabilities, self.consensus_posterior_probabilities
        m.sequence_weights = None if self.sequence_weights is None else np.array(self.sequence_weights, np.float64)
        m.cutoffs = dict(self.cutoffs)
        return m

    def textize(self):
        """``DigitalMSA.textize``: the alignment in text mode (gaps as ``-``)."""
        m = TextMSA(name=self.name, names=self.names, sequences=[self.alphabet.decode(r) for r in self.ax], accessions=self.accessions,
                    descriptions=self.descriptions, reference=self.reference, posterior_probabilities=self.posterior_probabilities,
                    consensus_posterior_probabilities=self.consensus_posterior_probabilities)
        m.description, m.accession, m.author = self.description, self.accession, self.author
        m.model_mask, m.secondary_structure = self.model_mask, self.secondary_structure
        return m
