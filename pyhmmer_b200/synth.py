"""Synthetic inputs of the shape BASELINE.json names: Pfam-like random profile HMMs and iid proteomes.

Used by bench.py and by the tests (seeded, so both the CUDA path and the CPU oracle see the
very same models after a round trip through the HMMER3 ASCII format).
"""
import numpy as np

from .easel import Alphabet, DigitalSequence, DigitalSequenceBlock
from .plan7 import HMM, Background


def random_hmm(alphabet, M, rng, name=None, sharpness=1.3):
    """A random but Pfam-like core model: peaked match emissions, background inserts, sparse indels."""
    K = alphabet.K
    bg = Background(alphabet).residue_frequencies.astype(np.float64)
    hmm = HMM(alphabet, M, (name or ("synth_M%d" % M)).encode() if not isinstance(name, bytes) else name)
    mat = bg[None, :] * np.exp(sharpness * rng.standard_normal((M, K)))
    boost = rng.integers(0, K, M)                     # one favoured residue per node
    mat[np.arange(M), boost] *= np.exp(rng.uniform(0.5, 2.5, M))
    mat /= mat.sum(1, keepdims=True)
    hmm.match_emissions[1:] = mat.astype(np.float32)
    hmm.match_emissions[0, 0] = 1.0
    hmm.insert_emissions[:] = bg.astype(np.float32)
    t = np.zeros((M + 1, 7))
    mi = rng.uniform(0.002, 0.03, M + 1)
    md = rng.uniform(0.002, 0.03, M + 1)
    im = rng.uniform(0.3, 0.7, M + 1)
    dm = rng.uniform(0.3, 0.8, M + 1)
    t[:, 0] = 1.0 - mi - md; t[:, 1] = mi; t[:, 2] = md
    t[:, 3] = im; t[:, 4] = 1.0 - im
    t[:, 5] = dm; t[:, 6] = 1.0 - dm
    t[0, 5], t[0, 6] = 1.0, 0.0                       # no D_0
    t[M, 0], t[M, 2] = 1.0 - t[M, 1], 0.0             # M_M -> E ; no D_{M+1}
    t[M, 5], t[M, 6] = 1.0, 0.0
    hmm.transition_probabilities[:] = t.astype(np.float32)
    cons = np.array(list(alphabet.symbols[:K]))[mat.argmax(1)]
    strong = mat.max(1) >= (0.5 if alphabet.is_amino() else 0.9)
    hmm.consensus = "".join(c.upper() if s else c.lower() for c, s in zip(cons, strong))
    hmm.set_composition()
    hmm.nseq = 1
    hmm.nseq_effective = 1.0
    return hmm


def random_sequences(alphabet, n, rng, mean_len=350, sd_len=100, lo=50, hi=1500, prefix="seq"):
    """iid residues from the background, lengths ~ N(mean, sd) clipped to [lo, hi]."""
    bg = Background(alphabet).residue_frequencies.astype(np.float64)
    bg /= bg.sum()
    lens = np.clip(np.rint(rng.normal(mean_len, sd_len, n)), lo, hi).astype(np.int64)
    res = rng.choice(alphabet.K, size=int(lens.sum()), p=bg).astype(np.uint8)
    block = DigitalSequenceBlock(alphabet)
    off = 0
    for i, L in enumerate(lens):
        block.append(DigitalSequence(alphabet, name=("%s%d" % (prefix, i)).encode(), sequence=res[off:off + L]))
        off += L
    return block


def emit_sequence(hmm, rng, alphabet=None):
    """Sample one sequence from the core model (match/insert/delete walk from B to E)."""
    abc = hmm.alphabet
    t = hmm.transition_probabilities.astype(np.float64)
    out = []
    k, state = 0, "M"
    while True:
        if state == "M":
            p = t[k, 0:3]
        elif state == "I":
            p = np.array([t[k, 3], t[k, 4], 0.0])
        else:
            p = np.array([t[k, 5], 0.0, t[k, 6]])
        p = p / p.sum()
        nxt = rng.choice(3, p=p)
        if nxt == 1:                                    # -> I_k
            state = "I"
            e = hmm.insert_emissions[k].astype(np.float64)
            out.append(rng.choice(abc.K, p=e / e.sum()))
            continue
        k += 1
        if k > hmm.M:
            break
        if nxt == 0:
            state = "M"
            e = hmm.match_emissions[k].astype(np.float64)
            out.append(rng.choice(abc.K, p=e / e.sum()))
        else:
            state = "D"
    return np.array(out, dtype=np.uint8)
