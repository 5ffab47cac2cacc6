"""Synthetic Pfam-like models and proteomes as plain numpy arrays -- no dependency on the rest of the package.

`pyhmmer_b200.synth` builds its `HMM` objects from these arrays; `bench_inputs.py` loads this file BY PATH, so that
`bench.py --impl reference` (the reference's CPU pipeline) gets the very same inputs without importing the package, i.e.
without loading libb2h.so.  Everything is seeded and written in explicit float32 / float64 steps: the arrays, and the
HMMER3 ASCII text written from them, are identical wherever they are produced.
"""
import ctypes
import math

import numpy as np

# amino-acid background frequencies (Swiss-Prot 50.8; the data table of p7_AminoFrequencies, hmmer.c:161)
AMINO_FREQ = np.array([0.0787945, 0.0151600, 0.0535222, 0.0668298, 0.0397062, 0.0695071, 0.0229198,
                       0.0590092, 0.0594422, 0.0963728, 0.0237718, 0.0414386, 0.0482904, 0.0395639,
                       0.0540978, 0.0683364, 0.0540687, 0.0673417, 0.0114135, 0.0304133], dtype=np.float32)
SYMBOLS = {20: "ACDEFGHIKLMNPQRSTVWY", 4: "ACGT"}
ALPH = {20: "amino", 4: "DNA"}


def background(K):
    return AMINO_FREQ.copy() if K == 20 else np.full(K, np.float32(1.0) / np.float32(K), dtype=np.float32)


def occupancy(t):
    """p7_hmm_CalculateOccupancy (p7_hmm.c:1338) in the reference's float arithmetic: (mocc, iocc), index 0 unused / node 0."""
    f32, f64 = np.float32, np.float64
    M = t.shape[0] - 1
    mocc, iocc = np.zeros(M + 1, f32), np.zeros(M + 1, f32)
    mocc[1] = f32(t[0, 1] + t[0, 0])
    for k in range(2, M + 1):
        mocc[k] = f32(f64(f32(mocc[k - 1] * f32(t[k - 1, 0] + t[k - 1, 1]))) + (1.0 - f64(mocc[k - 1])) * f64(t[k - 1, 5]))
    iocc[0] = f32(t[0, 1] / t[0, 3])
    for k in range(1, M + 1):
        iocc[k] = f32(f32(mocc[k] * t[k, 1]) / t[k, 3])
    return mocc, iocc


def composition(t, mat, ins):
    """p7_hmm_SetComposition (p7_hmm.c:1391): the occupancy-weighted mean emission distribution, float32 steps."""
    f32 = np.float32
    M, K = t.shape[0] - 1, mat.shape[1]
    mocc, iocc = occupancy(t)
    compo = np.zeros(K, f32)
    compo += ins[0] * iocc[0]
    for k in range(1, M + 1):
        compo += mat[k] * mocc[k]
        compo += ins[k] * iocc[k]
    s = c = f32(0.0)
    for v in compo:                                   # esl_vec_FNorm over a compensated sum
        y = f32(v - c)
        tt = f32(s + y)
        c = f32(f32(tt - s) - y)
        s = tt
    return (compo / s).astype(f32)


def model_arrays(K, M, rng, name, sharpness=1.3, with_composition=True):
    """A random but Pfam-like core model: peaked match emissions, background inserts, sparse indels.  Returns a dict of
    float32 arrays t [(M+1), 7] (MM MI MD IM II DM DD), mat / ins [(M+1), K], compo [K], the consensus string, placeholder
    statistics evparam [6] and the name."""
    bg = background(K).astype(np.float64)
    mat = bg[None, :] * np.exp(sharpness * rng.standard_normal((M, K)))
    boost = rng.integers(0, K, M)                     # one favoured residue per node
    mat[np.arange(M), boost] *= np.exp(rng.uniform(0.5, 2.5, M))
    mat /= mat.sum(1, keepdims=True)
    mat32 = np.zeros((M + 1, K), np.float32)
    mat32[1:] = mat.astype(np.float32)
    mat32[0, 0] = 1.0
    ins32 = np.empty((M + 1, K), np.float32)
    ins32[:] = bg.astype(np.float32)
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
    t32 = t.astype(np.float32)
    cons = np.array(list(SYMBOLS[K]))[mat.argmax(1)]
    strong = mat.max(1) >= (0.5 if K == 20 else 0.9)
    consensus = "".join(c.upper() if s else c.lower() for c, s in zip(cons, strong))
    mmu = -5.0 - math.log(M)
    return dict(name=name, M=int(M), K=K, t=t32, mat=mat32, ins=ins32, consensus=consensus,
                compo=composition(t32, mat32, ins32) if with_composition else None,
                evparam=np.array([mmu, 0.7, mmu - 0.7, 0.7, mmu + 5.1, 0.7], dtype=np.float32), max_length=-1)


def sequence_arrays(K, n, rng, mean_len=350, sd_len=100, lo=50, hi=1500):
    """iid residues from the background, lengths ~ N(mean, sd) clipped to [lo, hi]: a list of uint8 arrays
    (views into one buffer)."""
    bg = background(K).astype(np.float64)
    bg /= bg.sum()
    lens = np.clip(np.rint(rng.normal(mean_len, sd_len, n)), lo, hi).astype(np.int64)
    res = rng.choice(K, size=int(lens.sum()), p=bg).astype(np.uint8)
    out, off = [], 0
    for L in lens:
        out.append(res[off:off + L])
        off += L
    return out


def emit(model, rng):
    """Sample one sequence from the core model (match/insert/delete walk from B to E)."""
    t = model["t"].astype(np.float64)
    mat, ins, M, K = model["mat"], model["ins"], model["M"], model["K"]
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
            e = ins[k].astype(np.float64)
            out.append(rng.choice(K, p=e / e.sum()))
            continue
        k += 1
        if k > M:
            break
        if nxt == 0:
            state = "M"
            e = mat[k].astype(np.float64)
            out.append(rng.choice(K, p=e / e.sum()))
        else:
            state = "D"
    return np.array(out, dtype=np.uint8)


_logf = None


def _libm_logf(x):
    global _logf
    if _logf is None:
        m = ctypes.CDLL("libm.so.6")
        m.logf.restype, m.logf.argtypes = ctypes.c_float, [ctypes.c_float]
        _logf = m.logf
    return _logf(x)


def write_hmm(model, fh):
    """One model in HMMER3/f ASCII format (p7_hmmfile_WriteASCII, p7_hmmfile.c:560-700), exactly as `plan7.HMM.write` prints
    a model made from the same arrays (single-precision logf, %8.5f)."""
    K, M = model["K"], model["M"]

    def prob(p):
        return "*" if p == 0.0 else ("%.5f" % 0.0 if p == 1.0 else "%.5f" % (-_libm_logf(float(p))))

    row = lambda v: " ".join(prob(p).rjust(8) for p in v)
    out = ["HMMER3/f [3.4 | Aug 2023]\n", "NAME  %s\n" % model["name"], "LENG  %d\n" % M]
    if model.get("max_length", -1) > 0:
        out.append("MAXL  %d\n" % model["max_length"])
    out += ["ALPH  %s\n" % ALPH[K], "RF    no\n", "MM    no\n", "CONS  yes\n", "CS    no\n", "MAP   no\n", "NSEQ  1\n", "EFFN  %f\n" % 1.0]
    ev = model["evparam"]
    out += ["STATS LOCAL MSV      %8.4f %8.5f\n" % (ev[0], ev[1]), "STATS LOCAL VITERBI  %8.4f %8.5f\n" % (ev[2], ev[3]),
            "STATS LOCAL FORWARD  %8.4f %8.5f\n" % (ev[4], ev[5])]
    out.append("HMM     " + "".join("     %c   " % c for c in SYMBOLS[K]) + "\n")
    out.append("        %8s %8s %8s %8s %8s %8s %8s\n" % ("m->m", "m->i", "m->d", "i->m", "i->i", "d->m", "d->d"))
    if model.get("compo") is not None:
        out.append("  COMPO  " + row(model["compo"][:K]) + "\n")
    cons = model["consensus"]
    for k in range(0, M + 1):
        if k > 0:
            out.append(" %6d  " % k + row(model["mat"][k]) + " %6s %c - - -\n" % ("-", cons[k - 1]))
        out.append("         " + row(model["ins"][k]) + "\n")
        out.append("         " + row(model["t"][k]) + "\n")
    out.append("//\n")
    fh.write("".join(out).encode())
