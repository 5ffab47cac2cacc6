"""Seeded synthetic inputs of every BASELINE.json configuration, as plain numpy arrays.

This module imports neither `pyhmmer_b200` nor `oracle` at load time: `bench.py --impl reference` builds the reference's
models from these arrays (HMMER3 ASCII text written here, or `refm_from_arrays` for the big sets) without ever loading
libb2h.so, and the product arm wraps the very same arrays into `HMM` / `DigitalSequenceBlock` objects.

    C2  hmmsearch   100 Pfam-like profiles (M ~ 200) x 50 000 proteins            c2_inputs()
    C3  hmmsearch   20 000 Pfam-A-sized profiles x 100 000 proteins                pfam_like_models(), c3_sequences()
    C4  hmmscan     one 5 000-residue query x the same 20 000 profiles             c4_query()
    C5  nhmmer      one DNA profile (M ~ 1000) x a 100 Mb genome, both strands     c5_inputs()
"""
import importlib.util
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(ROOT, "tests", "golden")

_spec = importlib.util.spec_from_file_location("_b2h_synth_arrays", os.path.join(ROOT, "pyhmmer_b200", "_synth_arrays.py"))
sa = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(sa)

N_PROFILES = 100
N_SEQS = 50000
PLANT_FRAC = 0.01


# ------------------------------------------------------------------------------------------------ generic pieces
def make_models(n, seed, median_M=180.0, sigma=0.55, lo=30, hi=800, K=20, prefix="synPF", composition="exact"):
    """n random Pfam-like models, lengths ~ lognormal(median, sigma) clipped to [lo, hi].  composition: "exact" = the
    float32 steps of p7_hmm_SetComposition per model (slow: a Python loop over nodes), "batch" = one vectorised pass in
    double precision (the array is handed to both arms, so its last bits do not matter)."""
    rng = np.random.default_rng(seed)
    Ms = np.clip(np.rint(np.exp(rng.normal(np.log(median_M), sigma, n))), lo, hi).astype(int)
    models = [sa.model_arrays(K, int(M), rng, "%s%05d" % (prefix, i), with_composition=(composition == "exact")) for i, M in enumerate(Ms)]
    if composition != "exact":
        batch_composition(models)
    return models


def batch_composition(models):
    """Occupancy-weighted mean emissions of many models at once (p7_hmm_SetComposition's formula in float64, the node
    recursion vectorised over the models)."""
    n = len(models)
    Ms = np.array([m["M"] for m in models])
    Mx = int(Ms.max())
    T = np.zeros((n, Mx + 1, 7))
    for i, m in enumerate(models):
        T[i, :m["M"] + 1] = m["t"]
    T[:, :, 3][T[:, :, 3] == 0] = 1.0                  # padded nodes: avoid 0/0 (their occupancy is zeroed below)
    mocc = np.zeros((n, Mx + 1))
    mocc[:, 1] = T[:, 0, 1] + T[:, 0, 0]
    for k in range(2, Mx + 1):
        mocc[:, k] = mocc[:, k - 1] * (T[:, k - 1, 0] + T[:, k - 1, 1]) + (1.0 - mocc[:, k - 1]) * T[:, k - 1, 5]
    iocc = mocc * T[:, :, 1] / T[:, :, 3]
    iocc[:, 0] = T[:, 0, 1] / T[:, 0, 3]
    for i, m in enumerate(models):
        M = m["M"]
        c = (m["mat"][1:M + 1].astype(np.float64) * mocc[i, 1:M + 1, None]).sum(0) + (m["ins"][:M + 1].astype(np.float64) * iocc[i, :M + 1, None]).sum(0)
        m["compo"] = (c / c.sum()).astype(np.float32)


def make_sequences(n, seed, K=20, **kw):
    return sa.sequence_arrays(K, n, np.random.default_rng(seed), **kw)


def plant(seqs, models, count, seed, max_len=1500):
    """Insert a domain emitted from models[j % len(models)] into <count> randomly chosen sequences (in place)."""
    rng = np.random.default_rng(seed)
    where = rng.choice(len(seqs), count, replace=False)
    for j, t in enumerate(where):
        dom = sa.emit(models[j % len(models)], rng)
        s = seqs[int(t)]
        cut = int(rng.integers(0, len(s) + 1))
        seqs[int(t)] = np.concatenate([s[:cut], dom, s[cut:]])[:max_len]
    return where


def write_hmm_file(models, path):
    with open(path, "wb") as f:
        for m in models:
            sa.write_hmm(m, f)


def apply_stats(models, evparam):
    for m, ev in zip(models, evparam):
        m["evparam"] = np.asarray(ev, dtype=np.float32)


# ------------------------------------------------------------------------------------------------ adapters (lazy imports)
def to_hmms(models, alphabet):
    """`pyhmmer_b200.plan7.HMM` objects over the arrays (product arm)."""
    from pyhmmer_b200 import synth
    return [synth.hmm_from_arrays(alphabet, m) for m in models]


def to_block(seqs, alphabet, prefix="seq"):
    from pyhmmer_b200 import easel
    return easel.DigitalSequenceBlock(alphabet, [easel.DigitalSequence(alphabet, name="%s%d" % (prefix, i), sequence=s) for i, s in enumerate(seqs)])


def to_ref_models(models, nthreads=1, L=400):
    """Reference-side models (oracle/_ref) straight from the arrays: P7_HMM filled like `HMM(alphabet, M, name)` + array
    assignment, then p7_ProfileConfig + p7_oprofile_Convert (reference arm / tests)."""
    from oracle import refshim
    K = models[0]["K"]
    out = []
    for i0 in range(0, len(models), 2000):             # bounded staging buffers
        part = models[i0:i0 + 2000]
        t = np.concatenate([m["t"] for m in part])
        mat = np.concatenate([m["mat"] for m in part])
        ins = np.concatenate([m["ins"] for m in part])
        ev = np.stack([m["evparam"] for m in part])
        cons = "".join(m["consensus"] for m in part)
        out += refshim.models_from_arrays(3 if K == 20 else 2, [m["M"] for m in part], t, mat, ins, ev, [m["name"] for m in part],
                                          compo=np.stack([m["compo"] for m in part]), consensus=cons,
                                          max_length=[max(0, int(m.get("max_length", -1))) for m in part], L=L, nthreads=nthreads)
    return out


# ------------------------------------------------------------------------------------------------ C2 (the headline)
def c2_inputs(rank=0, n_profiles=N_PROFILES, n_seqs=N_SEQS):
    """BASELINE configs[1]: identical on every rank for the profiles, one 50 000-sequence shard per rank.  The RNG call
    order is that of round 1's generator, so models, sequences and planted homologs are unchanged."""
    prng = np.random.default_rng(20240901)
    Ms = np.clip(np.rint(np.exp(prng.normal(np.log(180.0), 0.55, n_profiles))), 30, 800).astype(int)
    models = [sa.model_arrays(20, int(M), prng, "synPF%05d" % i) for i, M in enumerate(Ms)]
    srng = np.random.default_rng(777 + rank)
    seqs = sa.sequence_arrays(20, n_seqs, srng)
    nplant = int(n_seqs * PLANT_FRAC)
    where = srng.choice(n_seqs, nplant, replace=False)
    for j, t in enumerate(where):
        dom = sa.emit(models[j % n_profiles], srng)
        s = seqs[int(t)]
        cut = int(srng.integers(0, len(s) + 1))
        seqs[int(t)] = np.concatenate([s[:cut], dom, s[cut:]])[:1500]
    try:
        st = json.load(open(os.path.join(GOLD, "bench_stats.json")))["evparam"]
        calibrated = len(st) >= n_profiles
        if calibrated:
            apply_stats(models, st)
    except Exception:
        calibrated = False
    return models, seqs, calibrated


# ------------------------------------------------------------------------------------------------ C3 / C4
PFAM_N = 20000


def pfam_like_models(n=PFAM_N):
    """BASELINE configs[2] / [3]: a Pfam-A-sized profile set -- lengths lognormal, median ~150, mean ~190, clipped to
    [20, 2300]; E-value statistics fitted once on a B200 by tools/calibrate_sets.py and committed (tests/golden/)."""
    models = make_models(n, seed=5, median_M=150.0, sigma=0.6, lo=20, hi=2300, prefix="pfl", composition="batch")
    calibrated = False
    path = os.path.join(GOLD, "bench_pfam_like_stats.npy")
    if os.path.exists(path):
        ev = np.load(path)
        if len(ev) >= n:
            apply_stats(models, ev[:n])
            calibrated = True
    return models, calibrated


def c3_sequences(n=100000, rank=0, world=1):
    """100 000 proteins (35 MB); with several ranks the database is cut into contiguous runs balanced by residues
    (the reference's parallel="targets" rule, _hmmsearch.py:153-171) -- every rank generates the whole set (seeded) and keeps
    its run."""
    seqs = make_sequences(n, seed=6)
    return seqs


def c4_query(models, L=5000):
    """One 5 000-residue query: iid background with two domains emitted from models of the set."""
    rng = np.random.default_rng(8)
    bg = sa.background(20).astype(np.float64)
    bg /= bg.sum()
    q = rng.choice(20, size=L, p=bg).astype(np.uint8)
    pos = L // 3
    for j in (3, 7):
        dom = sa.emit(models[j % len(models)], rng)
        q[pos:pos + len(dom)] = dom[:max(0, L - pos)]
        pos += len(dom) + 400
    return q


# ------------------------------------------------------------------------------------------------ C5
def c5_inputs(M=1000, megabases=100.0, plants_per_mb=1.0):
    """BASELINE configs[4]: a DNA profile of M nodes and an iid genome with homologs planted on both strands."""
    rng = np.random.default_rng(11)
    model = sa.model_arrays(4, M, rng, "synDNA%d" % M)
    path = os.path.join(GOLD, "bench_dna_stats.json")
    calibrated = False
    if os.path.exists(path):
        st = json.load(open(path))
        if str(M) in st:
            model["evparam"] = np.asarray(st[str(M)]["evparam"], dtype=np.float32)
            model["max_length"] = int(st[str(M)]["max_length"])
            calibrated = True
    n = int(megabases * 1000000)
    genome = rng.integers(0, 4, n).astype(np.uint8)
    nplant = max(4, int(plants_per_mb * megabases))
    comp = np.array([3, 2, 1, 0], np.uint8)
    for j in range(nplant):
        dom = sa.emit(model, rng)
        if j % 2:
            dom = comp[dom[::-1]]
        pos = int(rng.integers(0, n - len(dom)))
        genome[pos:pos + len(dom)] = dom
    return model, genome, nplant, calibrated
