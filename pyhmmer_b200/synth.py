"""Synthetic inputs of the shape BASELINE.json names: Pfam-like random profile HMMs and iid proteomes.

Used by bench.py and by the tests (seeded, so both the CUDA path and the CPU oracle see the
very same models after a round trip through the HMMER3 ASCII format).
"""
import math

import numpy as np

from . import _synth_arrays
from .easel import Alphabet, DigitalSequence, DigitalSequenceBlock
from .plan7 import HMM, Background


def hmm_from_arrays(alphabet, model):
    """An `HMM` from the arrays of `_synth_arrays.model_arrays`."""
    hmm = HMM(alphabet, model["M"], model["name"])
    hmm.match_emissions[:] = model["mat"]
    hmm.insert_emissions[:] = model["ins"]
    hmm.transition_probabilities[:] = model["t"]
    hmm.consensus = model["consensus"]
    if model.get("compo") is not None:
        hmm._compo[:] = 0.0
        hmm._compo[:alphabet.K] = model["compo"]
    else:
        hmm.set_composition()
    hmm.nseq = 1
    hmm.nseq_effective = 1.0
    hmm._evparam[:] = model["evparam"]
    if model.get("max_length", -1) > 0:
        hmm.max_length = int(model["max_length"])
    return hmm


def random_hmm(alphabet, M, rng, name=None, sharpness=1.3):
    """A random but Pfam-like core model: peaked match emissions, background inserts, sparse indels (the arrays come from
    `_synth_arrays.model_arrays`, which bench.py's reference arm uses without this package)."""
    return hmm_from_arrays(alphabet, _synth_arrays.model_arrays(alphabet.K, M, rng, name or ("synth_M%d" % M), sharpness))


def random_sequences(alphabet, n, rng, mean_len=350, sd_len=100, lo=50, hi=1500, prefix="seq"):
    """iid residues from the background, lengths ~ N(mean, sd) clipped to [lo, hi]."""
    block = DigitalSequenceBlock(alphabet)
    for i, res in enumerate(_synth_arrays.sequence_arrays(alphabet.K, n, rng, mean_len, sd_len, lo, hi)):
        block.append(DigitalSequence(alphabet, name="%s%d" % (prefix, i), sequence=res))
    return block


def emit_sequence(hmm, rng, alphabet=None):
    """Sample one sequence from the core model (match/insert/delete walk from B to E)."""
    return _synth_arrays.emit(dict(t=hmm.transition_probabilities, mat=hmm.match_emissions, ins=hmm.insert_emissions,
                                   M=hmm.M, K=hmm.alphabet.K), rng)


# ---------------------------------------------------------------------------------------------------
# E-value calibration of synthetic models on the GPU (the procedure of p7_Calibrate, evalues.c:58-150:
# lambda from the mean match relative entropy; MSV / Viterbi mu by ML Gumbel location fit on N=200 iid
# sequences of L=200; Forward tau from a complete Gumbel fit on N=200 sequences of L=100, tail mass 0.04).
# Random sequences come from numpy, not Easel's RNG, so the fitted numbers are statistically -- not
# bitwise -- those the reference's hmmbuild would write; both arms of the benchmark read the SAME file.
# ---------------------------------------------------------------------------------------------------
def _gumbel_fit_loc(x, lam):
    x = np.asarray(x, dtype=np.float64)
    return -np.log(np.mean(np.exp(-lam * x))) / lam


def _gumbel_fit_complete(x):
    """ML fit of (mu, lambda) (esl_gumbel_FitComplete: Newton-Raphson on lambda)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    lam = math.pi / math.sqrt(6.0 * x.var())
    for _ in range(100):
        e = np.exp(-lam * x)
        esum, xesum, xxesum = e.sum(), (x * e).sum(), (x * x * e).sum()
        fx = 1.0 / lam - x.mean() + xesum / esum
        dfx = (xesum / esum) ** 2 - xxesum / esum - 1.0 / (lam * lam)
        step = fx / dfx
        lam -= step
        if abs(step) < 1e-9:
            break
    mu = -np.log(np.exp(-lam * x).sum() / n) / lam
    return mu, lam


def calibrate(hmms, ctx=None, seed=42):
    """Set MSV/Viterbi/Forward statistics on each HMM in place, scoring random sequences on the GPU."""
    from . import _lib, plan7
    if not hmms:
        return
    ctx = ctx or _lib.context()
    abc = hmms[0].alphabet
    rng = np.random.default_rng(seed)
    bg = plan7.Background(abc)
    f = bg.residue_frequencies.astype(np.float64)
    f /= f.sum()

    def iid(n, L):
        res = rng.choice(abc.K, size=n * L, p=f).astype(np.uint8)
        return DigitalSequenceBlock(abc, [DigitalSequence(abc, name="r%d" % i, sequence=res[i * L:(i + 1) * L]) for i in range(n)])

    b200, b100 = iid(200, 200), iid(200, 100)
    db200, db100 = plan7.SequenceDatabase(ctx, b200), plan7.SequenceDatabase(ctx, b100)
    n1_200 = np.empty(200, np.float32)
    n1_100 = np.empty(200, np.float32)
    sc = np.empty(200, np.float32)
    st = np.empty(200, np.int32)
    LOG2 = math.log(2.0)
    first = True
    for hmm in hmms:
        mat = hmm.match_emissions[1:].astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            H = np.where(mat > 0, mat * np.log2(mat / f[None, :]), 0.0).sum(1).mean()
        lam = LOG2 + 1.44 / (hmm.M * H)
        om = plan7.Profile(hmm.M, abc).configure(hmm, bg, 200).to_optimized()
        h = om._device(ctx)
        if first:
            _lib.check(_lib.lib.b2h_null_scores(ctx.handle, h, db200.handle, _lib.ptr(n1_200), None), "null", ctx.handle)
            _lib.check(_lib.lib.b2h_null_scores(ctx.handle, h, db100.handle, _lib.ptr(n1_100), None), "null", ctx.handle)
            first = False
        _lib.check(_lib.lib.b2h_msv_filter(ctx.handle, h, db200.handle, _lib.ptr(sc), _lib.ptr(st)), "msv", ctx.handle)
        maxsc = (255 - om.base) / om.scale_b
        x = (np.where(np.isinf(sc), maxsc, sc) - n1_200) / LOG2
        mmu = _gumbel_fit_loc(x, lam)
        _lib.check(_lib.lib.b2h_viterbi_filter(ctx.handle, h, db200.handle, _lib.ptr(sc), _lib.ptr(st)), "vit", ctx.handle)
        maxsc = (32767.0 - om.base_w) / om.scale_w
        x = (np.where(np.isinf(sc), maxsc, sc) - n1_200) / LOG2
        vmu = _gumbel_fit_loc(x, lam)
        _lib.check(_lib.lib.b2h_forward_parser(ctx.handle, h, db100.handle, _lib.ptr(sc), _lib.ptr(st)), "fwd", ctx.handle)
        x = (sc.astype(np.float64) - n1_100) / LOG2
        gmu, glam = _gumbel_fit_complete(x)
        tailp = 0.04
        tau = (gmu - math.log(-math.log(1.0 - tailp)) / glam) + math.log(tailp) / lam
        hmm._evparam[:] = np.array([mmu, lam, vmu, lam, tau, lam], dtype=np.float32)
