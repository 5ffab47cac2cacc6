"""Model construction for single query sequences: what ``pyhmmer.plan7.Builder.build`` does for phmmer / nhmmer sequence
queries (p7_SingleBuilder, vendor/hmmer/src/p7_builder.c:440) -- a profile HMM from one sequence and a substitution score
matrix (p7_Seqmodel, seqmodel.c:49), its composition and consensus, and the calibration of its E-value parameters
(p7_Calibrate, evalues.c:58) on random sequences drawn with Easel's Mersenne Twister, scored by the GPU filters.

Models from multiple alignments (`Builder.build_msa`: weights, architecture, priors, effective sequence number) are built
by `msabuild`; both kinds share the calibration below.
"""
import math

import numpy as np


class Randomness:
    """Easel's Mersenne Twister (``ESL_RANDOMNESS``, vendor/easel/esl_random.c: MT19937 with Easel's own 69069 seeding)."""

    def __init__(self, seed=42):
        self.seed = int(seed) or 42
        self.reinit()

    def reinit(self):
        """esl_randomness_Init with the generator's own seed (what a reseeding builder does before every model)."""
        mt = np.empty(624, np.uint64)
        x = self.seed & 0xffffffff
        for z in range(624):
            mt[z] = x
            x = (69069 * x) & 0xffffffff
        self._mt = mt.astype(np.uint32)
        self._fill()

    def _fill(self):
        mt = self._mt.astype(np.uint64)
        mag = (0, 0x9908b0df)
        for z in range(227):
            y = (int(mt[z]) & 0x80000000) | (int(mt[z + 1]) & 0x7fffffff)
            mt[z] = int(mt[z + 397]) ^ (y >> 1) ^ mag[y & 1]
        for z in range(227, 623):
            y = (int(mt[z]) & 0x80000000) | (int(mt[z + 1]) & 0x7fffffff)
            mt[z] = int(mt[z - 227]) ^ (y >> 1) ^ mag[y & 1]
        y = (int(mt[623]) & 0x80000000) | (int(mt[0]) & 0x7fffffff)
        mt[623] = int(mt[396]) ^ (y >> 1) ^ mag[y & 1]
        self._mt = mt.astype(np.uint32)
        # the 624 tempered outputs of this table, as doubles in [0, 1)
        x = self._mt.copy()
        x ^= x >> np.uint32(11)
        x ^= (x << np.uint32(7)) & np.uint32(0x9d2c5680)
        x ^= (x << np.uint32(15)) & np.uint32(0xefc60000)
        x ^= x >> np.uint32(18)
        self._out = x.astype(np.float64) / 4294967296.0
        self._i = 0

    def random(self, n=None):
        """esl_random(): the next value(s), uniform on [0, 1)."""
        if n is None:
            if self._i >= 624:
                self._fill()
            v = self._out[self._i]
            self._i += 1
            return float(v)
        out = np.empty(n, np.float64)
        k = 0
        while k < n:
            if self._i >= 624:
                self._fill()
            take = min(n - k, 624 - self._i)
            out[k:k + take] = self._out[self._i:self._i + take]
            self._i += take
            k += take
        return out

    def iid(self, p, L):
        """esl_rsq_xfIID: L residues drawn from the float probabilities p (esl_rnd_FChoose: double sums, roll < sum / norm)."""
        p = np.asarray(p, np.float32).astype(np.float64)
        norm = 0.0
        for v in p:
            norm += v
        cdf = np.cumsum(p) / norm
        roll = self.random(L)
        return np.minimum(np.searchsorted(cdf, roll, side="right"), len(p) - 1).astype(np.uint8)


# -- substitution score systems (vendor/easel/esl_scorematrix.c: built-in matrices; canonical residues only) ----------------
_BLOSUM62 = """
 4  0 -2 -1 -2  0 -2 -1 -1 -1 -1 -2 -1 -1 -1  1  0  0 -3 -2
 0  9 -3 -4 -2 -3 -3 -1 -3 -1 -1 -3 -3 -3 -3 -1 -1 -1 -2 -2
-2 -3  6  2 -3 -1 -1 -3 -1 -4 -3  1 -1  0 -2  0 -1 -3 -4 -3
-1 -4  2  5 -3 -2  0 -3  1 -3 -2  0 -1  2  0  0 -1 -2 -3 -2
-2 -2 -3 -3  6 -3 -1  0 -3  0  0 -3 -4 -3 -3 -2 -2 -1  1  3
 0 -3 -1 -2 -3  6 -2 -4 -2 -4 -3  0 -2 -2 -2  0 -2 -3 -2 -3
-2 -3 -1  0 -1 -2  8 -3 -1 -3 -2  1 -2  0  0 -1 -2 -3 -2  2
-1 -1 -3 -3  0 -4 -3  4 -3  2  1 -3 -3 -3 -3 -2 -1  3 -3 -1
-1 -3 -1  1 -3 -2 -1 -3  5 -2 -1  0 -1  1  2  0 -1 -2 -3 -2
-1 -1 -4 -3  0 -4 -3  2 -2  4  2 -3 -3 -2 -2 -2 -1  1 -2 -1
-1 -1 -3 -2  0 -3 -2  1 -1  2  5 -2 -2  0 -1 -1 -1  1 -1 -1
-2 -3  1  0 -3  0  1 -3  0 -3 -2  6 -2  0  0  1  0 -3 -4 -2
-1 -3 -1 -1 -4 -2 -2 -3 -1 -3 -2 -2  7 -1 -2 -1 -1 -2 -4 -3
-1 -3  0  2 -3 -2  0 -3  1 -2  0  0 -1  5  1  0 -1 -2 -2 -1
-1 -3 -2  0 -3 -2  0 -3  2 -2 -1  0 -2  1  5 -1 -1 -3 -3 -2
 1 -1  0  0 -2  0 -1 -2  0 -2 -1  1 -1  0 -1  4  1 -2 -3 -2
 0 -1 -1 -1 -2 -2 -2 -1 -1 -1 -1  0 -1 -1 -1  1  5  0 -2 -2
 0 -1 -3 -2 -1 -3 -3  3 -2  1  1 -3 -2 -2 -3 -2  0  4 -3 -1
-3 -2 -4 -3  1 -2 -2 -3 -3 -2 -1 -4 -4 -2 -3 -3 -2 -3 11  2
-2 -2 -3 -2  3 -3  2 -1 -2 -1 -1 -2 -3 -1 -2 -2 -2 -1  2  7
"""
_DNA1 = """
 41 -32 -26 -26
-32  39 -38 -17
-26 -38  46 -31
-26 -17 -31  39
"""
SCORE_MATRICES = {"BLOSUM62": np.array(_BLOSUM62.split(), np.int64).reshape(20, 20), "DNA1": np.array(_DNA1.split(), np.int64).reshape(4, 4)}


_Q_CACHE = {}


def conditional_probabilities(alphabet, matrix, f):
    key = (alphabet.type, matrix, np.asarray(f, np.float32).tobytes())
    if key not in _Q_CACHE:
        _Q_CACHE[key] = _conditional_probabilities(alphabet, matrix, f)
    return _Q_CACHE[key]


def _conditional_probabilities(alphabet, matrix, f):
    """P(b | a) for every query residue code a (degenerate ones included) from a score matrix and background f:
    esl_scorematrix_ProbifyGivenBG (lambda by Newton/Raphson from the far side, esl_scorematrix.c) followed by
    esl_scorematrix_JointToConditionalOnQuery.  Returns a [Kp, K] float64 array."""
    S = SCORE_MATRICES[matrix].astype(np.float64)
    K, Kp = alphabet.K, alphabet.Kp
    if S.shape != (K, K):
        raise ValueError("score matrix %s does not fit the %s alphabet" % (matrix, alphabet.type))
    f = np.asarray(f, np.float32).astype(np.float64)[:K]
    ff = f[:, None] * f[None, :]

    def fdf(lam):
        fx = dfx = 0.0
        for i in range(K):
            for j in range(K):
                t = ff[i, j] * math.exp(lam * S[i, j])
                fx += t
                dfx += t * S[i, j]
        return fx - 1.0, dfx

    lam = 1.0 / float(S.max())
    fx = -1.0
    while lam < 50.0:
        fx, dfx = fdf(lam)
        if fx > 0:
            break
        lam *= 2.0
    if fx <= 0:
        raise ValueError("failed to bracket the root for lambda of score matrix %s" % matrix)
    x = lam
    fx, dfx = fdf(x)
    for _ in range(100):
        x0 = x
        x = x - fx / dfx
        fx, dfx = fdf(x)
        if fx == 0 or abs(x - x0) < 1e-15 + 1e-15 * x:
            break
    P = np.zeros((Kp, K), np.float64)
    for i in range(K):
        for j in range(K):
            P[i, j] = f[i] * f[j] * math.exp(x * S[i, j])
    for ip in range(K + 1, Kp - 2):                      # degenerate query residues: sums over what they stand for
        for j in range(K):
            s = 0.0
            for i in range(K):
                if alphabet.degen[ip, i]:
                    s += P[i, j]
            P[ip, j] = s
    Q = np.zeros_like(P)
    for a in range(Kp - 2):
        marg = 0.0                                       # P(a, X): the sum over all canonical b
        for j in range(K):
            marg += P[a, j]
        if marg != 0.0:
            Q[a] = P[a] / marg
    return Q


class FastRandomness:
    """Easel's "fast" generator, the one a builder owns (esl_randomness_CreateFast, p7_builder.c:127): Knuth's linear
    congruential x <- 69069 x + 1 on 32 bits, seeded through esl_mix3 (esl_random.c)."""

    def __init__(self, seed=42):
        self.seed = int(seed) or 42
        self.reinit()

    @staticmethod
    def _mix3(a, b, c):
        m = 0xffffffff
        for s1, s2, s3 in ((13, 8, 13), (12, 16, 5), (3, 10, 15)):
            a = (a - b - c) & m; a ^= c >> s1
            b = (b - c - a) & m; b ^= (a << s2) & m
            c = (c - a - b) & m; c ^= b >> s3
        return c

    def reinit(self):
        self._x = self._mix3(self.seed & 0xffffffff, 87654321, 12345678) or 42

    _A = np.array([1], np.uint64)                         # a^k and (a^k - 1) / (a - 1) mod 2^32, grown by doubling:
    _C = np.array([0], np.uint64)                         # x_k = A_k x_0 + C_k

    @classmethod
    def _tables(cls, n):
        m32 = np.uint64(0xffffffff)
        while len(cls._A) <= n:
            m = len(cls._A)
            Am = (cls._A[m - 1] * np.uint64(69069)) & m32                     # a^m
            Cm = (cls._C[m - 1] * np.uint64(69069) + np.uint64(1)) & m32     # C_m
            cls._A = np.concatenate([cls._A, (cls._A * Am) & m32])           # A_{m+i} = A_m A_i
            cls._C = np.concatenate([cls._C, (cls._A[:m] * Cm + cls._C) & m32])   # C_{m+i} = A_i C_m + C_i
        return cls._A, cls._C

    def random(self, n=None):
        if n is None:
            self._x = (self._x * 69069 + 1) & 0xffffffff
            return self._x / 4294967296.0
        A, C = self._tables(n)
        x = (A[1:n + 1] * np.uint64(self._x) + C[1:n + 1]) & np.uint64(0xffffffff)
        if n:
            self._x = int(x[-1])
        return x.astype(np.float64) / 4294967296.0

    iid = Randomness.iid


class Builder:
    """``pyhmmer.plan7.Builder`` for single query sequences (plan7.pyx:870-1260): `build` turns one sequence into a
    calibrated HMM with a substitution score matrix (phmmer / nhmmer sequence queries).  The calibration draws its random
    sequences with Easel's generator and scores them with the GPU filters."""

    def __init__(self, alphabet, *, architecture="fast", weighting="pb", effective_number="entropy", prior_scheme="alphabet",
                 symfrac=0.5, fragthresh=0.5, esigma=45.0, ere=None,
                 seed=42, popen=None, pextend=None, score_matrix=None, window_length=None, window_beta=None,
                 EmL=200, EmN=200, EvL=200, EvN=200, EfL=100, EfN=200, Eft=0.04):
        from .msabuild import Prior, ETARGET
        nucleotide = alphabet.K == 4
        self.alphabet = alphabet
        self.seed = int(seed)
        # alignment queries (build_msa): plan7.pyx:638-831
        if architecture not in ("fast", "hand"):
            raise ValueError("invalid architecture %r (expected 'fast' or 'hand')" % (architecture,))
        if weighting not in ("pb", "gsc", "blosum", "none", "given"):
            raise ValueError("invalid weighting %r (expected 'pb', 'gsc', 'blosum', 'none' or 'given')" % (weighting,))
        if isinstance(effective_number, str):
            if effective_number not in ("entropy", "exp", "clust", "none"):
                raise ValueError("invalid effective_number %r (expected 'entropy', 'exp', 'clust', 'none' or a number)" % (effective_number,))
        elif not isinstance(effective_number, (int, float)):
            raise TypeError("Expected str, int or float, found %s" % type(effective_number).__name__)
        self.architecture, self.weighting, self.effective_number, self.prior_scheme = architecture, weighting, effective_number, prior_scheme
        self.symfrac, self.fragthresh, self.esigma = float(symfrac), float(fragthresh), float(esigma)
        self.re_target = float(ere) if ere is not None else ETARGET.get(alphabet.type, 1.0)
        self.prior = Prior.for_alphabet(alphabet, prior_scheme)
        self.popen = (0.03125 if nucleotide else 0.02) if popen is None else float(popen)
        self.pextend = (0.75 if nucleotide else 0.4) if pextend is None else float(pextend)
        self.score_matrix = ("DNA1" if nucleotide else "BLOSUM62") if score_matrix is None else score_matrix
        if self.score_matrix not in SCORE_MATRICES:
            raise ValueError("no matrix named %s is available as a built-in" % self.score_matrix)
        self.window_length, self.window_beta = window_length, (1e-7 if window_beta is None else float(window_beta))
        self.EmL, self.EmN, self.EvL, self.EvN, self.EfL, self.EfN, self.Eft = EmL, EmN, EvL, EvN, EfL, EfN, Eft
        self.randomness = FastRandomness(self.seed)
        self._scorer = None                              # tests plug the reference's filters in here; None = the GPU

    def build(self, sequence, background):
        """(HMM, Profile, OptimizedProfile) for one digital query sequence (p7_SingleBuilder, p7_builder.c:440)."""
        from . import plan7
        import time
        abc = self.alphabet
        if sequence.alphabet != abc or background.alphabet != abc:
            raise plan7.AlphabetMismatch(abc, sequence.alphabet if sequence.alphabet != abc else background.alphabet)
        codes = np.asarray(sequence.sequence, np.uint8)
        M, K = len(codes), abc.K
        if M < 1:
            raise ValueError("cannot build a model from an empty sequence")
        f32 = np.float32
        bgf = np.asarray(background.residue_frequencies, f32)
        Q = conditional_probabilities(abc, self.score_matrix, bgf)
        name = sequence.name.decode() if isinstance(sequence.name, bytes) else str(sequence.name)
        hmm = plan7.HMM(abc, M, name)
        # p7_Seqmodel (seqmodel.c:49): rows of P(b|a) as match emissions, background inserts, gap-open / -extend transitions
        hmm.match_emissions[1:] = Q[codes][:, :K].astype(f32)
        hmm.match_emissions[0, 0] = 1.0
        hmm.insert_emissions[:] = bgf[:K]
        t = hmm.transition_probabilities
        t[:, 0], t[:, 1], t[:, 2] = f32(1.0 - 2 * self.popen), f32(self.popen), f32(self.popen)
        t[:, 3], t[:, 4], t[:, 5], t[:, 6] = f32(1.0 - self.pextend), f32(self.pextend), f32(1.0 - self.pextend), f32(self.pextend)
        t[M, 0], t[M, 2], t[M, 5], t[M, 6] = f32(1.0 - self.popen), 0.0, 1.0, 0.0
        hmm.nseq, hmm.command_line, hmm.creation_time = 1, "[HMM created from a query sequence]", time.asctime()
        hmm.set_composition()
        hmm.set_consensus(sequence)
        self.calibrate(hmm, background)
        if K == 4:
            if self.window_length:
                hmm.max_length = int(self.window_length)
            elif self.window_beta == 0.0:
                hmm.max_length = hmm.M * 4
            else:
                hmm.max_length = hmm.compute_max_length(self.window_beta)
        profile = plan7.Profile(M, abc).configure(hmm, background, self.EvL)
        return hmm, profile, profile.to_optimized()

    def build_msa(self, msa, background):
        """(HMM, Profile, OptimizedProfile) from a `DigitalMSA` (p7_Builder, p7_builder.c:415; see `msabuild`)."""
        from .msabuild import build_msa
        return build_msa(self, msa, background)

    def copy(self):
        b = Builder.__new__(Builder)
        b.__dict__.update(self.__dict__)
        b.randomness = FastRandomness(self.seed)
        return b

    # -- p7_Calibrate (evalues.c:58) ------------------------------------------------------------------------------------
    def _gpu_scores(self, om, seqs, which):
        from . import _lib, plan7, easel
        ctx = _lib.context()
        block = easel.DigitalSequenceBlock(self.alphabet, [easel.DigitalSequence(self.alphabet, name="r%d" % i, sequence=s) for i, s in enumerate(seqs)])
        db = plan7.SequenceDatabase(ctx, block)
        n = len(seqs)
        sc, st, n1 = np.empty(n, np.float32), np.empty(n, np.int32), np.empty(n, np.float32)
        h = om._device(ctx)
        _lib.check(_lib.lib.b2h_null_scores(ctx.handle, h, db.handle, _lib.ptr(n1), None), "b2h_null_scores", ctx.handle)
        fn = {"msv": _lib.lib.b2h_msv_filter, "vit": _lib.lib.b2h_viterbi_filter, "fwd": _lib.lib.b2h_forward_parser}[which]
        _lib.check(fn(ctx.handle, h, db.handle, _lib.ptr(sc), _lib.ptr(st)), "calibration filter", ctx.handle)
        return sc, n1

    def calibrate(self, hmm, background):
        """E-value parameters of ``hmm`` in place: lambda from the mean match relative entropy, the MSV and Viterbi Gumbel
        locations and the Forward exponential tail offset from random sequences (p7_Lambda, p7_MSVMu, p7_ViterbiMu, p7_Tau)."""
        from . import plan7
        abc, K = self.alphabet, self.alphabet.K
        bgf = np.asarray(background.residue_frequencies, np.float32)
        r = self.randomness
        if self.seed != 0:
            r.reinit()                                    # do_reseeding: the same random sequences for every model
        LOG2 = 0.69314718055994529
        lam = LOG2 + 1.44 / (float(hmm.M) * hmm.mean_match_relative_entropy(background))     # p7_Lambda
        om = plan7.Profile(hmm.M, abc).configure(hmm, background, self.EvL).to_optimized()
        scorer = self._scorer or self._gpu_scores
        mus = []
        for which, L, N, maxsc in (("msv", self.EmL, self.EmN, (255 - om.base) / om.scale_b),
                                   ("vit", self.EvL, self.EvN, (32767.0 - om.base_w) / om.scale_w)):
            seqs = [r.iid(bgf[:K], L) for _ in range(N)]
            sc, n1 = scorer(om, seqs, which)
            x = [float(np.float32((np.float32(maxsc) if math.isinf(v) else v) - n)) / LOG2 for v, n in zip(sc, n1)]
            esum = 0.0
            for v in x:
                esum += math.exp(-lam * v)
            mus.append(-math.log(esum / len(x)) / lam)    # esl_gumbel_FitCompleteLoc
        seqs = [r.iid(bgf[:K], self.EfL) for _ in range(self.EfN)]
        sc, n1 = scorer(om, seqs, "fwd")
        x = np.array([float(np.float32(v - n)) / LOG2 for v, n in zip(sc, n1)], np.float64)
        gmu, glam = _gumbel_fit_complete(x)
        tau = (gmu - math.log(-1.0 * math.log(1.0 - self.Eft)) / glam) + math.log(self.Eft) / lam
        hmm._evparam[:] = np.array([mus[0], lam, mus[1], lam, tau, lam], np.float32)
        return hmm


def _gumbel_fit_complete(x):
    """esl_gumbel_FitComplete (esl_gumbel.c:398): ML (mu, lambda) by Newton/Raphson on Lawless 4.1.6, tolerance 1e-5."""
    n = len(x)
    mean = x.sum() / n
    var = ((x - mean) ** 2).sum() / (n - 1)
    lam = math.pi / math.sqrt(6.0 * var)
    for _ in range(100):
        e = np.exp(-lam * x)
        esum, xesum, xxesum, xsum = e.sum(), (x * e).sum(), (x * x * e).sum(), x.sum()
        fx = 1.0 / lam - xsum / n + xesum / esum
        dfx = (xesum / esum) ** 2 - xxesum / esum - 1.0 / (lam * lam)
        if abs(fx) < 1e-5:
            break
        lam = lam - fx / dfx
        if lam <= 0.0:
            lam = 0.001
    esum = np.exp(-lam * x).sum()
    mu = -math.log(esum / n) / lam
    return mu, lam
