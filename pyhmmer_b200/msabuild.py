"""Model construction from multiple alignments: what ``pyhmmer.plan7.Builder.build_msa`` does (plan7.pyx:1018-1119 ->
p7_Builder, vendor/hmmer/src/p7_builder.c:415) for alignment queries and for every jackhmmer iteration after the first.

The steps and their reference functions, in p7_Builder's order:

  relative weights        esl_msaweight_PB_adv (vendor/easel/esl_msaweight.c:183): position-based weights over consensus columns
  fragment marking        esl_msa_MarkFragments_old (esl_msa.c): external gaps of short rows become missing data
  architecture + counts   p7_Fastmodelmaker / p7_Handmodelmaker (build.c:155, 81) -> matassign2hmm: faux traces
                          (p7_trace_FauxFromMSA, p7_trace.c:1277), p7_trace_Doctor (:1366), p7_trace_Count (:1454)
  effective seq number    p7_EntropyWeight (eweight.c:61): bisection on the mean match relative entropy
  parameters              p7_ParameterEstimation (p7_prior.c:278) with the mixture Dirichlet priors of p7_prior_CreateAmino /
                          CreateNucleic / CreateLaplace; posterior means by esl_mixdchlet_MPParameters (esl_mixdchlet.c)
  annotation              name / accession / description, composition, consensus, cutoffs, alignment map, checksum
  calibration             p7_Calibrate (the single-sequence builder's `Builder.calibrate`, GPU filters)

Host code by nature (the reference's Builder is CPU code that pyhmmer uses as is): numpy, with the
reference's float / double mix kept where it decides printed digits.
"""
import math

import numpy as np

f32, f64 = np.float32, np.float64
LOG2R = 1.44269504088896341
ETARGET = {"amino": 0.59, "DNA": 0.62, "RNA": 0.62}          # p7_ETARGET_AMINO / _DNA (hmmer.h), p7_ETARGET_OTHER = 1.0

# -- priors (p7_prior.c:39-275) ------------------------------------------------------------------------------------------
_AMINO_MQ = [0.178091, 0.056591, 0.0960191, 0.0781233, 0.0834977, 0.0904123, 0.114468, 0.0682132, 0.234585]
_AMINO_M = """
0.270671 0.039848 0.017576 0.016415 0.014268 0.131916 0.012391 0.022599 0.020358 0.030727 0.015315 0.048298 0.053803 0.020662 0.023612 0.216147 0.147226 0.065438 0.003758 0.009621
0.021465 0.010300 0.011741 0.010883 0.385651 0.016416 0.076196 0.035329 0.013921 0.093517 0.022034 0.028593 0.013086 0.023011 0.018866 0.029156 0.018153 0.036100 0.071770 0.419641
0.561459 0.045448 0.438366 0.764167 0.087364 0.259114 0.214940 0.145928 0.762204 0.247320 0.118662 0.441564 0.174822 0.530840 0.465529 0.583402 0.445586 0.227050 0.029510 0.121090
0.070143 0.011140 0.019479 0.094657 0.013162 0.048038 0.077000 0.032939 0.576639 0.072293 0.028240 0.080372 0.037661 0.185037 0.506783 0.073732 0.071587 0.042532 0.011254 0.028723
0.041103 0.014794 0.005610 0.010216 0.153602 0.007797 0.007175 0.299635 0.010849 0.999446 0.210189 0.006127 0.013021 0.019798 0.014509 0.012049 0.035799 0.180085 0.012744 0.026466
0.115607 0.037381 0.012414 0.018179 0.051778 0.017255 0.004911 0.796882 0.017074 0.285858 0.075811 0.014548 0.015092 0.011382 0.012696 0.027535 0.088333 0.944340 0.004373 0.016741
0.093461 0.004737 0.387252 0.347841 0.010822 0.105877 0.049776 0.014963 0.094276 0.027761 0.010040 0.187869 0.050018 0.110039 0.038668 0.119471 0.065802 0.025430 0.003215 0.018742
0.452171 0.114613 0.062460 0.115702 0.284246 0.140204 0.100358 0.550230 0.143995 0.700649 0.276580 0.118569 0.097470 0.126673 0.143634 0.278983 0.358482 0.661750 0.061533 0.199373
0.005193 0.004039 0.006722 0.006121 0.003468 0.016931 0.003647 0.002184 0.005019 0.005990 0.001473 0.004158 0.009055 0.003630 0.006583 0.003172 0.003690 0.002967 0.002772 0.002686
"""
_AMINO_EI = [681., 120., 623., 651., 313., 902., 241., 371., 687., 676., 143., 548., 647., 415., 551., 926., 623., 505., 102., 269.]


class Prior:
    """``P7_PRIOR``: five mixture Dirichlets (q[Q], alpha[Q][K]) -- match / insert / delete transitions, match / insert emissions."""

    def __init__(self, tm, ti, td, em, ei):
        mk = lambda q, a: (np.asarray(q, f64), np.atleast_2d(np.asarray(a, f64)))
        self.tm, self.ti, self.td, self.em, self.ei = mk(*tm), mk(*ti), mk(*td), mk(*em), mk(*ei)

    @classmethod
    def amino(cls):
        return cls(([1.0], [0.7939, 0.0278, 0.0135]), ([1.0], [0.1551, 0.1331]), ([1.0], [0.9002, 0.5630]),
                   (_AMINO_MQ, np.array(_AMINO_M.split(), f64).reshape(9, 20)), ([1.0], _AMINO_EI))

    @classmethod
    def nucleic(cls):
        return cls(([1.0], [2.0, 0.1, 0.1]), ([1.0], [0.12, 0.4]), ([1.0], [0.5, 1.0]),
                   ([0.24, 0.26, 0.08, 0.42], [[0.16, 0.45, 0.12, 0.39], [0.09, 0.03, 0.09, 0.04], [1.29, 0.40, 6.58, 0.51], [1.74, 1.49, 1.57, 1.95]]),
                   ([1.0], [1.0] * 4))

    @classmethod
    def laplace(cls, K):
        return cls(([1.0], [1.0] * 3), ([1.0], [1.0] * 2), ([1.0], [1.0] * 2), ([1.0], [1.0] * K), ([1.0], [1.0] * K))

    @classmethod
    def for_alphabet(cls, alphabet, scheme="alphabet"):
        if scheme is None:
            return None
        if scheme == "laplace":
            return cls.laplace(alphabet.K)
        if scheme != "alphabet":
            raise ValueError("invalid prior_scheme %r (expected 'laplace', 'alphabet' or None)" % (scheme,))
        return cls.amino() if alphabet.is_amino() else (cls.nucleic() if alphabet.is_nucleotide() else cls.laplace(alphabet.K))


_LG_COF = (4.694580336184385e+04, -1.560605207784446e+05, 2.065049568014106e+05, -1.388934775095388e+05, 5.031796415085709e+04,
           -9.601592329182778e+03, 8.785855930895250e+02, -3.155153906098611e+01, 2.908143421162229e-01, -2.319827630494973e-04,
           1.251639670050933e-10)


def log_gamma(x):
    """esl_stats_LogGamma (esl_stats.c): Lanczos' approximation, the same terms in the same order (element-wise on arrays)."""
    x = np.asarray(x, f64)
    xx = x - 1.0
    tx = xx + 11.0
    tmp = tx.copy()
    value = np.ones_like(x)
    for i in range(10, -1, -1):
        value = value + _LG_COF[i] / tmp
        tmp = tmp - 1.0
    value = np.log(value)
    tx = tx + 0.5
    return value + (0.918938533 + (xx + 0.5) * np.log(tx) - tx)


def _seqsum(a):
    """Left-to-right double sum along the last axis (numpy's own reduction is pairwise)."""
    return np.cumsum(a, axis=-1)[..., -1]


def _kahan(a):
    """esl_vec_DSum: compensated summation along the last axis."""
    a = np.asarray(a, f64)
    s = np.zeros(a.shape[:-1], f64)
    c = np.zeros(a.shape[:-1], f64)
    for i in range(a.shape[-1]):
        y = a[..., i] - c
        t = s + y
        c = (t - s) - y
        s = t
    return s


def mp_parameters(dchl, c):
    """esl_mixdchlet_MPParameters for a stack of count vectors c[n][K]: posterior component probabilities
    (mixdchlet_postq: log q + esl_dirichlet_logpdf_c, esl_vec_DLogNorm), then the mean posterior estimate, normalised."""
    q, alpha = dchl
    c = np.asarray(c, f64)
    n, K = c.shape
    Q = len(q)
    if Q == 1:
        postq = np.ones((n, 1), f64)                     # exp(x - x) normalised: exactly 1
    else:
        ca = c[:, None, :] + alpha[None, :, :]           # [n, Q, K]
        terms = log_gamma(ca) - log_gamma(c + 1.0)[:, None, :] - log_gamma(alpha)[None, :, :]
        logp = _seqsum(terms)
        sum1, sum2, sum3 = _seqsum(ca), _seqsum(alpha), _seqsum(c)
        logp = logp + (log_gamma(sum2)[None, :] + log_gamma(sum3 + 1.0)[:, None] - log_gamma(sum1))
        postq = np.log(q)[None, :] + logp
        mx = postq.max(axis=1, keepdims=True)
        e = np.where(postq > mx - 500.0, np.exp(postq - mx), 0.0)
        denom = np.log(_seqsum(e)) + mx[:, 0]
        postq = np.exp(postq - denom[:, None])
        postq = postq / _kahan(postq)[:, None]
    totc = _kahan(c)
    p = np.zeros((n, K), f64)
    for k in range(Q):
        totalpha = float(_kahan(alpha[k]))
        p = p + postq[:, k, None] * (c + alpha[k][None, :]) / (totc + totalpha)[:, None]
    s = _kahan(p)
    return np.where(s[:, None] != 0.0, p / np.where(s == 0.0, 1.0, s)[:, None], 1.0 / K)


def _fnorm(v):
    """esl_vec_FNorm on the rows of a float32 array."""
    v = np.asarray(v, f32)
    s = np.zeros(v.shape[:-1], f32)                      # esl_vec_FSum: compensated, in float
    c = np.zeros(v.shape[:-1], f32)
    for i in range(v.shape[-1]):
        y = v[..., i] - c
        t = s + y
        c = (t - s) - y
        s = t
    out = v / np.where(s == 0, f32(1.0), s)[..., None]
    out[s == 0] = f32(1.0 / v.shape[-1])
    return out.astype(f32)


def parameter_estimation(t, mat, ins, prior):
    """p7_ParameterEstimation (p7_prior.c:278): counts -> probabilities, in place on float32 arrays t[M+1][7] (MM MI MD IM II
    DM DD), mat[M+1][K], ins[M+1][K]."""
    M = t.shape[0] - 1
    if prior is None:                                    # p7_hmm_Renormalize
        mat[:], ins[:] = _fnorm(mat), _fnorm(ins)
        t[:, 0:3], t[:, 3:5], t[:, 5:7] = _fnorm(t[:, 0:3]), _fnorm(t[:, 3:5]), _fnorm(t[:, 5:7])
        t[M, 5], t[M, 6] = 1.0, 0.0
        if t[M, 2] > 0:
            t[M, 0], t[M, 1], t[M, 2] = 0.5, 0.5, 0.0
        return
    t[:, 0:3] = mp_parameters(prior.tm, t[:, 0:3]).astype(f32)
    t[M, 2] = 0.0
    t[M, 0:3] = _fnorm(t[M:M + 1, 0:3])[0]
    t[:, 3:5] = mp_parameters(prior.ti, t[:, 3:5]).astype(f32)
    if M > 1:
        t[1:M, 5:7] = mp_parameters(prior.td, t[1:M, 5:7]).astype(f32)
    t[0, 5] = t[M, 5] = 1.0
    t[0, 6] = t[M, 6] = 0.0
    mat[1:] = mp_parameters(prior.em, mat[1:]).astype(f32)
    mat[0] = 0.0
    mat[0, 0] = 1.0
    ins[:] = mp_parameters(prior.ei, ins).astype(f32)


def mean_match_relative_entropy(mat, bgf):
    """p7_MeanMatchRelativeEntropy (modelstats.c:95): esl_vec_FRelEntropy per node is a FLOAT sum of p * log2(p / q) terms
    computed in double; the mean over nodes is a double."""
    p = np.asarray(mat[1:], f32)
    q = np.asarray(bgf, f32)[: p.shape[1]]
    with np.errstate(divide="ignore", invalid="ignore"):
        term = p.astype(f64) * np.log2((p / q[None, :]).astype(f64))
    kl = np.zeros(p.shape[0], f32)
    for a in range(p.shape[1]):
        kl = np.where(p[:, a] > 0, (kl.astype(f64) + term[:, a]).astype(f32), kl)
    KL = 0.0
    for v in kl:
        KL += float(v)
    return KL / float(p.shape[0])


def entropy_weight(t, mat, ins, nseq, bgf, prior, etarget):
    """p7_EntropyWeight (eweight.c:61): the effective sequence number at which the parameterised model's mean match
    relative entropy equals ``etarget`` -- esl_root_Bisection on [0, nseq] to an absolute tolerance of 0.01."""
    def fx(neff):
        scale = f32(neff / float(nseq))                  # p7_hmm_Scale: esl_vec_FScale takes a float
        t2, m2, i2 = t * scale, mat * scale, ins * scale
        parameter_estimation(t2, m2, i2, prior)
        return mean_match_relative_entropy(m2, bgf) - etarget

    neff = float(nseq)
    if not fx(neff) > 0.0:
        return neff
    xl, xr = 0.0, float(nseq)
    fl, fr = fx(xl), fx(xr)
    if fl * fr >= 0:
        raise ValueError("internal failure in entropy weighting algorithm")      # "xl,xr do not bracket a root"
    x = neff
    for _ in range(100):
        x = (xl + xr) / 2.0
        f = fx(x)
        xmag = 0.0 if (xl < 0.0 and xr > 0.0) else x
        if f == 0.0:
            break
        if (xr - xl) < 0.01 + 1e-12 * xmag:
            break
        if fl > 0.0:
            if f > 0.0:
                xl, fl = x, f
            else:
                xr, fr = x, f
        else:
            if f < 0.0:
                xl, fl = x, f
            else:
                xr, fr = x, f
    else:
        raise ValueError("internal failure in entropy weighting algorithm")      # eslENOHALT
    return x


# -- relative weights ----------------------------------------------------------------------------------------------------
def _is_residue(ax, K, Kp):
    return (ax < K) | ((ax > K) & (ax < Kp - 2))


def pb_weights(ax, K, Kp, rf=None, fragthresh=0.5, symfrac=0.5):
    """esl_msaweight_PB_adv: Henikoff position-based weights over the consensus columns (the RF columns when given, else the
    columns with less than ``symfrac`` gaps among the counted symbols), external gaps of fragments not counted, each
    weight divided by the row's residue count, the set normalised to sum to nseq.  Doubles.  (The reference determines the
    consensus of alignments deeper than 50 000 rows from a random sample of 10 000; here all rows are always used.)"""
    nseq, alen = ax.shape
    if nseq == 1:
        return np.ones(1, f64)
    isres = _is_residue(ax, K, Kp)
    pos = np.arange(1, alen + 1)
    anyres = isres.any(axis=1)
    lpos = np.where(anyres, isres.argmax(axis=1) + 1, alen + 1)
    rpos = np.where(anyres, alen - isres[:, ::-1].argmax(axis=1), 0)
    minspan = int(math.ceil(float(f32(fragthresh) * f32(alen))))
    full = (rpos - lpos + 1) >= minspan
    lo = np.where(full, 1, lpos)
    hi = np.where(full, alen, rpos)
    counted = (pos[None, :] >= lo[:, None]) & (pos[None, :] <= hi[:, None])
    ct = np.zeros((alen, Kp), np.int64)
    for a in range(Kp):
        ct[:, a] = ((ax == a) & counted).sum(axis=0)
    if rf is not None:
        cons = np.array([c not in "-_." for c in rf], bool)
    else:
        tot = ct[:, : Kp - 2].sum(axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            cons = (ct[:, K].astype(f32) / tot.astype(f32)) < f32(symfrac)
    if not cons.any():
        cons[:] = True
    cols = np.nonzero(cons)[0]
    if rf is not None:                                    # collect_counts with known consensus columns counts only those; same numbers
        pass
    r = (ct[cols][:, :K] > 0).sum(axis=1)                 # distinct canonical residues per consensus column
    sub = ax[:, cols]
    canon = sub < K
    cnt = ct[cols[None, :], np.where(canon, sub, 0)]
    with np.errstate(divide="ignore"):
        contrib = np.where(canon, 1.0 / (r[None, :] * cnt).astype(f64), 0.0)
    wgt = _seqsum(contrib) if len(cols) else np.zeros(nseq, f64)
    rlen = canon.sum(axis=1)
    wgt = np.where(rlen > 0, wgt / np.where(rlen > 0, rlen, 1), wgt)
    s = float(_kahan(wgt))
    wgt = wgt / s if s != 0.0 else np.full(nseq, 1.0 / nseq)
    return wgt * float(nseq)


def mark_fragments(ax, K, Kp, fragthresh):
    """esl_msa_MarkFragments_old: rows with at most fragthresh * alen residues get their leading and trailing non-residue
    columns turned into missing data (in place)."""
    nseq, alen = ax.shape
    isres = _is_residue(ax, K, Kp)
    rlen = isres.sum(axis=1)
    for i in np.nonzero(rlen <= float(fragthresh) * alen)[0]:
        if rlen[i] == 0:
            ax[i, :] = Kp - 1
            continue
        first = int(isres[i].argmax())
        last = alen - 1 - int(isres[i, ::-1].argmax())
        ax[i, :first] = Kp - 1
        ax[i, last + 1:] = Kp - 1


# -- architecture and counts ---------------------------------------------------------------------------------------------
def fast_matassign(ax, wgt, K, Kp, symfrac):
    """p7_Fastmodelmaker's column rule: weighted residue fraction among residues + gaps >= symfrac (float accumulators)."""
    nseq, alen = ax.shape
    isres, isgap = _is_residue(ax, K, Kp), ax == K
    r = np.zeros(alen, f32)
    tot = np.zeros(alen, f32)
    for i in range(nseq):
        w = float(wgt[i])
        r = np.where(isres[i], (r.astype(f64) + w).astype(f32), r)
        tot = np.where(isres[i] | isgap[i], (tot.astype(f64) + w).astype(f32), tot)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (r > 0) & ((r / tot) >= f32(symfrac))


_B, _M, _D, _I, _X, _E = 0, 1, 2, 3, 4, 5


def _faux_trace(row, matassign, kcol, K, Kp):
    """p7_trace_FauxFromMSA + p7_trace_Doctor for one row: arrays (state, node, column index 0-based) between B and E."""
    res = (row < K) | ((row > K) & (row < Kp - 1))         # residues and '*' ("treat * as a residue")
    gap = row == K
    miss = row == Kp - 1
    st = np.where(matassign, np.where(res, _M, np.where(gap, _D, _X)), np.where(res, _I, np.where(miss, _X, -1)))
    keep = st >= 0
    st, k, col = st[keep], kcol[keep], np.nonzero(keep)[0]
    if len(st) > 1:                                       # "allow only one X in a row"
        dup = np.concatenate([[False], (st[1:] == _X) & (st[:-1] == _X)])
        if dup.any():
            st, k, col = st[~dup], k[~dup], col[~dup]
    # Doctor: D,I -> M (node of D, residue of I); I,D -> M (node of D, residue of I); left to right, pairs do not overlap
    pair = np.nonzero(((st[:-1] == _D) & (st[1:] == _I)) | ((st[:-1] == _I) & (st[1:] == _D)))[0] if len(st) > 1 else ()
    if len(pair):
        drop = np.zeros(len(st), bool)
        z_done = -1
        for z in pair:
            if z <= z_done:
                continue
            if st[z] == _D:
                col[z] = col[z + 1]
            else:
                k[z] = k[z + 1]
            st[z] = _M
            drop[z + 1] = True
            z_done = z + 1
        st, k, col = st[~drop], k[~drop], col[~drop]
    return st, k, col


def count_traces(ax, wgt, matassign, alphabet):
    """matassign2hmm's counting loop: weighted observed counts t[M+1][7], mat[M+1][K], ins[M+1][K] (float32, accumulated row
    by row like p7_trace_Count; degenerate residues spread over what they stand for, esl_abc_FCount)."""
    K, Kp = alphabet.K, alphabet.Kp
    nseq, alen = ax.shape
    M = int(matassign.sum())
    kcol = np.cumsum(matassign)                            # node of a match column / node an insert column follows
    t = np.zeros((M + 1, 7), f32)
    mat = np.zeros((M + 1, K), f32)
    ins = np.zeros((M + 1, K), f32)
    ndegen = alphabet.degen.sum(axis=1)
    # transition slot by (state, next state); E counts as M for the last node
    TR = {(_M, _M): 0, (_M, _I): 1, (_M, _D): 2, (_M, _E): 0, (_I, _M): 3, (_I, _I): 4, (_I, _E): 3, (_D, _M): 5, (_D, _D): 6, (_D, _E): 5}
    trlut = np.full((6, 6), -1, np.int64)
    for (a, b), v in TR.items():
        trlut[a, b] = v
    for idx in range(nseq):
        st, k, col = _faux_trace(ax[idx], matassign, kcol, K, Kp)
        wt = f32(wgt[idx])
        # full trace: B, ..., E
        S = np.concatenate([[_B], st, [_E]])
        Kk = np.concatenate([[0], k, [0]])
        C = np.concatenate([[0], col, [0]])
        N = len(S)
        z1, z2 = 0, N - 1
        if N > 1 and S[1] == _X:
            mpos = np.nonzero(S[2:N - 1] == _M)[0]
            if len(mpos):
                z1 = int(mpos[0]) + 2
        if N > 1 and S[N - 2] == _X:
            mpos = np.nonzero(S[1:N - 2] == _M)[0]
            if len(mpos):
                z2 = int(mpos[-1]) + 1
        if z2 <= z1:
            continue
        zs = np.arange(z1, z2)
        s1, s2, k1, k2 = S[zs], S[zs + 1], Kk[zs], Kk[zs + 1]
        ok = s1 != _X
        # emissions
        em = ok & ((s1 == _M) | (s1 == _I))
        x = ax[idx][C[zs]]
        for arr, sel in ((mat, em & (s1 == _M)), (ins, em & (s1 == _I))):
            canon = sel & (x < K)
            if canon.any():
                np.add.at(arr, (k1[canon], x[canon]), wt)
            deg = sel & (x > K) & (x < Kp - 2)
            for z in np.nonzero(deg)[0]:
                y = np.nonzero(alphabet.degen[x[z]])[0]
                arr[k1[z], y] += f32(wt / f32(ndegen[x[z]]))
        # transitions
        tr = ok & (s2 != _X)
        if z1 == 0 and tr[0]:                               # from B
            tr[0] = False
            if s2[0] == _M and k2[0] > 1:                   # wing-retracted B -> D..D -> Mk entry
                t[0, 2] += wt
                for kt in range(1, k2[0] - 1):
                    t[kt, 6] += wt
                t[k2[0] - 1, 5] += wt
            elif s2[0] == _M:
                t[0, 0] += wt
            elif s2[0] == _I:
                t[0, 1] += wt
            elif s2[0] == _D:
                t[0, 2] += wt
            elif s2[0] != _E:
                raise ValueError("bad transition in trace")
        slot = trlut[s1, s2]
        if (slot[tr] < 0).any():
            raise ValueError("bad transition in trace")
        np.add.at(t, (k1[tr], slot[tr]), wt)
    return t, mat, ins


def msa_checksum(ax):
    """esl_msa_Checksum over the digital rows (Jenkins' one-at-a-time hash, 32 bits)."""
    val = 0
    m = 0xffffffff
    for x in np.asarray(ax, np.uint8).ravel().tolist():
        val = (val + x) & m
        val = (val + (val << 10)) & m
        val ^= val >> 6
    val = (val + (val << 3)) & m
    val ^= val >> 11
    val = (val + (val << 15)) & m
    return val


def build_counts(msa, builder):
    """Everything of p7_Builder up to the weighted count model: returns (t, mat, ins, matassign, checksum).  Rewrites the
    alignment as the reference does -- sequence weights, fragment marks, the RF line."""
    abc = msa.alphabet
    K, Kp = abc.K, abc.Kp
    ax = msa.ax
    nseq, alen = ax.shape
    if nseq < 1 or alen < 1:
        raise ValueError("Could not build HMM: empty alignment")
    # validate_msa: missing-data symbols only at the edges of a row
    miss = ax == Kp - 1
    for i in np.nonzero(miss.any(axis=1))[0]:
        inner = np.nonzero(~miss[i])[0]
        if len(inner) and miss[i, inner[0]:inner[-1] + 1].any():
            raise ValueError("Could not build HMM: msa %s; sequence %s\nhas missing data chars (~) other than at fragment edges"
                             % (msa.name, msa.names[i]))
    checksum = msa_checksum(ax)
    hand = builder.architecture == "hand"
    if hand and msa.reference is None:
        raise ValueError("Could not build HMM: Alignment %s has no reference annotation line\n" % (msa.name or ""))
    if builder.weighting == "pb":
        msa.sequence_weights = pb_weights(ax, K, Kp, rf=msa.reference if hand else None)
    elif builder.weighting == "none":
        msa.sequence_weights = np.ones(nseq, f64)
    elif builder.weighting == "given":
        if msa.sequence_weights is None:
            msa.sequence_weights = np.ones(nseq, f64)
    else:
        raise NotImplementedError("relative weighting scheme %r" % (builder.weighting,))
    wgt = np.asarray(msa.sequence_weights, f64)
    mark_fragments(ax, K, Kp, builder.fragthresh)
    if hand:
        matassign = np.array([c not in "-_." for c in msa.reference], bool)
    else:
        matassign = fast_matassign(ax, wgt, K, Kp, builder.symfrac)
    if msa.model_mask is not None:                         # do_modelmask: masked columns count as the any-residue
        mm = np.array([c == "m" for c in msa.model_mask], bool)
        sel = mm[None, :] & (ax != K) & (ax != Kp - 1)
        ax[sel] = Kp - 3
    if not matassign.any():
        if hand:
            raise ValueError("Could not build HMM: Alignment %s has no annotated consensus columns - can't build a model.\n" % (msa.name or ""))
        raise ValueError("Could not build HMM: Alignment %s has no consensus columns w/ > %d%% residues - can't build a model.\n"
                         % (msa.name or "", int(100 * builder.symfrac)))
    t, mat, ins = count_traces(ax, wgt, matassign, abc)
    return t, mat, ins, matassign, checksum


def build_msa(builder, msa, background):
    """``Builder.build_msa``: (HMM, Profile, OptimizedProfile) from a digital alignment."""
    import time
    from . import plan7
    abc = builder.alphabet
    if background.alphabet != abc:
        raise plan7.AlphabetMismatch(abc, background.alphabet)
    if msa.alphabet != abc:
        raise plan7.AlphabetMismatch(abc, msa.alphabet)
    K = abc.K
    old_rf, old_mm = msa.reference, msa.model_mask
    t, mat, ins, matassign, checksum = build_counts(msa, builder)
    nseq = msa.ax.shape[0]
    M = t.shape[0] - 1
    bgf = np.asarray(background.residue_frequencies, f32)
    # effective sequence number
    eff = builder.effective_number
    if isinstance(eff, (int, float)) and not isinstance(eff, bool):
        neff = float(eff)
    elif eff == "none":
        neff = float(nseq)
    elif eff == "entropy":
        etarget = (builder.esigma - LOG2R * math.log(2.0 / (float(M) * float(M + 1)))) / float(M)
        etarget = max(builder.re_target, etarget)
        neff = entropy_weight(t, mat, ins, nseq, bgf, builder.prior, etarget)
    else:
        raise NotImplementedError("effective sequence number strategy %r" % (eff,))
    scale = f32(neff / float(nseq))
    t *= scale
    mat *= scale
    ins *= scale
    parameter_estimation(t, mat, ins, builder.prior)
    if not msa.name:
        raise ValueError("Could not build HMM: Unable to name the HMM.")
    txt = lambda v: None if v is None else (v.decode() if isinstance(v, (bytes, bytearray)) else str(v))
    hmm = plan7.HMM(abc, M, txt(msa.name))
    hmm.transition_probabilities[:] = t
    hmm.match_emissions[:] = mat
    hmm.insert_emissions[:] = ins
    hmm.accession = txt(msa.accession) or None
    hmm.description = txt(msa.description) or None
    hmm.nseq, hmm.nseq_effective = nseq, neff
    pick = lambda s: None if s is None else "".join(c for c, m in zip(s, matassign) if m)
    hmm.reference = pick(old_rf)
    if old_mm is not None:
        hmm.model_mask = "".join(("-" if c == "." else c) for c, m in zip(old_mm, matassign) if m)
    hmm.consensus_structure = pick(msa.secondary_structure)
    hmm.map = np.concatenate([[0], np.nonzero(matassign)[0] + 1]).astype(np.int64)
    msa.reference = "".join("x" if m else "." for m in matassign)     # "Reset #=RF line of alignment to reflect our assignment"
    hmm.creation_time = time.asctime()
    hmm.command_line = None
    hmm.set_composition()
    hmm.set_consensus(None)
    for tag, i in (("GA", 0), ("TC", 2), ("NC", 4)):
        c = (msa.cutoffs or {}).get(tag)
        if c is not None and c[0] is not None:
            hmm._cutoff[i] = c[0]
            if len(c) > 1 and c[1] is not None:
                hmm._cutoff[i + 1] = c[1]
    builder.calibrate(hmm, background)
    if hmm.model_mask is not None:                         # "force masked positions to background" (k = 1 .. M-1, as the reference loops)
        for k in range(1, M):
            if hmm.model_mask[k - 1] == "m":
                hmm.match_emissions[k, :K] = bgf[:K]
    if K == 4:
        if builder.window_length:
            hmm.max_length = int(builder.window_length)
        elif builder.window_beta == 0.0:
            hmm.max_length = M * 4
        else:
            hmm.max_length = hmm.compute_max_length(builder.window_beta)
    hmm.checksum = checksum
    profile = plan7.Profile(M, abc).configure(hmm, background, builder.EvL)
    return hmm, profile, profile.to_optimized()
