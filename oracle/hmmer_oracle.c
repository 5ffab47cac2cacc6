/* oracle/hmmer_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked or loaded by pyhmmer_b200/).
 *
 * A plain scalar C restatement of the reference's filter cascade kernels on one comparison,
 * written from the algorithm descriptions in SURVEY.md Appendix A and the reference sources
 * cited at each function.  No SIMD, no striping, nothing from the reference is linked: it is
 * the "port" half of the oracle.  It is PINNED: tests/test_oracle_cpu.py checks it bit-for-bit
 * (integer filters) / to 1e-4 nats (Forward) against oracle/_ref (the reference's own C code
 * compiled from its sources) and against the committed golden vectors produced by the reference
 * Python package (tests/golden/filters.json).
 *
 * Inputs are the node-major optimized-profile tables of include/b2h.h (node k at index k-1;
 * transition rows BM MM IM DM MD MI II DD), i.e. exactly what the CUDA kernels consume.
 *
 *   dsq[0..L-1] residue codes;  Kp residue rows of length M in each emission table.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define OK        0
#define ERANGE_  16
#define ENORESULT 19

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* p7_SSVFilter (impl_sse/ssvfilter.c:876-926), evaluated in wide integers.
 * cost[x*M + k-1] are the biased uint8 match costs (rbv).  Cells are kept relative to the constant
 * begin score, v(i,k) = sat(v(i-1,k-1) - sbv(x_i,k)) with sbv = min(cost - bias, 127), floor 0. */
int oracle_ssv(const uint8_t *dsq, int L, int M, const uint8_t *cost,
               int tbm, int tec, int tjb, int base, int bias, float scale_b, float *ret_sc)
{
  int *row = calloc((size_t)M + 1, sizeof(int)), *nxt = calloc((size_t)M + 1, sizeof(int));
  int i, k, maxw = 0, xE, xJ;
  *ret_sc = 0.f;
  if (tjb + tbm + tec + bias >= 127) { free(row); free(nxt); return ENORESULT; }
  for (i = 0; i < L; i++) {
    const uint8_t *c = cost + (size_t)dsq[i] * M;
    nxt[0] = 0;
    for (k = 1; k <= M; k++) {
      int s = -imin((int)c[k-1] - bias, 127);
      int v = imax(row[k-1] + s, 0);
      nxt[k] = v;
      if (v > maxw) maxw = v;
    }
    { int *t = row; row = nxt; nxt = t; }
  }
  free(row); free(nxt);
  if (maxw >= 127 - bias) { *ret_sc = INFINITY; return (base - tjb - tbm < 128) ? ENORESULT : ERANGE_; }
  xE = maxw + base - tjb - tbm;
  if (xE >= 255 - bias) { *ret_sc = INFINITY; return ERANGE_; }
  xJ = xE - tec;
  if (xJ > base) return ENORESULT;
  *ret_sc = ((float)(xJ - tjb) - (float)base);
  *ret_sc /= scale_b;
  *ret_sc -= 3.0f;
  return OK;
}

/* p7_MSVFilter's own recurrence (impl_sse/msvfilter.c:106-207): saturating uint8 arithmetic written out. */
int oracle_msv_full(const uint8_t *dsq, int L, int M, const uint8_t *cost,
                    int tbm, int tec, int tjb, int base, int bias, float scale_b, float *ret_sc)
{
  int *row = calloc((size_t)M + 1, sizeof(int)), *nxt = calloc((size_t)M + 1, sizeof(int));
  const int tjbm = (tjb + tbm) & 0xff;
  int i, k, xJ = 0, xB = imax(base - tjbm, 0);
  for (i = 0; i < L; i++) {
    const uint8_t *c = cost + (size_t)dsq[i] * M;
    int xE = 0;
    nxt[0] = 0;
    for (k = 1; k <= M; k++) {
      int sv = imax(row[k-1], xB);
      sv = imin(sv + bias, 255);
      sv = imax(sv - (int)c[k-1], 0);
      nxt[k] = sv;
      if (sv > xE) xE = sv;
    }
    if (xE + bias >= 255) { free(row); free(nxt); *ret_sc = INFINITY; return ERANGE_; }
    xE = imax(xE - tec, 0);
    xJ = imax(xJ, xE);
    xB = imax(imax(base, xJ) - tjbm, 0);
    { int *t = row; row = nxt; nxt = t; }
  }
  free(row); free(nxt);
  *ret_sc = ((float)(xJ - tjb) - (float)base);
  *ret_sc /= scale_b;
  *ret_sc -= 3.0f;
  return OK;
}

/* p7_MSVFilter (msvfilter.c:74-104): SSV first, the full recurrence only on eslENORESULT. */
int oracle_msv(const uint8_t *dsq, int L, int M, const uint8_t *cost,
               int tbm, int tec, int tjb, int base, int bias, float scale_b, float *ret_sc)
{
  int st = oracle_ssv(dsq, L, M, cost, tbm, tec, tjb, base, bias, scale_b, ret_sc);
  if (st != ENORESULT) return st;
  return oracle_msv_full(dsq, L, M, cost, tbm, tec, tjb, base, bias, scale_b, ret_sc);
}

/* p7_ViterbiFilter (impl_sse/vitfilter.c:83-248), scalar, full D->D closure every row (SURVEY A.4:
 * identical integers to the lazy-F evaluation).  rsc[x*M+k-1], tsc[t*M+k-1]. */
static int sat16(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }
int oracle_vit(const uint8_t *dsq, int L, int M, const int16_t *rsc, const int16_t *tsc,
               int xw_E_move, int xw_E_loop, int xw_move, int base_w, float scale_w, float *ret_sc)
{
  const int16_t *tBM = tsc, *tMM = tsc + M, *tIM = tsc + 2*M, *tDM = tsc + 3*M, *tMD = tsc + 4*M, *tMI = tsc + 5*M, *tII = tsc + 6*M, *tDD = tsc + 7*M;
  size_t n = (size_t)M + 2;
  int *Mp = malloc(n * sizeof(int)), *Ip = malloc(n * sizeof(int)), *Dp = malloc(n * sizeof(int));
  int *Mc = malloc(n * sizeof(int)), *Ic = malloc(n * sizeof(int)), *Dc = malloc(n * sizeof(int));
  int i, k, xN = base_w, xB = (int16_t)(xN + xw_move), xJ = -32768, xC = -32768, st = OK;
  for (k = 0; k <= M + 1; k++) Mp[k] = Ip[k] = Dp[k] = -32768;
  for (i = 0; i < L; i++) {
    const int16_t *r = rsc + (size_t)dsq[i] * M;
    int xE = -32768;
    Mc[0] = Ic[0] = Dc[0] = -32768; Dc[1] = -32768;
    for (k = 1; k <= M; k++) {
      int m = sat16(xB + tBM[k-1]);
      m = imax(m, sat16(Mp[k-1] + tMM[k-1]));
      m = imax(m, sat16(Ip[k-1] + tIM[k-1]));
      m = imax(m, sat16(Dp[k-1] + tDM[k-1]));
      m = sat16(m + r[k-1]);
      Mc[k] = m;
      if (m > xE) xE = m;
      Ic[k] = imax(sat16(Mp[k] + tMI[k-1]), sat16(Ip[k] + tII[k-1]));
      Dc[k+1] = sat16(m + tMD[k-1]);                      /* M->D partial for node k+1 (index M+1 is a spare) */
    }
    for (k = 2; k <= M; k++) Dc[k] = imax(Dc[k], sat16(Dc[k-1] + tDD[k-2]));
    if (xE >= 32767) { *ret_sc = INFINITY; st = ERANGE_; goto done; }
    xC = (int16_t)imax(xC, xE + xw_E_move);
    xJ = (int16_t)imax(xJ, xE + xw_E_loop);
    xB = (int16_t)imax(xJ + xw_move, xN + xw_move);
    { int *t; t = Mp; Mp = Mc; Mc = t; t = Ip; Ip = Ic; Ic = t; t = Dp; Dp = Dc; Dc = t; }
  }
  if (xC > -32768) { *ret_sc = (float)xC + (float)xw_move - (float)base_w; *ret_sc /= scale_w; *ret_sc -= 3.0f; }
  else *ret_sc = -INFINITY;
done:
  free(Mp); free(Ip); free(Dp); free(Mc); free(Ic); free(Dc);
  return st;
}

/* p7_ForwardParser (impl_sse/fwdback.c:256-463): odds space, sparse rescaling when xE > 1e4. */
int oracle_fwd(const uint8_t *dsq, int L, int M, const float *rsc, const float *tsc,
               float xf_E_move, float xf_E_loop, float pmove, float *ret_sc)
{
  const float *tBM = tsc, *tMM = tsc + M, *tIM = tsc + 2*M, *tDM = tsc + 3*M, *tMD = tsc + 4*M, *tMI = tsc + 5*M, *tII = tsc + 6*M, *tDD = tsc + 7*M;
  size_t n = (size_t)M + 2;
  float *Mp = calloc(n, sizeof(float)), *Ip = calloc(n, sizeof(float)), *Dp = calloc(n, sizeof(float));
  float *Mc = calloc(n, sizeof(float)), *Ic = calloc(n, sizeof(float)), *Dc = calloc(n, sizeof(float));
  const float ploop = 1.0f - pmove;
  float xN = 1.f, xJ = 0.f, xC = 0.f, xB = pmove, xE, totscale = 0.f;
  int i, k, st = OK;
  for (i = 0; i < L; i++) {
    const float *r = rsc + (size_t)dsq[i] * M;
    double e = 0.0;
    Dc[1] = 0.f;
    for (k = 1; k <= M; k++) {
      float m = xB * tBM[k-1];
      m += Mp[k-1] * tMM[k-1]; m += Ip[k-1] * tIM[k-1]; m += Dp[k-1] * tDM[k-1];
      m *= r[k-1];
      Mc[k] = m;
      Ic[k] = Mp[k] * tMI[k-1] + Ip[k] * tII[k-1];
      if (k < M) Dc[k+1] = m * tMD[k-1] + Dc[k] * tDD[k-1];
      e += m; e += Dc[k];
    }
    xE = (float)e;
    xN = xN * ploop;
    xC = xC * ploop + xE * xf_E_move;
    xJ = xJ * ploop + xE * xf_E_loop;
    xB = xJ * pmove + xN * pmove;
    if (xE > 1.0e4f) {
      const float inv = 1.0f / xE;
      xN /= xE; xC /= xE; xJ /= xE; xB /= xE;
      for (k = 1; k <= M; k++) { Mc[k] *= inv; Ic[k] *= inv; Dc[k] *= inv; }
      totscale = (float)((double)totscale + log((double)xE));
    }
    { float *t; t = Mp; Mp = Mc; Mc = t; t = Ip; Ip = Ic; Ic = t; t = Dp; Dp = Dc; Dc = t; }
  }
  if (isnan(xC) || (L > 0 && xC == 0.0f) || isinf(xC)) { st = ERANGE_; *ret_sc = 0.f; }
  else *ret_sc = (float)((double)totscale + log((double)(xC * pmove)));
  free(Mp); free(Ip); free(Dp); free(Mc); free(Ic); free(Dc);
  return st;
}

/* p7_bg_NullOne (p7_bg.c:357) and p7_bg_FilterScore -> esl_hmm_Forward (p7_bg.c:471, esl_hmm.c:353) */
float oracle_null1(int L)
{
  const float p1 = (float)L / (float)(L + 1);
  return (float)((double)(float)L * log((double)p1) + log(1. - (double)p1));
}
float oracle_bias(const uint8_t *dsq, int L, int M, const float *eo /* [Kp][2] */)
{
  const float p1 = (float)L / (float)(L + 1);
  const float L1 = (float)((double)(float)M / 8.0);
  const float t00 = p1, t01 = 1.0f - p1, t10 = 1.0f / (L1 + 1.0f), t11 = L1 / (L1 + 1.0f);
  float d0, d1, mx, logsc = 0.0f;
  int i;
  d0 = eo[dsq[0]*2] * 0.999f; d1 = eo[dsq[0]*2+1] * 0.001f;
  mx = 0.0f; if (d0 > mx) mx = d0; if (d1 > mx) mx = d1;
  d0 /= mx; d1 /= mx; logsc += (float)log((double)mx);
  for (i = 1; i < L; i++) {
    float n0 = 0.0f, n1 = 0.0f;
    n0 += d0 * t00; n0 += d1 * t10; n1 += d0 * t01; n1 += d1 * t11;
    n0 *= eo[dsq[i]*2]; n1 *= eo[dsq[i]*2+1];
    mx = 0.0f; if (n0 > mx) mx = n0; if (n1 > mx) mx = n1;
    d0 = n0 / mx; d1 = n1 / mx; logsc += (float)log((double)mx);
  }
  { float last = 0.0f; last += d0 * 1.0f; last += d1 * 1.0f; logsc += (float)log((double)last); }
  return logsc + (float)L * logf(p1) + logf((float)(1. - (double)p1));
}

/* p7_SSVFilter_longtarget (impl_sse/msvfilter.c:256-413), one chunk, scalar: the MSV recurrence at a constant begin score
 * with the byte saturations written out; when a cell of a row reaches sc_thresh the best cell is taken with the reference's
 * striped tie-break (Q = p7O_NQB(M) vectors of 16: scan order q outer, lane z inner, node k = q + Q*z + 1), the diagonal is
 * walked back to the begin level and extended while it keeps improving (5 non-improving steps end it), the window
 * {n, k, length, score} is recorded, the row is zeroed and the scan resumes behind the diagonal.
 * dsq[0..L-1]; cost[x*M + k-1]; tjb for the model's max_length.  win [cap][3] = n (1-based), k, length.  Returns the count. */
int oracle_ssv_longtarget(const uint8_t *dsq, int L, int M, const uint8_t *cost, int tbm, int tec, int tjb, int base, int bias,
                          float scale_b, int sc_thresh, int cap, int64_t *win, float *win_sc)
{
  int *row = calloc((size_t)M + 1, sizeof(int)), *nxt = calloc((size_t)M + 1, sizeof(int));
  const int Q = imax(2, (M - 1) / 16 + 1);
  const int tjbm = (tjb + tbm) & 0xff;
  const int xB = imax(base - tjbm, 0);
  int i, k, q, z, nwin = 0;
  (void)tec;
  for (i = 1; i <= L; i++) {
    const uint8_t *c = cost + (size_t)dsq[i-1] * M;
    int hit = 0;
    nxt[0] = 0;
    for (k = 1; k <= M; k++) {
      int sv = imax(row[k-1], xB);
      sv = imin(sv + bias, 255);
      sv = imax(sv - (int)c[k-1], 0);
      nxt[k] = sv;
      if (sv >= sc_thresh) hit = 1;
    }
    { int *t = row; row = nxt; nxt = t; }
    if (hit) {
      int end = -1, rem_sc = -1, start, target_start, target_end, sc, n, max_end, max_sc, pos_since_max;
      float ret_sc;
      for (q = 0; q < Q; q++)
        for (z = 0; z < 16; z++) {
          k = q + Q * z + 1;
          if (k <= M && row[k] >= sc_thresh && row[k] > rem_sc) { end = k; rem_sc = row[k]; }
        }
      for (k = 0; k <= M; k++) row[k] = 0;
      start = end; target_end = target_start = i; sc = rem_sc;
      while (rem_sc > base - tjb - tbm) {
        rem_sc -= bias - (int)cost[(size_t)dsq[target_start-1] * M + (start-1)];
        --start; --target_start;
      }
      start++; target_start++;
      k = end + 1; n = target_end + 1; max_end = target_end; max_sc = sc; pos_since_max = 0;
      while (k < M && n <= L) {
        sc += bias - (int)cost[(size_t)dsq[n-1] * M + (k-1)];
        if (sc >= max_sc) { max_sc = sc; max_end = n; pos_since_max = 0; }
        else if (++pos_since_max == 5) break;
        k++; n++;
      }
      end += max_end - target_end;
      target_end = max_end;
      ret_sc = ((float)(max_sc - tjb) - (float)base);
      ret_sc /= scale_b;
      ret_sc -= 3.0f;
      if (nwin < cap) { win[nwin*3+0] = target_start; win[nwin*3+1] = end; win[nwin*3+2] = end - start + 1; win_sc[nwin] = ret_sc; }
      nwin++;
      i = target_end;
    }
  }
  free(row); free(nxt);
  return nwin;
}

/* p7_ViterbiFilter_longtarget (impl_sse/vitfilter.c:292-497), one window, scalar.  The row recurrence of p7_ViterbiFilter
 * without its overflow test; when the row's best match cell xE reaches sc_thresh, every node k holding xE is recorded as a
 * landmark (i, k) in the reference's striped order (Q = p7O_NQW(M) vectors of 8: q outer, lane z inner, k = q + Q*z + 1), the
 * three rows are reset to -32768 and the special states keep their values from the previous row (:411-425).  Otherwise the
 * specials are updated and the D->D paths are closed only when Dmax + ddbound_w > xB (the lazy-F test, :450); when they are
 * not, the row keeps its M->D values alone (:483-487).  dsq[0..L-1]; rsc[x*M+k-1], tsc[t*M+k-1]; xw_move for the length the
 * caller configured (p7_oprofile_ReconfigRestLength).  hit [cap][2] = i (1-based), k.  Returns the number of landmarks. */
int oracle_vit_longtarget(const uint8_t *dsq, int L, int M, const int16_t *rsc, const int16_t *tsc,
                          int xw_E_move, int xw_E_loop, int xw_move, int base_w, int ddbound_w, int sc_thresh, int cap, int64_t *hit)
{
  const int16_t *tBM = tsc, *tMM = tsc + M, *tIM = tsc + 2*M, *tDM = tsc + 3*M, *tMD = tsc + 4*M, *tMI = tsc + 5*M, *tII = tsc + 6*M, *tDD = tsc + 7*M;
  const int Q = imax(2, (M - 1) / 8 + 1);
  size_t n = (size_t)M + 2;
  int *Mp = malloc(n * sizeof(int)), *Ip = malloc(n * sizeof(int)), *Dp = malloc(n * sizeof(int));
  int *Mc = malloc(n * sizeof(int)), *Ic = malloc(n * sizeof(int)), *Dc = malloc(n * sizeof(int));
  int i, k, q, z, nhit = 0, xN = base_w, xB = (int16_t)(xN + xw_move), xJ = -32768, xC = -32768;
  for (k = 0; k <= M + 1; k++) Mp[k] = Ip[k] = Dp[k] = -32768;
  for (i = 1; i <= L; i++) {
    const int16_t *r = rsc + (size_t)dsq[i-1] * M;
    int xE = -32768, Dmax = -32768;
    Mc[0] = Ic[0] = Dc[0] = -32768; Dc[1] = -32768;
    for (k = 1; k <= M; k++) {
      int m = sat16(xB + tBM[k-1]);
      m = imax(m, sat16(Mp[k-1] + tMM[k-1]));
      m = imax(m, sat16(Ip[k-1] + tIM[k-1]));
      m = imax(m, sat16(Dp[k-1] + tDM[k-1]));
      m = sat16(m + r[k-1]);
      Mc[k] = m;
      if (m > xE) xE = m;
      Ic[k] = imax(sat16(Mp[k] + tMI[k-1]), sat16(Ip[k] + tII[k-1]));
      Dc[k+1] = sat16(m + tMD[k-1]);
    }
    /* Dmaxv covers the M->D value of every striped cell, node M's (tMD = -32768 there) and the padding cells' included */
    for (k = 2; k <= M + 1; k++) if (Dc[k] > Dmax) Dmax = Dc[k];
    if (xE >= sc_thresh) {
      for (q = 0; q < Q; q++)
        for (z = 0; z < 8; z++) {
          k = q + Q * z + 1;
          if (k <= M && Mc[k] == xE) { if (nhit < cap) { hit[nhit*2] = i; hit[nhit*2+1] = k; } nhit++; }
        }
      for (k = 0; k <= M + 1; k++) Mc[k] = Ic[k] = Dc[k] = -32768;
    } else {
      xC = (int16_t)imax(xC, xE + xw_E_move);
      xJ = (int16_t)imax(xJ, xE + xw_E_loop);
      xB = (int16_t)imax(xJ + xw_move, xN + xw_move);
      if (Dmax + ddbound_w > xB)
        for (k = 2; k <= M; k++) Dc[k] = imax(Dc[k], sat16(Dc[k-1] + tDD[k-2]));
    }
    { int *t; t = Mp; Mp = Mc; Mc = t; t = Ip; Ip = Ic; Ic = t; t = Dp; Dp = Dc; Dc = t; }
  }
  free(Mp); free(Ip); free(Dp); free(Mc); free(Ic); free(Dc);
  return nhit;
}
