// b2h_domaindef.cpp -- host-side domain definition for the comparisons that passed the Forward filter.
//
// The GPU cascade hands over, per survivor, the Forward and Backward parser special-state rows.
// This file restates (in plain node-major scalar C++, nothing striped, nothing copied) what
// p7_Pipeline does after p7_BackwardParser (p7_pipeline.c:767-933):
//
//   p7_DomainDecoding (impl_sse/decoding.c:160)           -> btot / etot / mocc
//   region finding + is_multidomain_region (p7_domaindef.c:384-493, 531)
//   multidomain regions: multihit p7_Forward on the region, 200 stochastic tracebacks
//     (impl_sse/stotrace.c:71) with Easel's LCG stream (esl_random.c knuth()), p7_Null2_ByTrace
//     (impl_sse/null2.c:131), single-linkage clustering of segment pairs (p7_spensemble.c:281)
//   every envelope: unihit p7_Forward / p7_Backward / p7_Decoding (fwdback.c, decoding.c:76),
//     p7_OptimalAccuracy + p7_OATrace (impl_sse/optacc.c:58, 225), p7_Null2_ByExpectation (null2.c:44),
//     alignment display (p7_alidisplay.c:92)
//   per-sequence and per-domain scores, null2 corrections, P-values (p7_pipeline.c:776-933)
//
// Survivors are ~1e-5 of all comparisons; this is "row 11-12, host orchestration" of SURVEY 8(a).
// Moving the envelope DP to the GPU is SURVEY 8(f) rank 1 (next round).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#if defined(__SSE2__)
#include <xmmintrin.h>
#include <pmmintrin.h>
#endif
#include "b2h_internal.h"
#include "b2h_domaindef.h"

namespace {

const float NEGINF = -std::numeric_limits<float>::infinity();
enum { XE = 0, XN = 1, XJ = 2, XB = 3, XC = 4, XSC = 5, NX = 6 };
enum { sM = 0, sD = 1, sI = 2 };
enum { T_BM = 0, T_MM, T_IM, T_DM, T_MD, T_MI, T_II, T_DD };
enum { ST_M = 1, ST_D, ST_I, ST_S, ST_N, ST_B, ST_E, ST_C, ST_T, ST_J };   // p7t_statetype_e values are irrelevant; only identity matters

// p7_FLogsum's lookup table (logsum.c:58-113)
struct Logsum {
  float tbl[16000];
  Logsum() { for (int i = 0; i < 16000; i++) tbl[i] = (float)log(1. + exp((double)-i / 1000.f)); }
  float operator()(float a, float b) const {
    const float mx = a > b ? a : b, mn = a > b ? b : a;
    return (mn == NEGINF || (mx - mn) >= 15.7f) ? mx : mx + tbl[(int)((mx - mn) * 1000.f)];
  }
};
const Logsum &flogsum() { static Logsum L; return L; }

// esl_vec_FSum: Kahan summation (esl_vectorops.c)
float kahan_sum(const float *v, int n) {
  float sum = 0.f, c = 0.f;
  for (int i = 0; i < n; i++) { float y = v[i] - c; float t = sum + y; c = (t - sum) - y; sum = t; }
  return sum;
}
void fnorm(float *v, int n) {
  float s = kahan_sum(v, n);
  if (s != 0.0f) for (int i = 0; i < n; i++) v[i] /= s;
  else           for (int i = 0; i < n; i++) v[i] = (float)(1. / (float)n);
}

// Easel's "fast" generator as the pipeline creates it (esl_randomness_CreateFast; esl_random.c knuth())
struct FastRng {
  uint32_t seed, x;
  static uint32_t mix3(uint32_t a, uint32_t b, uint32_t c) {
    a -= b; a -= c; a ^= (c >> 13); b -= c; b -= a; b ^= (a << 8);  c -= a; c -= b; c ^= (b >> 13);
    a -= b; a -= c; a ^= (c >> 12); b -= c; b -= a; b ^= (a << 16); c -= a; c -= b; c ^= (b >> 5);
    a -= b; a -= c; a ^= (c >> 3);  b -= c; b -= a; b ^= (a << 10); c -= a; c -= b; c ^= (b >> 15);
    return c;
  }
  void init(uint32_t s) { seed = s; x = mix3(s, 87654321u, 12345678u); if (x == 0) x = 42; }
  double next() { x *= 69069u; x += 1u; return (double)x / 4294967296.0; }
  int fchoose(const float *p, int n) {
    double norm = 0.0, sum = 0.0; const double roll = next();
    for (int i = 0; i < n; i++) norm += p[i];
    for (int i = 0; i < n; i++) { sum += (double)p[i]; if (roll < (sum / norm)) return i; }
    return n - 1;
  }
};

struct Model {               // host view of one profile in a given uni/multihit + length configuration
  int M, K, Kp;
  const float *rsc;          // [Kp][M]
  const float *t[8];         // node-major transition rows, index k-1
  const float *zrow;         // M zeros: the emission row of residue codes outside the table
  const float *pf_up, *pf_dn;// blocked prefix products of tDD for the D->D chains (index n = node, see chain_up / chain_down)
  float eM, eL;              // xf[E][MOVE], xf[E][LOOP]
  float pmove, ploop;        // xf[N|C|J][MOVE|LOOP]
  const uint8_t *degen;
  float r(int x, int k) const { return (x < Kp) ? rsc[(size_t)x * M + (k - 1)] : 0.0f; }
  const float *rrow(int x) const { return (x < Kp) ? rsc + (size_t)x * M : zrow; }      // index k-1
};

void configure(Model &m, bool multihit, int L) {
  const float nj = multihit ? 1.0f : 0.0f;
  m.eM = multihit ? 0.5f : 1.0f; m.eL = multihit ? 0.5f : 0.0f;
  m.pmove = (2.0f + nj) / ((float)L + 2.0f + nj);
  m.ploop = 1.0f - m.pmove;
}

// Full DP matrix: rows 0..L; every row is three planes {M, D, I} of S floats (node k at index k, node 0 and node M+1
// are guards), so that the recurrences are unit-stride loops the compiler vectorises.  Specials per row in x.
struct Mx {
  int M = 0, L = 0; size_t S = 0;
  std::vector<float> dp, x, rs;          // rs: per-row factor of a posterior matrix (see backward_decode)
  float totscale = 0.f; bool own_scales = false;
  // grow-only; only row 0 and the two guard columns (node 0 and node M+1) of every row are cleared: every other
  // cell that a recurrence reads has been written by the same pass before
  void resize(int M_, int L_, bool with_cells = true) {
    M = M_; L = L_; S = ((size_t)M + 2 + 15) & ~(size_t)15;
    const size_t need = with_cells ? (size_t)(L + 1) * S * 3 : 0, needx = (size_t)(L + 1) * NX;
    if (dp.size() < need) dp.resize(need);
    if (x.size() < needx) x.resize(needx);
    if (with_cells) {
      std::fill(dp.begin(), dp.begin() + S * 3, 0.0f);
      for (int i = 1; i <= L; i++) { float *r = row(i); for (int s = 0; s < 3; s++) { r[s * S] = 0.0f; r[s * S + M + 1] = 0.0f; } }
    }
    std::fill(x.begin(), x.begin() + needx, 0.0f);
  }
  float *row(int i) { return dp.data() + (size_t)i * S * 3; }
  const float *row(int i) const { return dp.data() + (size_t)i * S * 3; }
  float &X(int i, int s) { return x[(size_t)i * NX + s]; }
  float X(int i, int s) const { return x[(size_t)i * NX + s]; }
};
#define C3(rowp, k, s) (rowp)[(size_t)(s) * SS + (size_t)(k)]     /* SS = the matrix' plane stride, in scope */

// ---- D->D chains.  y[n+1] += t[n] * y[n] is a serial mul+add chain (8 cycles per node); cut into blocks of CB
// nodes it becomes (1) independent local chains per block, which the core overlaps, and (2) one vectorisable pass
// adding the block's incoming value times the prefix products of t (precomputed per model). ----
constexpr int CB = 16;
// The O(M*L) host loops below are unit-stride and written for the compiler's vectoriser.  B2H_SIMD_CLONES compiles each of them
// twice -- an AVX2 clone and a baseline x86-64 clone -- and the dynamic loader picks one per CPU (GCC function multiversioning),
// so the library does not require AVX2 of its host.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define B2H_SIMD_CLONES __attribute__((target_clones("avx2", "default")))
#else
#define B2H_SIMD_CLONES
#endif

B2H_SIMD_CLONES
void chain_prefix(const float *tDD /* index k-1 */, int M, std::vector<float> &up, std::vector<float> &dn)
{
  up.assign(M + 2, 0.f); dn.assign(M + 2, 0.f);
  for (int n0 = 1; n0 <= M - 1; n0 += CB) {                       // forward: n = 1..M-1, t[n] = tDD[n-1]
    float p = 1.0f;
    for (int n = n0; n <= std::min(M - 1, n0 + CB - 1); n++) { p *= tDD[n - 1]; up[n] = p; }
  }
  for (int k0 = M; k0 >= 1; k0 -= CB) {                           // backward: k = M..1, t[k] = tDD[k-1]
    float p = 1.0f;
    for (int k = k0; k >= std::max(1, k0 - CB + 1); k--) { p *= tDD[k - 1]; dn[k] = p; }
  }
}
// y[n+1] = y[n+1] + t[n]*y[n], n = 1..N ascending (y[1] final on entry); t[n] = tDD[n-1]
inline void chain_up(float *__restrict__ y, const float *__restrict__ tDD, const float *__restrict__ pf, int N)
{
  for (int n0 = 1; n0 <= N; n0 += CB) {
    const int n1 = std::min(N, n0 + CB - 1);
    for (int n = n0 + 1; n <= n1; n++) y[n + 1] += tDD[n - 1] * y[n];
  }
  for (int n0 = 1; n0 <= N; n0 += CB) {
    const int n1 = std::min(N, n0 + CB - 1);
    const float yin = y[n0];
    if (yin == 0.0f) continue;
#pragma omp simd
    for (int n = n0; n <= n1; n++) y[n + 1] += pf[n] * yin;
  }
}
// y[k] = y[k] + t[k]*y[k+1], k = N..1 descending (y[N+1] final on entry); t[k] = tDD[k-1]
inline void chain_down(float *__restrict__ y, const float *__restrict__ tDD, const float *__restrict__ pf, int N)
{
  for (int k0 = N; k0 >= 1; k0 -= CB) {
    const int k1 = std::max(1, k0 - CB + 1);
    for (int k = k0 - 1; k >= k1; k--) y[k] += tDD[k - 1] * y[k + 1];
  }
  for (int k0 = N; k0 >= 1; k0 -= CB) {
    const int k1 = std::max(1, k0 - CB + 1);
    const float yin = y[k0 + 1];
    if (yin == 0.0f) continue;
#pragma omp simd
    for (int k = k1; k <= k0; k++) y[k] += pf[k] * yin;
  }
}

// forward_engine, do_full (fwdback.c:256-463).  dsq[0..L-1].  Returns false on eslERANGE.
B2H_SIMD_CLONES
bool forward_full(const Model &m, const uint8_t *dsq, int L, Mx &ox, float *sc)
{
  const int M = m.M;
  ox.resize(M, L);
  const size_t SS = ox.S;
  const float *__restrict__ tBM = m.t[T_BM], *__restrict__ tMM = m.t[T_MM], *__restrict__ tIM = m.t[T_IM], *__restrict__ tDM = m.t[T_DM],
              *__restrict__ tMD = m.t[T_MD], *__restrict__ tMI = m.t[T_MI], *__restrict__ tII = m.t[T_II];
  float xE = 0.f, xN = 1.f, xJ = 0.f, xB = m.pmove, xC = 0.f;
  ox.X(0, XE) = 0.f; ox.X(0, XN) = 1.f; ox.X(0, XJ) = 0.f; ox.X(0, XB) = xB; ox.X(0, XC) = 0.f; ox.X(0, XSC) = 1.f;
  ox.totscale = 0.f;
  for (int i = 1; i <= L; i++) {
    const float *pp = ox.row(i - 1); float *cp = ox.row(i);
    const float *__restrict__ pM = pp + sM * SS, *__restrict__ pD = pp + sD * SS, *__restrict__ pI = pp + sI * SS;
    float *__restrict__ cM = cp + sM * SS, *__restrict__ cD = cp + sD * SS, *__restrict__ cI = cp + sI * SS;
    const float *__restrict__ rs = m.rrow(dsq[i - 1]);
    cD[1] = 0.f;
#pragma omp simd
    for (int c = 0; c < M; c++) {                              // node k = c + 1
      float v = xB * tBM[c];
      v += pM[c] * tMM[c];
      v += pI[c] * tIM[c];
      v += pD[c] * tDM[c];
      v *= rs[c];
      cM[c + 1] = v;
      cI[c + 1] = pM[c + 1] * tMI[c] + pI[c + 1] * tII[c];
      cD[c + 2] = v * tMD[c];                                  // M(k) -> D(k+1); the D->D term is added by chain_up
    }
    cD[M + 1] = 0.f;                                           // guard column
    chain_up(cD, m.t[T_DD], m.pf_up, M - 1);
    float em = 0.f, ed = 0.f;
#pragma omp simd reduction(+:em,ed)
    for (int k = 1; k <= M; k++) { em += cM[k]; ed += cD[k]; }
    xE = em + ed;
    xN = xN * m.ploop;
    xC = (xC * m.ploop) + (xE * m.eM);
    xJ = (xJ * m.ploop) + (xE * m.eL);
    xB = (xJ * m.pmove) + (xN * m.pmove);
    if (xE > 1.0e4f) {
      xN = xN / xE; xC = xC / xE; xJ = xJ / xE; xB = xB / xE;
      const float inv = 1.0f / xE;
#pragma omp simd
      for (int k = 1; k <= M; k++) { cM[k] *= inv; cD[k] *= inv; cI[k] *= inv; }
      ox.X(i, XSC) = xE;
      ox.totscale = (float)((double)ox.totscale + log((double)xE));
      xE = 1.0f;
    } else ox.X(i, XSC) = 1.0f;
    ox.X(i, XE) = xE; ox.X(i, XN) = xN; ox.X(i, XJ) = xJ; ox.X(i, XB) = xB; ox.X(i, XC) = xC;
  }
  if (std::isnan(xC) || (L > 0 && xC == 0.0f) || std::isinf(xC)) return false;
  if (sc) *sc = (float)((double)ox.totscale + log((double)(xC * m.pmove)));
  return true;
}

// backward_engine, do_full (fwdback.c:468-733) fused with p7_Decoding (decoding.c:76-134): the Backward rows are
// consumed as they are produced (two rolling rows in <brow>), and what is stored is the posterior matrix
//   pp(i,k) = f(i,k) * b(i,k) * rs[i],   rs[i] = scaleproduct(i) * fwd scale(i)
// with the per-row factor kept apart (it is only known once Backward has reached row 0).  bk receives the Backward
// specials.  Returns false on eslERANGE from either routine.
B2H_SIMD_CLONES
bool backward_decode(const Model &m, const uint8_t *dsq, int L, const Mx &fwd, Mx &bk, Mx &pp, std::vector<float> &brow)
{
  const int M = m.M;
  bk.resize(M, L, false);
  pp.resize(M, L);
  const size_t SS = fwd.S;
  if (brow.size() < 6 * SS) brow.resize(6 * SS);
  std::fill(brow.begin(), brow.begin() + 6 * SS, 0.0f);
  const float *__restrict__ tBM = m.t[T_BM], *__restrict__ tMM = m.t[T_MM], *__restrict__ tIM = m.t[T_IM], *__restrict__ tDM = m.t[T_DM],
              *__restrict__ tMD = m.t[T_MD], *__restrict__ tMI = m.t[T_MI], *__restrict__ tII = m.t[T_II], *__restrict__ tDD = m.t[T_DD];
  bk.own_scales = false;
  float xJ = 0.f, xB = 0.f, xN = 0.f, xC = m.pmove, xE = xC * m.eM;
  auto emit_pp = [&](int i, const float *__restrict__ bM, const float *__restrict__ bI) {
    const float *fr = fwd.row(i); float *pr = pp.row(i);
    const float *__restrict__ fM = fr + sM * SS, *__restrict__ fI = fr + sI * SS;
    float *__restrict__ qM = pr + sM * SS, *__restrict__ qI = pr + sI * SS;
#pragma omp simd
    for (int k = 1; k <= M; k++) { qM[k] = fM[k] * bM[k]; qI[k] = fI[k] * bI[k]; }
  };
  {
    float *cp = brow.data() + (size_t)(L & 1) * 3 * SS;
    float *__restrict__ cM = cp + sM * SS, *__restrict__ cD = cp + sD * SS, *__restrict__ cI = cp + sI * SS;
    cD[M + 1] = 0.f;
#pragma omp simd
    for (int k = 1; k <= M; k++) { cD[k] = xE; cI[k] = 0.f; }
    chain_down(cD, tDD, m.pf_dn, M);
#pragma omp simd
    for (int k = 1; k <= M; k++) cM[k] = xE + tMD[k - 1] * cD[k + 1];
    const float scL = fwd.X(L, XSC);
    if (scL > 1.0f) {
      xE = xE / scL; xN = xN / scL; xC = xC / scL; xJ = xJ / scL; xB = xB / scL;
      const float inv = 1.0f / scL;
#pragma omp simd
      for (int k = 1; k <= M; k++) { cM[k] *= inv; cD[k] *= inv; cI[k] *= inv; }
    }
    bk.X(L, XSC) = scL;
    bk.totscale = (float)log((double)scL);
    bk.X(L, XE) = xE; bk.X(L, XN) = xN; bk.X(L, XJ) = xJ; bk.X(L, XB) = xB; bk.X(L, XC) = xC;
    if (L >= 1) emit_pp(L, cM, cI);
  }
  for (int i = L - 1; i >= 1; i--) {
    const float *pr = brow.data() + (size_t)((i + 1) & 1) * 3 * SS; float *cp = brow.data() + (size_t)(i & 1) * 3 * SS;
    const float *__restrict__ pM = pr + sM * SS, *__restrict__ pI = pr + sI * SS;
    float *__restrict__ cM = cp + sM * SS, *__restrict__ cD = cp + sD * SS, *__restrict__ cI = cp + sI * SS;
    const float *__restrict__ rs = m.rrow(dsq[i]);             // x_{i+1}, index k-1
    float bsum = 0.f;
    // mpv(k) = M(i+1,k+1) * e(k+1) (0 for k = M): node k+1 is index c+1 of the transition rows
#pragma omp simd reduction(+:bsum)
    for (int k = 1; k < M; k++) {
      const float mpv = pM[k + 1] * rs[k];
      const float ipv = pI[k];
      cI[k] = (ipv * tII[k - 1]) + (mpv * tIM[k]);
      cD[k] = mpv * tDM[k];
      cM[k] = (ipv * tMI[k - 1]) + (mpv * tMM[k]);
      bsum += (pM[k] * rs[k - 1]) * tBM[k - 1];
    }
    { const float ipv = pI[M];                                 // k = M: mpv = 0, and the transitions towards M+1 count as 0
      cI[M] = (ipv * tII[M - 1]) + (0.f * 0.f);
      cD[M] = 0.f * 0.f;
      cM[M] = (ipv * tMI[M - 1]) + (0.f * 0.f);
      bsum += (pM[M] * rs[M - 1]) * tBM[M - 1]; }
    xB = bsum;
    xC = xC * m.ploop;
    xJ = (xB * m.pmove) + (xJ * m.ploop);
    xN = (xB * m.pmove) + (xN * m.ploop);
    xE = (xC * m.eM) + (xJ * m.eL);
    cD[M + 1] = 0.f;
#pragma omp simd
    for (int k = 1; k <= M; k++) cD[k] = cD[k] + xE;
    chain_down(cD, tDD, m.pf_dn, M);
#pragma omp simd
    for (int k = 1; k <= M; k++) cM[k] = (cM[k] + xE) + tMD[k - 1] * cD[k + 1];
    if (xB > 1.0e16f) bk.own_scales = true;
    const float scale = bk.own_scales ? ((xB > 1.0e4f) ? xB : 1.0f) : fwd.X(i, XSC);
    bk.X(i, XSC) = scale;
    if (scale > 1.0f) {
      xE /= scale; xN /= scale; xJ /= scale; xB /= scale; xC /= scale;
      const float inv = 1.0f / scale;
#pragma omp simd
      for (int k = 1; k <= M; k++) { cM[k] *= inv; cD[k] *= inv; cI[k] *= inv; }
      bk.totscale = (float)((double)bk.totscale + log((double)scale));
    }
    bk.X(i, XE) = xE; bk.X(i, XN) = xN; bk.X(i, XJ) = xJ; bk.X(i, XB) = xB; bk.X(i, XC) = xC;
    emit_pp(i, cM, cI);
  }
  {
    float bsum = 0.f;
    if (L >= 1) {
      const float *pr = brow.data() + (size_t)(1 & 1) * 3 * SS; const float *__restrict__ pM = pr + sM * SS;
      const float *__restrict__ rs = m.rrow(dsq[0]);
#pragma omp simd reduction(+:bsum)
      for (int k = 1; k <= M; k++) bsum += (pM[k] * rs[k - 1]) * tBM[k - 1];
    }
    xB = bsum;
    xN = (xB * m.pmove) + (xN * m.ploop);
    bk.X(0, XB) = xB; bk.X(0, XC) = 0.f; bk.X(0, XJ) = 0.f; bk.X(0, XN) = xN; bk.X(0, XE) = 0.f; bk.X(0, XSC) = 1.f;
  }
  if (std::isnan(xN) || (L > 0 && xN == 0.0f) || std::isinf(xN)) { /* p7_Backward's status is ignored by the caller too */ }
  // p7_Decoding's row factors and special-state posteriors
  if (pp.rs.size() < (size_t)L + 1) pp.rs.resize((size_t)L + 1);
  float scaleproduct = 1.0f / bk.X(0, XN);
  pp.rs[0] = 0.f;
  for (int i = 1; i <= L; i++) {
    pp.rs[i] = scaleproduct * fwd.X(i, XSC);
    pp.X(i, XE) = 0.f;
    pp.X(i, XN) = fwd.X(i - 1, XN) * bk.X(i, XN) * m.ploop * scaleproduct;
    pp.X(i, XJ) = fwd.X(i - 1, XJ) * bk.X(i, XJ) * m.ploop * scaleproduct;
    pp.X(i, XC) = fwd.X(i - 1, XC) * bk.X(i, XC) * m.ploop * scaleproduct;
    pp.X(i, XB) = 0.f;
    if (bk.own_scales) scaleproduct *= fwd.X(i, XSC) / bk.X(i, XSC);
  }
  return !std::isinf(scaleproduct);
}

// p7_OptimalAccuracy (optacc.c:58-176).  "impossible transition" contributes 0.0 (the AND-mask trick), not -inf.
// The D(k) <- D(k-1) dependency is a max/mask chain: exact under any re-association, so it is blocked like the
// Forward chain (local chains per block, then the incoming value gated by the block's running AND of the masks).
B2H_SIMD_CLONES
float optimal_accuracy(const Model &m, const Mx &pp, Mx &ox, std::vector<float> &pmask)
{
  const int M = m.M, L = pp.L;
  ox.resize(M, L);
  const size_t SS = ox.S;
  const float *__restrict__ tBM = m.t[T_BM], *__restrict__ tMM = m.t[T_MM], *__restrict__ tIM = m.t[T_IM], *__restrict__ tDM = m.t[T_DM],
              *__restrict__ tMD = m.t[T_MD], *__restrict__ tMI = m.t[T_MI], *__restrict__ tII = m.t[T_II], *__restrict__ tDD = m.t[T_DD];
  // pm[k] = 1 if every D(j-1)->D(j) transition from the start of k's block up to k is possible (k = 2..M; tDD[j-2])
  if (pmask.size() < (size_t)M + 2) pmask.resize((size_t)M + 2);
  float *__restrict__ pm = pmask.data();
  for (int k0 = 2; k0 <= M; k0 += CB) { bool ok = true; for (int k = k0; k <= std::min(M, k0 + CB - 1); k++) { ok = ok && (tDD[k - 2] > 0.0f); pm[k] = ok ? 1.0f : 0.0f; } }
  { float *r0 = ox.row(0); for (int k = 0; k <= M + 1; k++) { C3(r0, k, sM) = NEGINF; C3(r0, k, sD) = NEGINF; C3(r0, k, sI) = NEGINF; } }
  ox.X(0, XE) = NEGINF; ox.X(0, XN) = 0.f; ox.X(0, XJ) = NEGINF; ox.X(0, XB) = 0.f; ox.X(0, XC) = NEGINF;
  for (int i = 1; i <= L; i++) {
    const float *pr = ox.row(i - 1); float *cr = ox.row(i); const float *ppr = pp.row(i);
    const float *__restrict__ pM = pr + sM * SS, *__restrict__ pD = pr + sD * SS, *__restrict__ pI = pr + sI * SS;
    float *__restrict__ cM = cr + sM * SS, *__restrict__ cD = cr + sD * SS, *__restrict__ cI = cr + sI * SS;
    const float *__restrict__ qM = ppr + sM * SS, *__restrict__ qI = ppr + sI * SS;
    const float xB = ox.X(i - 1, XB), ps = pp.rs[i];
    cM[0] = cD[0] = cI[0] = NEGINF;
    cD[1] = NEGINF;
    float xE = NEGINF;
#pragma omp simd reduction(max:xE)
    for (int c = 0; c < M; c++) {                              // node k = c + 1
      float sv = (tBM[c] > 0.0f) ? xB : 0.0f;
      sv = std::max(sv, (tMM[c] > 0.0f) ? pM[c] : 0.0f);
      sv = std::max(sv, (tIM[c] > 0.0f) ? pI[c] : 0.0f);
      sv = std::max(sv, (tDM[c] > 0.0f) ? pD[c] : 0.0f);
      sv = sv + qM[c + 1] * ps;
      xE = std::max(xE, sv);
      cM[c + 1] = sv;
      cD[c + 2] = (tMD[c] > 0.0f) ? sv : 0.0f;                 // value entering D(k+1) from M(k)
      float iv = (tMI[c] > 0.0f) ? pM[c + 1] : 0.0f;
      iv = std::max(iv, (tII[c] > 0.0f) ? pI[c + 1] : 0.0f);
      cI[c + 1] = iv + qI[c + 1] * ps;
    }
    // D(k) = max(D(k), gate(tDD[k-2], D(k-1))), k = 2..M
    for (int k0 = 2; k0 <= M; k0 += CB) {
      const int k1 = std::min(M, k0 + CB - 1);
      if (!(tDD[k0 - 2] > 0.0f)) cD[k0] = std::max(cD[k0], 0.0f);
      for (int k = k0 + 1; k <= k1; k++) cD[k] = std::max(cD[k], (tDD[k - 2] > 0.0f) ? cD[k - 1] : 0.0f);
    }
    for (int k0 = 2; k0 <= M; k0 += CB) {
      const int k1 = std::min(M, k0 + CB - 1);
      const float yin = cD[k0 - 1];
#pragma omp simd
      for (int k = k0; k <= k1; k++) cD[k] = std::max(cD[k], (pm[k] > 0.0f) ? yin : cD[k]);
    }
    cD[M + 1] = 0.f;
#pragma omp simd reduction(max:xE)
    for (int k = 1; k <= M; k++) xE = std::max(xE, cD[k]);
    ox.X(i, XE) = xE;
    float t1 = (m.ploop == 0.0f) ? 0.0f : ox.X(i - 1, XJ) + pp.X(i, XJ);
    float t2 = (m.eL == 0.0f) ? 0.0f : ox.X(i, XE);
    ox.X(i, XJ) = std::max(t1, t2);
    t1 = (m.ploop == 0.0f) ? 0.0f : ox.X(i - 1, XC) + pp.X(i, XC);
    t2 = (m.eM == 0.0f) ? 0.0f : ox.X(i, XE);
    ox.X(i, XC) = std::max(t1, t2);
    ox.X(i, XN) = (m.ploop == 0.0f) ? 0.0f : ox.X(i - 1, XN) + pp.X(i, XN);
    t1 = (m.pmove == 0.0f) ? 0.0f : ox.X(i, XN);
    t2 = (m.pmove == 0.0f) ? 0.0f : ox.X(i, XJ);
    ox.X(i, XB) = std::max(t1, t2);
  }
  return ox.X(L, XC);
}

struct Trace {
  std::vector<int8_t> st; std::vector<int> k, i; std::vector<float> pp;
  // domain index (p7_trace_Index)
  std::vector<int> tfrom, tto, sqfrom, sqto, hmmfrom, hmmto;
  void clear() { st.clear(); k.clear(); i.clear(); pp.clear(); tfrom.clear(); tto.clear(); sqfrom.clear(); sqto.clear(); hmmfrom.clear(); hmmto.clear(); }
  void push(int s, int kk, int ii, float p = 0.f) {
    // p7_trace_AppendWithPP: only emitting states keep their i, only main states their k (p7_trace.c)
    int ki = 0, ii2 = 0;
    switch (s) {
      case ST_N: case ST_C: case ST_J: ii2 = (!st.empty() && st.back() == s) ? ii : 0; break;
      case ST_M: ki = kk; ii2 = ii; break;
      case ST_I: ki = kk; ii2 = ii; break;
      case ST_D: ki = kk; break;
      default: break;
    }
    st.push_back((int8_t)s); k.push_back(ki); i.push_back(ii2); pp.push_back((ii2 > 0) ? p : 0.f);
  }
  void reverse() {
    // p7_trace_Reverse (p7_trace.c): traces built backwards hold C-,Cx,Cx; pull residues back by one before reversing
    for (size_t z = 0; z + 1 < st.size(); z++)
      if (st[z] == st[z + 1] && (st[z] == ST_N || st[z] == ST_C || st[z] == ST_J) && i[z] == 0 && i[z + 1] > 0) {
        i[z] = i[z + 1]; i[z + 1] = 0; pp[z] = pp[z + 1]; pp[z + 1] = 0.f;
      }
    std::reverse(st.begin(), st.end()); std::reverse(k.begin(), k.end()); std::reverse(i.begin(), i.end()); std::reverse(pp.begin(), pp.end()); }
  void index() {
    tfrom.clear(); tto.clear(); sqfrom.clear(); sqto.clear(); hmmfrom.clear(); hmmto.clear();
    for (size_t z = 0; z < st.size(); z++) {
      if (st[z] == ST_B) { tfrom.push_back((int)z); sqfrom.push_back(0); hmmfrom.push_back(0); tto.push_back(0); sqto.push_back(0); hmmto.push_back(0); }
      else if (st[z] == ST_M && !tfrom.empty()) {
        const size_t d = tfrom.size() - 1;
        if (sqfrom[d] == 0) sqfrom[d] = i[z];
        if (hmmfrom[d] == 0) hmmfrom[d] = k[z];
        sqto[d] = i[z]; hmmto[d] = k[z];
      } else if (st[z] == ST_E && !tfrom.empty()) tto[tfrom.size() - 1] = (int)z;
    }
  }
  int ndom() const { return (int)tfrom.size(); }
};
// NB on traceback bookkeeping: the reference appends (state, k, i) during the walk from T back to S where an
// N/C/J state "emits on transition": p7_trace_Append stores i only for the second and later of a run in
// traceback order; after p7_trace_Reverse the FIRST state of each run is the non-emitting one.  We build the
// trace directly in traceback order with the same rule and then reverse, exactly as the reference does.

inline int Qf(int M) { return std::max(2, (M - 1) / 4 + 1); }     // p7O_NQF

// p7_OATrace (optacc.c:225-268) with its select_* helpers
bool oa_trace(const Model &m, const Mx &pp, const Mx &ox, Trace &tr)
{
  const int M = m.M, L = ox.L, Q = Qf(M);
  const size_t SS = ox.S;
  int i = L, k = 0;
  tr.clear();
  tr.push(ST_T, k, i); tr.push(ST_C, k, i);
  int s0 = ST_C;
  auto path = [](float t, float v) -> float { return (t == 0.0f) ? NEGINF : v; };
  size_t guard = (size_t)(L + M + 8) * 4 + 64;
  while (s0 != ST_S) {
    if (guard-- == 0) return false;
    int s1 = -1;
    switch (s0) {
      case ST_M: {
        const float *pr = ox.row(i - 1); const int c = k - 1;
        float p[4] = { path(m.t[T_MM][c], C3(pr, k - 1, sM)), path(m.t[T_IM][c], C3(pr, k - 1, sI)),
                       path(m.t[T_DM][c], C3(pr, k - 1, sD)), path(m.t[T_BM][c], ox.X(i - 1, XB)) };
        if (k == 1) { p[0] = path(m.t[T_MM][c], 0.0f); p[1] = path(m.t[T_IM][c], 0.0f); p[2] = path(m.t[T_DM][c], 0.0f); }   // rightshiftz: zeros shift in
        static const int state[4] = { ST_M, ST_I, ST_D, ST_B };
        int best = 0; for (int j = 1; j < 4; j++) if (p[j] > p[best]) best = j;
        s1 = state[best]; k--; i--; break; }
      case ST_D: {
        const float *cr = ox.row(i); const int c = k - 1;
        const float p0 = (k > 1) ? path(m.t[T_MD][c - 1], C3(cr, k - 1, sM)) : NEGINF;
        const float p1 = (k > 1) ? path(m.t[T_DD][c - 1], C3(cr, k - 1, sD)) : NEGINF;
        s1 = (p0 >= p1) ? ST_M : ST_D; k--; break; }
      case ST_I: {
        const float *pr = ox.row(i - 1); const int c = k - 1;
        const float p0 = path(m.t[T_MI][c], C3(pr, k, sM)), p1 = path(m.t[T_II][c], C3(pr, k, sI));
        s1 = (p0 >= p1) ? ST_M : ST_I; i--; break; }
      case ST_N: s1 = (i == 0) ? ST_S : ST_N; break;
      case ST_C: {
        const float p0 = (m.ploop == 0.0f) ? NEGINF : ox.X(i - 1, XC) + pp.X(i, XC);
        const float p1 = (m.eM == 0.0f) ? NEGINF : ox.X(i, XE);
        s1 = (p0 > p1) ? ST_C : ST_E; break; }
      case ST_J: {
        const float p0 = (m.ploop == 0.0f) ? NEGINF : ox.X(i - 1, XJ) + pp.X(i, XJ);
        const float p1 = (m.eL == 0.0f) ? NEGINF : ox.X(i, XE);
        s1 = (p0 > p1) ? ST_J : ST_E; break; }
      case ST_E: {
        // the reference scans its striped row: q outer, lane r inner, M cells win ties (>=), D cells need > (optacc.c:404-421)
        const float *cr = ox.row(i);
        float mx = NEGINF; int smax = -1, kmax = 0;
        for (int q = 0; q < Q; q++) {
          for (int r = 0; r < 4; r++) { const int kk = r * Q + q + 1; if (kk <= M && C3(cr, kk, sM) >= mx) { mx = C3(cr, kk, sM); smax = ST_M; kmax = kk; } }
          for (int r = 0; r < 4; r++) { const int kk = r * Q + q + 1; if (kk <= M && C3(cr, kk, sD) >  mx) { mx = C3(cr, kk, sD); smax = ST_D; kmax = kk; } }
        }
        k = kmax; s1 = smax; break; }
      case ST_B: {
        const float p0 = (m.pmove == 0.0f) ? NEGINF : ox.X(i, XN);
        const float p1 = (m.pmove == 0.0f) ? NEGINF : ox.X(i, XJ);
        s1 = (p0 > p1) ? ST_N : ST_J; break; }
      default: return false;
    }
    if (s1 == -1) return false;
    float postprob = 0.0f;
    switch (s1) {
      case ST_M: postprob = C3(pp.row(i), k, sM) * pp.rs[i]; break;
      case ST_I: postprob = C3(pp.row(i), k, sI) * pp.rs[i]; break;
      case ST_N: if (s0 == s1) postprob = pp.X(i, XN); break;
      case ST_C: if (s0 == s1) postprob = pp.X(i, XC); break;
      case ST_J: if (s0 == s1) postprob = pp.X(i, XJ); break;
      default: break;
    }
    tr.push(s1, k, i, postprob);
    if ((s1 == ST_N || s1 == ST_J || s1 == ST_C) && s1 == s0) i--;
    s0 = s1;
  }
  tr.reverse();
  return true;
}

// p7_StochasticTrace (stotrace.c:71-113)
bool stochastic_trace(FastRng &rng, const Model &m, int L, const Mx &ox, Trace &tr)
{
  const int M = m.M, Q = Qf(M);
  const size_t SS = ox.S;
  int i = L, k = 0;
  tr.clear();
  tr.push(ST_T, k, i); tr.push(ST_C, k, i);
  int s0 = ST_C;
  size_t guard = (size_t)(L + M + 8) * 8 + 64;
  while (s0 != ST_S) {
    if (guard-- == 0) return false;
    int s1 = -1;
    switch (s0) {
      case ST_M: {
        const float *pr = ox.row(i - 1); const int c = k - 1;
        float p[4] = { ox.X(i - 1, XB) * m.t[T_BM][c], C3(pr, k - 1, sM) * m.t[T_MM][c], C3(pr, k - 1, sI) * m.t[T_IM][c], C3(pr, k - 1, sD) * m.t[T_DM][c] };
        if (k == 1) { p[1] = 0.0f * m.t[T_MM][c]; p[2] = 0.0f * m.t[T_IM][c]; p[3] = 0.0f * m.t[T_DM][c]; }
        static const int state[4] = { ST_B, ST_M, ST_I, ST_D };
        fnorm(p, 4); s1 = state[rng.fchoose(p, 4)]; k--; i--; break; }
      case ST_D: {
        const float *cr = ox.row(i); const int c = k - 1;
        float p[2] = { (k > 1) ? C3(cr, k - 1, sM) * m.t[T_MD][c - 1] : 0.0f, (k > 1) ? C3(cr, k - 1, sD) * m.t[T_DD][c - 1] : 0.0f };
        fnorm(p, 2); s1 = (rng.fchoose(p, 2) == 0) ? ST_M : ST_D; k--; break; }
      case ST_I: {
        const float *pr = ox.row(i - 1); const int c = k - 1;
        float p[2] = { C3(pr, k, sM) * m.t[T_MI][c], C3(pr, k, sI) * m.t[T_II][c] };
        fnorm(p, 2); s1 = (rng.fchoose(p, 2) == 0) ? ST_M : ST_I; i--; break; }
      case ST_N: s1 = (i == 0) ? ST_S : ST_N; break;
      case ST_C: {
        float p[2] = { ox.X(i - 1, XC) * m.ploop, ox.X(i, XE) * m.eM * ox.X(i, XSC) };
        fnorm(p, 2); s1 = (rng.fchoose(p, 2) == 0) ? ST_C : ST_E; break; }
      case ST_J: {
        float p[2] = { ox.X(i - 1, XJ) * m.ploop, ox.X(i, XE) * m.eL * ox.X(i, XSC) };
        fnorm(p, 2); s1 = (rng.fchoose(p, 2) == 0) ? ST_J : ST_E; break; }
      case ST_E: {
        double sum = 0.0; const double roll = rng.next();
        const float norm = (float)(1.0 / ox.X(i, XE));
        const float *cr = ox.row(i);
        bool found = false;
        for (int rep = 0; rep < 4 && !found; rep++)
          for (int q = 0; q < Q && !found; q++) {
            for (int r = 0; r < 4 && !found; r++) { const int kk = r * Q + q + 1; const float v = (kk <= M) ? C3(cr, kk, sM) * norm : 0.0f; sum += v; if (roll < sum) { k = kk; s1 = ST_M; found = true; } }
            for (int r = 0; r < 4 && !found; r++) { const int kk = r * Q + q + 1; const float v = (kk <= M) ? C3(cr, kk, sD) * norm : 0.0f; sum += v; if (roll < sum) { k = kk; s1 = ST_D; found = true; } }
          }
        if (!found) return false;
        break; }
      case ST_B: {
        float p[2] = { ox.X(i, XN) * m.pmove, ox.X(i, XJ) * m.pmove };
        fnorm(p, 2); s1 = (rng.fchoose(p, 2) == 0) ? ST_N : ST_J; break; }
      default: return false;
    }
    if (s1 == -1) return false;
    tr.push(s1, k, i);
    if ((s1 == ST_N || s1 == ST_J || s1 == ST_C) && s1 == s0) i--;
    s0 = s1;
  }
  tr.reverse();
  return true;
}

void avg_degenerate(const Model &m, float *null2)      // esl_abc_FAvgScVec + the three special codes
{
  for (int x = m.K + 1; x <= m.Kp - 3; x++) {
    float result = 0.f; int nd = 0;
    for (int y = 0; y < m.K; y++) if (m.degen && m.degen[(size_t)x * m.K + y]) { result += null2[y]; nd++; }
    null2[x] = nd ? result / (float)nd : 0.0f;
  }
  null2[m.K] = 1.0f; null2[m.Kp - 2] = 1.0f; null2[m.Kp - 1] = 1.0f;
}

// p7_Null2_ByExpectation (null2.c:44-110)
B2H_SIMD_CLONES
void null2_by_expectation(const Model &m, const Mx &pp, float *null2)
{
  const int M = m.M, Ld = pp.L;
  const size_t SS = pp.S;
  std::vector<float> em(M + 1, 0.f), ei(M + 1, 0.f);
  float xn = pp.X(1, XN), xc = pp.X(1, XC), xj = pp.X(1, XJ);
  { const float *r = pp.row(1); const float ps = pp.rs[1]; for (int k = 1; k <= M; k++) { em[k] = C3(r, k, sM) * ps; ei[k] = C3(r, k, sI) * ps; } }
  for (int i = 2; i <= Ld; i++) {
    const float *r = pp.row(i); const float ps = pp.rs[i];
    const float *__restrict__ qM = r + sM * SS, *__restrict__ qI = r + sI * SS;
    float *__restrict__ am = em.data(), *__restrict__ ai = ei.data();
#pragma omp simd
    for (int k = 1; k <= M; k++) { am[k] = qM[k] * ps + am[k]; ai[k] = qI[k] * ps + ai[k]; }
    xn += pp.X(i, XN); xc += pp.X(i, XC); xj += pp.X(i, XJ);
  }
  const float norm = (float)(1.0 / (float)Ld);
  for (int k = 1; k <= M; k++) { em[k] *= norm; ei[k] *= norm; }
  xn *= norm; xc *= norm; xj *= norm;
  const float xfactor = xn + xc + xj;
  for (int x = 0; x < m.K; x++) {
    float sv = 0.f;
    for (int k = 1; k <= M; k++) { sv += em[k] * m.r(x, k); sv += ei[k]; }
    null2[x] = sv + xfactor;
  }
  avg_degenerate(m, null2);
}

// p7_Null2_ByTrace (null2.c:131-205).  Reference quirk kept: I states are counted in the M slot of their node.
void null2_by_trace(const Model &m, const Trace &tr, int zstart, int zend, float *null2)
{
  const int M = m.M;
  std::vector<float> em(M + 1, 0.f);
  float xn = 0.f, xc = 0.f, xj = 0.f; int Ld = 0;
  for (int z = zstart; z <= zend; z++) {
    if (tr.i[z] == 0) continue;
    Ld++;
    if (tr.k[z] > 0) em[tr.k[z]] += 1.0f;
    else switch (tr.st[z]) { case ST_N: xn += 1.0f; break; case ST_C: xc += 1.0f; break; case ST_J: xj += 1.0f; break; default: break; }
  }
  const float norm = (float)(1.0 / (float)Ld);
  for (int k = 1; k <= M; k++) em[k] *= norm;
  xn *= norm; xc *= norm; xj *= norm;
  const float xfactor = xn + xc + xj;
  for (int x = 0; x < m.K; x++) {
    float sv = 0.f;
    for (int k = 1; k <= M; k++) sv += em[k] * m.r(x, k);
    null2[x] = sv + xfactor;
  }
  avg_degenerate(m, null2);
}

// ---- segment-pair ensemble clustering (p7_spensemble.c) ----
struct SegPair { int idx, i, j, k, m; float prob; };

bool sp_link(const SegPair &h1, const SegPair &h2)      // link_spsamples with min_overlap 0.8, of_smaller, max_diagdiff 4
{
  int nov = std::min(h1.j, h2.j) - std::max(h1.i, h2.i) + 1;
  int n = std::min(h1.j - h1.i + 1, h2.j - h2.i + 1);
  if ((float)nov / (float)n < 0.8f) return false;
  nov = std::min(h1.m, h2.m) - std::max(h1.k, h2.k);
  n = std::min(h1.m - h1.k + 1, h2.m - h2.k + 1);
  if ((float)nov / (float)n < 0.8f) return false;
  int d1 = h1.i - h1.k, d2 = h2.i - h2.k; if (std::abs(d1 - d2) <= 4) return true;
  d1 = h1.j - h1.m; d2 = h2.j - h2.m;     if (std::abs(d1 - d2) <= 4) return true;
  return false;
}

void sp_cluster(const std::vector<SegPair> &sp, int nsamples, std::vector<SegPair> &sigc)
{
  const int n = (int)sp.size();
  std::vector<int> a(n), b(n), assign(n, 0);
  // esl_cluster_SingleLinkage (esl_cluster.c)
  for (int v = 0; v < n; v++) a[v] = n - v - 1;
  int na = n, nb = 0, nc = 0;
  while (na > 0) {
    int v = a[na - 1]; na--; b[nb++] = v;
    while (nb > 0) {
      v = b[nb - 1]; nb--; assign[v] = nc;
      for (int i = na - 1; i >= 0; i--)
        if (sp_link(sp[v], sp[a[i]])) { const int w = a[i]; a[i] = a[na - 1]; na--; b[nb++] = w; }
    }
    nc++;
  }
  sigc.clear();
  std::vector<int> epc;
  for (int c = 0; c < nc; c++) {
    int ninc = 0, idx_of_last = -1;
    for (int h = 0; h < n; h++) if (assign[h] == c) { if (sp[h].idx != idx_of_last) ninc++; idx_of_last = sp[h].idx; }
    if ((float)ninc / (float)nsamples < 0.25f) continue;
    int imin = 0, imax = 0, jmin = 0, jmax = 0, kmin = 0, kmax = 0, mmin = 0, mmax = 0;
    for (int h = 0; h < n; h++) if (assign[h] == c) {
      if (imin == 0) { imin = imax = sp[h].i; jmin = jmax = sp[h].j; kmin = kmax = sp[h].k; mmin = mmax = sp[h].m; }
      else { imin = std::min(imin, sp[h].i); imax = std::max(imax, sp[h].i); jmin = std::min(jmin, sp[h].j); jmax = std::max(jmax, sp[h].j);
             kmin = std::min(kmin, sp[h].k); kmax = std::max(kmax, sp[h].k); mmin = std::min(mmin, sp[h].m); mmax = std::max(mmax, sp[h].m); }
    }
    const int thr = (int)ceilf((float)ninc * 0.02f);
    auto argmax = [&](int w) { int best = 0; for (int z = 1; z < w; z++) if (epc[z] > epc[best]) best = z; return best; };
    auto leftmost = [&](int lo, int hi, int SegPair::*fld) {
      epc.assign(hi - lo + 1, 0);
      for (int h = 0; h < n; h++) if (assign[h] == c) epc[sp[h].*fld - lo]++;
      int best; for (best = lo; best <= hi; best++) if (epc[best - lo] >= thr) break;
      if (best > hi) best = lo + argmax(hi - lo + 1);
      return best; };
    auto rightmost = [&](int lo, int hi, int SegPair::*fld) {
      epc.assign(hi - lo + 1, 0);
      for (int h = 0; h < n; h++) if (assign[h] == c) epc[sp[h].*fld - lo]++;
      int best; for (best = hi; best >= lo; best--) if (epc[best - lo] >= thr) break;
      if (best < lo) best = lo + argmax(hi - lo + 1);
      return best; };
    const int best_i = leftmost(imin, imax, &SegPair::i), best_k = leftmost(kmin, kmax, &SegPair::k);
    const int best_j = rightmost(jmin, jmax, &SegPair::j), best_m = rightmost(mmin, mmax, &SegPair::m);
    if (best_i > best_j || best_k > best_m) continue;
    SegPair s; s.i = best_i; s.j = best_j; s.k = best_k; s.m = best_m; s.idx = c; s.prob = (float)ninc / (float)nsamples;
    sigc.push_back(s);
  }
  // qsort by start; qsort is not stable, but equal starts within one region are vanishingly rare
  std::stable_sort(sigc.begin(), sigc.end(), [](const SegPair &x, const SegPair &y) { return x.i < y.i; });
}

struct DomOut { b2h_domain d; std::string text; };
struct HitOut { bool valid = false; b2h_hit hit; std::vector<DomOut> doms; };

char encode_pp(float p) { return (p + 0.05 >= 1.0) ? '*' : (char)((char)((p + 0.05) * 10.0) + '0'); }

struct Worker {
  Mx fwd, bck, pp, oa;
  Trace tr;
  std::vector<float> btot, etot, mocc, n2sc;
  std::vector<float> brow, pmask, pf_up, pf_dn, zrow;     // rolling Backward rows, D-chain tables of the current model
  FastRng rng;
};

// One envelope of one survivor.  Phase A (regions) fills i, j, null2_done; phase B (the O(M*Ld) numeric part of
// rescore_isolated_domain: Forward, Backward/Decoding, OptimalAccuracy + trace, Null2_ByExpectation) fills the rest,
// on the host (rescore_numeric) or on the GPU (b2h_envelope.cu); phase C renders the alignment and the scores.
struct EnvRec {
  int i = 0, j = 0;                // envelope, 1-based in the full sequence
  bool null2_done = false;         // the region's null2 odds were already set by the trace ensemble
  bool ok = false;                 // phase B succeeded (false = the reference's eslFAIL / eslERANGE: domain dropped)
  float envsc = 0.f, oasc = 0.f;
  float null2[B2H_NCODE];          // only when !null2_done
  Trace tr;                        // optimal-accuracy trace; i coordinates relative to the envelope (1..Ld)
};

// rescore_isolated_domain (p7_domaindef.c:814-982), protein (non long-target) branch: the numeric part.
bool rescore_numeric(Worker &w, Model &m, const uint8_t *dsq, EnvRec &e)
{
  const int i = e.i, Ld = e.j - e.i + 1;
  e.ok = false;
  if (!forward_full(m, dsq + i - 1, Ld, w.fwd, &e.envsc)) e.envsc = std::numeric_limits<float>::infinity();   // p7_Forward's status is ignored by the caller
  if (!backward_decode(m, dsq + i - 1, Ld, w.fwd, w.bck, w.pp, w.brow)) return false;   // eslERANGE from p7_Decoding -> domain dropped (eslFAIL)
  e.oasc = optimal_accuracy(m, w.pp, w.oa, w.pmask);
  if (!oa_trace(m, w.pp, w.oa, e.tr)) return false;
  if (!e.null2_done) null2_by_expectation(m, w.pp, e.null2);
  e.ok = true;
  return true;
}

// the rest of rescore_isolated_domain: alignment display (p7_alidisplay_Create, p7_alidisplay.c:92-273) from first M to
// last M of the (single) domain, null2 correction of the envelope.  n2sc is the survivor's per-residue null2 score vector.
bool render_domain(const Model &m, const b2h_profile *prof, const uint8_t *dsq, EnvRec &e, std::vector<float> &n2sc, DomOut &out)
{
  if (!e.ok) return false;
  const int i = e.i, j = e.j;
  Trace &tr = e.tr;
  for (size_t z = 0; z < tr.st.size(); z++) if (tr.i[z] > 0) tr.i[z] += i - 1;
  int z1 = -1, z2 = -1;
  for (size_t z = 0; z < tr.st.size(); z++) if (tr.st[z] == ST_M) { if (z1 < 0) z1 = (int)z; z2 = (int)z; }
  if (z1 < 0) return false;
  const int N = z2 - z1 + 1;
  const bool has_rf = !prof->rf.empty(), has_cs = !prof->cs.empty();
  std::string model(N, ' '), mline(N, ' '), aseq(N, ' '), ppline(N, ' '), rfline, csline;
  if (has_rf) rfline.assign(N, ' ');
  if (has_cs) csline.assign(N, ' ');
  const std::string &sym = prof->symbols;
  int code_of[256];                                        // residue code of an (upper-case) consensus character, -1: none
  for (int c = 0; c < 256; c++) code_of[c] = -1;
  for (size_t c = sym.size(); c-- > 0;) code_of[(unsigned char)sym[c]] = (int)c;
  auto cons = [&](int k) -> char { return (k >= 1 && k <= (int)prof->consensus.size()) ? prof->consensus[k - 1] : 'x'; };
  for (int z = z1; z <= z2; z++) {
    const int k = tr.k[z], ii = tr.i[z], s = tr.st[z], a = z - z1;
    const int x = (ii > 0) ? dsq[ii - 1] : 0;
    if (has_rf) rfline[a] = (s == ST_I) ? '.' : prof->rf[k - 1];
    if (has_cs) csline[a] = (s == ST_I) ? '.' : prof->cs[k - 1];
    ppline[a] = (s == ST_D) ? '.' : encode_pp(tr.pp[z]);
    if (s == ST_M) {
      model[a] = cons(k);
      const char cu = (char)toupper((unsigned char)cons(k));
      if (code_of[(unsigned char)cu] == x) mline[a] = model[a];
      else if (m.r(x, k) > 1.0f) mline[a] = '+';
      else mline[a] = ' ';
      aseq[a] = (char)toupper((unsigned char)sym[x]);
    } else if (s == ST_I) {
      model[a] = '.'; mline[a] = ' '; aseq[a] = (char)tolower((unsigned char)sym[x]);
    } else {
      model[a] = cons(k); mline[a] = ' '; aseq[a] = '-';
    }
  }
  b2h_domain &d = out.d;
  memset(&d, 0, sizeof d);
  d.hmmfrom = tr.k[z1]; d.hmmto = tr.k[z2]; d.sqfrom = tr.i[z1]; d.sqto = tr.i[z2]; d.N = N;
  d.has_rf = has_rf; d.has_cs = has_cs;
  out.text.clear();
  for (const std::string *sp : { &model, &mline, &aseq, &ppline }) { out.text += *sp; out.text.push_back('\0'); }
  if (has_rf) { out.text += rfline; out.text.push_back('\0'); }
  if (has_cs) { out.text += csline; out.text.push_back('\0'); }

  float domcorrection = 0.0f;
  if (!e.null2_done) {                                     // (one logf per residue CODE, not per position: same values)
    float lg[B2H_NCODE];
    for (int x = 0; x < m.Kp; x++) lg[x] = logf(e.null2[x]);
    for (int pos = i; pos <= j; pos++) n2sc[pos] = lg[dsq[pos - 1] < m.Kp ? dsq[pos - 1] : m.Kp - 1];
  }
  for (int pos = i; pos <= j; pos++) domcorrection += n2sc[pos];
  d.domcorrection = domcorrection;
  d.iali = d.sqfrom; d.jali = d.sqto; d.ienv = i; d.jenv = j; d.envsc = e.envsc; d.oasc = e.oasc;
  return true;
}

// p7_Null2_ByExpectation's second half from the column sums of the posterior matrix (the GPU returns those)
B2H_SIMD_CLONES
void null2_from_sums(const Model &m, int Ld, const float *em_in, const float *ei_in, float xn, float xc, float xj, float *null2)
{
  const int M = m.M;
  const float norm = (float)(1.0 / (float)Ld);
  xn *= norm; xc *= norm; xj *= norm;
  const float xfactor = xn + xc + xj;
  for (int x = 0; x < m.K; x++) {
    float sv = 0.f;
    for (int k = 1; k <= M; k++) { sv += (em_in[k - 1] * norm) * m.r(x, k); sv += ei_in[k - 1] * norm; }
    null2[x] = sv + xfactor;
  }
  avg_degenerate(m, null2);
}

// rebuild the Trace from the records of a traceback done elsewhere (same push/reverse rules as oa_trace)
bool trace_from_records(const std::vector<int32_t> &rec, int Ld, Trace &tr)
{
  tr.clear();
  tr.push(ST_T, 0, Ld); tr.push(ST_C, 0, Ld);
  for (size_t z = 0; z + 3 < rec.size(); z += 4) {
    float p; memcpy(&p, &rec[z + 3], 4);
    tr.push(rec[z], rec[z + 1], rec[z + 2], p);
  }
  if (tr.st.empty() || tr.st.back() != ST_S) return false;
  tr.reverse();
  return true;
}

// What phase A leaves behind for one survivor.
struct TaskState {
  bool dead = true;                // p7_DomainDecoding failed (eslERANGE): no hit
  std::vector<float> n2sc;         // per-residue null2 scores, 0..L
  float nexpected = 0.f;
  int nregions = 0, nclustered = 0;
  std::vector<EnvRec> envs;        // in sequence order
  std::vector<int> region_of;      // region id of each envelope (overlap bookkeeping is per clustered region)
};

void model_of(Worker &w, const b2h_profile *prof, Model &m)
{
  m.M = prof->M; m.K = prof->K; m.Kp = prof->Kp; m.rsc = prof->h_fwd_rsc.data();
  for (int q = 0; q < 8; q++) m.t[q] = prof->h_fwd_tsc.data() + (size_t)q * prof->M;
  m.degen = prof->h_degen.empty() ? nullptr : prof->h_degen.data();
  chain_prefix(m.t[T_DD], m.M, w.pf_up, w.pf_dn);
  if (w.zrow.size() < (size_t)m.M) w.zrow.assign((size_t)m.M, 0.0f);
  m.pf_up = w.pf_up.data(); m.pf_dn = w.pf_dn.data(); m.zrow = w.zrow.data();
}

// Phase A: p7_DomainDecoding on the parser specials, region finding, trace-ensemble clustering of multidomain
// regions (p7_domaindef.c:384-493, 531, 597-678) -> the list of envelopes to rescore.
void ddef_regions(Worker &w, const b2h_ddef_task &t, const b2h_search_params *prm, TaskState &ts)
{
  ts.dead = true; ts.envs.clear(); ts.region_of.clear(); ts.nregions = ts.nclustered = 0;
  const b2h_profile *prof = t.prof;
  const int L = t.L;
  Model m; model_of(w, prof, m);
  configure(m, true, L);
  const float ploop_multi = m.ploop;

  // p7_DomainDecoding (decoding.c:160-193) on the parser specials
  auto FX = [&](int i, int s) { return t.fx[(size_t)i * NX + s]; };
  auto BX = [&](int i, int s) { return t.bx[(size_t)i * NX + s]; };
  w.btot.assign(L + 1, 0.f); w.etot.assign(L + 1, 0.f); w.mocc.assign(L + 1, 0.f); ts.n2sc.assign(L + 1, 0.f);
  {
    float scaleproduct = 1.0f / BX(0, XN);
    for (int i = 1; i <= L; i++) {
      w.btot[i] = w.btot[i - 1] + (FX(i - 1, XB) * BX(i - 1, XB) * FX(i - 1, XSC) * scaleproduct);
      if (t.bck_own_scales) scaleproduct *= FX(i - 1, XSC) / BX(i - 1, XSC);
      w.etot[i] = w.etot[i - 1] + (FX(i, XE) * BX(i, XE) * FX(i, XSC) * scaleproduct);
      float njcp = FX(i - 1, XN) * BX(i, XN) * ploop_multi * scaleproduct;
      njcp += FX(i - 1, XJ) * BX(i, XJ) * ploop_multi * scaleproduct;
      njcp += FX(i - 1, XC) * BX(i, XC) * ploop_multi * scaleproduct;
      w.mocc[i] = (float)(1. - (double)njcp);
    }
    if (std::isinf(scaleproduct)) return;                 // eslERANGE from p7_DomainDecoding: the reference's pipeline fails here
  }
  ts.nexpected = w.btot[L];
  ts.dead = false;
  int &nregions = ts.nregions, &nclustered = ts.nclustered;
  const float rt1 = 0.25f, rt2 = 0.10f, rt3 = 0.20f;
  const int nsamples = 200;

  configure(m, false, L);                                 // p7_oprofile_ReconfigUnihit(om, saveL)
  int i = -1; bool triggered = false;
  for (int j = 1; j <= L; j++) {
    if (!triggered) {
      if (w.mocc[j] - (w.btot[j] - w.btot[j - 1]) < rt2) i = j;
      else if (i == -1) i = j;
      if (w.mocc[j] >= rt1) triggered = true;
    } else if (w.mocc[j] - (w.etot[j] - w.etot[j - 1]) < rt2) {
      nregions++;
      // is_multidomain_region (p7_domaindef.c:531)
      float mx = -1.0f;
      for (int z = i; z <= j; z++) { const float en = std::min(w.etot[z] - w.etot[i - 1], w.btot[j] - w.btot[z - 1]); mx = std::max(mx, en); }
      if (mx >= rt3) {
        nclustered++;
        configure(m, true, L);                            // ReconfigMultihit(om, saveL)
        const int Lr = j - i + 1;
        forward_full(m, t.dsq + i - 1, Lr, w.fwd, nullptr);
        // region_trace_ensemble (p7_domaindef.c:597-678)
        for (int pos = i; pos <= j; pos++) ts.n2sc[pos] = 0.0f;
        if (prm->seed != 0) w.rng.init(prm->seed);        // do_reseeding
        std::vector<SegPair> sp;
        float null2[B2H_NCODE];
        for (int smp = 0; smp < nsamples; smp++) {
          if (!stochastic_trace(w.rng, m, Lr, w.fwd, w.tr)) break;
          w.tr.index();
          int pos = 1;
          for (int d = 0; d < w.tr.ndom(); d++) {
            SegPair s; s.idx = smp; s.i = w.tr.sqfrom[d] + i - 1; s.j = w.tr.sqto[d] + i - 1; s.k = w.tr.hmmfrom[d]; s.m = w.tr.hmmto[d]; s.prob = 0.f;
            sp.push_back(s);
            null2_by_trace(m, w.tr, w.tr.tfrom[d], w.tr.tto[d], null2);
            for (; pos <= w.tr.sqfrom[d]; pos++) ts.n2sc[i + pos - 1] += 1.0f;
            for (; pos <= w.tr.sqto[d];   pos++) { const int x = t.dsq[i + pos - 2]; ts.n2sc[i + pos - 1] += null2[x < m.Kp ? x : m.Kp - 1]; }
          }
          for (; pos <= Lr; pos++) ts.n2sc[i + pos - 1] += 1.0f;
        }
        for (int pos = i; pos <= j; pos++) ts.n2sc[pos] = logf(ts.n2sc[pos] / (float)nsamples);
        std::vector<SegPair> sigc;
        sp_cluster(sp, nsamples, sigc);
        // remove dominated clusters (p7_domaindef.c:647-676)
        std::vector<char> dominated(sigc.size(), 0);
        for (size_t d = 0; d < sigc.size(); d++)
          for (size_t d2 = d + 1; d2 < sigc.size(); d2++) {
            const int nov = std::min(sigc[d].j, sigc[d2].j) - std::max(sigc[d].i, sigc[d2].i) + 1;
            if (nov == 0) break;
            const int n = std::min(sigc[d].j - sigc[d].i + 1, sigc[d2].j - sigc[d2].i + 1);
            if ((float)nov / (float)n >= 0.8f) { if (sigc[d].prob > sigc[d2].prob) dominated[d2] = 1; else dominated[d] = 1; }
          }
        for (size_t d = 0; d < sigc.size(); d++) {
          if (dominated[d]) continue;
          EnvRec e; e.i = sigc[d].i; e.j = sigc[d].j; e.null2_done = true;
          ts.envs.push_back(std::move(e)); ts.region_of.push_back(nregions);
        }
      } else {
        EnvRec e; e.i = i; e.j = j; e.null2_done = false;
        ts.envs.push_back(std::move(e)); ts.region_of.push_back(nregions);
      }
      i = -1; triggered = false;
    }
  }
}

// Phase C: alignments, overlap bookkeeping, per-sequence and per-domain scores (p7_pipeline.c:776-865).
void ddef_finish(const b2h_ddef_task &t, const b2h_search_params *prm, TaskState &ts, HitOut &out)
{
  out.valid = false; out.doms.clear();
  if (ts.dead) return;
  const b2h_profile *prof = t.prof;
  const int L = t.L;
  Model m; m.M = prof->M; m.K = prof->K; m.Kp = prof->Kp; m.rsc = prof->h_fwd_rsc.data();
  for (int q = 0; q < 8; q++) m.t[q] = prof->h_fwd_tsc.data() + (size_t)q * prof->M;
  m.degen = nullptr; m.zrow = nullptr; m.pf_up = m.pf_dn = nullptr;
  const float nexpected = ts.nexpected;
  const int nregions = ts.nregions, nclustered = ts.nclustered;
  int noverlaps = 0, nenvelopes = 0;
  {
    int last_j2 = 0, cur_region = -1;
    for (size_t d = 0; d < ts.envs.size(); d++) {
      EnvRec &e = ts.envs[d];
      if (ts.region_of[d] != cur_region) { cur_region = ts.region_of[d]; last_j2 = 0; }
      if (e.null2_done && e.i <= last_j2) noverlaps++;
      nenvelopes++;
      DomOut dom;
      if (render_domain(m, prof, t.dsq, e, ts.n2sc, dom)) { out.doms.push_back(std::move(dom)); if (e.null2_done) last_j2 = e.j; }
    }
  }
  if (nregions == 0 || nenvelopes == 0 || out.doms.empty()) return;

  // per-sequence and per-domain scores (p7_pipeline.c:776-865)
  b2h_len_params lp; b2h_length_params(L, 1.0f, &lp);
  const float nullsc = lp.null1, fwdsc = t.surv.fwdsc;
  const float omega = (float)(1. / 256.);
  const double LOG2 = 0.69314718055994529;
  float seqbias;
  if (prm->do_null2) {
    seqbias = kahan_sum(ts.n2sc.data(), L + 1);
    seqbias = flogsum()(0.0f, (float)(log((double)omega) + (double)seqbias));
  } else seqbias = 0.0f;
  float pre_score = (float)((double)(fwdsc - nullsc) / LOG2);
  float seq_score = (float)((double)(fwdsc - (nullsc + seqbias)) / LOG2);
  float sum_score = 0.0f; seqbias = 0.0f; int Ld = 0;
  if (prm->do_null2) {
    for (auto &dm : out.doms) if (dm.d.envsc - dm.d.domcorrection > 0.0f) { sum_score += dm.d.envsc; Ld += dm.d.jenv - dm.d.ienv + 1; seqbias += dm.d.domcorrection; }
    seqbias = flogsum()(0.0f, (float)(log((double)omega) + (double)seqbias));
  } else {
    for (auto &dm : out.doms) if (dm.d.envsc > 0.0f) { sum_score += dm.d.envsc; Ld += dm.d.jenv - dm.d.ienv + 1; }
    seqbias = 0.0f;
  }
  sum_score = (float)((double)sum_score + (double)(L - Ld) * log((double)((float)L / (float)(L + 3))));
  const float pre2_score = (float)((double)(sum_score - nullsc) / LOG2);
  sum_score = (float)((double)(sum_score - (nullsc + seqbias)) / LOG2);
  if (Ld > 0 && sum_score > seq_score) { seq_score = sum_score; pre_score = pre2_score; }
  const double tau = prof->evparam[4], lam = prof->evparam[5];
  auto logsurv = [&](double x) { return (x < tau) ? 0.0 : -lam * (x - tau); };
  b2h_hit &h = out.hit;
  memset(&h, 0, sizeof h);
  h.profile = t.surv.profile; h.seq = t.surv.seq;
  h.score = seq_score; h.pre_score = pre_score; h.sum_score = sum_score;
  h.lnP = logsurv((double)seq_score); h.pre_lnP = logsurv((double)pre_score); h.sum_lnP = logsurv((double)sum_score);
  h.nexpected = nexpected; h.nregions = nregions; h.nclustered = nclustered; h.noverlaps = noverlaps; h.nenvelopes = nenvelopes;
  h.ndom = (int)out.doms.size(); h.best_domain = 0;
  for (int d = 0; d < h.ndom; d++) {
    b2h_domain &dd = out.doms[d].d;
    const int Ldd = dd.jenv - dd.ienv + 1;
    float bs = (float)((double)dd.envsc + (double)(L - Ldd) * log((double)((float)L / (float)(L + 3))));
    dd.dombias = prm->do_null2 ? flogsum()(0.0f, (float)(log((double)omega) + (double)dd.domcorrection)) : 0.0f;
    bs = (float)((double)(bs - (nullsc + dd.dombias)) / LOG2);
    dd.bitscore = bs;
    dd.lnP = logsurv((double)bs);
    if (dd.bitscore > out.doms[h.best_domain].d.bitscore) h.best_domain = d;
  }
  out.valid = true;
}

// =====================================================================================================
// Long-target (nhmmer) branch: what p7_pli_postViterbi_LongTarget does with a window that passed the Forward gate
// (p7_pipeline.c:1113-1280) and the long_target paths of rescore_isolated_domain (p7_domaindef.c:814-982).
// =====================================================================================================

// reparameterize_model (p7_domaindef.c:715-750) + p7_oprofile_UpdateFwdEmissionScores (impl_sse/p7_oprofile.c:438-487):
// the background becomes a mixture of the model's own and the composition of the envelope i..i+Ld-1 of the window, and the
// match odds are rebuilt from the emission probabilities p7_oprofile_GetFwdEmissionArray recovered (odds x background).
// rsc [Kp][M] receives the new odds; nodes the model mask marks ('m') keep a zero score.  (The reference also leaves them at
// zero when it reverts to the original background, i.e. it overwrites a masked node's own odds after the first envelope; a
// masked column emits the background in every model hmmbuild writes, so its odds are 1 to begin with and nothing changes.)
void lt_reparameterize(const b2h_profile *prof, const uint8_t *dsq, int wlen, int i, int Ld, std::vector<float> &rsc)
{
  const int M = prof->M, K = prof->K, Kp = prof->Kp;
  const uint8_t *degen = prof->h_degen.data();
  float cnt[B2H_MAXABET], bgn[B2H_MAXABET];
  for (int x = 0; x < K; x++) cnt[x] = 0.f;
  for (int pos = i; pos < i + Ld; pos++) {                  // esl_sq_CountResidues / esl_abc_FCount
    const int x = dsq[pos - 1];
    if (x < K) cnt[x] += 1.0f;
    else if (x == K || x >= Kp - 2) continue;               // gap, nonresidue, missing data
    else {
      int nd = 0;
      for (int y = 0; y < K; y++) nd += degen[(size_t)x * K + y] ? 1 : 0;
      for (int y = 0; y < K; y++) if (degen[(size_t)x * K + y]) cnt[y] += 1.0f / (float)nd;
    }
  }
  fnorm(cnt, K);
  const float bg_smooth = (float)(25.0 / (double)std::min(100, std::max(50, wlen)));
  for (int x = 0; x < K; x++) bgn[x] = (float)((double)(bg_smooth * prof->bgf[x]) + ((1.0 - (double)bg_smooth) * (double)cnt[x]));
  rsc.resize((size_t)Kp * M);
  const float *orig = prof->h_fwd_rsc.data();
  float sc[B2H_MAXCODE];
  for (int k = 1; k <= M; k++) {
    const bool masked = (size_t)k <= prof->mm.size() && prof->mm[k - 1] == 'm';
    for (int x = 0; x < K; x++) {
      const float em = orig[(size_t)x * M + (k - 1)] * prof->bgf[x];      // p7_oprofile_GetFwdEmissionArray
      sc[x] = masked ? 0.0f : (float)log((double)em / bgn[x]);
    }
    sc[K] = sc[Kp - 2] = sc[Kp - 1] = NEGINF;
    for (int x = K + 1; x <= Kp - 3; x++) {                  // esl_abc_FExpectScVec
      float result = 0.f, denom = 0.f;
      for (int y = 0; y < K; y++) if (degen[(size_t)x * K + y]) { result += sc[y] * bgn[y]; denom += bgn[y]; }
      sc[x] = result / denom;
    }
    for (int x = 0; x < Kp; x++) rsc[(size_t)x * M + (k - 1)] = b2h_cephes_expf(sc[x]);
  }
}

struct LtWindow {
  const uint8_t *dsq; int L; const float *fx, *bx;
  int64_t window_start, seq_start; int complement; int seq; bool bck_own_scales;
};

// p7_pli_postViterbi_LongTarget's arithmetic for one rescored envelope (p7_pipeline.c:1150-1262): the score corrections to the
// model's max_length window, the coordinates mapped to the target, one hit per domain appended to <outs>.
void lt_finish_hit(const b2h_profile *prof, const LtWindow &lw, int i, int j, float envsc, float domcorrection, DomOut &dom, std::vector<HitOut> &outs)
{
  const int maxL = prof->max_length;
  const double LOG2 = 0.69314718055994529;
  const double tau = prof->evparam[4], lam = prof->evparam[5];
  auto logsurv = [&](double x) { return (x < tau) ? 0.0 : -lam * (x - tau); };
  if (domcorrection < envsc) envsc = domcorrection;
  b2h_domain &dd = dom.d;
  dd.domcorrection = domcorrection - envsc;
  dd.envsc = envsc; dd.ienv = i; dd.jenv = j;
  const int env_len = dd.jenv - dd.ienv + 1, ali_len = dd.jali - dd.iali + 1;
  if (ali_len < 8) return;
  float bitscore = dd.envsc;
  bitscore = (float)((double)bitscore - 2 * log(2. / (env_len + 2)));
  bitscore = (float)((double)bitscore + 2 * log(2. / (maxL + 2)));
  bitscore = (float)((double)bitscore - (env_len - ali_len) * log((double)((float)env_len / (float)(env_len + 2))));
  bitscore = (float)((double)bitscore + (std::max(maxL, env_len) - ali_len) * log((double)((float)maxL / (float)(maxL + 2))));
  const float dom_bias = dd.domcorrection;
  b2h_len_params lp; b2h_length_params(std::max(maxL, env_len), 1.0f, &lp);
  const float dom_score = (float)((double)(bitscore - lp.null1) / LOG2);
  const double dom_lnP = logsurv((double)dom_score);
  // positions in the target: x + seq_start + window_start - 2 on the top strand, seq_start - (window_start + x) + 2 on the other
  auto map  = [&](int x) -> int32_t { return (int32_t)((int64_t)x + lw.seq_start + lw.window_start - 2); };
  auto mapc = [&](int x) -> int32_t { return (int32_t)(lw.seq_start - (lw.window_start + (int64_t)x) + 2); };
  if (lw.complement) { dd.ienv = mapc(dd.ienv); dd.jenv = mapc(dd.jenv); dd.iali = mapc(dd.iali); dd.jali = mapc(dd.jali); dd.sqfrom = mapc(dd.sqfrom); dd.sqto = mapc(dd.sqto); }
  else               { dd.ienv = map(dd.ienv);  dd.jenv = map(dd.jenv);  dd.iali = map(dd.iali);  dd.jali = map(dd.jali);  dd.sqfrom = map(dd.sqfrom);  dd.sqto = map(dd.sqto); }
  dd.dombias = dom_bias; dd.bitscore = dom_score; dd.lnP = dom_lnP;
  HitOut out;
  b2h_hit &h = out.hit;
  memset(&h, 0, sizeof h);
  h.profile = 0; h.seq = lw.seq;
  h.pre_score = (float)((double)bitscore / LOG2); h.pre_lnP = logsurv((double)h.pre_score);
  h.score = h.sum_score = dom_score; h.lnP = h.sum_lnP = dom_lnP;
  h.ndom = 1; h.best_domain = 0;
  out.doms.push_back(std::move(dom));
  out.valid = true;
  outs.push_back(std::move(out));
}

// One window: regions and envelopes exactly as for proteins (ddef_regions), then per envelope the long-target rescoring and
// the hit the reference builds from it.  Hits are appended in the order of ddef->dcl.
void lt_window(Worker &w, const b2h_profile *prof, const LtWindow &lw, const b2h_search_params *prm, std::vector<HitOut> &outs)
{
  b2h_ddef_task t;
  t.surv.profile = 0; t.surv.seq = lw.seq; t.surv.fwdsc = 0.f; t.surv.filtersc = 0.f;
  t.prof = prof; t.dsq = lw.dsq; t.L = lw.L; t.fx = lw.fx; t.bx = lw.bx; t.bck_own_scales = lw.bck_own_scales;
  TaskState ts;
  ddef_regions(w, t, prm, ts);
  if (ts.dead || ts.nregions == 0 || ts.envs.empty()) return;
  const int max_env_extra = 20;
  std::vector<float> rsc, n2sc((size_t)lw.L + 1, 0.f);
  for (size_t d = 0; d < ts.envs.size(); d++) {
    EnvRec &e = ts.envs[d];
    int i = e.i, j = e.j;
    Model m; model_of(w, prof, m);
    DomOut dom;
    bool ok = true;
    for (int pass = 0; pass < 2; pass++) {
      const int Ld = j - i + 1;
      configure(m, false, Ld);                              // p7_oprofile_ReconfigRestLength(om, env_len), unihit
      if (prm->do_null2) { lt_reparameterize(prof, lw.dsq, lw.L, i, Ld, rsc); m.rsc = rsc.data(); }
      e.i = i; e.j = j; e.null2_done = true;               // (no null2 by expectation on this path)
      if (!rescore_numeric(w, m, lw.dsq, e)) { ok = false; break; }
      std::fill(n2sc.begin(), n2sc.end(), 0.f);
      if (!render_domain(m, prof, lw.dsq, e, n2sc, dom)) { ok = false; break; }
      if (pass == 0 && (i < dom.d.sqfrom - max_env_extra || j > dom.d.sqto + max_env_extra)) {
        i = std::max(i, dom.d.sqfrom - max_env_extra);     // trim the envelope around the alignment and do it again
        j = std::min(j, dom.d.sqto + max_env_extra);
        continue;
      }
      break;
    }
    if (!ok) continue;
    const int Ld = j - i + 1;
    float envsc = e.envsc, domcorrection = envsc;
    if (prm->do_null2) {                                    // what the score would have been without reparameterisation
      Model mo; model_of(w, prof, mo);
      configure(mo, false, Ld);
      float sc = domcorrection;
      if (forward_full(mo, lw.dsq + i - 1, Ld, w.fwd, &sc)) domcorrection = sc;
    }
    lt_finish_hit(prof, lw, i, j, envsc, domcorrection, dom, outs);
  }
}

} // namespace

struct FtzScope {
#if defined(__SSE2__)
  unsigned int saved;
  FtzScope() : saved(_mm_getcsr()) { _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON); _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON); }
  ~FtzScope() { _mm_setcsr(saved); }
#endif
};

// A process-wide pool of persistent worker threads (created on first use, grown on demand).  Each thread keeps its
// Worker (DP matrices, trace buffers) across searches, so steady-state searches allocate nothing.
namespace {

class ThreadPool {
 public:
  static ThreadPool &get() { static ThreadPool p; return p; }
  // run fn(worker, index) for index in [0, n) on up to nthreads threads (the caller participates)
  template <typename F> void parallel_for(size_t n, int nthreads, F fn) {
    if (n == 0) return;
    std::unique_lock<std::mutex> run_lock(run_mu_);                  // one parallel_for at a time
    nthreads = (int)std::min<size_t>((size_t)std::max(1, nthreads), n);
    grow(nthreads - 1);
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = [&](Worker &w, size_t i) { fn(w, i); };
      n_ = n; next_.store(0); active_ = nthreads - 1; done_ = 0; generation_++;
    }
    cv_.notify_all();
    drain(self_);
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return done_ == active_; });
    job_ = nullptr;
  }

 private:
  ThreadPool() {}
  ~ThreadPool() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto &t : threads_) t.join();
  }
  void grow(int want) {
    while ((int)threads_.size() < want) {
      const int id = (int)threads_.size();
      threads_.emplace_back([this, id] { loop(id); });
    }
  }
  void drain(Worker &w) { FtzScope ftz; for (;;) { const size_t i = next_.fetch_add(1); if (i >= n_) break; job_(w, i); } }
  void loop(int id) {
    Worker w; w.rng.init(42u);
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || (generation_ != seen && id < active_); });
        if (stop_) return;
        seen = generation_;
      }
      drain(w);
      { std::lock_guard<std::mutex> lk(mu_); done_++; }
      done_cv_.notify_one();
    }
  }
  std::mutex run_mu_, mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> threads_;
  std::function<void(Worker &, size_t)> job_;
  std::atomic<size_t> next_{0};
  size_t n_ = 0; int active_ = 0, done_ = 0; uint64_t generation_ = 0; bool stop_ = false;
  Worker self_;
};

} // namespace

b2h_ddef_pool::b2h_ddef_pool(int n)
{
  if (n <= 0) {
    // all hardware threads but two (the thread that feeds the GPU needs a core), shared fairly between the ranks
    // of a one-process-per-GPU job on this node (torchrun exports LOCAL_WORLD_SIZE)
    n = (int)std::thread::hardware_concurrency();
    int ranks = 1;
    if (const char *ev = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(ev));
    // (the thread that feeds the GPU and the one that drives this pool mostly wait on events: only a single-rank job, which
    //  has cores to spare, leaves two of them alone)
    n = std::max(1, ranks > 1 ? n / ranks : n - 2);
  }
  nthreads = std::max(1, std::min(n, 128));
}

int b2h_ddef_pool::run(std::vector<b2h_ddef_task> &tasks, const b2h_search_params *prm, b2h_results *res, b2h_env_backend *backend)
{
  const size_t n = tasks.size();
  std::vector<HitOut> outs(n);
  std::vector<TaskState> states(n);
  // longest-processing-time-first: the cost of a task grows with model length x sequence length
  std::vector<uint32_t> order(n);
  for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    return (int64_t)tasks[a].prof->M * tasks[a].L > (int64_t)tasks[b].prof->M * tasks[b].L; });
  // phase A: regions and envelopes
  const bool trace = getenv("B2H_TRACE") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tA0 = now();
  ThreadPool::get().parallel_for(n, nthreads, [&](Worker &w, size_t i) { const uint32_t e = order[i]; ddef_regions(w, tasks[e], prm, states[e]); });
  const double tA1 = now();
  // phase B: the O(M*Ld) numeric rescoring of every envelope, largest first
  std::vector<std::pair<uint32_t, uint32_t>> envs;
  for (size_t e = 0; e < n; e++) for (size_t d = 0; d < states[e].envs.size(); d++) envs.emplace_back((uint32_t)e, (uint32_t)d);
  std::sort(envs.begin(), envs.end(), [&](const std::pair<uint32_t, uint32_t> &a, const std::pair<uint32_t, uint32_t> &b) {
    const EnvRec &x = states[a.first].envs[a.second], &y = states[b.first].envs[b.second];
    return (int64_t)tasks[a.first].prof->M * (x.j - x.i + 1) > (int64_t)tasks[b.first].prof->M * (y.j - y.i + 1); });
  std::vector<char> on_host(envs.size(), 1);
  if (backend && !envs.empty()) {
    std::vector<b2h_env_job> jobs(envs.size());
    for (size_t q = 0; q < envs.size(); q++) {
      const EnvRec &e = states[envs[q].first].envs[envs[q].second];
      jobs[q].task = (int)envs[q].first; jobs[q].i = e.i; jobs[q].j = e.j;
    }
    const int st = backend->run(tasks, jobs);
    if (st != B2H_OK) return st;
    ThreadPool::get().parallel_for(envs.size(), nthreads, [&](Worker &, size_t q) {
      const b2h_env_job &jb = jobs[q];
      if (jb.status != 0) return;                            // left to the host below
      const b2h_ddef_task &t = tasks[envs[q].first];
      EnvRec &e = states[envs[q].first].envs[envs[q].second];
      const int Ld = e.j - e.i + 1;
      e.envsc = jb.envsc; e.oasc = jb.oasc;
      e.ok = trace_from_records(jb.trace, Ld, e.tr);
      if (e.ok && !e.null2_done) {
        Model m; m.M = t.prof->M; m.K = t.prof->K; m.Kp = t.prof->Kp; m.rsc = t.prof->h_fwd_rsc.data();
        m.degen = t.prof->h_degen.empty() ? nullptr : t.prof->h_degen.data();
        null2_from_sums(m, Ld, jb.em.data(), jb.ei.data(), jb.xn, jb.xc, jb.xj, e.null2);
      }
      on_host[q] = 0;
    });
  }
  ThreadPool::get().parallel_for(envs.size(), nthreads, [&](Worker &w, size_t i) {
    if (!on_host[i]) return;
    const b2h_ddef_task &t = tasks[envs[i].first];
    Model m; model_of(w, t.prof, m);
    configure(m, false, t.L);                               // p7_oprofile_ReconfigUnihit(om, saveL)
    rescore_numeric(w, m, t.dsq, states[envs[i].first].envs[envs[i].second]);
  });
  // phase C: alignments and scores
  const double tC0 = now();
  ThreadPool::get().parallel_for(n, nthreads, [&](Worker &, size_t i) { ddef_finish(tasks[i], prm, states[i], outs[i]); });
  if (trace) fprintf(stderr, "[b2h_ddef]   %zu survivors on %d threads: regions %.2f ms, envelopes (backend + unpack + host leftovers) %.2f ms, alignments + scores %.2f ms\n",
                     n, nthreads, tA1 - tA0, tC0 - tA1, now() - tC0);
  for (size_t e = 0; e < n; e++) {
    if (!outs[e].valid) continue;
    b2h_hit h = outs[e].hit;
    h.dom_offset = (int64_t)res->doms.size();
    for (auto &dm : outs[e].doms) {
      b2h_domain d = dm.d;
      d.text_offset = (int64_t)res->text.size();
      res->text.insert(res->text.end(), dm.text.begin(), dm.text.end());
      res->doms.push_back(d);
    }
    res->hits.push_back(h);
  }
  return B2H_OK;
}

// The same with the O(M*Ld) rescoring of the envelopes on a backend (the GPU envelope kernels), batched over ALL windows:
//   A   (host threads) regions and envelopes of every window (ddef_regions);
//   B0  (backend)      every envelope: Forward, Backward + decoding, optimal accuracy + trace, with the emission odds
//                      re-estimated from the envelope (lt_reparameterize) and the profile configured for the envelope length;
//   C0  (host threads) alignments; an envelope reaching more than 20 residues beyond its alignment is trimmed ...
//   B1  (backend)      ... and rescored the same way; then one Forward-only pass of every final envelope with the model's OWN
//                      emissions (the bias of the reference = the score without the re-estimation);
//   C1  (host threads) the hit arithmetic (lt_finish_hit).
// An envelope the backend cannot do falls back to the host code, one by one.  windows[w] is task w: task.surv.seq must be the
// index of the window in the backend's database.
int b2h_longtarget_domains_backend(const b2h_profile *p, const b2h_lt_window *wins, const int32_t *db_index, size_t n, const b2h_search_params *prm,
                                   int nthreads, b2h_env_backend *backend, b2h_results *res)
{
  const int max_env_extra = 20;
  nthreads = std::max(1, nthreads);
  const bool trace = getenv("B2H_TRACE") != nullptr;
  auto tnow = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_prev = tnow();
  auto lap = [&](const char *what, size_t count) { if (trace) { const double t = tnow(); fprintf(stderr, "[b2h_longtarget_hits]   %s (%zu): %.2f ms\n", what, count, t - t_prev); t_prev = t; } };
  std::vector<b2h_ddef_task> tasks(n);
  std::vector<TaskState> states(n);
  std::vector<LtWindow> lws(n);
  for (size_t q = 0; q < n; q++) {
    const b2h_lt_window &x = wins[q];
    LtWindow &lw = lws[q];
    lw.dsq = x.dsq; lw.L = x.L; lw.fx = x.fwd_xmx; lw.bx = x.bck_xmx; lw.window_start = x.window_start; lw.seq_start = x.seq_start;
    lw.complement = x.complement; lw.seq = x.seq; lw.bck_own_scales = x.bck_own_scales != 0;
    b2h_ddef_task &t = tasks[q];
    t.surv.profile = 0; t.surv.seq = db_index[q]; t.surv.fwdsc = 0.f; t.surv.filtersc = 0.f;
    t.prof = p; t.dsq = x.dsq; t.L = x.L; t.fx = x.fwd_xmx; t.bx = x.bck_xmx; t.bck_own_scales = lw.bck_own_scales;
  }
  ThreadPool::get().parallel_for(n, nthreads, [&](Worker &w, size_t q) { ddef_regions(w, tasks[q], prm, states[q]); });
  lap("regions of the windows, host", n);
  struct Env { uint32_t win, d; int i, j; std::vector<float> rsc; DomOut dom; bool ok = false, trimmed = false; float envsc = 0.f, domcorrection = 0.f; };
  std::vector<Env> envs;
  for (size_t q = 0; q < n; q++) {
    if (states[q].dead || states[q].nregions == 0) continue;
    for (size_t d = 0; d < states[q].envs.size(); d++) { Env e; e.win = (uint32_t)q; e.d = (uint32_t)d; e.i = states[q].envs[d].i; e.j = states[q].envs[d].j; envs.push_back(std::move(e)); }
  }
  const size_t ne = envs.size();
  // one rescoring pass over the envelopes listed in <which>: backend first, the host for what it leaves
  auto rescore = [&](const std::vector<uint32_t> &which) -> int {
    const size_t m = which.size();
    if (m == 0) return B2H_OK;
    ThreadPool::get().parallel_for(m, nthreads, [&](Worker &, size_t z) {
      Env &e = envs[which[z]];
      if (prm->do_null2) lt_reparameterize(p, lws[e.win].dsq, lws[e.win].L, e.i, e.j - e.i + 1, e.rsc);
    });
    lap("re-estimated emission tables, host", m);
    std::vector<b2h_env_job> jobs(m);
    for (size_t z = 0; z < m; z++) {
      Env &e = envs[which[z]];
      jobs[z].task = (int)e.win; jobs[z].i = e.i; jobs[z].j = e.j; jobs[z].cfg_len = e.j - e.i + 1;
      jobs[z].rsc = prm->do_null2 ? e.rsc.data() : nullptr;
    }
    if (backend) { const int st = backend->run(tasks, jobs); if (st != B2H_OK) return st; }
    lap("envelope Forward / Backward / OA on the device", m);
    ThreadPool::get().parallel_for(m, nthreads, [&](Worker &w, size_t z) {
      Env &e = envs[which[z]];
      const LtWindow &lw = lws[e.win];
      EnvRec &er = states[e.win].envs[e.d];
      const int Ld = e.j - e.i + 1;
      Model mm; model_of(w, p, mm);
      configure(mm, false, Ld);
      if (prm->do_null2) mm.rsc = e.rsc.data();
      er.i = e.i; er.j = e.j; er.null2_done = true;
      bool ok;
      if (backend && jobs[z].status == 0) { er.envsc = jobs[z].envsc; er.oasc = jobs[z].oasc; ok = er.ok = trace_from_records(jobs[z].trace, Ld, er.tr); }
      else ok = rescore_numeric(w, mm, lw.dsq, er);
      std::vector<float> n2sc((size_t)lw.L + 1, 0.f);
      e.ok = ok && render_domain(mm, p, lw.dsq, er, n2sc, e.dom);
      e.envsc = er.envsc;
    });
    lap("traces, null2, alignment rendering, host", m);
    return B2H_OK;
  };
  std::vector<uint32_t> all(ne);
  for (size_t z = 0; z < ne; z++) all[z] = (uint32_t)z;
  int st = rescore(all);
  if (st != B2H_OK) return st;
  std::vector<uint32_t> again;
  for (size_t z = 0; z < ne; z++) {
    Env &e = envs[z];
    if (e.ok && (e.i < e.dom.d.sqfrom - max_env_extra || e.j > e.dom.d.sqto + max_env_extra)) {
      e.i = std::max(e.i, e.dom.d.sqfrom - max_env_extra);     // trim the envelope around the alignment and do it again
      e.j = std::min(e.j, e.dom.d.sqto + max_env_extra);
      e.trimmed = true;
      again.push_back((uint32_t)z);
    }
  }
  if ((st = rescore(again)) != B2H_OK) return st;
  // the score without the re-estimated background: a plain Forward of every final envelope
  for (size_t z = 0; z < ne; z++) envs[z].domcorrection = envs[z].envsc;
  if (prm->do_null2) {
    std::vector<uint32_t> live;
    for (size_t z = 0; z < ne; z++) if (envs[z].ok) live.push_back((uint32_t)z);
    std::vector<b2h_env_job> jobs(live.size());
    for (size_t z = 0; z < live.size(); z++) {
      const Env &e = envs[live[z]];
      jobs[z].task = (int)e.win; jobs[z].i = e.i; jobs[z].j = e.j; jobs[z].cfg_len = e.j - e.i + 1; jobs[z].rsc = nullptr; jobs[z].fwd_only = true;
    }
    if (backend && !jobs.empty()) { if ((st = backend->run(tasks, jobs)) != B2H_OK) return st; }
    lap("plain Forward of the final envelopes on the device", jobs.size());
    ThreadPool::get().parallel_for(live.size(), nthreads, [&](Worker &w, size_t z) {
      Env &e = envs[live[z]];
      if (backend && jobs[z].status == 0) { e.domcorrection = jobs[z].envsc; return; }
      if (backend && std::isinf(jobs[z].envsc)) return;        // (p7_Forward's range error: the reference keeps envsc)
      Model mo; model_of(w, p, mo);
      const int Ld = e.j - e.i + 1;
      configure(mo, false, Ld);
      float sc = e.domcorrection;
      if (forward_full(mo, lws[e.win].dsq + e.i - 1, Ld, w.fwd, &sc)) e.domcorrection = sc;
    });
  }
  std::vector<std::vector<HitOut>> outs(n);
  for (size_t z = 0; z < ne; z++) {                         // in the order of ddef->dcl
    Env &e = envs[z];
    if (!e.ok) continue;
    lt_finish_hit(p, lws[e.win], e.i, e.j, e.envsc, e.domcorrection, e.dom, outs[e.win]);
  }
  for (size_t q = 0; q < n; q++)
    for (auto &o : outs[q]) {
      b2h_hit h = o.hit;
      h.profile = (int32_t)q;                              // the window the hit came from
      h.dom_offset = (int64_t)res->doms.size();
      for (auto &dm : o.doms) {
        b2h_domain d = dm.d;
        d.text_offset = (int64_t)res->text.size();
        res->text.insert(res->text.end(), dm.text.begin(), dm.text.end());
        res->doms.push_back(d);
      }
      res->hits.push_back(h);
    }
  return B2H_OK;
}

// Hits of the Forward-surviving windows of a long target (host side; see lt_window).
int b2h_longtarget_domains_host(const b2h_profile *p, const b2h_lt_window *wins, size_t n, const b2h_search_params *prm, int nthreads, b2h_results *res)
{
  std::vector<std::vector<HitOut>> outs(n);
  std::vector<uint32_t> order(n);
  for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return wins[a].L > wins[b].L; });
  ThreadPool::get().parallel_for(n, std::max(1, nthreads), [&](Worker &w, size_t q) {
    const b2h_lt_window &x = wins[order[q]];
    LtWindow lw; lw.dsq = x.dsq; lw.L = x.L; lw.fx = x.fwd_xmx; lw.bx = x.bck_xmx;
    lw.window_start = x.window_start; lw.seq_start = x.seq_start; lw.complement = x.complement; lw.seq = x.seq; lw.bck_own_scales = x.bck_own_scales != 0;
    lt_window(w, p, lw, prm, outs[order[q]]);
  });
  for (size_t e = 0; e < n; e++)
    for (auto &o : outs[e]) {
      b2h_hit h = o.hit;
      h.profile = (int32_t)e;                              // the window the hit came from
      h.dom_offset = (int64_t)res->doms.size();
      for (auto &dm : o.doms) {
        b2h_domain d = dm.d;
        d.text_offset = (int64_t)res->text.size();
        res->text.insert(res->text.end(), dm.text.begin(), dm.text.end());
        res->doms.push_back(d);
      }
      res->hits.push_back(h);
    }
  return B2H_OK;
}
