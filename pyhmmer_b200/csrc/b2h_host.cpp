// b2h_host.cpp -- host-side model preparation for libb2h.so.
//
// Everything here is per-query scalar work (micro- to milliseconds) whose outputs feed the
// integer filters, so it must agree BIT-FOR-BIT with the reference's libm call sites.  It is
// therefore written in C++ against the same glibc libm (not numpy, whose SIMD log/exp differ
// in the last ulp) and compiled with -ffp-contract=off.  Each function cites the reference
// code whose arithmetic (operand types, evaluation order) it restates; none of it is copied.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <limits>
#include <vector>
#include <atomic>
#include <thread>
#include "b2h.h"

namespace {

const float  kNegInf  = -std::numeric_limits<float>::infinity();
const double kLog2    = 0.69314718055994529;    // eslCONST_LOG2  (easel.h)
const double kLog2R   = 1.44269504088896341;    // eslCONST_LOG2R (easel.h)

// HMM transition order (hmmer.h:129 p7h_transitions_e) and profile order (hmmer.h:222 p7p_tsc_e)
enum { H_MM = 0, H_MI, H_MD, H_IM, H_II, H_DM, H_DD };
enum { P_MM = 0, P_IM, P_DM, P_BM, P_MD, P_DD, P_MI, P_II, P_NT };
enum { P_E = 0, P_N, P_J, P_C };            // gm->xsc rows
enum { P_LOOP = 0, P_MOVE = 1 };            // gm->xsc cols  (NB: the optimized profile uses MOVE=0, LOOP=1)
enum { O_E = 0, O_N, O_J, O_C };
enum { O_MOVE = 0, O_LOOP = 1 };
enum { O_BM = 0, O_MM, O_IM, O_DM, O_MD, O_MI, O_II, O_DD };

// ---- byte / word quantisers (p7_oprofile.c:667-705) ----
inline uint8_t unbiased_byteify(float scale_b, float sc) {
  sc = -1.0f * roundf(scale_b * sc);
  return (sc > 255.) ? 255 : (uint8_t)(int)sc;
}
inline uint8_t biased_byteify(float scale_b, uint8_t bias_b, float sc) {
  sc = -1.0f * roundf(scale_b * sc);
  // the reference writes `(uint8_t) sc + bias_b` into a uint8_t: negative costs wrap mod 256
  if (sc > (float)(255 - bias_b)) return 255;
  return (uint8_t)((int)sc + (int)bias_b);
}
inline int16_t wordify(float scale_w, float sc) {
  sc = roundf(scale_w * sc);
  if (sc >= 32767.0) return 32767;
  if (sc <= -32768.0) return -32768;
  return (int16_t)sc;
}

// Scalar restatement of the 4-lane Cephes expf the reference uses to build Forward odds ratios
// (vendor/easel/esl_sse.c:182-246): identical constants, identical operation order, one lane.
inline float cephes_expf(float x) {
  const float p0 = 1.9875691500E-4f, p1 = 1.3981999507E-3f, p2 = 8.3334519073E-3f,
              p3 = 4.1665795894E-2f, p4 = 1.6666665459E-1f, p5 = 5.0000001201E-1f;
  const float c0 = 0.693359375f, c1 = -2.12194440e-4f;
  const float maxlogf = 88.3762626647949f, minlogf = -88.3762626647949f;
  const bool  over = (x > maxlogf), under = (x <= minlogf);   // NaN compares false, as cmpgt/cmple do
  float fx = x * (float)kLog2R;
  fx = fx + 0.5f;
  int   k   = (int)fx;                  // cvttps: truncation
  float tmp = (float)k;
  if (tmp > fx) tmp = tmp - 1.0f;       // floorf without a branch in the original
  fx = tmp;
  k  = (int)fx;
  tmp = fx * c0;
  float z = fx * c1;
  x = x - tmp;
  x = x - z;
  z = x * x;
  float y = p0;       y = y * x;
  y = y + p1;         y = y * x;
  y = y + p2;         y = y * x;
  y = y + p3;         y = y * x;
  y = y + p4;         y = y * x;
  y = y + p5;         y = y * z;
  y = y + x;
  y = y + 1.0f;
  uint32_t bits = (uint32_t)(k + 127) << 23;
  float pow2k; std::memcpy(&pow2k, &bits, 4);
  y = y * pow2k;
  if (over)  y = std::numeric_limits<float>::infinity();
  if (under) y = 0.0f;
  return y;
}

} // namespace

float b2h_cephes_expf(float x) { return cephes_expf(x); }

extern "C" {

int b2h_hmm_decode_probs(const double *neglog, float *out, size_t n)
{
  for (size_t i = 0; i < n; i++)
    out[i] = std::isinf(neglog[i]) ? 0.0f : expf((float)(-1.0 * neglog[i]));   // p7_hmmfile.c:1486
  return B2H_OK;
}

// The node table of a HMMER3 ASCII model (read_asc30hmm, p7_hmmfile.c:1411-1500): <text> holds everything between the
// "HMM ..." column header lines and the closing "//" -- an optional COMPO line, the node-0 insert and transition lines, then
// per node a match line (k, K scores, <nanno> annotation fields: MAP [CONS] RF [MM] CS), an insert line (K) and a transition
// line (7).  Scores are "-log p" or "*"; they become probabilities as the reference makes them, expf(-atof(tok)).
// mat / ins [(M+1)*K] (row 0 of mat is left alone), t [(M+1)*7], compo [K] (has_compo says whether the line was there),
// map [M+1], anno [(nanno-1)*M]: the first character of every annotation field after MAP, field-major.
static inline bool fast_field(const char *p, const char *e, double *v)
{
  // "d.ddddd" with at most 15 significant digits: integer / power of ten is correctly rounded, i.e. what strtod returns
  uint64_t num = 0; int digits = 0, frac = 0; bool dot = false;
  if (p == e) return false;
  for (; p < e; p++) {
    if (*p >= '0' && *p <= '9') { num = num * 10 + (uint64_t)(*p - '0'); digits++; if (dot) frac++; }
    else if (*p == '.' && !dot) dot = true;
    else return false;
  }
  if (digits == 0 || digits > 15 || frac > 15) return false;
  static const double p10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
  *v = (double)num / p10[frac];
  return true;
}

int b2h_hmm_parse_body(const char *text, size_t len, int M, int K, int nanno,
                       float *compo, int32_t *has_compo, float *mat, float *ins, float *t, int64_t *map, char *anno)
{
  if (!text || M < 1 || K < 1 || K > B2H_MAXABET || nanno < 1 || nanno > 8 || !mat || !ins || !t || !has_compo) return B2H_EINVAL;
  const char *p = text, *end = text + len;
  const char *tb = nullptr, *te = nullptr;
  auto next = [&]() -> bool {                              // next whitespace-separated field; '#' starts a comment line
    for (;;) {
      while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++;
      if (p >= end) return false;
      if (*p == '#') { while (p < end && *p != '\n') p++; continue; }
      tb = p;
      while (p < end && !(*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++;
      te = p;
      return true;
    }
  };
  auto prob = [&](float *out) -> bool {
    if (!next()) return false;
    if (te - tb == 1 && *tb == '*') { *out = 0.0f; return true; }
    double v;
    if (!fast_field(tb, te, &v)) {
      char buf[64]; const size_t n = (size_t)(te - tb);
      if (n >= sizeof buf) return false;
      memcpy(buf, tb, n); buf[n] = 0;
      char *ep = nullptr;
      v = strtod(buf, &ep);
      if (ep == buf || *ep) return false;
    }
    *out = std::isinf(v) ? 0.0f : expf((float)(-1.0 * v));
    return true;
  };
  *has_compo = 0;
  { const char *save = p;
    if (next() && te - tb == 5 && memcmp(tb, "COMPO", 5) == 0) {
      *has_compo = 1;
      for (int x = 0; x < K; x++) { float v; if (!prob(&v)) return B2H_EINVAL; if (compo) compo[x] = v; }
    } else p = save; }
  for (int x = 0; x < K; x++) if (!prob(&ins[x])) return B2H_EINVAL;
  for (int x = 0; x < 7; x++) if (!prob(&t[x])) return B2H_EINVAL;
  for (int k = 1; k <= M; k++) {
    if (!next()) return B2H_EINVAL;
    long kk = 0;
    for (const char *q = tb; q < te; q++) { if (*q < '0' || *q > '9') return B2H_EINVAL; kk = kk * 10 + (*q - '0'); }
    if (kk != k) return B2H_ERANGE;                        // "expected match line to start with k"
    for (int x = 0; x < K; x++) if (!prob(&mat[(size_t)k * K + x])) return B2H_EINVAL;
    for (int a = 0; a < nanno; a++) {
      if (!next()) return B2H_EINVAL;
      if (a == 0) {
        long v = 0; bool num = true;
        for (const char *q = tb; q < te; q++) { if (*q < '0' || *q > '9') { num = false; break; } v = v * 10 + (*q - '0'); }
        if (map) map[k] = num ? v : 0;
      } else if (anno) anno[(size_t)(a - 1) * M + (k - 1)] = *tb;
    }
    for (int x = 0; x < K; x++) if (!prob(&ins[(size_t)k * K + x])) return B2H_EINVAL;
    for (int x = 0; x < 7; x++) if (!prob(&t[(size_t)k * 7 + x])) return B2H_EINVAL;
  }
  if (next()) return B2H_EINVAL;                           // trailing fields
  return B2H_OK;
}

// p7_ProfileConfig, local modes (modelconfig.c:48-187) + p7_hmm_CalculateOccupancy (p7_hmm.c:1338)
int b2h_profile_config(int M, int K, int Kp, const uint8_t *degen,
                       const float *t, const float *mat, const float *bgf,
                       int L, int multihit, float *tsc, float *msc, float *xsc)
{
  if (M < 1 || K < 1 || Kp < K + 3) return B2H_EINVAL;
  auto T = [&](int k, int s) -> float { return t[(size_t)k * 7 + s]; };

  for (size_t i = 0; i < (size_t)M * P_NT; i++) tsc[i] = kNegInf;      // p7_profile_Create edge init
  // occupancy
  std::vector<float> occ(M + 1);
  occ[0] = 0.f;
  occ[1] = T(0, H_MI) + T(0, H_MM);
  for (int k = 2; k <= M; k++) {
    float a = occ[k-1] * (T(k-1, H_MM) + T(k-1, H_MI));                 // float * float
    occ[k]  = (float)((double)a + (1.0 - (double)occ[k-1]) * (double)T(k-1, H_DM));   // `1.0-x` promotes to double
  }
  float Z = 0.f;
  for (int k = 1; k <= M; k++) Z += occ[k] * (float)(M - k + 1);
  for (int k = 1; k <= M; k++) tsc[(size_t)(k-1) * P_NT + P_BM] = (float)log((double)(occ[k] / Z));

  for (int k = 1; k < M; k++) {
    float *tp = tsc + (size_t)k * P_NT;
    tp[P_MM] = (float)log((double)T(k, H_MM));
    tp[P_MI] = (float)log((double)T(k, H_MI));
    tp[P_MD] = (float)log((double)T(k, H_MD));
    tp[P_IM] = (float)log((double)T(k, H_IM));
    tp[P_II] = (float)log((double)T(k, H_II));
    tp[P_DM] = (float)log((double)T(k, H_DM));
    tp[P_DD] = (float)log((double)T(k, H_DD));
  }

  // match emission log-odds; degenerate residues by expected score (esl_alphabet.c:1474,1582)
  std::vector<float> sc(Kp);
  for (int x = 0; x < Kp; x++) msc[(size_t)x * (M + 1)] = kNegInf;     // node 0
  for (int k = 1; k <= M; k++) {
    for (int x = 0; x < K; x++) sc[x] = (float)log((double)mat[(size_t)k * K + x] / (double)bgf[x]);
    sc[K] = kNegInf; sc[Kp-2] = kNegInf; sc[Kp-1] = kNegInf;
    for (int x = K + 1; x <= Kp - 3; x++) {
      float result = 0.f, denom = 0.f;
      for (int i = 0; i < K; i++)
        if (degen[(size_t)x * K + i]) { result += sc[i] * bgf[i]; denom += bgf[i]; }
      sc[x] = result / denom;
    }
    for (int x = 0; x < Kp; x++) msc[(size_t)x * (M + 1) + k] = sc[x];
  }

  // specials (modelconfig.c:110-118, 223-230)
  float nj;
  if (multihit) { xsc[P_E*2+P_MOVE] = (float)-kLog2; xsc[P_E*2+P_LOOP] = (float)-kLog2; nj = 1.0f; }
  else          { xsc[P_E*2+P_MOVE] = 0.0f;          xsc[P_E*2+P_LOOP] = kNegInf;       nj = 0.0f; }
  float pmove = (2.0f + nj) / ((float)L + 2.0f + nj);
  float ploop = 1.0f - pmove;
  float lloop = (float)log((double)ploop), lmove = (float)log((double)pmove);
  xsc[P_N*2+P_LOOP] = xsc[P_C*2+P_LOOP] = xsc[P_J*2+P_LOOP] = lloop;
  xsc[P_N*2+P_MOVE] = xsc[P_C*2+P_MOVE] = xsc[P_J*2+P_MOVE] = lmove;
  return B2H_OK;
}

// p7_oprofile_Convert = mf_conversion + vf_conversion + fb_conversion (p7_oprofile.c:773-990),
// written straight into node-major tables (node k at index k-1) instead of SSE stripes.
int b2h_oprofile_convert(int M, int K, int Kp, int L, int multihit,
                         const float *tsc, const float *msc, const float *xsc,
                         uint8_t *msv_cost, int16_t *vit_rsc, int16_t *vit_tsc,
                         float *fwd_rsc, float *fwd_tsc, b2h_oprofile_desc *d)
{
  if (M < 1) return B2H_EINVAL;
  auto MSC = [&](int k, int x) -> float { return msc[(size_t)x * (M + 1) + k]; };
  auto TSC = [&](int k, int s) -> float { return tsc[(size_t)k * P_NT + s]; };   // k in 0..M-1

  d->M = M; d->K = K; d->Kp = Kp; d->L = L; d->mode_multihit = multihit;

  // ---- MSV: third-bit units, offset 190 ----
  float maxsc = 0.0f;     // the reference scans rsc[x][(M+1)*2] for x<K, i.e. match AND insert scores (inserts are 0 / -inf)
  for (int x = 0; x < K; x++) for (int k = 0; k <= M; k++) { float v = MSC(k, x); if (v > maxsc) maxsc = v; }
  d->scale_b = (float)(3.0 / kLog2);
  d->base_b  = 190;
  d->bias_b  = unbiased_byteify(d->scale_b, (float)(-1.0 * maxsc));
  for (int x = 0; x < Kp; x++)
    for (int k = 1; k <= M; k++)
      msv_cost[(size_t)x * M + (k-1)] = biased_byteify(d->scale_b, d->bias_b, MSC(k, x));
  d->tbm_b = unbiased_byteify(d->scale_b, logf(2.0f / ((float)M * (float)(M + 1))));
  d->tec_b = unbiased_byteify(d->scale_b, logf(0.5f));
  d->tjb_b = unbiased_byteify(d->scale_b, logf(3.0f / (float)(L + 3)));

  // ---- ViterbiFilter: 1/500-bit units, offset 12000 ----
  d->scale_w = (float)(500.0 / kLog2);
  d->base_w  = 12000;
  for (int x = 0; x < Kp; x++)
    for (int k = 1; k <= M; k++)
      vit_rsc[(size_t)x * M + (k-1)] = wordify(d->scale_w, MSC(k, x));
  for (int k = 1; k <= M; k++) {
    // BM,MM,IM,DM come from node k-1 (valid while k-1 < M, always); MD,MI,II,DD from node k (valid while k < M)
    auto w = [&](int node, int s, int16_t maxval) -> int16_t {
      int16_t val = (node < M) ? wordify(d->scale_w, TSC(node, s)) : (int16_t)-32768;
      return (val <= maxval) ? val : maxval;
    };
    vit_tsc[(size_t)O_BM * M + (k-1)] = w(k-1, P_BM, 0);
    vit_tsc[(size_t)O_MM * M + (k-1)] = w(k-1, P_MM, 0);
    vit_tsc[(size_t)O_IM * M + (k-1)] = w(k-1, P_IM, 0);
    vit_tsc[(size_t)O_DM * M + (k-1)] = w(k-1, P_DM, 0);
    vit_tsc[(size_t)O_MD * M + (k-1)] = w(k,   P_MD, 0);
    vit_tsc[(size_t)O_MI * M + (k-1)] = w(k,   P_MI, 0);
    vit_tsc[(size_t)O_II * M + (k-1)] = w(k,   P_II, -1);    // never a free II loop (p7_oprofile.c:877)
    vit_tsc[(size_t)O_DD * M + (k-1)] = (k < M) ? wordify(d->scale_w, TSC(k, P_DD)) : (int16_t)-32768;
  }
  d->xw[O_E][O_LOOP] = wordify(d->scale_w, xsc[P_E*2+P_LOOP]);
  d->xw[O_E][O_MOVE] = wordify(d->scale_w, xsc[P_E*2+P_MOVE]);
  d->xw[O_N][O_MOVE] = wordify(d->scale_w, xsc[P_N*2+P_MOVE]);
  d->xw[O_N][O_LOOP] = 0;
  d->xw[O_C][O_MOVE] = wordify(d->scale_w, xsc[P_C*2+P_MOVE]);
  d->xw[O_C][O_LOOP] = 0;
  d->xw[O_J][O_MOVE] = wordify(d->scale_w, xsc[P_J*2+P_MOVE]);
  d->xw[O_J][O_LOOP] = 0;
  {
    int ddbound = -32768;
    for (int k = 2; k < M - 1; k++) {
      int dd = (int)wordify(d->scale_w, TSC(k, P_DD));
      dd += (int)wordify(d->scale_w, TSC(k+1, P_DM));
      dd -= (int)wordify(d->scale_w, TSC(k+1, P_BM));
      if (dd > ddbound) ddbound = dd;
    }
    d->ddbound_w = (int16_t)ddbound;
  }

  // ---- Forward/Backward: odds ratios through the Cephes polynomial ----
  for (int x = 0; x < Kp; x++)
    for (int k = 1; k <= M; k++)
      fwd_rsc[(size_t)x * M + (k-1)] = cephes_expf(MSC(k, x));
  for (int k = 1; k <= M; k++) {
    auto f = [&](int node, int s) -> float { return cephes_expf((node < M) ? TSC(node, s) : kNegInf); };
    fwd_tsc[(size_t)O_BM * M + (k-1)] = f(k-1, P_BM);
    fwd_tsc[(size_t)O_MM * M + (k-1)] = f(k-1, P_MM);
    fwd_tsc[(size_t)O_IM * M + (k-1)] = f(k-1, P_IM);
    fwd_tsc[(size_t)O_DM * M + (k-1)] = f(k-1, P_DM);
    fwd_tsc[(size_t)O_MD * M + (k-1)] = f(k,   P_MD);
    fwd_tsc[(size_t)O_MI * M + (k-1)] = f(k,   P_MI);
    fwd_tsc[(size_t)O_II * M + (k-1)] = f(k,   P_II);
    fwd_tsc[(size_t)O_DD * M + (k-1)] = f(k,   P_DD);
  }
  d->xf[O_E][O_LOOP] = expf(xsc[P_E*2+P_LOOP]);
  d->xf[O_E][O_MOVE] = expf(xsc[P_E*2+P_MOVE]);
  d->xf[O_N][O_LOOP] = expf(xsc[P_N*2+P_LOOP]);
  d->xf[O_N][O_MOVE] = expf(xsc[P_N*2+P_MOVE]);
  d->xf[O_C][O_LOOP] = expf(xsc[P_C*2+P_LOOP]);
  d->xf[O_C][O_MOVE] = expf(xsc[P_C*2+P_MOVE]);
  d->xf[O_J][O_LOOP] = expf(xsc[P_J*2+P_LOOP]);
  d->xf[O_J][O_MOVE] = expf(xsc[P_J*2+P_MOVE]);
  return B2H_OK;
}

// k = q + z*Q + 1  <->  vector q, lane z   (p7_oprofile.c:800, 856, 949)
int b2h_destripe_oprofile(int M, int Kp,
                          const uint8_t *rbv, const int16_t *rwv, const int16_t *twv,
                          const float *rfv, const float *tfv,
                          uint8_t *msv_cost, int16_t *vit_rsc, int16_t *vit_tsc,
                          float *fwd_rsc, float *fwd_tsc)
{
  const int Q16 = (M - 1) / 16 + 1 > 2 ? (M - 1) / 16 + 1 : 2;   // p7O_NQB: ESL_MAX(2, ...)
  const int Q8  = (M - 1) / 8  + 1 > 2 ? (M - 1) / 8  + 1 : 2;   // p7O_NQW
  const int Q4  = (M - 1) / 4  + 1 > 2 ? (M - 1) / 4  + 1 : 2;   // p7O_NQF
  for (int x = 0; x < Kp; x++)
    for (int k = 1; k <= M; k++) {
      if (msv_cost) { int q = (k-1) % Q16, z = (k-1) / Q16; msv_cost[(size_t)x*M + k-1] = rbv[((size_t)x*Q16 + q)*16 + z]; }
      if (vit_rsc)  { int q = (k-1) % Q8,  z = (k-1) / Q8;  vit_rsc [(size_t)x*M + k-1] = rwv[((size_t)x*Q8  + q)*8  + z]; }
      if (fwd_rsc)  { int q = (k-1) % Q4,  z = (k-1) / Q4;  fwd_rsc [(size_t)x*M + k-1] = rfv[((size_t)x*Q4  + q)*4  + z]; }
    }
  for (int k = 1; k <= M; k++) {
    if (vit_tsc) {
      int q = (k-1) % Q8, z = (k-1) / Q8;
      for (int t = 0; t < 7; t++) vit_tsc[(size_t)t*M + k-1] = twv[((size_t)q*7 + t)*8 + z];
      vit_tsc[(size_t)7*M + k-1] = twv[((size_t)7*Q8 + q)*8 + z];
    }
    if (fwd_tsc) {
      int q = (k-1) % Q4, z = (k-1) / Q4;
      for (int t = 0; t < 7; t++) fwd_tsc[(size_t)t*M + k-1] = tfv[((size_t)q*7 + t)*4 + z];
      fwd_tsc[(size_t)7*M + k-1] = tfv[((size_t)7*Q4 + q)*4 + z];
    }
  }
  return B2H_OK;
}

// p7_oprofile_ReconfigMSVLength/RestLength (p7_oprofile.c:1095-1136), p7_bg_SetLength/NullOne/FilterScore tail (p7_bg.c:189,357,479)
int b2h_length_params(int L, float nj, b2h_len_params *o)
{
  const float scale_b = (float)(3.0 / kLog2);
  const float scale_w = (float)(500.0 / kLog2);
  o->tjb_b   = unbiased_byteify(scale_b, logf(3.0f / (float)(L + 3)));
  o->pmove   = (2.0f + nj) / ((float)L + 2.0f + nj);
  o->ploop   = 1.0f - o->pmove;
  o->xw_move = wordify(scale_w, logf(o->pmove));
  o->p1      = (float)L / (float)(L + 1);
  o->null1   = (float)((double)(float)L * log((double)o->p1) + log(1. - (double)o->p1));
  o->flt_len_a = (float)L * logf(o->p1);
  o->flt_len_b = logf((float)(1. - (double)o->p1));
  return B2H_OK;
}

// p7_Builder_MaxLength (vendor/hmmer/src/p7_builder.c:651-755): the window length beyond which the model emits less than
// emit_thresh of its probability mass -- a DP over (state, emitted length) in double precision on the HMM's transition
// probabilities t [(M+1)*7] = MM MI MD IM II DM DD.  Capped at max(M, min(20 M, 100000)).
int b2h_hmm_max_length(int M, const float *t, double emit_thresh, int32_t *max_length)
{
  if (M < 1 || !t || !max_length) return B2H_EINVAL;
  enum { hMM = 0, hMI = 1, hMD = 2, hIM = 3, hII = 4, hDM = 5, hDD = 6 };
  auto T = [&](int k, int x) -> double { return (double)t[(size_t)k * 7 + x]; };
  const int length_bound = std::max(M, std::min(20 * M, 100000));
  if (M == 1) { *max_length = 1; return B2H_OK; }
  *max_length = length_bound;
  std::vector<double> Iv((size_t)(M + 1) * 2, 0.0), Mv((size_t)(M + 1) * 2, 0.0), Dv((size_t)(M + 1) * 2, 0.0);
  auto I = [&](int k, int c) -> double & { return Iv[(size_t)k * 2 + c]; };
  auto Mm = [&](int k, int c) -> double & { return Mv[(size_t)k * 2 + c]; };
  auto D = [&](int k, int c) -> double & { return Dv[(size_t)k * 2 + c]; };
  Mm(1, 0) = 1.0;
  I(1, 0) = D(1, 0) = Mm(2, 0) = I(2, 0) = 0;
  D(2, 0) = T(1, hMD);
  for (int k = 3; k <= M; k++) { Mm(k, 0) = I(k, 0) = 0; D(k, 0) = T(k - 1, hDD) * D(k - 1, 0); }
  Mm(1, 1) = D(1, 1) = D(2, 1) = I(2, 1) = 0;
  I(1, 1) = T(1, hMI) * Mm(1, 0);
  Mm(2, 1) = T(1, hMM) * Mm(1, 0);
  for (int k = 3; k <= M; k++) {
    Mm(k, 1) = T(k - 1, hDM) * D(k - 1, 0);
    I(k, 1) = 0;
    D(k, 1) = T(k - 1, hMD) * Mm(k - 1, 1) + T(k - 1, hDD) * D(k - 1, 1);
  }
  double p_sum = Mm(M, 0) + Mm(M, 1) + D(M, 0) + D(M, 1);
  int cp = 0;
  for (int col = 3; col <= length_bound; col++) {
    const int pp = 1 - cp;
    double surv = 0.0;
    Mm(1, cp) = D(1, cp) = 0;
    I(1, cp) = T(1, hII) * I(1, pp);
    surv += I(1, cp);
    for (int k = 2; k <= M; k++) {
      Mm(k, cp) = T(k - 1, hMM) * Mm(k - 1, pp) + T(k - 1, hDM) * D(k - 1, pp) + T(k - 1, hIM) * I(k - 1, pp);
      I(k, cp) = T(k, hMI) * Mm(k, pp) + T(k, hII) * I(k, pp);
      D(k, cp) = T(k - 1, hMD) * Mm(k - 1, cp) + T(k - 1, hDD) * D(k - 1, cp);
      surv += I(k, cp) + Mm(k, cp) * (1 - T(k, hMD)) + D(k, cp) * (1 - T(k, hDD));
    }
    surv += Mm(M, cp) * T(M, hMD) + D(M, cp) * T(M, hDD) - I(M, cp);
    p_sum += Mm(M, cp) + D(M, cp);
    surv /= surv + p_sum;
    if (surv < emit_thresh) { *max_length = col; break; }
    cp = 1 - cp;
  }
  return B2H_OK;
}

int b2h_pack_windows(const uint8_t *const *targets, size_t nwin, const int32_t *win_target, const int64_t *win_offset,
                     const int64_t *win_len, const int32_t *win_comp, const uint8_t *comp_table, int Kp,
                     uint8_t *out, const int64_t *out_off, int nthreads)
{
  if ((!targets || !win_target || !win_offset || !win_len || !win_comp || !out || !out_off) && nwin) return B2H_EINVAL;
  if (comp_table && (Kp < 1 || Kp > 256)) return B2H_EINVAL;
  uint8_t table[256];
  for (int x = 0; x < 256; x++) table[x] = (comp_table && x < Kp) ? comp_table[x] : (uint8_t)x;
  for (size_t w = 0; w < nwin; w++) if (win_comp[w] && !comp_table) return B2H_EINVAL;
  auto work = [&](size_t w) {
    const uint8_t *src = targets[win_target[w]] + win_offset[w];
    uint8_t *dst = out + out_off[w];
    const int64_t n = win_len[w];
    if (!win_comp[w]) { memcpy(dst, src, (size_t)n); return; }
    for (int64_t q = 0; q < n; q++) dst[q] = table[src[n - 1 - q]];
  };
  int T = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
  int64_t total = 0; for (size_t w = 0; w < nwin; w++) total += win_len[w];
  T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(T, 32), std::min<int64_t>((int64_t)nwin, total / (1 << 20) + 1)));
  if (T <= 1) { for (size_t w = 0; w < nwin; w++) work(w); return B2H_OK; }
  std::atomic<size_t> next{0};
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++) th.emplace_back([&]() { for (size_t w = next.fetch_add(1); w < nwin; w = next.fetch_add(1)) work(w); });
  for (auto &t : th) t.join();
  return B2H_OK;
}

} // extern "C"
