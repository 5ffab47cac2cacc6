// b2h_generic.cu -- the generic (unstriped, log-space) reference DP of HMMER on the P7_PROFILE:
//
//   p7_GMSV      (vendor/hmmer/src/generic_msv.c:56)       p7_GViterbi  (generic_viterbi.c:64)
//   p7_GForward  (generic_fwdback.c:48)                     p7_GBackward (generic_fwdback.c:164)
//
// for one profile against every sequence of a resident database.  These are what pyhmmer reaches through
// Profile.msv_filter (plan7.pyx:8212) and what HMMER's own unit tests use as the yardstick for the vector code.
//
// Design.  Forward/Backward sum probabilities with p7_FLogsum (logsum.c:105): max + table[(int)((max-min)*1000)], a
// 16 000-entry table of log(1+e^-x).  That operation is NOT associative, and the reference chains it along the model
// (the D->D path and the E-state accumulation run over k in order), so the only evaluation order that reproduces the
// reference's floats is the reference's own.  The parallelism therefore comes from the batch, not from the matrix:
// ONE THREAD PER COMPARISON walks its L x M matrix in the reference's order, sequences are taken in the length-sorted
// order of the database so that the threads of a warp finish together, the two live DP rows of a thread are kept in a
// global scratch buffer interleaved by thread (word k of all threads of a warp is one coalesced 128-byte line), the
// logsum table is staged in shared memory, transition scores are warp-uniform loads and emission scores L1 gathers.
// Results are bit-identical to the reference on the same host (the table and the length-dependent logs are computed on
// the host with the libm the reference uses; the kernels only add, compare and look up).
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>
#include "b2h_internal.h"

namespace {

constexpr int LOGSUM_TBL = 16000;          // p7_LOGSUM_TBL, logsum.c:59
enum { T_MM = 0, T_IM = 1, T_DM = 2, T_BM = 3, T_MD = 4, T_DD = 5, T_MI = 6, T_II = 7 };   // p7p_tsc_e

struct GenArgs {
  int M, K, Kp;
  const float *tsc;       // [M][8]      TSC(s,k) = tsc[k*8+s], k = 0..M-1
  const float *msc;       // [Kp][M+1]   match scores; insert scores are 0 (modelconfig.c:152-167)
  float xE_loop, xE_move; // xsc[E][LOOP], xsc[E][MOVE]
  float tbmk, tej, tec;   // p7_GMSV's own transitions (generic_msv.c:62-66)
  const float *lenp;      // [n][4] per sequence: GMSV tloop, tmove; profile xsc[NCJ][LOOP], xsc[NCJ][MOVE] (p7_ReconfigLength)
  const float *tbl;       // flogsum_lookup
  float *scratch;         // [2 rows][3 states][M+1][nthreads]
  SeqDev sd;
  float *gmsv, *gvit, *gfwd, *gbck;     // indexed by sequence; any may be null
};

__device__ __forceinline__ float LS(const float *tbl, float a, float b)
{
  const float mx = (a > b) ? a : b, mn = (a < b) ? a : b;              // ESL_MAX / ESL_MIN
  return (mn == -INFINITY || (mx - mn) >= 15.7f) ? mx : mx + tbl[(int)((mx - mn) * 1000.f)];
}
__device__ __forceinline__ float MX2(float a, float b) { return (a > b) ? a : b; }

// one thread's view of its two DP rows
struct Rows {
  float *base; size_t stride; int M1;
  __device__ __forceinline__ float &at(int row, int st, int k) const { return base[(((size_t)row * 3 + st) * M1 + k) * stride]; }
};

__device__ __forceinline__ float isc(const GenArgs &a, int x, int k)     // p7P_ISC (modelconfig.c:160-167)
{ return (k >= a.M || x == a.K || x >= a.Kp - 2) ? -INFINITY : 0.0f; }

__device__ float g_msv(const GenArgs &a, const float *tbl, const Rows &R, const uint8_t *seq, int L, float tloop, float tmove)
{
  const int M = a.M;
  float xN = 0.f, xB = tmove, xJ = -INFINITY, xC = -INFINITY;
  for (int k = 0; k <= M; k++) R.at(0, 0, k) = -INFINITY;
  for (int i = 1; i <= L; i++) {
    const int cur = i & 1, prv = cur ^ 1;
    const float *rsc = a.msc + (size_t)seq[i - 1] * (M + 1);
    float xE = -INFINITY;
    float diag = R.at(prv, 0, 0);                                       // MMX(i-1,k-1)
    R.at(cur, 0, 0) = -INFINITY;
    const float bin = xB + a.tbmk;
    for (int k = 1; k <= M; k++) {
      const float up = R.at(prv, 0, k);
      const float m = __ldg(rsc + k) + MX2(diag, bin);
      R.at(cur, 0, k) = m;
      xE = MX2(xE, m);
      diag = up;
    }
    xJ = MX2(xJ + tloop, xE + a.tej);
    xC = MX2(xC + tloop, xE + a.tec);
    xN = xN + tloop;
    xB = MX2(xN + tmove, xJ + tmove);
  }
  return xC + tmove;
}

// p7_GViterbi (VIT = true) and p7_GForward (VIT = false) share their shape: a "pull" over rows
// FULL: keep every row (row index = i) and the special states of every row in xout [(L+1)][5] = E N J B C, for decoding
template <bool VIT, bool FULL = false>
__device__ float g_fwd(const GenArgs &a, const float *tbl, const Rows &R, const uint8_t *seq, int L, float xloop, float xmove, float *xout = nullptr)
{
  const int M = a.M;
  const float *tsc = a.tsc;
  auto OP = [&](float p, float q) { return VIT ? MX2(p, q) : LS(tbl, p, q); };
  float xN = 0.f, xB = xmove, xJ = -INFINITY, xC = -INFINITY;
  for (int k = 0; k <= M; k++) { R.at(0, 0, k) = -INFINITY; R.at(0, 1, k) = -INFINITY; R.at(0, 2, k) = -INFINITY; }
  if (FULL) { xout[0] = -INFINITY; xout[1] = xN; xout[2] = xJ; xout[3] = xB; xout[4] = xC; }
  for (int i = 1; i <= L; i++) {
    const int cur = FULL ? i : (i & 1), prv = FULL ? i - 1 : (cur ^ 1);
    const int x = seq[i - 1];
    const float *rsc = a.msc + (size_t)x * (M + 1);
    float xE = -INFINITY;
    float mpd = R.at(prv, 0, 0), ipd = R.at(prv, 1, 0), dpd = R.at(prv, 2, 0);    // row i-1, node k-1
    float mc1 = -INFINITY, dc1 = -INFINITY;                                         // row i,   node k-1
    R.at(cur, 0, 0) = -INFINITY; R.at(cur, 1, 0) = -INFINITY; R.at(cur, 2, 0) = -INFINITY;
    for (int k = 1; k <= M; k++) {
      const float *t1 = tsc + (size_t)(k - 1) * 8;
      const float mpk = R.at(prv, 0, k), ipk = R.at(prv, 1, k), dpk = R.at(prv, 2, k);
      float sc;
      if (VIT) {
        sc = MX2(mpd + t1[T_MM], ipd + t1[T_IM]);
        sc = MX2(sc, dpd + t1[T_DM]);
        sc = MX2(sc, xB + t1[T_BM]);
      } else {
        sc = LS(tbl, LS(tbl, mpd + t1[T_MM], ipd + t1[T_IM]), LS(tbl, xB + t1[T_BM], dpd + t1[T_DM]));
      }
      const float mck = sc + __ldg(rsc + k);
      float ick = -INFINITY;
      if (k < M) {
        const float *t0 = tsc + (size_t)k * 8;
        ick = OP(mpk + t0[T_MI], ipk + t0[T_II]) + isc(a, x, k);
      }
      const float dck = OP(mc1 + t1[T_MD], dc1 + t1[T_DD]);
      R.at(cur, 0, k) = mck; R.at(cur, 1, k) = ick; R.at(cur, 2, k) = dck;
      if (VIT) {
        if (k < M) xE = MX2(xE, mck);                                     // esc = 0: local
        else { const float s2 = MX2(xE, mck); xE = MX2(s2, dck); }
      } else {
        xE = LS(tbl, LS(tbl, mck, dck), xE);                             // (+ esc = 0 for k < M)
      }
      mpd = mpk; ipd = ipk; dpd = dpk; mc1 = mck; dc1 = dck;
    }
    if (VIT) {
      xJ = MX2(xJ + xloop, xE + a.xE_loop);
      xC = MX2(xC + xloop, xE + a.xE_move);
      xN = xN + xloop;
      xB = MX2(xN + xmove, xJ + xmove);
    } else {
      xJ = LS(tbl, xJ + xloop, xE + a.xE_loop);
      xC = LS(tbl, xC + xloop, xE + a.xE_move);
      xN = xN + xloop;
      xB = LS(tbl, xN + xmove, xJ + xmove);
    }
    if (FULL) { float *q = xout + (size_t)i * 5; q[0] = xE; q[1] = xN; q[2] = xJ; q[3] = xB; q[4] = xC; }
  }
  return xC + xmove;
}

template <bool FULL = false>
__device__ float g_bck(const GenArgs &a, const float *tbl, const Rows &R, const uint8_t *seq, int L, float xloop, float xmove, float *xout = nullptr)
{
  const int M = a.M;
  const float *tsc = a.tsc;
  // row L
  float xJ = -INFINITY, xN = -INFINITY, xC = xmove, xE = xC + a.xE_move;
  {
    const int r = FULL ? L : (L & 1);
    if (FULL) { float *q = xout + (size_t)L * 5; q[0] = xE; q[1] = xN; q[2] = xJ; q[3] = -INFINITY; q[4] = xC; }
    R.at(r, 0, M) = xE; R.at(r, 2, M) = xE; R.at(r, 1, M) = -INFINITY;
    float dn = xE;                                                        // DMX(L,k+1)
    for (int k = M - 1; k >= 1; k--) {
      const float *t0 = tsc + (size_t)k * 8;
      const float m = LS(tbl, xE, dn + t0[T_MD]);
      const float d = LS(tbl, xE, dn + t0[T_DD]);
      R.at(r, 0, k) = m; R.at(r, 2, k) = d; R.at(r, 1, k) = -INFINITY;
      dn = d;
    }
  }
  for (int i = L - 1; i >= 1; i--) {
    const int cur = FULL ? i : (i & 1), nxt = FULL ? i + 1 : (cur ^ 1);
    const int x1 = seq[i];                                               // residue x_{i+1} (0-based array)
    const float *rsc = a.msc + (size_t)x1 * (M + 1);
    float xB = R.at(nxt, 0, 1) + tsc[T_BM] + __ldg(rsc + 1);
    for (int k = 2; k <= M; k++) xB = LS(tbl, xB, R.at(nxt, 0, k) + tsc[(size_t)(k - 1) * 8 + T_BM] + __ldg(rsc + k));
    xJ = LS(tbl, xJ + xloop, xB + xmove);
    xC = xC + xloop;
    xE = LS(tbl, xJ + a.xE_loop, xC + a.xE_move);
    xN = LS(tbl, xN + xloop, xB + xmove);
    if (FULL) { float *q = xout + (size_t)i * 5; q[0] = xE; q[1] = xN; q[2] = xJ; q[3] = xB; q[4] = xC; }
    R.at(cur, 0, M) = xE; R.at(cur, 2, M) = xE; R.at(cur, 1, M) = -INFINITY;
    float dn = xE;                                                        // DMX(i,k+1)
    for (int k = M - 1; k >= 1; k--) {
      const float *t0 = tsc + (size_t)k * 8;
      const float mn1 = R.at(nxt, 0, k + 1), in0 = R.at(nxt, 1, k);
      const float e1 = __ldg(rsc + k + 1), ie = isc(a, x1, k);
      const float m = LS(tbl, LS(tbl, mn1 + t0[T_MM] + e1, in0 + t0[T_MI] + ie), LS(tbl, xE, dn + t0[T_MD]));
      const float iv = LS(tbl, mn1 + t0[T_IM] + e1, in0 + t0[T_II] + ie);
      const float d = LS(tbl, mn1 + t0[T_DM] + e1, LS(tbl, dn + t0[T_DD], xE));
      R.at(cur, 0, k) = m; R.at(cur, 1, k) = iv; R.at(cur, 2, k) = d;
      dn = d;
    }
  }
  // i = 0
  if (L >= 1) {
    const float *rsc = a.msc + (size_t)seq[0] * (M + 1);
    float xB = R.at(1, 0, 1) + tsc[T_BM] + __ldg(rsc + 1);
    for (int k = 2; k <= M; k++) xB = LS(tbl, xB, R.at(1, 0, k) + tsc[(size_t)(k - 1) * 8 + T_BM] + __ldg(rsc + k));
    xN = LS(tbl, xN + xloop, xB + xmove);
    if (FULL) { xout[0] = -INFINITY; xout[1] = xN; xout[2] = -INFINITY; xout[3] = xB; xout[4] = -INFINITY;
                for (int k = 0; k <= M; k++) { R.at(0, 0, k) = -INFINITY; R.at(0, 1, k) = -INFINITY; R.at(0, 2, k) = -INFINITY; } }
  }
  return xN;
}

// p7_GDecoding (generic_decoding.c:77-140) needs the full Forward and Backward matrices of ONE comparison: one thread
// fills them in the reference's order (the same functions as above, FULL), then one thread per row turns them into
// posterior probabilities, accumulating the row's normaliser in the reference's order.
struct DecArgs { GenArgs g; const uint8_t *seq; int L; float xloop, xmove; float *F, *B, *fx, *bx, *pp, *xpp, *sc; };

__global__ void generic_full_kernel(const DecArgs d)
{
  extern __shared__ float s_tbl[];
  for (int i = threadIdx.x; i < LOGSUM_TBL; i += blockDim.x) s_tbl[i] = d.g.tbl[i];
  __syncthreads();
  if (threadIdx.x != 0) return;
  Rows RF; RF.base = d.F; RF.stride = 1; RF.M1 = d.g.M + 1;
  Rows RB; RB.base = d.B; RB.stride = 1; RB.M1 = d.g.M + 1;
  d.sc[0] = g_fwd<false, true>(d.g, s_tbl, RF, d.seq, d.L, d.xloop, d.xmove, d.fx);
  d.sc[1] = g_bck<true>(d.g, s_tbl, RB, d.seq, d.L, d.xloop, d.xmove, d.bx);
}

__global__ void generic_decode_kernel(const DecArgs d)
{
  const int M = d.g.M, L = d.L, M1 = M + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > L) return;
  float *pp = d.pp + (size_t)i * M1 * 3, *xp = d.xpp + (size_t)i * 5;
  if (i == 0) { for (int k = 0; k <= M; k++) { pp[k*3] = pp[k*3+1] = pp[k*3+2] = 0.f; } for (int c = 0; c < 5; c++) xp[c] = 0.f; return; }
  const float overall = d.fx[(size_t)L * 5 + 4] + d.xmove;                 // XMX_fwd(L,C) + xsc[C][MOVE]
  auto F = [&](int st, int k) { return d.F[(((size_t)i * 3 + st) * M1 + k)]; };
  auto B = [&](int st, int k) { return d.B[(((size_t)i * 3 + st) * M1 + k)]; };
  float denom = 0.0f;
  pp[0] = pp[1] = pp[2] = 0.f;
  for (int k = 1; k < M; k++) {
    const float m = expf(F(0, k) + B(0, k) - overall); denom += m;
    const float iv = expf(F(1, k) + B(1, k) - overall); denom += iv;
    pp[k*3] = m; pp[k*3+1] = iv; pp[k*3+2] = 0.f;
  }
  { const float m = expf(F(0, M) + B(0, M) - overall); denom += m; pp[M*3] = m; pp[M*3+1] = 0.f; pp[M*3+2] = 0.f; }
  const float *fp = d.fx + (size_t)(i - 1) * 5, *bp = d.bx + (size_t)i * 5;
  float xn = expf(fp[1] + bp[1] + d.xloop - overall), xj = expf(fp[2] + bp[2] + d.xloop - overall), xc = expf(fp[4] + bp[4] + d.xloop - overall);
  denom += xn + xj + xc;
  denom = (float)(1.0 / (double)denom);
  for (int k = 1; k < M; k++) { pp[k*3] *= denom; pp[k*3+1] *= denom; }
  pp[M*3] *= denom;
  xp[0] = 0.f; xp[1] = xn * denom; xp[2] = xj * denom; xp[3] = 0.f; xp[4] = xc * denom;
}


__global__ void __launch_bounds__(128) generic_kernel(const GenArgs a)
{
  extern __shared__ float s_tbl[];
  for (int i = threadIdx.x; i < LOGSUM_TBL; i += blockDim.x) s_tbl[i] = a.tbl[i];
  __syncthreads();
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  Rows R; R.base = a.scratch + tid; R.stride = nthreads; R.M1 = a.M + 1;
  for (size_t e = tid; e < (size_t)a.sd.n; e += nthreads) {
    const int s = a.sd.order[e];
    const int L = a.sd.len[s];
    const uint8_t *seq = a.sd.res + a.sd.off[s];
    const float *lp = a.lenp + (size_t)s * 4;
    if (a.gmsv) a.gmsv[s] = g_msv(a, s_tbl, R, seq, L, lp[0], lp[1]);
    if (a.gvit) a.gvit[s] = g_fwd<true>(a, s_tbl, R, seq, L, lp[2], lp[3]);
    if (a.gfwd) a.gfwd[s] = g_fwd<false>(a, s_tbl, R, seq, L, lp[2], lp[3]);
    if (a.gbck) a.gbck[s] = (L >= 1) ? g_bck<false>(a, s_tbl, R, seq, L, lp[2], lp[3]) : -INFINITY;
  }
}

} // namespace

extern "C" int b2h_generic_scores(b2h_ctx *ctx, int M, int K, int Kp, const float *tsc, const float *msc, const float *xsc, float nj,
                                  const b2h_seqdb *db, float nu, float *gmsv, float *gviterbi, float *gforward, float *gbackward)
{
  if (!ctx || !db || db->ctx != ctx || !tsc || !msc || !xsc || M < 1 || Kp < 1 || Kp > B2H_NCODE - 1 || K < 1 || K >= Kp || !(nu > 0.f)) return B2H_EINVAL;
  const size_t n = db->n;
  if (n == 0) return B2H_OK;
  B2H_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // host-side constants with the reference's libm: the logsum table (logsum.c:87-88) and the length-dependent transitions
  std::vector<float> tbl(LOGSUM_TBL);
  for (int i = 0; i < LOGSUM_TBL; i++) tbl[i] = (float)log(1. + exp((double)-i / 1000.f));
  std::vector<float> lenp(n * 4);
  {
    std::map<int, size_t> seen;
    for (size_t s = 0; s < n; s++) {
      const int L = db->h_len[s];
      auto it = seen.find(L);
      if (it != seen.end()) { for (int c = 0; c < 4; c++) lenp[s * 4 + c] = lenp[it->second * 4 + c]; continue; }
      seen[L] = s;
      lenp[s * 4 + 0] = logf((float)L / (float)(L + 3));                       // generic_msv.c:62-63
      lenp[s * 4 + 1] = logf(3.0f / (float)(L + 3));
      const float pmove = (2.0f + nj) / ((float)L + 2.0f + nj), ploop = 1.0f - pmove;   // p7_ReconfigLength, modelconfig.c:228-231
      lenp[s * 4 + 2] = (float)log((double)ploop);    // C's log(double), as the reference calls it -- in a .cu file log(float) is logf
      lenp[s * 4 + 3] = (float)log((double)pmove);
    }
  }
  GenArgs a;
  a.M = M; a.K = K; a.Kp = Kp;
  a.xE_loop = xsc[0]; a.xE_move = xsc[1];
  a.tbmk = logf(2.0f / ((float)M * (float)(M + 1)));
  a.tej = logf((nu - 1.0f) / nu);
  a.tec = logf(1.0f / nu);
  a.sd = b2h_seqdev(db);
  const int threads = 128;
  int blocks = (int)std::min<size_t>((n + threads - 1) / threads, (size_t)ctx->sm_count * 8);
  const size_t nthreads = (size_t)blocks * threads;
  std::vector<void *> keep;
  auto dalloc = [&](void **p, size_t bytes) { cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 4, st); if (e == cudaSuccess) keep.push_back(*p); return e; };
  auto release = [&]() { for (void *p : keep) cudaFreeAsync(p, st); };
  float *d_tsc = nullptr, *d_msc = nullptr, *d_tbl = nullptr, *d_lenp = nullptr, *d_scratch = nullptr, *d_out[4] = {nullptr, nullptr, nullptr, nullptr};
  float *host_out[4] = {gmsv, gviterbi, gforward, gbackward};
  cudaError_t e = cudaSuccess;
  if ((e = dalloc((void **)&d_tsc, (size_t)M * 8 * 4)) != cudaSuccess || (e = dalloc((void **)&d_msc, (size_t)Kp * (M + 1) * 4)) != cudaSuccess ||
      (e = dalloc((void **)&d_tbl, LOGSUM_TBL * 4)) != cudaSuccess || (e = dalloc((void **)&d_lenp, n * 16)) != cudaSuccess ||
      (e = dalloc((void **)&d_scratch, (size_t)6 * (M + 1) * nthreads * 4)) != cudaSuccess) {
    ctx->err = std::string("generic DP: ") + cudaGetErrorString(e); release(); return B2H_EMEM;
  }
  for (int c = 0; c < 4; c++) if (host_out[c] && (e = dalloc((void **)&d_out[c], n * 4)) != cudaSuccess) { ctx->err = cudaGetErrorString(e); release(); return B2H_EMEM; }
  cudaMemcpyAsync(d_tsc, tsc, (size_t)M * 8 * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_msc, msc, (size_t)Kp * (M + 1) * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_tbl, tbl.data(), LOGSUM_TBL * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_lenp, lenp.data(), n * 16, cudaMemcpyHostToDevice, st);
  a.tsc = d_tsc; a.msc = d_msc; a.tbl = d_tbl; a.lenp = d_lenp; a.scratch = d_scratch;
  a.gmsv = d_out[0]; a.gvit = d_out[1]; a.gfwd = d_out[2]; a.gbck = d_out[3];
  const size_t smem = LOGSUM_TBL * sizeof(float);
  int occ = 1;
  { const int rc = b2h_kernel_occupancy(ctx, (const void *)generic_kernel, threads, smem, &occ); if (rc != B2H_OK) { release(); return rc; } }
  generic_kernel<<<blocks, threads, smem, st>>>(a);
  ctx->launches++;
  for (int c = 0; c < 4; c++) if (host_out[c]) cudaMemcpyAsync(host_out[c], d_out[c], n * 4, cudaMemcpyDeviceToHost, st);
  e = cudaStreamSynchronize(st);
  release();
  if (e != cudaSuccess) { ctx->err = std::string("generic DP kernel: ") + cudaGetErrorString(e); return B2H_ECUDA; }
  return B2H_OK;
}

extern "C" int b2h_generic_decoding(b2h_ctx *ctx, int M, int K, int Kp, const float *tsc, const float *msc, const float *xsc, float nj,
                                    const uint8_t *residues, int L, float *pp_dp, float *pp_xmx, float *fwdsc, float *bcksc,
                                    float *dom_btot, float *dom_etot, float *dom_mocc)
{
  if (!ctx || !tsc || !msc || !xsc || !residues || !pp_dp || !pp_xmx || M < 1 || L < 1 || Kp < 1 || Kp > B2H_NCODE - 1 || K < 1 || K >= Kp) return B2H_EINVAL;
  B2H_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  std::vector<float> tbl(LOGSUM_TBL);
  for (int i = 0; i < LOGSUM_TBL; i++) tbl[i] = (float)log(1. + exp((double)-i / 1000.f));
  const float pmove = (2.0f + nj) / ((float)L + 2.0f + nj), ploop = 1.0f - pmove;
  DecArgs d;
  d.g.M = M; d.g.K = K; d.g.Kp = Kp; d.g.xE_loop = xsc[0]; d.g.xE_move = xsc[1]; d.g.tbmk = d.g.tej = d.g.tec = 0.f;
  d.g.lenp = nullptr; d.g.scratch = nullptr; d.g.gmsv = d.g.gvit = d.g.gfwd = d.g.gbck = nullptr;
  d.L = L; d.xloop = (float)log((double)ploop); d.xmove = (float)log((double)pmove);
  const size_t cells = (size_t)(L + 1) * 3 * (M + 1), xc = (size_t)(L + 1) * 5;
  std::vector<void *> keep;
  auto dalloc = [&](void **p, size_t bytes) { cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 4, st); if (e == cudaSuccess) keep.push_back(*p); return e; };
  auto release = [&]() { for (void *p : keep) cudaFreeAsync(p, st); };
  float *d_tsc, *d_msc, *d_tbl; uint8_t *d_seq;
  cudaError_t e;
  if ((e = dalloc((void **)&d_tsc, (size_t)M * 32)) != cudaSuccess || (e = dalloc((void **)&d_msc, (size_t)Kp * (M + 1) * 4)) != cudaSuccess ||
      (e = dalloc((void **)&d_tbl, LOGSUM_TBL * 4)) != cudaSuccess || (e = dalloc((void **)&d_seq, (size_t)L)) != cudaSuccess ||
      (e = dalloc((void **)&d.F, cells * 4)) != cudaSuccess || (e = dalloc((void **)&d.B, cells * 4)) != cudaSuccess ||
      (e = dalloc((void **)&d.fx, xc * 4)) != cudaSuccess || (e = dalloc((void **)&d.bx, xc * 4)) != cudaSuccess ||
      (e = dalloc((void **)&d.pp, cells * 4)) != cudaSuccess || (e = dalloc((void **)&d.xpp, xc * 4)) != cudaSuccess ||
      (e = dalloc((void **)&d.sc, 8)) != cudaSuccess) {
    ctx->err = std::string("generic decoding: ") + cudaGetErrorString(e); release(); return B2H_EMEM;
  }
  cudaMemcpyAsync(d_tsc, tsc, (size_t)M * 32, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_msc, msc, (size_t)Kp * (M + 1) * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_tbl, tbl.data(), LOGSUM_TBL * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_seq, residues, (size_t)L, cudaMemcpyHostToDevice, st);
  d.g.tsc = d_tsc; d.g.msc = d_msc; d.g.tbl = d_tbl; d.seq = d_seq;
  const size_t smem = LOGSUM_TBL * sizeof(float);
  int occ = 1;
  { const int rc = b2h_kernel_occupancy(ctx, (const void *)generic_full_kernel, 128, smem, &occ); if (rc != B2H_OK) { release(); return rc; } }
  generic_full_kernel<<<1, 128, smem, st>>>(d);
  generic_decode_kernel<<<(L + 1 + 127) / 128, 128, 0, st>>>(d);
  ctx->launches += 2;
  // the reference indexes its matrices [i][k][s]; ours are [i][s][k] on the device and are transposed by the decode kernel's output layout
  cudaMemcpyAsync(pp_dp, d.pp, cells * 4, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(pp_xmx, d.xpp, xc * 4, cudaMemcpyDeviceToHost, st);
  float sc[2] = {0.f, 0.f};
  cudaMemcpyAsync(sc, d.sc, 8, cudaMemcpyDeviceToHost, st);
  std::vector<float> hfx, hbx;
  if (dom_btot || dom_etot || dom_mocc) {
    hfx.resize(xc); hbx.resize(xc);
    cudaMemcpyAsync(hfx.data(), d.fx, xc * 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(hbx.data(), d.bx, xc * 4, cudaMemcpyDeviceToHost, st);
  }
  e = cudaStreamSynchronize(st);
  release();
  if (e != cudaSuccess) { ctx->err = std::string("generic decoding kernels: ") + cudaGetErrorString(e); return B2H_ECUDA; }
  if (fwdsc) *fwdsc = sc[0];
  if (bcksc) *bcksc = sc[1];
  if (!hfx.empty()) {
    // p7_GDomainDecoding (generic_decoding.c:207-230) from the special-state rows, on the host with the reference's libm calls
    const float *fx = hfx.data(), *bx = hbx.data();
    const float overall = fx[(size_t)L * 5 + 4] + d.xmove;
    float bt = 0.f, et = 0.f;
    if (dom_btot) dom_btot[0] = 0.f;
    if (dom_etot) dom_etot[0] = 0.f;
    if (dom_mocc) dom_mocc[0] = 0.f;
    for (int i = 1; i <= L; i++) {
      bt = (float)((double)bt + exp((double)(fx[(size_t)(i - 1) * 5 + 3] + bx[(size_t)(i - 1) * 5 + 3] - overall)));
      et = (float)((double)et + exp((double)(fx[(size_t)i * 5 + 0] + bx[(size_t)i * 5 + 0] - overall)));
      float njcp = expf(fx[(size_t)(i - 1) * 5 + 1] + bx[(size_t)i * 5 + 1] + d.xloop - overall);
      njcp += expf(fx[(size_t)(i - 1) * 5 + 2] + bx[(size_t)i * 5 + 2] + d.xloop - overall);
      njcp += expf(fx[(size_t)(i - 1) * 5 + 4] + bx[(size_t)i * 5 + 4] + d.xloop - overall);
      if (dom_btot) dom_btot[i] = bt;
      if (dom_etot) dom_etot[i] = et;
      if (dom_mocc) dom_mocc[i] = (float)(1. - (double)njcp);
    }
  }
  return B2H_OK;
}
