// b2h_ltvit.cu -- the Viterbi stage of the long-target (nhmmer) pipeline: p7_ViterbiFilter_longtarget
// (vendor/hmmer/src/impl_sse/vitfilter.c:292-497) over the windows that survived the SSV / MSV / bias gates, followed by
// what p7_pli_postSSV_LongTarget does with its landmarks (p7_pipeline.c:1385-1412): p7_pli_ExtendAndMergeWindows(.., 0.5)
// and the cut of windows longer than 80 kb into overlapping pieces.
//
// The reference scans a window row by row with the ViterbiFilter recurrence (16-bit, lazy-F D->D) and, instead of
// returning one score, records a landmark (i, k) for every node k whose match cell equals the row maximum xE on a row
// where xE reaches the score threshold of P-value F2 -- then resets the three DP rows and scans on, the special states
// keeping their values.  The reset makes a window one sequential scan; the parallelism is over windows (thousands per
// genome) and over the model: the register-resident tile of rvit_kernel (b2h_dpreg.cu) -- lane z of a W-warp group owns
// C consecutive nodes with their cells and transition scores in registers, the emission row of the current residue is
// one conflict-free vector load from a table staged in shared memory by a TMA bulk copy, the (i-1, k-1) dependency is
// three shuffles, the D->D closure a max-plus scan -- with the threshold test between the row and its special states.
// Values stay below the int16 ceiling because a row at or above the threshold is reset (no overflow test in the
// reference either), so the int32 arithmetic with a floor at -32768 equals the saturating SSE arithmetic.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "b2h_internal.h"

namespace {

constexpr int NEG16 = -32768;
constexpr uint32_t FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
    "{\n .reg .pred p;\n"
    "WAIT_%=:\n"
    " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    " @p bra DONE_%=;\n"
    " bra WAIT_%=;\n"
    "DONE_%=:\n}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// The C emission scores of this lane for residue x; table [32 residues][W warps][C/G groups][32 lanes][G] as built by
// b2h_profile_upload for the register-resident DP kernels (G = 4, 2 or 1 by the divisibility of C).
template <int C>
__device__ __forceinline__ void load_emis(const int *tab, int stride, int x, int lane, int (&r)[C])
{
  const int *row = tab + (size_t)x * stride;
  if (C % 2 != 0) {
#pragma unroll
    for (int g = 0; g < C; g++) r[g] = row[g * 32 + lane];
  } else if (C % 4 != 0) {
#pragma unroll
    for (int g = 0; g < C / 2; g++) {
      const int2 v = *reinterpret_cast<const int2 *>(row + g * 64 + lane * 2);
      r[2*g+0] = v.x; r[2*g+1] = v.y;
    }
  } else {
#pragma unroll
    for (int g = 0; g < C / 4; g++) {
      const int4 v = *reinterpret_cast<const int4 *>(row + g * 128 + lane * 4);
      r[4*g+0] = v.x; r[4*g+1] = v.y; r[4*g+2] = v.z; r[4*g+3] = v.w;
    }
  }
}

template <int W>
__device__ __forceinline__ void group_sync(int grp)
{
  if (W > 1) asm volatile("bar.sync %0, %1;" :: "r"(grp + 1), "r"(W * 32) : "memory");
}

struct LtMark { int32_t win, i, k; };

struct LtVitArgs {
  ProfDev P;
  SeqDev sd;                    // the windows, one "sequence" each
  const int32_t *order;         // [nwork] window indices, longest first
  const int32_t *thresh;        // [n] score threshold per window (an int16 value)
  const int32_t *xwmove;        // [n] N/C/J move score of the length the window's profile is configured for
  int nwork;
  int *counter;
  LtMark *marks; int *nmarks; int cap;
};

template <int C, int W>
__global__ void __launch_bounds__(256) lt_vit_kernel(const LtVitArgs a)
{
  extern __shared__ __align__(128) int s_rsc[];            // [32][W][32*C] int32 emission scores
  __shared__ uint64_t s_bar;
  __shared__ int s_w[8];
  __shared__ int s_x[2][4][6][W];                           // per row parity: M, I, D, M+tMD of each warp's last node; xE, Dmax partials
  __shared__ int s_y[4][2][W];                              // max-plus composite (A, T) of each warp's D chain
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = warp / W, wi = warp % W;
  const int gl = wi * 32 + lane;                            // lane index inside the group: owns nodes gl*C .. gl*C+C-1 (0-based)
  constexpr int STRIDE = 32 * C * W;
  constexpr uint32_t TAB_BYTES = 32u * STRIDE * 4u;
  const ProfDev &P = a.P;
  // W == 8 (models of 1537 .. 3072 nodes): the lane-grouped table would not fit in shared memory (393 KB); the emission scores
  // are read from the node-major int16 table in global memory instead (L1 / L2 resident: 2 * Mpad bytes per residue row)
  constexpr bool GLOBAL_TAB = (W == 8);
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (!GLOBAL_TAB && threadIdx.x == 0) { mbar_expect_tx(&s_bar, TAB_BYTES); tma_load_1d(s_rsc, P.vit_rsc32, TAB_BYTES, &s_bar); }
  int tBM[C], tMM[C], tIM[C], tDM[C], tMD[C], tMI[C], tII[C], tDD[C];
  int tDDin;
  {
    const int16_t *ts = P.vit_tsc; const int Mp = P.Mpad;
#pragma unroll
    for (int c = 0; c < C; c++) {
      const int k0 = gl * C + c;
      const bool in = k0 < Mp;
      tBM[c] = in ? ts[0 * Mp + k0] : NEG16; tMM[c] = in ? ts[1 * Mp + k0] : NEG16; tIM[c] = in ? ts[2 * Mp + k0] : NEG16;
      tDM[c] = in ? ts[3 * Mp + k0] : NEG16; tMD[c] = in ? ts[4 * Mp + k0] : NEG16; tMI[c] = in ? ts[5 * Mp + k0] : NEG16;
      tII[c] = in ? ts[6 * Mp + k0] : NEG16; tDD[c] = in ? ts[7 * Mp + k0] : NEG16;
    }
    tDDin = __shfl_up_sync(FULL, tDD[C - 1], 1);
    if (lane == 0) tDDin = (W > 1 && wi > 0 && gl * C - 1 < Mp) ? (int)ts[7 * Mp + gl * C - 1] : NEG16;
  }
  if (!GLOBAL_TAB) mbar_wait(&s_bar, 0);
  const int xwEm = P.xw_E_move, xwEl = P.xw_E_loop, base_w = P.base_w, ddbound = P.ddbound_w, Mnodes = P.M;
  const int *my_rsc = s_rsc + wi * 32 * C;

  for (;;) {
    int wk;
    if (W == 1) { wk = (lane == 0) ? atomicAdd(a.counter, 1) : 0; wk = __shfl_sync(FULL, wk, 0); }
    else {
      group_sync<W>(grp);                                   // everyone is done with the previous window's buffers
      if (wi == 0 && lane == 0) s_w[grp] = atomicAdd(a.counter, 1);
      group_sync<W>(grp);
      wk = s_w[grp];
    }
    if (wk >= a.nwork) break;
    const int s = a.order[wk];
    const int L = a.sd.len[s];
    const uint8_t *res = a.sd.res + a.sd.off[s];
    const int xw_move = a.xwmove[s], thresh = a.thresh[s];
    int M[C], I[C], D[C];
#pragma unroll
    for (int c = 0; c < C; c++) { M[c] = NEG16; I[c] = NEG16; D[c] = NEG16; }
    const int xN = base_w;
    int xB = (int16_t)(xN + xw_move), xJ = NEG16, xC = NEG16;
    int cM = NEG16, cI = NEG16, cD = NEG16;                // W > 1: previous row's cells of the left warp's last node
    int myres = 0;

    for (int i = 0; i < L; i++) {
      if ((i & 31) == 0) myres = (i + lane < L) ? (int)__ldg(res + i + lane) : 0;
      const int x = __shfl_sync(FULL, myres, i & 31) & 31;
      int r[C];
      if (GLOBAL_TAB) {
        const int16_t *row = P.vit_rsc + (size_t)x * P.Mpad + gl * C;
#pragma unroll
        for (int c = 0; c < C; c++) r[c] = (gl * C + c < P.Mpad) ? (int)__ldg(row + c) : NEG16;
      } else load_emis<C>(my_rsc, STRIDE, x, lane, r);
      int mp = __shfl_up_sync(FULL, M[C - 1], 1), ip = __shfl_up_sync(FULL, I[C - 1], 1), dp = __shfl_up_sync(FULL, D[C - 1], 1);
      if (lane == 0) { mp = cM; ip = cI; dp = cD; }
      int xEm = NEG16;
#pragma unroll
      for (int c = C - 1; c >= 0; c--) {
        const int pm = (c == 0) ? mp : M[c - 1], pi = (c == 0) ? ip : I[c - 1], pd = (c == 0) ? dp : D[c - 1];
        const int inew = __viaddmax_s32(M[c], tMI[c], __viaddmax_s32(I[c], tII[c], NEG16));
        int m = __viaddmax_s32(xB, tBM[c], NEG16);
        m = __viaddmax_s32(pm, tMM[c], m);
        m = __viaddmax_s32(pi, tIM[c], m);
        m = __viaddmax_s32(pd, tDM[c], m);
        m = __viaddmax_s32(m, r[c], NEG16);
        M[c] = m; I[c] = inew;
        xEm = max(xEm, m);
      }
      int xE = __reduce_max_sync(FULL, xEm);
      // M->D partials: D[c] is the value entering node c from M of node c-1
      const int mdl = __viaddmax_s32(M[C - 1], tMD[C - 1], NEG16);
      int dleft = __shfl_up_sync(FULL, mdl, 1);
      if (lane == 0) dleft = NEG16;
      int Dm = max(dleft, (W == 1 && lane == 31) ? mdl : NEG16);      // the last lane's own M->D value counts for Dmax too
      D[0] = dleft;
#pragma unroll
      for (int c = 1; c < C; c++) { D[c] = __viaddmax_s32(M[c - 1], tMD[c - 1], NEG16); Dm = max(Dm, D[c]); }
      int Dmax = __reduce_max_sync(FULL, Dm);
      if (W > 1) {
        int (*X)[W] = s_x[i & 1][grp];
        if (lane == 31) { X[0][wi] = M[C - 1]; X[1][wi] = I[C - 1]; X[2][wi] = D[C - 1]; X[3][wi] = mdl; }
        if (lane == 0)  { X[4][wi] = xE; X[5][wi] = Dmax; }
        group_sync<W>(grp);
        xE = X[4][0]; Dmax = X[5][0];
#pragma unroll
        for (int w = 1; w < W; w++) { xE = max(xE, X[4][w]); Dmax = max(Dmax, max(X[5][w], X[3][w - 1])); }
        Dmax = max(Dmax, X[3][W - 1]);
        if (wi > 0) {
          cM = X[0][wi - 1]; cI = X[1][wi - 1]; cD = X[2][wi - 1];
          if (lane == 0) D[0] = X[3][wi - 1];
        }
      }
      if (xE >= thresh) {
        // a landmark for every node whose match cell carries the row maximum, then the rows start over (vitfilter.c:411-425)
#pragma unroll
        for (int c = 0; c < C; c++) {
          const int k = gl * C + c + 1;
          if (M[c] == xE && k <= Mnodes) {
            const int idx = atomicAdd(a.nmarks, 1);
            if (idx < a.cap) { LtMark mk; mk.win = s; mk.i = i + 1; mk.k = k; a.marks[idx] = mk; }
          }
          M[c] = NEG16; I[c] = NEG16; D[c] = NEG16;
        }
        cM = NEG16; cI = NEG16; cD = NEG16;
        continue;
      }
      xC = (int16_t)max(xC, xE + xwEm);
      xJ = (int16_t)max(xJ, xE + xwEl);
      xB = (int16_t)max(xJ + xw_move, xN + xw_move);
      if (Dmax + ddbound > xB) {
        // close the D->D chain: serial inside the lane, then a max-plus scan over the 32 lane composites
        int T[C];
        T[0] = tDDin;
#pragma unroll
        for (int c = 1; c < C; c++) { D[c] = __viaddmax_s32(D[c - 1], tDD[c - 1], D[c]); T[c] = T[c - 1] + tDD[c - 1]; }
        int A = D[C - 1], Tt = T[C - 1];                     // lane composite: d -> max(A, d + Tt)
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
          const int A2 = __shfl_up_sync(FULL, A, dlt), T2 = __shfl_up_sync(FULL, Tt, dlt);
          if (lane >= dlt) { A = max(A, A2 + Tt); Tt = max(T2 + Tt, -(1 << 29)); }
        }
        int din = __shfl_up_sync(FULL, A, 1);                // D of the previous lane's last node
        if (lane == 0) din = NEG16;
        if (W > 1) {
          int (*Y)[W] = s_y[grp];
          if (lane == 31) { Y[0][wi] = A; Y[1][wi] = Tt; }
          const int Tp = __shfl_up_sync(FULL, Tt, 1);
          group_sync<W>(grp);
          int d = NEG16;                                     // D of the last node of the warp to the left, closed
#pragma unroll
          for (int w = 0; w < W - 1; w++) if (w < wi) d = max(Y[0][w], max(d + Y[1][w], NEG16));
          if (wi > 0) { cD = d; din = (lane == 0) ? d : max(din, max(d + Tp, NEG16)); }
        }
#pragma unroll
        for (int c = 0; c < C; c++) D[c] = max(D[c], max(din + T[c], NEG16));
      }
    }
    (void)xC;
  }
}

template <int C, int W>
int launch_ltvit(b2h_ctx *ctx, const LtVitArgs &a, cudaStream_t strm)
{
  const size_t smem = (W == 8) ? 0 : (size_t)32 * 32 * C * W * 4;
  int occ = 1;
  { const int st = b2h_kernel_occupancy(ctx, (const void *)lt_vit_kernel<C, W>, 256, smem, &occ); if (st != B2H_OK) return st; }
  const int groups = 8 / W;
  int grid = ctx->sm_count * occ;
  grid = std::min(grid, (a.nwork + groups - 1) / groups);
  if (grid < 1) grid = 1;
  lt_vit_kernel<C, W><<<grid, 256, smem, strm>>>(a);
  ctx->launches++;
  B2H_CUDA(cudaGetLastError());
  return B2H_OK;
}

// esl_gumbel_invsurv (vendor/easel/esl_gumbel.c:185)
double gumbel_invsurv(double p, double mu, double lambda)
{
  const double log_part = (p < 5e-9) ? (pow(p, p) - 1) / p : log(-1. * log(1 - p));
  return mu - (log_part / lambda);
}

} // namespace

// The score threshold p7_ViterbiFilter_longtarget derives from a P-value and the window's null / bias score
// (vitfilter.c:330-346: float invP, the sum in double, truncation to int16), and the N/C/J move score of cfg_len.
extern "C" int b2h_longtarget_vit_threshold(const b2h_profile *p, int cfg_len, float filtersc, double F2, int32_t *thresh, int32_t *xw_move)
{
  if (!p || !thresh || !xw_move || cfg_len < 1) return B2H_EINVAL;
  b2h_len_params lp;
  b2h_length_params(cfg_len, 1.0f, &lp);
  const float invP = (float)gumbel_invsurv(F2, (double)p->evparam[2], (double)p->evparam[3]);     // p7_VMU, p7_VLAMBDA
  const double t = ceil((((double)filtersc + (0.69314718055994529 * (double)invP) + 3.0) * (double)p->scale_w)
                        - (double)(float)p->xw[0][0] - (double)(float)lp.xw_move + (double)(float)p->base_w);
  *thresh = (int32_t)(int16_t)(int)t;
  *xw_move = lp.xw_move;
  return B2H_OK;
}

// What p7_pli_postSSV_LongTarget does with the landmarks of its windows (p7_pipeline.c:1385-1406), host only: <marks> {seq =
// window, n = row, k, length 1} are put into the reference's order in place -- window by window, row by row, and inside a row
// the striped scan of its 8-lane vectors (q outer, lane z inner, node k = q + Q*z + 1 with Q = p7O_NQW(M)) -- then extended and
// merged (p7_pli_ExtendAndMergeWindows(.., 0.5)) and cut at 80 kb.  window_len[w] = length of window w.
extern "C" int b2h_longtarget_vit_finish(const b2h_profile *p, b2h_window *marks, size_t nm, const int32_t *window_len, size_t nwindows,
                                         b2h_window **out, size_t *nout)
{
  if (!p || (!marks && nm) || (!window_len && nwindows) || !out || !nout || p->max_length <= 0) return B2H_EINVAL;
  *out = nullptr; *nout = 0;
  for (size_t i = 0; i < nm; i++)
    if (marks[i].seq < 0 || (size_t)marks[i].seq >= nwindows || marks[i].k < 1 || marks[i].k > p->M || marks[i].length != 1) return B2H_EINVAL;
  const int Q = std::max(2, (p->M - 1) / 8 + 1);
  auto key = [Q](int k) { return ((k - 1) % Q) * 8 + (k - 1) / Q; };
  std::sort(marks, marks + nm, [&](const b2h_window &x, const b2h_window &y) {
    if (x.seq != y.seq) return x.seq < y.seq;
    if (x.n != y.n) return x.n < y.n;
    return key(x.k) < key(y.k);
  });
  std::vector<b2h_window> w(marks, marks + nm);
  std::vector<int64_t> tlen(nm);
  for (size_t i = 0; i < nm; i++) tlen[i] = window_len[marks[i].seq];
  const size_t nmerged = b2h_extend_merge(p, w.data(), nm, tlen.data(), 0.5f);
  w.resize(nmerged);
  // windows above 80 kb are cut into 80 kb pieces that overlap by min(40 kb, max_length); the pieces go to the END of their
  // window's list, as p7_hmmwindow_new appends them
  const int max_window_len = 80000, overlap_len = std::min(40000, p->max_length);
  std::vector<b2h_window> res;
  res.reserve(nmerged + 8);
  for (size_t lo = 0; lo < nmerged;) {
    size_t hi = lo;
    while (hi < nmerged && w[hi].seq == w[lo].seq) hi++;
    std::vector<b2h_window> extra;
    for (size_t i = lo; i < hi; i++) {
      b2h_window x = w[i];
      if (x.length > max_window_len) {
        int64_t new_n = x.n; uint32_t new_len = (uint32_t)x.length;
        x.length = max_window_len;
        do {
          const int shift = max_window_len - overlap_len;
          new_n += shift; new_len -= shift;
          b2h_window y; y.seq = x.seq; y.k = 0; y.n = new_n; y.length = (int32_t)std::min<uint32_t>((uint32_t)max_window_len, new_len); y.score = 0.f;
          extra.push_back(y);
        } while (new_len > (uint32_t)max_window_len);
      }
      res.push_back(x);
    }
    // (the reference's loop also visits the pieces it has just appended; none of them is above 80 kb)
    res.insert(res.end(), extra.begin(), extra.end());
    lo = hi;
  }
  b2h_window *o = (b2h_window *)malloc(std::max<size_t>(1, res.size()) * sizeof(b2h_window));
  if (!o) return B2H_EMEM;
  if (!res.empty()) memcpy(o, res.data(), res.size() * sizeof(b2h_window));
  *out = o; *nout = res.size();
  return B2H_OK;
}

extern "C" int b2h_longtarget_viterbi_windows(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *windows, const float *filtersc,
                                              const uint8_t *active, double F2,
                                              b2h_window **marks_out, size_t *nmarks_out, b2h_window **out, size_t *nout)
{
  if (!ctx || !p || !windows || p->ctx != ctx || windows->ctx != ctx || !filtersc || !marks_out || !nmarks_out || !out || !nout) return B2H_EINVAL;
  *marks_out = *out = nullptr; *nmarks_out = *nout = 0;
  if (p->max_length <= 0) { ctx->err = "long-target search needs the model's max_length (MAXL)"; return B2H_EINVAL; }
  if (!p->regC && p->M > 3072) { ctx->err = "long-target Viterbi: models above 3072 nodes are not supported"; return B2H_EINVAL; }
  const int cls = p->regC ? p->regW * 64 + p->regC : 8 * 64 + 12;     // above 1536 nodes: 8 warps x 32 lanes x 12 nodes, table in global memory
  const size_t n = windows->n;
  if (n == 0) return B2H_OK;
  B2H_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // per-window threshold and length model: min(window, max_length) (p7_pipeline.c:1369-1385)
  std::vector<int32_t> hbuf(3 * n);
  int32_t *thr = hbuf.data(), *xwm = thr + n, *ord = xwm + n;
  size_t nwork = 0;
  int64_t work_res = 0;
  for (size_t i = 0; i < n; i++) {
    thr[i] = 32767; xwm[i] = 0;
    if (active && !active[i]) continue;
    const int64_t len = windows->h_len[i];
    if (len < 1) continue;
    const int rc = b2h_longtarget_vit_threshold(p, (int)std::min<int64_t>(len, p->max_length), filtersc[i], F2, &thr[i], &xwm[i]);
    if (rc != B2H_OK) return rc;
    ord[nwork++] = (int32_t)i;
    work_res += len;
  }
  std::vector<LtMark> hm;
  if (nwork > 0) {
    std::stable_sort(ord, ord + nwork, [&](int32_t x, int32_t y) { return windows->h_len[x] > windows->h_len[y]; });
    LtVitArgs a;
    a.P = b2h_profdev(p); a.sd = b2h_seqdev(windows); a.nwork = (int)nwork; a.counter = ctx->d_counters;
    a.cap = (int)std::min<int64_t>((int64_t)1 << 26, std::max<int64_t>(4096, work_res / 8 + 4096));
    int32_t *d_buf = nullptr; int *d_nm = nullptr;
    cudaError_t e;
    if ((e = cudaMallocAsync((void **)&d_buf, 3 * n * sizeof(int32_t), st)) != cudaSuccess ||
        (e = cudaMallocAsync((void **)&d_nm, sizeof(int), st)) != cudaSuccess) {
      ctx->err = cudaGetErrorString(e);
      if (d_buf) cudaFreeAsync(d_buf, st);
      return B2H_EMEM;
    }
    cudaMemcpyAsync(d_buf, hbuf.data(), 3 * n * sizeof(int32_t), cudaMemcpyHostToDevice, st);
    a.thresh = d_buf; a.xwmove = d_buf + n; a.order = d_buf + 2 * n; a.nmarks = d_nm;
    int rc = B2H_OK;
    // the scan is deterministic: when the landmark list overflows, the count it reports sizes the second attempt
    for (int attempt = 0; attempt < 2 && rc == B2H_OK; attempt++) {
      LtMark *d_marks = nullptr;
      if ((e = cudaMallocAsync((void **)&d_marks, (size_t)a.cap * sizeof(LtMark), st)) != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = B2H_EMEM; break; }
      a.marks = d_marks;
      cudaMemsetAsync(d_nm, 0, sizeof(int), st);
      cudaMemsetAsync(a.counter, 0, sizeof(int), st);
      rc = B2H_EINVAL;
      switch (cls) {
#define CASE(CC, WW) case (WW) * 64 + (CC): rc = launch_ltvit<CC, WW>(ctx, a, st); break;
        CASE(2, 1) CASE(3, 1) CASE(4, 1) CASE(5, 1) CASE(6, 1) CASE(7, 1) CASE(8, 1)
        CASE(9, 1) CASE(10, 1) CASE(11, 1) CASE(12, 1) CASE(14, 1) CASE(16, 1)
        CASE(9, 2) CASE(10, 2) CASE(11, 2) CASE(12, 2) CASE(14, 2) CASE(16, 2)
        CASE(10, 4) CASE(12, 4) CASE(12, 8)
#undef CASE
      }
      int nm = 0;
      bool again = false;
      if (rc == B2H_OK) {
        cudaMemcpyAsync(&nm, d_nm, sizeof(int), cudaMemcpyDeviceToHost, st);
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { ctx->err = std::string("long-target Viterbi kernel: ") + cudaGetErrorString(e); rc = B2H_ECUDA; }
        else if (nm > a.cap) {
          if (attempt == 0) { a.cap = nm; again = true; }
          else { ctx->err = "long-target Viterbi: landmark list overflow"; rc = B2H_ERANGE; }
        }
        else if (nm > 0) {
          hm.resize(nm);
          cudaMemcpyAsync(hm.data(), d_marks, (size_t)nm * sizeof(LtMark), cudaMemcpyDeviceToHost, st);
          if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = B2H_ECUDA; }
        }
      }
      cudaFreeAsync(d_marks, st);
      if (!again) break;
    }
    cudaFreeAsync(d_buf, st); cudaFreeAsync(d_nm, st);
    if (rc != B2H_OK) return rc;
  }
  const size_t nm = hm.size();
  b2h_window *marks = (b2h_window *)malloc(std::max<size_t>(1, nm) * sizeof(b2h_window));
  if (!marks) return B2H_EMEM;
  for (size_t i = 0; i < nm; i++) { b2h_window x; x.seq = hm[i].win; x.k = hm[i].k; x.n = hm[i].i; x.length = 1; x.score = 0.f; marks[i] = x; }
  const int rc = b2h_longtarget_vit_finish(p, marks, nm, windows->h_len.data(), n, out, nout);
  if (rc != B2H_OK) { free(marks); return rc; }
  *marks_out = marks; *nmarks_out = nm;
  return B2H_OK;
}
