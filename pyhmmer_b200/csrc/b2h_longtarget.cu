// b2h_longtarget.cu -- first stage of the long-target (nhmmer) pipeline: p7_SSVFilter_longtarget
// (vendor/hmmer/src/impl_sse/msvfilter.c:256-413) followed by p7_pli_ExtendAndMergeWindows (p7_pipeline.c:323-400) with
// the prefix/suffix lengths of p7_hmm_ScoreDataComputeRest (p7_scoredata.c:313-381), as p7_Pipeline_LongTarget runs
// them for every chunk of a long target (p7_pipeline.c:1535-1565).
//
// The reference scans a chunk row by row with the MSV recurrence at a constant begin score (no J state); the first row
// on which a cell reaches the significance threshold ends a "diagonal": the best cell is located, the diagonal is
// walked back to where it left the begin score and forward while it keeps improving (a scalar walk over the byte costs),
// a window {start, model end, length, score} is recorded, the DP row is zeroed and the scan resumes behind the diagonal.
// The reset makes the scan order part of the result, so a chunk is one sequential scan here too -- the parallelism is
// over chunks (a 100 Mb genome on both strands is ~800 chunks of 262 144) and over the model (the SSV register tiles):
// every group of G lanes scans one chunk with the cell arithmetic of rmsv_kernel (fp16x2, exact for byte values), keeps
// its own row pointer, and the rare hit handling runs on the group's first lane.
#include <chrono>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>
#include "b2h_internal.h"

namespace {

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
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v;
}
template <int G, int NR>
__device__ __forceinline__ void load_row(uint32_t addr, uint32_t addr_rem, uint32_t (&e)[NR])
{
  constexpr int FULLQ = NR / 4, REM = NR % 4;
#pragma unroll
  for (int g = 0; g < FULLQ; g++) { uint4 v = lds128(addr + g * (G * 16)); e[4*g] = v.x; e[4*g+1] = v.y; e[4*g+2] = v.z; e[4*g+3] = v.w; }
#pragma unroll
  for (int r = 0; r < REM; r++) e[4*FULLQ + r] = lds32(addr_rem + r * 128);
}
__device__ __forceinline__ uint32_t hfma2_relu_add(uint32_t m, uint32_t e)
{
  uint32_t r;
  asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(m), "r"(0x3c003c00u), "r"(e));
  return r;
}
__device__ __forceinline__ int half_bits_to_int(uint32_t bits) { return (int)__half2float(__ushort_as_half((unsigned short)bits)); }

struct LtWin { int32_t seq, idx, k, length; long long n; float score; int32_t item; };

// A stretch of one chunk for one lane group: rows [scan_from, row_end] are computed from a zero row, diagonals are detected
// from row detect_from on (the rows before it only warm the DP row up).  Without items a group scans a whole chunk.
struct LtItem { int32_t seq, scan_from, detect_from, row_end, id; };

struct LtArgs {
  ProfDev P; SeqDev sd;
  int sc_thresh, tjb, Q;
  LtWin *win; int *nwin; int cap; int *counter;
  const LtItem *items = nullptr; int nitems = 0;
};

// FAST: cells kept RELATIVE to the begin floor, c = max(m, xB) - xB, which turns the recurrence into the plain SSV one,
// c' = relu(c + (bias - cost)): ONE instruction per cell pair (HFMA2.RELU) instead of three, and the saturation of
// adds_epu8(.., bias) left out.  Exact when xB < threshold <= 256 - bias: every cell a scan keeps is below the threshold (a row
// that reaches it is recorded and zeroed; the warm-up rows of a stretch hold lower bounds of rows in which the reference found
// nothing), hence <= 255 - bias, so the min() never binds; max(relu(x), xB) = max(x, xB) for xB >= 0; a cell >= threshold > xB
// has max(m, xB) = m; all values are integers below 2048, exact in fp16.  The host picks the path per search (launch_lt).
template <int G, int NR> struct LtShape {
  static constexpr uint32_t TAB_BYTES = (uint32_t)B2H_NCODE * ((uint32_t)(NR / 4) * G * 16 + (uint32_t)(NR % 4) * 128);
  static constexpr int THREADS = TAB_BYTES > 76 * 1024 ? 512 : 256;          // warps share the CTA's copy of the table
  static constexpr int MINB = TAB_BYTES > 76 * 1024 ? 1 : 3;                  // 24 warps per SM: 80 registers
};
template <int G, int NR, bool FAST>
__global__ void __launch_bounds__(LtShape<G, NR>::THREADS, LtShape<G, NR>::MINB) lt_ssv_kernel(const LtArgs a)
{
  extern __shared__ __align__(128) uint32_t s_tab[];
  __shared__ uint64_t s_bar;
  constexpr int FULLQ = NR / 4, REM = NR % 4, NG = 32 / G;
  constexpr uint32_t ROWB = (uint32_t)FULLQ * G * 16 + (uint32_t)REM * 128;
  constexpr uint32_t TAB_BYTES = (uint32_t)B2H_NCODE * ROWB;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1), grp = lane / G;
  const int src_lane = (lane & ~(G - 1)) | ((gl + G - 1) & (G - 1));
  const uint32_t tab_lane = smem_u32(s_tab) + gl * 16;
  const uint32_t tab_rem  = smem_u32(s_tab) + FULLQ * G * 16 + lane * 4;
  const ProfDev &P = a.P;

  if (threadIdx.x == 0) { mbar_init(&s_bar, 1); mbar_expect_tx(&s_bar, TAB_BYTES); tma_load_1d(s_tab, P.ssv_emis, TAB_BYTES, &s_bar); }
  __syncthreads();
  mbar_wait(&s_bar, 0);

  const int M = P.M, bias = P.bias, base = P.base, Mpad = P.Mpad;
  const int tjbm = (a.tjb + P.tbm) & 0xff;                    // (int8)tjb + (int8)tbm splatted into bytes (msvfilter.c:338)
  const int xB = max(base - tjbm, 0);                          // constant: no J state in SSV
  const int floor_sc = base - a.tjb - P.tbm;                   // the level a diagonal left the begin state at (msvfilter.c:384)
  const __half2 xb2 = __float2half2_rn((float)xB), cap2 = __float2half2_rn((float)(255 - bias)), th2 = __float2half2_rn((float)a.sc_thresh);
  const uint32_t xBh = *reinterpret_cast<const uint32_t *>(&xb2), caph = *reinterpret_cast<const uint32_t *>(&cap2);
  const __half2 thf2 = __float2half2_rn((float)(a.sc_thresh - xB));
  const int thr_bits = (int)(*reinterpret_cast<const uint32_t *>(FAST ? &thf2 : &th2) & 0xffffu);     // FAST: on the scale of c

  for (;;) {
    int e0 = 0;
    if (lane == 0) e0 = atomicAdd(a.counter, NG);
    e0 = __shfl_sync(FULL, e0, 0);
    const int nwork = a.items ? a.nitems : a.sd.n;
    if (e0 >= nwork) break;
    const int e = e0 + grp;
    const bool valid = e < nwork;
    LtItem it; it.seq = 0; it.scan_from = 1; it.detect_from = 1; it.row_end = 0; it.id = 0;
    if (valid) { if (a.items) it = a.items[e]; else { it.seq = a.sd.order[e]; it.row_end = a.sd.len[it.seq]; } }
    const int s = it.seq;
    const int L = valid ? a.sd.len[s] : 0;                    // the chunk's true length: the diagonal walks may leave the item's rows
    const int row_end = valid ? it.row_end : 0, detect_from = it.detect_from;
    const uint8_t *seq = a.sd.res + a.sd.off[s];
    const uint32_t *seqw = reinterpret_cast<const uint32_t *>(seq);
    const int nwords = (L + 3) >> 2;

    const uint32_t zero_row = 0u;
    uint32_t m[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) m[j] = zero_row;
    int i = it.scan_from, nemit = 0;                           // next row of this group's scan (1-based), windows emitted so far
    int w0 = -G; uint32_t myw = 0;

    while (__any_sync(FULL, i <= row_end)) {
      const bool active = i <= row_end;
      uint32_t x = B2H_PAD_CODE;
      {
        const int w = active ? ((i - 1) >> 2) : w0;            // word holding residue i
        if (active && (w < w0 || w >= w0 + G)) { w0 = w & ~(G - 1); myw = (w0 + gl < nwords) ? __ldg(seqw + w0 + gl) : 0x1f1f1f1fu; }
        const uint32_t wr = __shfl_sync(FULL, myw, max(w - w0, 0), G);
        if (active) x = (wr >> (((i - 1) & 3) * 8)) & 0xffu;
      }
      uint32_t ev[NR];
      load_row<G, NR>(tab_lane + x * ROWB, tab_rem + x * ROWB, ev);
      const uint32_t t  = __shfl_sync(FULL, m[NR-1], src_lane);
      const uint32_t s0 = __byte_perm(t, m[NR-1], 0x5432u);
      // sv = subs_epu8(adds_epu8(max(mpv, xB), bias), cost) == relu(min(max(mpv, xB), 255 - bias) + (bias - cost))
      if (FAST) {
#pragma unroll
        for (int j = NR - 1; j >= 1; j--) m[j] = hfma2_relu_add(m[j-1], ev[j]);
        m[0] = hfma2_relu_add(s0, ev[0]);
      } else {
#pragma unroll
        for (int j = NR - 1; j >= 1; j--) m[j] = hfma2_relu_add(__vimin3_s16x2(__vimax3_s16x2(m[j-1], xBh, xBh), caph, caph), ev[j]);
        m[0] = hfma2_relu_add(__vimin3_s16x2(__vimax3_s16x2(s0, xBh, xBh), caph, caph), ev[0]);
      }
      uint32_t xq[4] = {0u, 0u, 0u, 0u};                       // four independent chains, then one join
#pragma unroll
      for (int j = 0; j + 1 < NR; j += 2) xq[(j >> 1) & 3] = __vimax3_s16x2(xq[(j >> 1) & 3], m[j], m[j+1]);
      if (NR & 1) xq[3] = __vimax3_s16x2(xq[3], m[NR-1], m[NR-1]);
      const uint32_t xe = __vimax3_s16x2(__vimax3_s16x2(xq[0], xq[1], xq[2]), xq[3], xq[3]);
      const bool lanehit = active && i >= detect_from && (max((int)(xe & 0xffffu), (int)(xe >> 16)) >= thr_bits);
      if (__any_sync(FULL, lanehit)) {
        // the cell the reference picks: the largest value >= threshold among the model's nodes, first in its striped scan
        // order (q outer, z inner; node k = q + Q*z + 1) -- packed as (value bits << 16) | (0xffff - (q*16 + z))
        int best = 0;
        if (active && i >= detect_from) {                      // (another group of the warp may have triggered this block)
#pragma unroll
          for (int j = 0; j < NR; j++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const int k = gl * 2 * NR + j + h * NR + 1;
              const int vb = (int)((m[j] >> (16 * h)) & 0xffffu);
              if (k <= M && vb >= thr_bits) {
                const int key = ((k - 1) % a.Q) * 16 + (k - 1) / a.Q;
                best = max(best, (vb << 16) | (0xffff - key));
              }
            }
          }
        }
#pragma unroll
        for (int o = G / 2; o >= 1; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
        int new_i = i;
        if (best > 0 && gl == 0) {
          const int key = 0xffff - (best & 0xffff);
          int end = (key >> 4) + a.Q * (key & 15) + 1;
          int rem_sc = half_bits_to_int((uint32_t)best >> 16) + (FAST ? xB : 0);
          int sc = rem_sc;
          int start = end, target_start = i, target_end = i;
          while (rem_sc > floor_sc && start >= 1 && target_start >= 1) {          // walk the diagonal back to the begin level
            rem_sc -= bias - (int)P.msv_cost8[(size_t)seq[target_start - 1] * Mpad + (start - 1)];
            --start; --target_start;
          }
          start++; target_start++;
          int k = end + 1, n = target_end + 1, max_end = target_end, max_sc = sc, pos_since_max = 0;
          while (k < M && n <= L) {                                                 // single-diagonal extension (msvfilter.c:392-408)
            sc += bias - (int)P.msv_cost8[(size_t)seq[n - 1] * Mpad + (k - 1)];
            if (sc >= max_sc) { max_sc = sc; max_end = n; pos_since_max = 0; }
            else if (++pos_since_max == 5) break;
            k++; n++;
          }
          end += max_end - target_end;
          target_end = max_end;
          float ret_sc = ((float)(max_sc - a.tjb) - (float)base);
          ret_sc /= P.scale_b;
          ret_sc -= 3.0f;
          const int slot = atomicAdd(a.nwin, 1);
          if (slot < a.cap) { LtWin w; w.seq = s; w.idx = nemit; w.k = end; w.length = end - start + 1; w.n = target_start; w.score = ret_sc; w.item = it.id; a.win[slot] = w; }
          new_i = target_end;                                                       // skip forward (msvfilter.c:411)
        }
        best = __shfl_sync(FULL, best, 0, G);
        new_i = __shfl_sync(FULL, new_i, 0, G);
        if (best > 0) {                                                             // this group recorded a window: reset its row
#pragma unroll
          for (int j = 0; j < NR; j++) m[j] = zero_row;
          i = new_i; nemit++;
        }
      }
      if (active) i++;
    }
  }
}

static bool lt_fast_cells(int base, int tbm, int bias, int tjb, int sc_thresh)
{
  const int xB = std::max(base - ((tjb + tbm) & 0xff), 0);
  return xB < sc_thresh && sc_thresh <= 256 - bias && !getenv("B2H_LT_FULL_CELLS");       // (the variable: tests of the general path)
}

template <int G, int NR, bool FAST>
int launch_lt_path(b2h_ctx *ctx, const LtArgs &a, cudaStream_t strm)
{
  // One scan is a chain of dependent rows, ~100-135 instructions each: what an SM delivers is set by the warps it holds.  The
  // CTA shares one copy of the table, so the warps per CTA follow the table size (LtShape): 8 (three 64 KB tables per SM = 24
  // warps), 16 when only one or two tables fit.
  const size_t smem = (size_t)B2H_NCODE * b2h_ssv_row_bytes(G, NR);
  const int threads = LtShape<G, NR>::THREADS, wpc = threads / 32;
  int occ = 1;
  { const int st = b2h_kernel_occupancy(ctx, (const void *)lt_ssv_kernel<G, NR, FAST>, threads, smem, &occ); if (st != B2H_OK) return st; }
  const int NG = 32 / G;
  const int nwork = a.items ? a.nitems : a.sd.n;
  int grid = std::min(ctx->sm_count * occ, std::max(1, (nwork + wpc * NG - 1) / (wpc * NG)));
  lt_ssv_kernel<G, NR, FAST><<<grid, threads, smem, strm>>>(a);
  ctx->launches++;
  B2H_CUDA(cudaGetLastError());
  return B2H_OK;
}
template <int G, int NR>
int launch_lt(b2h_ctx *ctx, const LtArgs &a, cudaStream_t strm)
{
  const bool fast = lt_fast_cells(a.P.base, a.P.tbm, a.P.bias, a.tjb, a.sc_thresh);
  return fast ? launch_lt_path<G, NR, true>(ctx, a, strm) : launch_lt_path<G, NR, false>(ctx, a, strm);
}

// esl_gumbel_invsurv (vendor/easel/esl_gumbel.c:185)
double gumbel_invsurv(double p, double mu, double lambda)
{
  const double log_part = (p < 5e-9) ? (pow(p, p) - 1) / p : log(-1. * log(1 - p));
  return mu - (log_part / lambda);
}

// prefix / suffix lengths (p7_hmm_ScoreDataComputeRest, p7_scoredata.c:356-378) from the Forward M->I and I->I odds; suffix has M+2 slots
void window_lengths(const b2h_profile *p, float *pre, float *suf)
{
  const int M = p->M;
  const float *t_mi = p->h_fwd_tsc.data() + (size_t)5 * M, *t_ii = p->h_fwd_tsc.data() + (size_t)6 * M;   // node k at [k-1]
  for (int k = 0; k <= M; k++) pre[k] = suf[k] = 0.f;
  float sum = 0;
  for (int k = 1; k < M; k++) {
    if (t_mi[k - 1] == 0) pre[k] = 1;
    else pre[k] = (float)(1 + (int)(log(1e-7 / t_mi[k - 1]) / log((double)t_ii[k - 1])));
    sum += pre[k];
  }
  pre[0] = pre[M] = 0;
  for (int k = 1; k < M; k++) pre[k] /= sum;
  suf[M] = (M >= 1) ? pre[M - 1] : 0.f;
  for (int k = M - 1; k >= 1; k--) suf[k] = suf[k + 1] + pre[k - 1];
  for (int k = 2; k < M; k++) pre[k] += pre[k - 1];
}

// p7_pli_ExtendAndMergeWindows (p7_pipeline.c:323-400), in place: every diagonal {n, k, length} becomes the window the
// model could reach around it (max_length x the prefix / suffix share of the nodes before / after it, plus 10 %), then
// windows of the same chunk that overlap by more than <pct_overlap> of the shorter one are fused.  Returns the new count.
size_t extend_merge(const b2h_profile *p, b2h_window *w, size_t n, const int64_t *target_len, float pct_overlap)
{
  const int M = p->M, maxlen = p->max_length;
  std::vector<float> pre(M + 1, 0.f), suf(M + 2, 0.f);
  window_lengths(p, pre.data(), suf.data());
  size_t nm = 0;
  for (size_t i = 0; i < n; i++) {
    b2h_window x = w[i];
    const long long ws = (long long)std::max<double>(1.0, (double)x.n - ((double)maxlen * (0.1 + (double)pre[x.k - x.length + 1])));
    const long long we = (long long)std::min<double>((double)target_len[i], (double)x.n + (double)x.length + ((double)maxlen * (0.1 + (double)suf[x.k])));
    x.n = ws; x.length = (int32_t)(we - ws + 1);
    if (nm > 0 && w[nm - 1].seq == x.seq) {
      b2h_window &pv = w[nm - 1];
      const long long os = std::max<long long>(pv.n, x.n), oe = std::min<long long>(pv.n + pv.length - 1, x.n + x.length - 1), ol = oe - os + 1;
      if ((float)ol / (float)std::min(pv.length, x.length) > pct_overlap) {
        const long long ms = std::min<long long>(pv.n, x.n), me = std::max<long long>(pv.n + pv.length - 1, x.n + x.length - 1);
        pv.n = ms; pv.length = (int32_t)(me - ms + 1);
        continue;
      }
    }
    w[nm++] = x;
  }
  return nm;
}

} // namespace

size_t b2h_extend_merge(const b2h_profile *p, b2h_window *w, size_t n, const int64_t *target_len, float pct_overlap)
{ return extend_merge(p, w, n, target_len, pct_overlap); }

extern "C" void b2h_free(void *p) { free(p); }

extern "C" int b2h_extend_merge_windows(const b2h_profile *p, b2h_window *windows, size_t n, const int64_t *target_len, float pct_overlap, size_t *nout)
{
  if (!p || (!windows && n) || (!target_len && n) || !nout || p->max_length <= 0) return B2H_EINVAL;
  for (size_t i = 0; i < n; i++) if (windows[i].k < 1 || windows[i].k > p->M || windows[i].length < 1 || windows[i].k - windows[i].length + 1 < 0) return B2H_EINVAL;
  *nout = extend_merge(p, windows, n, target_len, pct_overlap);
  return B2H_OK;
}

extern "C" int b2h_window_lengths(const b2h_profile *p, float *prefix, float *suffix)
{
  if (!p || !prefix || !suffix) return B2H_EINVAL;
  std::vector<float> suf(p->M + 2, 0.f);
  window_lengths(p, prefix, suf.data());
  for (int k = 0; k <= p->M; k++) suffix[k] = suf[k];
  return B2H_OK;
}

// threshold on the byte scale for P-value F1 with the length model of max_length (msvfilter.c:289-327)
static int lt_threshold(const b2h_profile *p, double F1, b2h_len_params *lp)
{
  b2h_length_params(p->max_length, 1.0f, lp);
  const float invP = (float)gumbel_invsurv(F1, (double)p->evparam[0], (double)p->evparam[1]);
  return (int)(uint8_t)(int)ceil((((double)lp->null1 + ((double)invP * 0.69314718055994529) + 3.0) * (double)p->scale_b)
                                 + (double)p->base_b + (double)p->tec_b + (double)lp->tjb_b);
}

// What a scan with this profile and F1 would use: the byte threshold and whether the two-instruction cell applies (tests).
extern "C" int b2h_longtarget_scan_info(const b2h_profile *p, double F1, int *sc_thresh, int *fast_cells)
{
  if (!p || p->max_length <= 0) return B2H_EINVAL;
  b2h_len_params lp;
  const int thr = lt_threshold(p, F1, &lp);
  if (sc_thresh) *sc_thresh = thr;
  if (fast_cells) *fast_cells = lt_fast_cells(p->base_b, p->tbm_b, p->bias_b, lp.tjb_b, thr) ? 1 : 0;
  return B2H_OK;
}

extern "C" int b2h_longtarget_windows(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *db, double F1,
                                      b2h_window **raw_out, size_t *nraw_out, b2h_window **merged_out, size_t *nmerged_out)
{
  if (!ctx || !p || !db || p->ctx != ctx || db->ctx != ctx || !raw_out || !nraw_out || !merged_out || !nmerged_out) return B2H_EINVAL;
  *raw_out = *merged_out = nullptr; *nraw_out = *nmerged_out = 0;
  if (p->max_length <= 0) { ctx->err = "long-target search needs the model's max_length (MAXL)"; return B2H_EINVAL; }
  const size_t n = db->n;
  if (n == 0) return B2H_OK;
  B2H_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  b2h_len_params lp;
  const int sc_thresh = lt_threshold(p, F1, &lp);
  LtArgs a;
  a.P = b2h_profdev(p); a.sd = b2h_seqdev(db); a.sc_thresh = sc_thresh; a.tjb = lp.tjb_b;
  a.Q = std::max(2, (p->M - 1) / 16 + 1);                     // p7O_NQB(M): the striping the reference's tie-break follows
  a.cap = (int)std::min<size_t>((size_t)1 << 24, std::max<size_t>(1024, (size_t)(db->nres / 16) + 1024));
  a.counter = ctx->d_counters;
  int *d_nwin = nullptr;
  cudaError_t e;
  if ((e = cudaMallocAsync((void **)&a.win, (size_t)a.cap * sizeof(LtWin), st)) != cudaSuccess ||
      (e = cudaMallocAsync((void **)&d_nwin, sizeof(int), st)) != cudaSuccess) { ctx->err = cudaGetErrorString(e); return B2H_EMEM; }
  a.nwin = d_nwin;
  struct Release { LtArgs &a; int *nw; cudaStream_t st; ~Release() { cudaFreeAsync(a.win, st); cudaFreeAsync(nw, st); } } release{a, d_nwin, st};
  // one launch over <items> (nullptr: every chunk whole); its diagonals are appended to <out>
  auto launch = [&](const std::vector<LtItem> *items, std::vector<LtWin> &out) -> int {
    LtItem *d_items = nullptr;
    if (items) {
      if (items->empty()) return B2H_OK;
      if ((e = cudaMallocAsync((void **)&d_items, items->size() * sizeof(LtItem), st)) != cudaSuccess) { ctx->err = cudaGetErrorString(e); return B2H_EMEM; }
      cudaMemcpyAsync(d_items, items->data(), items->size() * sizeof(LtItem), cudaMemcpyHostToDevice, st);
      a.items = d_items; a.nitems = (int)items->size();
    } else { a.items = nullptr; a.nitems = 0; }
    cudaMemsetAsync(d_nwin, 0, sizeof(int), st);
    cudaMemsetAsync(a.counter, 0, sizeof(int), st);
    static const bool trace = getenv("B2H_TRACE") != nullptr;
    const auto t_launch = std::chrono::steady_clock::now();
    int rc = B2H_EINVAL;
    switch (p->G * 64 + p->NR) {
#define CASE(g, n_) case (g) * 64 + (n_): rc = launch_lt<g, n_>(ctx, a, st); break;
#define CASE8(g, n_) CASE(g, n_) CASE(g, n_ + 1) CASE(g, n_ + 2) CASE(g, n_ + 3) CASE(g, n_ + 4) CASE(g, n_ + 5) CASE(g, n_ + 6) CASE(g, n_ + 7)
      CASE8(8, 1) CASE8(8, 9) CASE8(8, 17) CASE8(8, 25)
      CASE8(16, 17) CASE8(16, 25)
      CASE(32, 18) CASE(32, 20) CASE(32, 22) CASE(32, 24) CASE(32, 26) CASE(32, 28) CASE(32, 30) CASE(32, 32) CASE(32, 40) CASE(32, 48)
#undef CASE8
#undef CASE
    }
    int nwin = 0;
    if (rc == B2H_OK) {
      cudaMemcpyAsync(&nwin, d_nwin, sizeof(int), cudaMemcpyDeviceToHost, st);
      if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { ctx->err = std::string("long-target SSV kernel: ") + cudaGetErrorString(e); rc = B2H_ECUDA; }
      else if (nwin > a.cap) { ctx->err = "long-target SSV: window list overflow"; rc = B2H_ERANGE; }
      else if (nwin > 0) {
        const size_t at = out.size();
        out.resize(at + nwin);
        cudaMemcpyAsync(out.data() + at, a.win, (size_t)nwin * sizeof(LtWin), cudaMemcpyDeviceToHost, st);
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = B2H_ECUDA; }
      }
    }
    if (d_items) cudaFreeAsync(d_items, st);
    if (trace) fprintf(stderr, "[b2h_longtarget_windows] scan of %zu %s: %d diagonals, %.2f ms\n", items ? items->size() : n, items ? "stretches" : "chunks",
                       nwin, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_launch).count());
    return rc;
  };
  std::vector<LtWin> hw;
  // A chunk is ONE dependent scan in the reference (the DP row is reset behind every diagonal), which leaves a 100 Mb genome with
  // ~800 chains of 262 144 rows: latency bound.  Speculative stretches: the chunk is cut into stretches of S rows, each scanned
  // from a zero row O = M rows early (a cell depends on the < M rows of its diagonal only, so the row is exact when detection
  // starts) -- provided no diagonal of the chunk ended within those O rows, which is decided from the PREVIOUS stretch's own,
  // exact, diagonals: if its scan resumed (target end + 1) later than O rows before this stretch, this stretch is scanned again
  // from that row.  Stretches of a chunk are confirmed in order; a round rescans, in parallel over the chunks, the first
  // unconfirmed stretch of each; after 6 rounds what is left of a chunk is scanned in one piece.
  static const int stretch = getenv("B2H_LT_STRETCH") ? atoi(getenv("B2H_LT_STRETCH")) : 8192;
  const int S = std::max(256, stretch), O = p->M;
  bool whole = stretch <= 0;
  if (!whole) { size_t nst = 0; for (size_t s2 = 0; s2 < n; s2++) nst += (db->h_len[s2] + S - 1) / S; whole = nst <= n + n / 2; }   // chunks no longer than a stretch
  if (whole) { const int rc = launch(nullptr, hw); if (rc != B2H_OK) return rc; }
  else {
    std::vector<LtItem> items;
    for (size_t s2 = 0; s2 < n; s2++)
      for (int c = 0, L2 = db->h_len[s2]; c * S < L2; c++)
        items.push_back(LtItem{(int32_t)s2, std::max(1, c * S + 1 - O), c * S + 1, std::min(L2, (c + 1) * S), c});
    std::vector<LtWin> spec;
    { const int rc = launch(&items, spec); if (rc != B2H_OK) return rc; }
    auto by_scan = [](const LtWin &x, const LtWin &y) { return x.seq != y.seq ? x.seq < y.seq : x.item != y.item ? x.item < y.item : x.idx < y.idx; };
    std::sort(spec.begin(), spec.end(), by_scan);
    std::vector<size_t> first(n + 1, 0);                      // spec[first[s] .. first[s+1]) = diagonals of chunk s
    { size_t q = 0; for (size_t s2 = 0; s2 < n; s2++) { first[s2] = q; while (q < spec.size() && (size_t)spec[q].seq == s2) q++; } first[n] = spec.size(); }
    struct Cursor { int c = 0; long long resume = 0; size_t q = 0; bool waiting = false; };     // resume: row the reference's scan resumed at behind its last diagonal
    std::vector<Cursor> cur(n);
    for (size_t s2 = 0; s2 < n; s2++) cur[s2].q = first[s2];
    auto take = [&](size_t s2, const LtWin *w, size_t cnt) {   // confirmed diagonals of the current stretch of chunk s2
      for (size_t z = 0; z < cnt; z++) { LtWin x = w[z]; x.idx = (int32_t)hw.size(); hw.push_back(x); cur[s2].resume = x.n + x.length; }   // target_end + 1
    };
    for (int round = 0;; round++) {
      std::vector<LtItem> redo;
      std::vector<size_t> redo_of;
      for (size_t s2 = 0; s2 < n; s2++) {
        Cursor &k = cur[s2];
        const int L2 = db->h_len[s2];
        while (!k.waiting && (long long)k.c * S < L2) {
          const long long a0 = (long long)k.c * S + 1, end = std::min<long long>(L2, (long long)(k.c + 1) * S);
          size_t q1 = k.q; while (q1 < first[s2 + 1] && spec[q1].item == k.c) q1++;
          if (k.resume <= a0 - O || k.resume <= 1) { take(s2, spec.data() + k.q, q1 - k.q); k.q = q1; k.c++; continue; }   // the speculation held
          k.q = q1;                                            // its speculative diagonals are void
          if (k.resume > end) { k.c++; continue; }             // the scan skipped this stretch altogether
          const bool rest = round >= 6;                        // a dense chunk: finish it in one sequential piece
          redo.push_back(LtItem{(int32_t)s2, (int32_t)k.resume, (int32_t)std::max(a0, k.resume), (int32_t)(rest ? L2 : end), k.c});
          redo_of.push_back(s2);
          if (rest) { k.c = (L2 + S - 1) / S; k.q = first[s2 + 1]; }
          k.waiting = true;
        }
      }
      if (redo.empty()) break;
      std::vector<LtWin> got;
      { const int rc = launch(&redo, got); if (rc != B2H_OK) return rc; }
      std::sort(got.begin(), got.end(), by_scan);
      size_t g = 0;
      for (size_t z = 0; z < redo_of.size(); z++) {            // redo is in chunk order, so is got
        const size_t s2 = redo_of[z];
        size_t g1 = g; while (g1 < got.size() && (size_t)got[g1].seq == s2) g1++;
        take(s2, got.data() + g, g1 - g);
        g = g1;
        Cursor &k = cur[s2];
        k.waiting = false;
        if ((long long)k.c * S < db->h_len[s2]) k.c++;       // (a "rest" item has already closed its chunk)
      }
    }
  }
  // every chunk's diagonals in the order its scan produced them
  std::stable_sort(hw.begin(), hw.end(), [](const LtWin &x, const LtWin &y) { return x.seq != y.seq ? x.seq < y.seq : x.idx < y.idx; });   // (idx: scan order inside a chunk in both modes)
  b2h_window *raw = (b2h_window *)malloc(std::max<size_t>(1, hw.size()) * sizeof(b2h_window));
  b2h_window *mer = (b2h_window *)malloc(std::max<size_t>(1, hw.size()) * sizeof(b2h_window));
  if (!raw || !mer) { free(raw); free(mer); return B2H_EMEM; }
  for (size_t i = 0; i < hw.size(); i++) { raw[i].seq = hw[i].seq; raw[i].k = hw[i].k; raw[i].n = hw[i].n; raw[i].length = hw[i].length; raw[i].score = hw[i].score; }
  std::vector<int64_t> tlen(hw.size());
  for (size_t i = 0; i < hw.size(); i++) { mer[i] = raw[i]; tlen[i] = db->h_len[raw[i].seq]; }
  const size_t nm = extend_merge(p, mer, hw.size(), tlen.data(), 0.0f);
  *raw_out = raw; *nraw_out = hw.size(); *merged_out = mer; *nmerged_out = nm;
  return B2H_OK;
}
