// b2h_dpreg.cu -- register-resident ViterbiFilter / Forward parser / Backward parser for models with
// M <= 32*C*W: C = 2..16 nodes per lane, W = 1, 2 or 4 warps per comparison (B2H_REG_CLASSES: W = 1 for M <= 512 in
// steps of 32 or 64 nodes, ~95 % of Pfam;  W = 2: M <= 1024;  W = 4: M <= 1536); a model runs in the smallest class
// that holds it, so at most 32 (64) nodes per comparison are padding.
// With W > 1 the W warps of a group own consecutive 32*C-node segments of the model and meet at one named
// barrier per row (two for Backward and for Viterbi rows that need the D->D closure): the per-warp partial
// results (xE maxima / sums, the affine or max-plus composite of the warp's D chain, the M/I/D cells of its
// last node) are exchanged through a few words of shared memory and combined by every warp in the same
// order, so all warps of a group take identical decisions.
//
// Same recurrences and the same numerical semantics as the generic kernels of b2h_dp.cu (which remain
// the path for longer models), but organised so that almost nothing is re-read per row:
//   * lane z owns the C consecutive nodes z*C+1 .. z*C+C; its M/I/D cells AND its 8*C transition
//     scores live in registers for the whole comparison;
//   * the only per-row memory traffic is one conflict-free vector load of the C emission scores of
//     the current residue from shared memory (table staged per CTA per profile by a TMA bulk copy);
//   * the (i-1,k-1) dependency across lanes is three warp shuffles per row;
//   * the D->D chain is closed in two levels: serially inside the lane (C steps), then one warp scan
//     of the 32 per-lane composites (max-plus for Viterbi, affine for Forward/Backward).
// The generic kernels execute ~88 instructions per cell and row; these execute ~0.5.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdlib>
#include <algorithm>
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

struct Item { int p, e_begin, e_end; };
// SUB > 1: a CTA takes 1/SUB of a work item at a time.  The multi-warp classes keep only 4 (W = 2) or 2 (W = 4) comparisons
// in flight per CTA and a comparison is one dependent chain of L rows, so whole items (16 entries) serialise 4 to 8 such
// chains per group while other SMs idle at the end of the stage; with sub-items every group holds one comparison.
template <int SUB = 1>
__device__ __forceinline__ bool next_item(const WorkList &wl, int *s_item, Item &it)
{
  __syncthreads();
  if (threadIdx.x == 0) *s_item = atomicAdd(wl.counter, 1);
  __syncthreads();
  const int raw = *s_item;
  const int item = raw / SUB + wl.itemoff[wl.plo];
  if (item >= wl.itemoff[wl.phi]) return false;
  int lo = wl.plo, hi = wl.phi;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (wl.itemoff[mid] <= item) lo = mid; else hi = mid; }
  it.p = lo;
  constexpr int PER = B2H_ITEM_ENTRIES / SUB;
  const int e0 = wl.poff[lo] + (item - wl.itemoff[lo]) * B2H_ITEM_ENTRIES;
  const int e1 = min(wl.poff[lo + 1], e0 + B2H_ITEM_ENTRIES);
  it.e_begin = e0 + (raw % SUB) * PER;
  it.e_end   = min(e1, it.e_begin + PER);
  return true;
}
template <int W> struct ItemSplit { static constexpr int SUB = (W >= 4) ? 8 : (W == 2) ? 4 : 1; };

// The C emission values of this lane for residue x.  Table layout: [32 residues][W warps][C/G groups][32 lanes][G]
// with G = 4 (C % 4 == 0), 2 (C even) or 1, so every access is one conflict-free LDS.128 / LDS.64 / LDS.32 per lane; <tab> already
// points at this warp's segment and <stride> = 32*C*W is the distance between residues.
template <int C, typename T>
__device__ __forceinline__ void load_emis(const T *tab, int stride, int x, int lane, T (&r)[C])
{
  const T *row = tab + (size_t)x * stride;
  if (C % 2 != 0) {
#pragma unroll
    for (int g = 0; g < C; g++) r[g] = row[g * 32 + lane];
  } else if (C % 4 != 0) {
#pragma unroll
    for (int g = 0; g < C / 2; g++) {
      const float2 v = *reinterpret_cast<const float2 *>(row + g * 64 + lane * 2);
      r[2*g+0] = *(const T *)&v.x; r[2*g+1] = *(const T *)&v.y;
    }
  } else {
#pragma unroll
    for (int g = 0; g < C / 4; g++) {
      const float4 v = *reinterpret_cast<const float4 *>(row + g * 128 + lane * 4);
      r[4*g+0] = *(const T *)&v.x; r[4*g+1] = *(const T *)&v.y; r[4*g+2] = *(const T *)&v.z; r[4*g+3] = *(const T *)&v.w;
    }
  }
}

// named barrier of one W-warp group (barrier 0 stays the CTA barrier of next_item)
template <int W>
__device__ __forceinline__ void group_sync(int grp)
{
  if (W > 1) asm volatile("bar.sync %0, %1;" :: "r"(grp + 1), "r"(W * 32) : "memory");
}
constexpr int MAXGRP = 4;            // 256 threads / (32 * W), W >= 2

// residues: each lane keeps one 4-residue word of the current 128-row window (as the SSV kernel does)
struct SeqWin {
  const uint32_t *seqw; int nwords; uint32_t myw; int w0;
  __device__ __forceinline__ void init(const uint8_t *seq, int L, int lane) { seqw = reinterpret_cast<const uint32_t *>(seq); nwords = (L + 3) >> 2; w0 = -32; myw = 0; (void)lane; }
  __device__ __forceinline__ int get(int i0, int lane) {      // residue at 0-based position i0 (monotonically increasing calls)
    const int w = i0 >> 2;
    if (w >= w0 + 32) { w0 = w & ~31; myw = (w0 + lane < nwords) ? __ldg(seqw + w0 + lane) : 0x1f1f1f1fu; }
    const uint32_t wr = __shfl_sync(FULL, myw, w - w0);
    return (wr >> ((i0 & 3) * 8)) & 0xff;
  }
};

// =================================================================================================
// ViterbiFilter, register resident
// =================================================================================================
template <int C, int W>
__global__ void __launch_bounds__(256) rvit_kernel(const WorkList wl, const SeqDev sd, const StageOut out)
{
  extern __shared__ __align__(128) int s_rsc[];            // [32][W][32*C] int32 emission scores
  __shared__ uint64_t s_bar;
  __shared__ int s_item;
  __shared__ int s_x[2][MAXGRP][6][W];                      // per row parity: M, I, D, M+tMD of each warp's last node; xE, Dmax partials
  __shared__ int s_y[MAXGRP][2][W];                         // max-plus composite (A, T) of each warp's D chain
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int grp = warp / W, wi = warp % W, ngrp = nwarps / W;
  const int gl = wi * 32 + lane;                            // lane index inside the group: owns nodes gl*C .. gl*C+C-1
  constexpr int STRIDE = 32 * C * W;
  constexpr uint32_t TAB_BYTES = 32u * STRIDE * 4u;
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  uint32_t phase = 0;
  int cur_p = -1;
  int tBM[C], tMM[C], tIM[C], tDM[C], tMD[C], tMI[C], tII[C], tDD[C];
  int tDDin = NEG16;                                        // D_{k-1}->D_k for this lane's first node (the left lane's last tDD)
  Item it;
  while (next_item<ItemSplit<W>::SUB>(wl, &s_item, it)) {
    const ProfDev &P = wl.profs[it.p];
    if (it.p != cur_p) {
      cur_p = it.p;
      if (threadIdx.x == 0) { mbar_expect_tx(&s_bar, TAB_BYTES); tma_load_1d(s_rsc, P.vit_rsc32, TAB_BYTES, &s_bar); }
      const int16_t *ts = P.vit_tsc; const int Mp = P.Mpad;
#pragma unroll
      for (int c = 0; c < C; c++) {
        const int k0 = gl * C + c;                          // 0-based node column
        const bool in = k0 < Mp;
        tBM[c] = in ? ts[0 * Mp + k0] : NEG16; tMM[c] = in ? ts[1 * Mp + k0] : NEG16; tIM[c] = in ? ts[2 * Mp + k0] : NEG16;
        tDM[c] = in ? ts[3 * Mp + k0] : NEG16; tMD[c] = in ? ts[4 * Mp + k0] : NEG16; tMI[c] = in ? ts[5 * Mp + k0] : NEG16;
        tII[c] = in ? ts[6 * Mp + k0] : NEG16; tDD[c] = in ? ts[7 * Mp + k0] : NEG16;
      }
      tDDin = __shfl_up_sync(FULL, tDD[C - 1], 1);
      if (lane == 0) tDDin = (W > 1 && wi > 0 && gl * C - 1 < Mp) ? (int)ts[7 * Mp + gl * C - 1] : NEG16;
      mbar_wait(&s_bar, phase); phase ^= 1;
    }
    const int xwEm = P.xw_E_move, xwEl = P.xw_E_loop, base_w = P.base_w, ddbound = P.ddbound_w;
    const int *my_rsc = s_rsc + wi * 32 * C;

    for (int e = it.e_begin + grp; e < it.e_end; e += ngrp) {
      if (out.redo_only && out.status[e] != B2H_REDO) continue;      // second pass behind the packed kernel (uniform per group)
      const int s = wl.ent_s[e];
      const int L = sd.len[s];
      const int xw_move = sd.xwmove[s];
      SeqWin sw; sw.init(sd.res + sd.off[s], L, lane);
      int M[C], I[C], D[C];
#pragma unroll
      for (int c = 0; c < C; c++) { M[c] = NEG16; I[c] = NEG16; D[c] = NEG16; }
      int xN = base_w, xB = (int16_t)(xN + xw_move), xJ = NEG16, xC = NEG16;
      int cM = NEG16, cI = NEG16, cD = NEG16;              // W > 1: previous row's cells of the left warp's last node
      bool overflow = false;
      group_sync<W>(grp);                                   // the exchange buffers of the previous comparison are free

      for (int i = 0; i < L; i++) {
        const int x = sw.get(i, lane);
        int r[C];
        load_emis<C, int>(my_rsc, STRIDE, x, lane, r);
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
        }
#pragma unroll
        for (int c = 0; c + 1 < C; c += 2) xEm = __vimax3_s32(xEm, M[c], M[c + 1]);      // one VIMNMX3 per two cells
        if (C & 1) xEm = max(xEm, M[C - 1]);
        int xE = __reduce_max_sync(FULL, xEm);
        // M->D partials: D[c] is the value entering node c from M of node c-1
        const int mdl = __viaddmax_s32(M[C - 1], tMD[C - 1], NEG16);
        int dleft = __shfl_up_sync(FULL, mdl, 1);
        if (lane == 0) dleft = NEG16;
        int Dm = dleft;
        D[0] = dleft;
#pragma unroll
        for (int c = 1; c < C; c++) D[c] = __viaddmax_s32(M[c - 1], tMD[c - 1], NEG16);
#pragma unroll
        for (int c = 1; c + 1 < C; c += 2) Dm = __vimax3_s32(Dm, D[c], D[c + 1]);
        if (!(C & 1)) Dm = max(Dm, D[C - 1]);
        int Dmax = __reduce_max_sync(FULL, Dm);
        if (W > 1) {
          int (*X)[W] = s_x[i & 1][grp];
          if (lane == 31) { X[0][wi] = M[C - 1]; X[1][wi] = I[C - 1]; X[2][wi] = D[C - 1]; X[3][wi] = mdl; }
          if (lane == 0)  { X[4][wi] = xE; X[5][wi] = Dmax; }
          group_sync<W>(grp);
          xE = X[4][0]; Dmax = X[5][0];
#pragma unroll
          for (int w = 1; w < W; w++) { xE = max(xE, X[4][w]); Dmax = max(Dmax, max(X[5][w], X[3][w - 1])); }
          if (wi > 0) {
            cM = X[0][wi - 1]; cI = X[1][wi - 1]; cD = X[2][wi - 1];
            if (lane == 0) D[0] = X[3][wi - 1];
          }
        }
        if (xE >= 32767) { overflow = true; break; }
        xC = (int16_t)max(xC, xE + xwEm);
        xJ = (int16_t)max(xJ, xE + xwEl);
        xB = (int16_t)max(xJ + xw_move, xN + xw_move);
        if (Dmax + ddbound > xB) {
          // close the D->D chain: serial inside the lane, then a max-plus scan over the 32 lane composites
          int T[C];
          T[0] = tDDin;
#pragma unroll
          for (int c = 1; c < C; c++) { D[c] = __viaddmax_s32(D[c - 1], tDD[c - 1], D[c]); T[c] = T[c - 1] + tDD[c - 1]; }
          int A = D[C - 1], Tt = T[C - 1];                   // lane composite: d -> max(A, d + Tt)
#pragma unroll
          for (int dlt = 1; dlt < 32; dlt <<= 1) {
            const int A2 = __shfl_up_sync(FULL, A, dlt), T2 = __shfl_up_sync(FULL, Tt, dlt);
            if (lane >= dlt) { A = max(A, A2 + Tt); Tt = max(T2 + Tt, -(1 << 29)); }
          }
          int din = __shfl_up_sync(FULL, A, 1);              // D of the previous lane's last node
          if (lane == 0) din = NEG16;
          if (W > 1) {
            int (*Y)[W] = s_y[grp];
            if (lane == 31) { Y[0][wi] = A; Y[1][wi] = Tt; }
            const int Tp = __shfl_up_sync(FULL, Tt, 1);
            group_sync<W>(grp);
            int d = NEG16;                                   // D of the last node of the warp to the left, closed
#pragma unroll
            for (int w = 0; w < W - 1; w++) if (w < wi) d = max(Y[0][w], max(d + Y[1][w], NEG16));
            if (wi > 0) { cD = d; din = (lane == 0) ? d : max(din, max(d + Tp, NEG16)); }
          }
#pragma unroll
          for (int c = 0; c < C; c++) D[c] = max(D[c], max(din + T[c], NEG16));
        }
      }
      if (gl == 0) {
        float sc; int st = B2H_OK;
        if (overflow) { sc = INFINITY; st = B2H_ERANGE; }
        else if (xC > NEG16) { sc = (float)xC + (float)xw_move - (float)base_w; sc /= P.scale_w; sc -= 3.0f; }
        else sc = -INFINITY;
        out.sc[e] = sc;
        if (out.status) out.status[e] = st;
      }
    }
  }
}

// =================================================================================================
// ViterbiFilter in packed 16-bit lanes (single-warp classes): two model nodes per 32-bit register, every term of the
// recurrence ONE VIADDMNMX.S16x2 for two cells -- half the ALU work of rvit_kernel, which is ALU-pipe bound.
//
// Lane z owns the C consecutive nodes z*C .. z*C+C-1; register j (0 <= j < H = C/2) packs node j (low half) and node
// j+H (high half), so "the cell of node k-1" is register j-1 for j >= 1 and one shuffle + PRMT for j = 0 (as in the SSV
// kernel).  DPX adds do not saturate, so exactness is obtained differently from the reference's _mm_adds_epi16:
//   * state cells hold (true value + V2_SIG) and are clamped from below at V2_FLOOR, table values at V2_TF, and the
//     row maximum is kept below V2_HI; with these constants no 16-bit sum wraps (b2h_internal.h);
//   * every true value >= V2_LO is computed exactly; values below V2_LO are only known to be below V2_LO.  Each M cell's
//     four-way maximum contains the begin term xB + tBM >= V2_LO (checked per comparison), so inexact low terms never
//     win there, and they stay low in the I and D chains because transitions are <= 0 (proof in DESIGN.md);
//   * a comparison that leaves this regime (xE >= V2_HI: every strong hit; begin floor below V2_LO; final xC below
//     V2_LO) is flagged B2H_REDO and decided by rvit_kernel in a second launch on the same stream (redo_only).
// The D->D closure (rare on the random-sequence bulk) stays packed as well; only the scan of the 32 lane composites is 32-bit.
// =================================================================================================
__device__ __forceinline__ uint32_t pack2(int lo, int hi) { return ((uint32_t)hi << 16) | ((uint32_t)lo & 0xffffu); }
__device__ __forceinline__ int lo16(uint32_t v) { return (int)(int16_t)(v & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t v) { return ((int)v) >> 16; }

__device__ __forceinline__ uint32_t dp_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr)); return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v;
}

template <int H>
__device__ __forceinline__ void load_emis2(const uint32_t *tab, int x, int lane, uint32_t (&r)[H])
{
  const uint32_t *row = tab + (size_t)x * (32 * H);
  if (H % 4 == 0) {
#pragma unroll
    for (int g = 0; g < H / 4; g++) { const uint4 v = *reinterpret_cast<const uint4 *>(row + g * 128 + lane * 4); r[4*g] = v.x; r[4*g+1] = v.y; r[4*g+2] = v.z; r[4*g+3] = v.w; }
  } else if (H % 2 == 0) {
#pragma unroll
    for (int g = 0; g < H / 2; g++) { const uint2 v = *reinterpret_cast<const uint2 *>(row + g * 64 + lane * 2); r[2*g] = v.x; r[2*g+1] = v.y; }
  } else {
#pragma unroll
    for (int g = 0; g < H; g++) r[g] = row[g * 32 + lane];
  }
}

template <int C>
__global__ void __launch_bounds__(256, (C <= 4) ? 4 : (C <= 12) ? 2 : 1) rvit2_kernel(const WorkList wl, const SeqDev sd, const StageOut out)
{
  constexpr int H = C / 2;
  extern __shared__ __align__(128) uint32_t s_rsc2[];       // [32][32*H] packed emission scores
  __shared__ uint64_t s_bar;
  __shared__ int s_item;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  constexpr uint32_t TAB_BYTES = 32u * 32u * H * 4u;
  constexpr uint32_t ROWB2 = 32u * H * 4u;                   // bytes of one residue's row of packed emission scores
  constexpr int LW = (H % 4 == 0) ? 16 : (H % 2 == 0) ? 8 : 4;   // bytes a lane reads per load (LDS.128 / .64 / .32)
  constexpr uint32_t FLOOR2 = ((uint32_t)(uint16_t)(int16_t)B2H_V2_FLOOR << 16) | (uint32_t)(uint16_t)(int16_t)B2H_V2_FLOOR;
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  uint32_t phase = 0;
  int cur_p = -1;
  const uint32_t tab_lane = dp_smem_u32(s_rsc2) + lane * LW;
  uint32_t tBM[H], tMM[H], tIM[H], tDM[H], tMD[H], tMI[H], tII[H], tDD[H];
  uint32_t tlink[H], tfull[H];                             // (0, T_hi[j]) and (T[j], T[H+j]): see the D->D closure
  int tDDin0 = B2H_V2_TF;
  Item it;
  while (next_item(wl, &s_item, it)) {
    const ProfDev &P = wl.profs[it.p];
    if (it.p != cur_p) {
      cur_p = it.p;
      if (threadIdx.x == 0) { mbar_expect_tx(&s_bar, TAB_BYTES); tma_load_1d(s_rsc2, P.vit_rsc2, TAB_BYTES, &s_bar); }
      const int16_t *ts = P.vit_tsc; const int Mp = P.Mpad;
      auto tv = [&](int t, int k0) -> int { return (k0 < Mp) ? max((int)ts[t * Mp + k0], B2H_V2_TF) : B2H_V2_TF; };
#pragma unroll
      for (int j = 0; j < H; j++) {
        const int ka = lane * C + j, kb = ka + H;           // 0-based node columns of the two halves
        tBM[j] = pack2(tv(0, ka), tv(0, kb)); tMM[j] = pack2(tv(1, ka), tv(1, kb)); tIM[j] = pack2(tv(2, ka), tv(2, kb));
        tDM[j] = pack2(tv(3, ka), tv(3, kb)); tMD[j] = pack2(tv(4, ka), tv(4, kb)); tMI[j] = pack2(tv(5, ka), tv(5, kb));
        tII[j] = pack2(tv(6, ka), tv(6, kb)); tDD[j] = pack2(tv(7, ka), tv(7, kb));
      }
      if (lane == 0) {                                          // nothing enters the model's first node from "node 0" (see row())
        constexpr uint32_t TF_LO = (uint32_t)(uint16_t)(int16_t)B2H_V2_TF;
        tMM[0] = (tMM[0] & 0xffff0000u) | TF_LO; tIM[0] = (tIM[0] & 0xffff0000u) | TF_LO; tDM[0] = (tDM[0] & 0xffff0000u) | TF_LO;
      }
      tDDin0 = __shfl_up_sync(FULL, hi16(tDD[H - 1]), 1);      // D_{k-1} -> D_k for this lane's first node
      if (lane == 0) tDDin0 = B2H_V2_TF;
      {                                                        // prefix sums of tDD (clamped at V2_TF: anything lower only yields values below V2_LO)
        int tlo = tDDin0, thi = 0;
#pragma unroll
        for (int j = 0; j < H; j++) { tlink[j] = pack2(0, thi); thi = max(thi + hi16(tDD[j]), B2H_V2_TF); }
        int tl2[H];
#pragma unroll
        for (int j = 0; j < H; j++) { tl2[j] = tlo; tlo = max(tlo + lo16(tDD[j]), B2H_V2_TF); }
        thi = tlo;                                             // tlo is now the sum up to the first node of the high half
#pragma unroll
        for (int j = 0; j < H; j++) { tfull[j] = pack2(tl2[j], thi); thi = max(thi + hi16(tDD[j]), B2H_V2_TF); }
      }
      mbar_wait(&s_bar, phase); phase ^= 1;
    }
    const int xwEm = P.xw_E_move, xwEl = P.xw_E_loop, base_w = P.base_w, ddbound = P.ddbound_w;

    for (int e = it.e_begin + warp; e < it.e_end; e += nwarps) {
      const int s = wl.ent_s[e];
      const int L = sd.len[s];
      const int xw_move = sd.xwmove[s];
      bool redo = !P.v2_ok || (base_w + xw_move + P.tbm_min < B2H_V2_LO);
      const uint32_t *seqw = reinterpret_cast<const uint32_t *>(sd.res + sd.off[s]);
      const int nwords = (L + 3) >> 2;
      uint32_t M[H], I[H], D[H];
#pragma unroll
      for (int j = 0; j < H; j++) { M[j] = FLOOR2; I[j] = FLOOR2; D[j] = FLOOR2; }
      // specials in true coordinates.  The reference keeps them in int16 with saturating adds; here they are plain ints:
      // xE < V2_HI on every row that is kept, the loop / move scores are <= 0 and every maximum below contains a term that
      // is representable, so no value ever leaves the int16 range and the two arithmetics agree.
      const int xNm = base_w + xw_move;                       // xN + tNB: N never changes (tNN = 0 in the filter)
      int xB = xNm, xJ = NEG16, xC = NEG16;

      // One DP row for residue code x (byte address of its table row = x * ROWB2).  Returns false when the packed
      // arithmetic cannot decide this comparison (row maximum at or above V2_HI).
      auto row = [&](uint32_t x) -> bool {
        uint32_t r[H];
        {
          const uint32_t a = tab_lane + x * ROWB2;
          if (H % 4 == 0) {
#pragma unroll
            for (int g = 0; g < H / 4; g++) { const uint4 v = lds128(a + g * 512); r[4*g] = v.x; r[4*g+1] = v.y; r[4*g+2] = v.z; r[4*g+3] = v.w; }
          } else if (H % 2 == 0) {
#pragma unroll
            for (int g = 0; g < H / 2; g++) { const uint2 v = lds64(a + g * 256); r[2*g] = v.x; r[2*g+1] = v.y; }
          } else {
#pragma unroll
            for (int g = 0; g < H; g++) r[g] = lds32(a + g * 128);
          }
        }
        // cells of node k-1 for this lane's first halves: low <- last node of the lane to the left, high <- own node H-1.
        // Lane 0 gets its own last node from the shuffle: harmless, the transitions INTO the model's first node are V2_TF
        // (set with the transition registers above), so those terms stay below V2_LO and never win against the begin term.
        const uint32_t mt = __shfl_up_sync(FULL, M[H - 1], 1), itt = __shfl_up_sync(FULL, I[H - 1], 1), dt = __shfl_up_sync(FULL, D[H - 1], 1);
        const uint32_t m0 = __byte_perm(mt, M[H - 1], 0x5432u), i0 = __byte_perm(itt, I[H - 1], 0x5432u), d0 = __byte_perm(dt, D[H - 1], 0x5432u);
        const uint32_t xB2 = __byte_perm((uint32_t)(xB + B2H_V2_SIG), 0u, 0x1010u);
#pragma unroll
        for (int j = H - 1; j >= 0; j--) {
          const uint32_t pm = (j == 0) ? m0 : M[j - 1], pi = (j == 0) ? i0 : I[j - 1], pd = (j == 0) ? d0 : D[j - 1];
          const uint32_t inew = __viaddmax_s16x2(M[j], tMI[j], __viaddmax_s16x2(I[j], tII[j], FLOOR2));
          uint32_t m = __viaddmax_s16x2(xB2, tBM[j], FLOOR2);
          m = __viaddmax_s16x2(pm, tMM[j], m);
          m = __viaddmax_s16x2(pi, tIM[j], m);
          m = __viaddmax_s16x2(pd, tDM[j], m);
          m = __viaddmax_s16x2(m, r[j], FLOOR2);
          M[j] = m; I[j] = inew;
        }
        uint32_t xe2 = FLOOR2;
#pragma unroll
        for (int j = 0; j + 1 < H; j += 2) xe2 = __vimax3_s16x2(xe2, M[j], M[j + 1]);
        if (H & 1) xe2 = __vimax3_s16x2(xe2, M[H - 1], M[H - 1]);
        xe2 = __vmaxs2(xe2, __byte_perm(xe2, 0u, 0x1032u));        // both halves = the lane's maximum
        const int xE = __reduce_max_sync(FULL, (int)(int16_t)xe2) - B2H_V2_SIG;
        // M->D partials: D of node k is what enters from M of node k-1
        const uint32_t mdl = __viaddmax_s16x2(M[H - 1], tMD[H - 1], FLOOR2);
        uint32_t mdt = __shfl_up_sync(FULL, mdl, 1);
        if (lane == 0) mdt = FLOOR2;
#pragma unroll
        for (int j = H - 1; j >= 1; j--) D[j] = __viaddmax_s16x2(M[j - 1], tMD[j - 1], FLOOR2);
        D[0] = __byte_perm(mdt, mdl, 0x5432u);
        uint32_t dm2 = FLOOR2;
#pragma unroll
        for (int j = 0; j + 1 < H; j += 2) dm2 = __vimax3_s16x2(dm2, D[j], D[j + 1]);
        if (H & 1) dm2 = __vimax3_s16x2(dm2, D[H - 1], D[H - 1]);
        dm2 = __vmaxs2(dm2, __byte_perm(dm2, 0u, 0x1032u));
        const int Dmax = __reduce_max_sync(FULL, (int)(int16_t)dm2) - B2H_V2_SIG;
        if (xE >= B2H_V2_HI) return false;
        xC = max(xC, xE + xwEm);
        xJ = max(xJ, xE + xwEl);
        xB = max(xJ + xw_move, xNm);
        if (Dmax + ddbound > xB) {
          // Close the D->D chain in packed form.  (1) the two half-chains of the lane side by side; (2) the end of the
          // low half-chain enters the high one; (3) lane-to-lane: the closed D of the left lane's last node enters a lane
          // through its first node and runs down its prefix sums of tDD.  Delete runs are short, so instead of a full
          // 5-step scan the lanes hand their last node to the right neighbour until no first node improves any more
          // (if no lane's first node is improved by its neighbour's current value, the closure is complete: a change
          // would have to enter the leftmost changed lane through its first node) -- one or two rounds in practice.
#pragma unroll
          for (int j = 1; j < H; j++) D[j] = __viaddmax_s16x2(D[j - 1], tDD[j - 1], D[j]);
          {
            const int xin = max(lo16(D[H - 1]) + lo16(tDD[H - 1]), B2H_V2_FLOOR);
            const uint32_t X2 = pack2(B2H_V2_FLOOR, xin);
#pragma unroll
            for (int j = 0; j < H; j++) D[j] = __viaddmax_s16x2(X2, tlink[j], D[j]);
          }
          for (;;) {
            int din = __shfl_up_sync(FULL, hi16(D[H - 1]), 1);
            if (lane == 0) din = B2H_V2_FLOOR;
            const bool improves = din + tDDin0 > lo16(D[0]);
            if (!__any_sync(FULL, improves)) break;
            if (improves) {
              const uint32_t DIN2 = pack2(din, din);
#pragma unroll
              for (int j = 0; j < H; j++) D[j] = __viaddmax_s16x2(DIN2, tfull[j], D[j]);
            }
          }
        }
        return true;
      };

      // residues come as 32-bit words of four (the arena pads every sequence to 16 bytes): each lane keeps one word of the
      // current 128-row window, a word is broadcast once and its four rows are unrolled with constant byte selectors
      uint32_t myw = 0;
      const int nfull = L >> 2;
      int w = 0;
      if (!redo) {
        for (; w < nfull; w++) {
          if ((w & 31) == 0) myw = (w + lane < nwords) ? __ldg(seqw + w + lane) : 0x1f1f1f1fu;
          const uint32_t wr = __shfl_sync(FULL, myw, w & 31);
          if (!row(__byte_perm(wr, 0u, 0x4440u))) { redo = true; break; }
          if (!row(__byte_perm(wr, 0u, 0x4441u))) { redo = true; break; }
          if (!row(__byte_perm(wr, 0u, 0x4442u))) { redo = true; break; }
          if (!row(__byte_perm(wr, 0u, 0x4443u))) { redo = true; break; }
        }
        if (!redo && (L & 3)) {
          if ((w & 31) == 0) myw = (w + lane < nwords) ? __ldg(seqw + w + lane) : 0x1f1f1f1fu;
          const uint32_t wr = __shfl_sync(FULL, myw, w & 31);
          for (int rr = 0; rr < (L & 3); rr++)
            if (!row((wr >> (8 * rr)) & 0xffu)) { redo = true; break; }
        }
      }
      if (!redo && xC < B2H_V2_LO) redo = true;             // nothing exact reached E (e.g. a target of impossible residues)
      if (lane == 0) {
        if (redo) out.status[e] = B2H_REDO;
        else {
          float sc = (float)xC + (float)xw_move - (float)base_w; sc /= P.scale_w; sc -= 3.0f;
          out.sc[e] = sc;
          out.status[e] = B2H_OK;
        }
      }
    }
  }
}

// =================================================================================================
// Forward parser, register resident
// =================================================================================================
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

template <int C, int W>
__global__ void __launch_bounds__(256) rfwd_kernel(const WorkList wl, const SeqDev sd, const StageOut out)
{
  extern __shared__ __align__(128) float s_rscf[];          // [32][W][32*C] fp32 emission odds
  __shared__ uint64_t s_bar;
  __shared__ int s_item;
  __shared__ float s_x[2][MAXGRP][8][W];                    // per row parity and warp: A, T, aout, tDDlast, M_last, I_last, S1, S2
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int grp = warp / W, wi = warp % W, ngrp = nwarps / W;
  const int gl = wi * 32 + lane;
  constexpr int STRIDE = 32 * C * W;
  constexpr uint32_t TAB_BYTES = 32u * STRIDE * 4u;
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  uint32_t phase = 0;
  int cur_p = -1;
  float tBM[C], tMM[C], tIM[C], tDM[C], tMD[C], tMI[C], tII[C], tDD[C];
  float tDDin = 0.f;
  Item it;
  while (next_item<ItemSplit<W>::SUB>(wl, &s_item, it)) {
    const ProfDev &P = wl.profs[it.p];
    if (it.p != cur_p) {
      cur_p = it.p;
      if (threadIdx.x == 0) { mbar_expect_tx(&s_bar, TAB_BYTES); tma_load_1d(s_rscf, P.fwd_rscr, TAB_BYTES, &s_bar); }
      const float *ts = P.fwd_tsc; const int Mp = P.Mpad;
#pragma unroll
      for (int c = 0; c < C; c++) {
        const int k0 = gl * C + c;
        const bool in = k0 < Mp;
        tBM[c] = in ? ts[0 * Mp + k0] : 0.f; tMM[c] = in ? ts[1 * Mp + k0] : 0.f; tIM[c] = in ? ts[2 * Mp + k0] : 0.f;
        tDM[c] = in ? ts[3 * Mp + k0] : 0.f; tMD[c] = in ? ts[4 * Mp + k0] : 0.f; tMI[c] = in ? ts[5 * Mp + k0] : 0.f;
        tII[c] = in ? ts[6 * Mp + k0] : 0.f; tDD[c] = in ? ts[7 * Mp + k0] : 0.f;
      }
      tDDin = __shfl_up_sync(FULL, tDD[C - 1], 1);
      // W > 1: the value entering the first node of warp wi > 0 is handed over complete (o_in below), hence factor 1
      if (lane == 0) tDDin = (W > 1 && wi > 0) ? 1.0f : 0.f;
      mbar_wait(&s_bar, phase); phase ^= 1;
    }
    const float tEC = P.xf_E_move, tEJ = P.xf_E_loop;
    const float *my_rsc = s_rscf + wi * 32 * C;

    for (int e = it.e_begin + grp; e < it.e_end; e += ngrp) {
      const int s = wl.ent_s[e];
      const int L = sd.len[s];
      const float pmove = sd.pmove[s], ploop = 1.0f - pmove;
      float *xout = (out.fwd_xmx && out.xoff[e] >= 0) ? out.fwd_xmx + out.xoff[e] * 6 : nullptr;
      SeqWin sw; sw.init(sd.res + sd.off[s], L, lane);
      float M[C], I[C], D[C];
#pragma unroll
      for (int c = 0; c < C; c++) { M[c] = 0.f; I[c] = 0.f; D[c] = 0.f; }
      float xN = 1.0f, xJ = 0.0f, xC = 0.0f, xB = pmove, xE = 0.0f, totscale = 0.0f;
      float cM = 0.f, cI = 0.f, cD = 0.f;                   // W > 1: previous row's cells of the left warp's last node
      if (xout && gl == 0) { xout[0] = 0.f; xout[1] = 1.f; xout[2] = 0.f; xout[3] = xB; xout[4] = 0.f; xout[5] = 1.f; }
      group_sync<W>(grp);

      for (int i = 1; i <= L; i++) {
        const int x = sw.get(i - 1, lane);
        float r[C];
        load_emis<C, float>(my_rsc, STRIDE, x, lane, r);
        float mp = __shfl_up_sync(FULL, M[C - 1], 1), ip = __shfl_up_sync(FULL, I[C - 1], 1), dp = __shfl_up_sync(FULL, D[C - 1], 1);
        if (lane == 0) { mp = cM; ip = cI; dp = cD; }
        float esum = 0.f;
#pragma unroll
        for (int c = C - 1; c >= 0; c--) {
          const float pm = (c == 0) ? mp : M[c - 1], pi = (c == 0) ? ip : I[c - 1], pd = (c == 0) ? dp : D[c - 1];
          const float inew = M[c] * tMI[c] + I[c] * tII[c];
          float m = xB * tBM[c];
          m += pm * tMM[c];
          m += pi * tIM[c];
          m += pd * tDM[c];
          m *= r[c];
          esum += m;
          M[c] = m; I[c] = inew;
        }
        // D chain: D(k) = M(k-1)*tMD(k-1) + D(k-1)*tDD(k-1); two-level affine scan
        const float aout = M[C - 1] * tMD[C - 1];
        float aleft = __shfl_up_sync(FULL, aout, 1);
        if (lane == 0) aleft = 0.f;
        float T[C];
        D[0] = aleft; T[0] = tDDin;
#pragma unroll
        for (int c = 1; c < C; c++) { D[c] = M[c - 1] * tMD[c - 1] + D[c - 1] * tDD[c - 1]; T[c] = T[c - 1] * tDD[c - 1]; }
        float A = D[C - 1], Tt = T[C - 1];
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
          const float A2 = __shfl_up_sync(FULL, A, dlt), T2 = __shfl_up_sync(FULL, Tt, dlt);
          if (lane >= dlt) { A = A + A2 * Tt; Tt = T2 * Tt; }
        }
        float din = __shfl_up_sync(FULL, A, 1);
        if (lane == 0) din = 0.f;
        if (W == 1) {
#pragma unroll
          for (int c = 0; c < C; c++) { D[c] = D[c] + din * T[c]; esum += D[c]; }
          xE = warp_sum(esum);
        } else {
          float Tp = __shfl_up_sync(FULL, Tt, 1);             // product of tDD from the warp's entry point to this lane's entry point
          if (lane == 0) Tp = 1.0f;
          float sT = 0.f;
#pragma unroll
          for (int c = 0; c < C; c++) { D[c] = D[c] + din * T[c]; esum += D[c]; sT += T[c]; }
          const float S1 = warp_sum(esum), S2 = warp_sum(Tp * sT);
          float (*X)[W] = s_x[i & 1][grp];
          if (lane == 31) { X[0][wi] = A; X[1][wi] = Tt; X[2][wi] = aout; X[3][wi] = tDD[C - 1]; X[4][wi] = M[C - 1]; X[5][wi] = I[C - 1]; }
          if (lane == 0)  { X[6][wi] = S1; X[7][wi] = S2; }
          group_sync<W>(grp);
          float o = 0.f, o_in = 0.f;                          // o: D entering the first node of the next warp
          xE = 0.f;
#pragma unroll
          for (int w = 0; w < W; w++) {
            xE += X[6][w] + o * X[7][w];
            if (w == wi) o_in = o;
            const float Dl = X[0][w] + X[1][w] * o;           // D of warp w's last node
            if (w == wi - 1) { cM = X[4][w]; cI = X[5][w]; cD = Dl; }
            o = X[2][w] + Dl * X[3][w];
          }
          if (wi > 0) {
            const float oi = o_in * Tp;
#pragma unroll
            for (int c = 0; c < C; c++) D[c] = D[c] + oi * T[c];
          }
        }
        xN = xN * ploop;
        xC = (xC * ploop) + (xE * tEC);
        xJ = (xJ * ploop) + (xE * tEJ);
        xB = (xJ * pmove) + (xN * pmove);
        float scale = 1.0f;
        if (xE > 1.0e4f) {
          xN = xN / xE; xC = xC / xE; xJ = xJ / xE; xB = xB / xE;
          const float inv = 1.0f / xE;
#pragma unroll
          for (int c = 0; c < C; c++) { M[c] *= inv; I[c] *= inv; D[c] *= inv; }
          if (W > 1) { cM *= inv; cI *= inv; cD *= inv; }
          scale = xE;
          totscale = (float)((double)totscale + log((double)xE));
          xE = 1.0f;
        }
        if (xout && gl == 0) { float *q = xout + (size_t)i * 6; q[0] = xE; q[1] = xN; q[2] = xJ; q[3] = xB; q[4] = xC; q[5] = scale; }
      }
      if (gl == 0) {
        int st = B2H_OK; float sc;
        if (isnan(xC) || (L > 0 && xC == 0.0f) || isinf(xC)) { st = B2H_ERANGE; sc = isnan(xC) ? NAN : (xC == 0.0f ? -INFINITY : INFINITY); }
        else sc = (float)((double)totscale + log((double)(xC * pmove)));
        out.sc[e] = sc;
        if (out.status) out.status[e] = st;
      }
    }
  }
}

// =================================================================================================
// Backward parser, register resident
// =================================================================================================
template <int C, int W>
__global__ void __launch_bounds__(256) rbck_kernel(const WorkList wl, const SeqDev sd, const StageOut out)
{
  extern __shared__ __align__(128) float s_rscf[];
  __shared__ uint64_t s_bar;
  __shared__ int s_item;
  __shared__ float s_x[MAXGRP][4][W];                       // per warp: bsum partial, composite (A, T) of the D chain, M of its first node
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int grp = warp / W, wi = warp % W, ngrp = nwarps / W;
  const int gl = wi * 32 + lane;
  constexpr int STRIDE = 32 * C * W;
  constexpr uint32_t TAB_BYTES = 32u * STRIDE * 4u;
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  uint32_t phase = 0;
  int cur_p = -1;
  // per node k (own): tBM[k], tMD[k], tMI[k], tII[k], tDD[k]; and towards k+1: tMMn = tMM[k+1], tIMn = tIM[k+1], tDMn = tDM[k+1]
  float tBM[C], tMD[C], tMI[C], tII[C], tDD[C], tMMn[C], tIMn[C], tDMn[C];
  Item it;
  while (next_item<ItemSplit<W>::SUB>(wl, &s_item, it)) {
    const ProfDev &P = wl.profs[it.p];
    const int M = P.M;
    if (it.p != cur_p) {
      cur_p = it.p;
      if (threadIdx.x == 0) { mbar_expect_tx(&s_bar, TAB_BYTES); tma_load_1d(s_rscf, P.fwd_rscr, TAB_BYTES, &s_bar); }
      const float *ts = P.fwd_tsc; const int Mp = P.Mpad;
#pragma unroll
      for (int c = 0; c < C; c++) {
        const int k0 = gl * C + c;
        const bool in = k0 < M, inn = (k0 + 1) < M;
        tBM[c] = in ? ts[0 * Mp + k0] : 0.f; tMD[c] = in ? ts[4 * Mp + k0] : 0.f; tMI[c] = in ? ts[5 * Mp + k0] : 0.f;
        tII[c] = in ? ts[6 * Mp + k0] : 0.f; tDD[c] = in ? ts[7 * Mp + k0] : 0.f;
        tMMn[c] = inn ? ts[1 * Mp + k0 + 1] : 0.f; tIMn[c] = inn ? ts[2 * Mp + k0 + 1] : 0.f; tDMn[c] = inn ? ts[3 * Mp + k0 + 1] : 0.f;
      }
      mbar_wait(&s_bar, phase); phase ^= 1;
    }
    const float tEC = P.xf_E_move, tEJ = P.xf_E_loop;
    const float *my_rsc = s_rscf + wi * 32 * C;
    float (*X)[W] = s_x[grp];

    for (int e = it.e_begin + grp; e < it.e_end; e += ngrp) {
      const int s = wl.ent_s[e];
      const int L = sd.len[s];
      const uint8_t *seq = sd.res + sd.off[s];
      const float pmove = sd.pmove[s], ploop = 1.0f - pmove;
      const float *fx = out.fwd_xmx + out.xoff[e] * 6;
      float *bx = out.bck_xmx ? out.bck_xmx + out.xoff[e] * 6 : nullptr;
      float Mv[C], Iv[C], Dv[C];
      float xJ = 0.0f, xB = 0.0f, xN = 0.0f, xC = pmove, xE = xC * tEC;
      bool own_scales = false;
      float totscale;
      float dext = 0.f;                                     // W > 1: closed D of the first node of the warp to the right

      // reverse two-level affine closure:  D(k) = a(k) + tDD(k) * D(k+1)   (a given in Dv, result in Dv)
      auto close_dd = [&](void) {
        float T[C];
        T[C - 1] = tDD[C - 1];
#pragma unroll
        for (int c = C - 2; c >= 0; c--) { Dv[c] = Dv[c] + tDD[c] * Dv[c + 1]; T[c] = tDD[c] * T[c + 1]; }
        // note: Dv[C-1] still lacks the contribution of the next lane's first node
        float A = Dv[0], Tt = T[0];                        // lane composite acting on D(first node of next lane): D(first) = A + Tt * d
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
          const float A2 = __shfl_down_sync(FULL, A, dlt), T2 = __shfl_down_sync(FULL, Tt, dlt);
          if (lane + dlt < 32) { A = A + Tt * A2; Tt = Tt * T2; }
        }
        float din = __shfl_down_sync(FULL, A, 1);          // D of the next lane's first node
        if (lane == 31) din = 0.f;
        if (W > 1) {
          if (lane == 0) { X[1][wi] = A; X[2][wi] = Tt; }
          const float Tn = __shfl_down_sync(FULL, Tt, 1);
          group_sync<W>(grp);
          float d = 0.f;
#pragma unroll
          for (int w = W - 1; w >= 1; w--) if (w > wi) d = X[1][w] + X[2][w] * d;
          dext = d;
          din = (lane == 31) ? d : din + Tn * d;
        }
#pragma unroll
        for (int c = 0; c < C; c++) Dv[c] = Dv[c] + T[c] * din;
      };
      // sum of the warps' partials of xB (every warp adds them in the same order)
      auto group_sum = [&](float v) -> float {
        if (W == 1) return v;
        if (lane == 0) X[0][wi] = v;
        group_sync<W>(grp);
        float t = X[0][0];
#pragma unroll
        for (int w = 1; w < W; w++) t += X[0][w];
        return t;
      };
      group_sync<W>(grp);

      // row L
#pragma unroll
      for (int c = 0; c < C; c++) { const bool in = (gl * C + c) < M; Dv[c] = in ? xE : 0.f; Iv[c] = 0.f; }
      close_dd();
      {
        float dnext = __shfl_down_sync(FULL, Dv[0], 1); if (lane == 31) dnext = (W > 1) ? dext : 0.f;
#pragma unroll
        for (int c = 0; c < C; c++) { const bool in = (gl * C + c) < M; const float dn = (c == C - 1) ? dnext : Dv[c + 1]; Mv[c] = in ? xE + tMD[c] * dn : 0.f; }
        const float scL = fx[(size_t)L * 6 + 5];
        if (scL > 1.0f) {
          xE = xE / scL; xN = xN / scL; xC = xC / scL; xJ = xJ / scL; xB = xB / scL;
          const float inv = 1.0f / scL;
#pragma unroll
          for (int c = 0; c < C; c++) { Mv[c] *= inv; Dv[c] *= inv; }
        }
        totscale = (float)log((double)scL);
        if (bx && gl == 0) { float *q = bx + (size_t)L * 6; q[0] = xE; q[1] = xN; q[2] = xJ; q[3] = xB; q[4] = xC; q[5] = scL; }
        if (W > 1 && lane == 0) X[3][wi] = Mv[0];
      }

      for (int i = L - 1; i >= 1; i--) {
        const int x = seq[i];                               // x_{i+1}
        float r[C];
        load_emis<C, float>(my_rsc, STRIDE, x, lane, r);
        // mpv(k) = M(i+1,k+1) * e(k+1): own nodes shifted down by one
        float me[C];
#pragma unroll
        for (int c = 0; c < C; c++) me[c] = Mv[c] * r[c];   // M(i+1,k) e(k)
        float bsum = 0.f;
#pragma unroll
        for (int c = 0; c < C; c++) bsum += me[c] * tBM[c];
        xB = group_sum(warp_sum(bsum));                     // (W > 1: also orders the X[3] hand-over of the previous row)
        float menext = __shfl_down_sync(FULL, me[0], 1);
        if (lane == 31) menext = (W > 1 && wi < W - 1) ? X[3][wi + 1] * s_rscf[(size_t)x * STRIDE + (wi + 1) * 32 * C] : 0.f;
#pragma unroll
        for (int c = 0; c < C; c++) {
          const float mpv = (c == C - 1) ? menext : me[c + 1];
          const float ipv = Iv[c];
          Iv[c] = (ipv * tII[c]) + (mpv * tIMn[c]);
          Dv[c] = mpv * tDMn[c];
          Mv[c] = (ipv * tMI[c]) + (mpv * tMMn[c]);
        }
        xC = xC * ploop;
        xJ = (xB * pmove) + (xJ * ploop);
        xN = (xB * pmove) + (xN * ploop);
        xE = (xC * tEC) + (xJ * tEJ);
#pragma unroll
        for (int c = 0; c < C; c++) { const bool in = (gl * C + c) < M; Dv[c] = in ? Dv[c] + xE : 0.f; }
        close_dd();
        float dnext = __shfl_down_sync(FULL, Dv[0], 1); if (lane == 31) dnext = (W > 1) ? dext : 0.f;
#pragma unroll
        for (int c = 0; c < C; c++) { const bool in = (gl * C + c) < M; const float dn = (c == C - 1) ? dnext : Dv[c + 1]; Mv[c] = in ? (Mv[c] + xE) + tMD[c] * dn : 0.f; }
        if (xB > 1.0e16f) own_scales = true;
        const float scale = own_scales ? ((xB > 1.0e4f) ? xB : 1.0f) : fx[(size_t)i * 6 + 5];
        if (scale > 1.0f) {
          xE /= scale; xN /= scale; xJ /= scale; xB /= scale; xC /= scale;
          const float inv = 1.0f / scale;
#pragma unroll
          for (int c = 0; c < C; c++) { Mv[c] *= inv; Dv[c] *= inv; Iv[c] *= inv; }
          totscale = (float)((double)totscale + log((double)scale));
        }
        if (bx && gl == 0) { float *q = bx + (size_t)i * 6; q[0] = xE; q[1] = xN; q[2] = xJ; q[3] = xB; q[4] = xC; q[5] = scale; }
        if (W > 1 && lane == 0) X[3][wi] = Mv[0];
      }
      {
        float r[C];
        load_emis<C, float>(my_rsc, STRIDE, (int)seq[0], lane, r);
        float bsum = 0.f;
        if (L >= 1) {
#pragma unroll
          for (int c = 0; c < C; c++) bsum += (Mv[c] * r[c]) * tBM[c];
        }
        xB = group_sum(warp_sum(bsum));
        xN = (xB * pmove) + (xN * ploop);
        if (gl == 0) {
          if (bx) { bx[0] = 0.f; bx[1] = xN; bx[2] = 0.f; bx[3] = xB; bx[4] = 0.f; bx[5] = 1.0f; }
          int st = B2H_OK; float sc;
          if (isnan(xN) || (L > 0 && xN == 0.0f) || isinf(xN)) { st = B2H_ERANGE; sc = isnan(xN) ? NAN : (xN == 0.0f ? -INFINITY : INFINITY); }
          else sc = (float)((double)totscale + log((double)xN));
          out.sc[e] = sc;
          if (out.status) out.status[e] = own_scales ? (st | 0x100) : st;
        }
      }
    }
  }
}

template <typename K>
int launch_reg(b2h_ctx *ctx, K kernel, int C, int W, const WorkList &wl, const SeqDev &sd, int nitems_hint, const StageOut &out, cudaStream_t strm)
{
  const size_t smem = (size_t)32 * 32 * C * W * 4;
  int occ = 1;
  { const int st = b2h_kernel_occupancy(ctx, (const void *)kernel, 256, smem, &occ); if (st != B2H_OK) return st; }
  static const int occ_cap = getenv("B2H_DP_OCC") ? atoi(getenv("B2H_DP_OCC")) : 0;         // experiments: resident CTAs per SM
  if (occ_cap > 0 && occ > occ_cap) occ = occ_cap;
  int grid = ctx->sm_count * occ;
  const int sub = (W >= 4) ? 8 : (W == 2) ? 4 : 1;            // ItemSplit<W>::SUB: CTAs of the multi-warp classes pull sub-items
  if (nitems_hint > 0 && grid > nitems_hint * sub) grid = nitems_hint * sub;
  if (grid < 1) grid = 1;
  B2H_CUDA(cudaMemsetAsync(wl.counter, 0, sizeof(int), strm));
  kernel<<<grid, 256, smem, strm>>>(wl, sd, out);
  ctx->launches++;
  B2H_CUDA(cudaGetLastError());
  return B2H_OK;
}

} // namespace

// Packed 16-bit ViterbiFilter for a single-warp class of C nodes per lane (C2 = C rounded up to even): leaves B2H_REDO in
// out.status for the comparisons the caller must hand to rvit_kernel (redo_only).
int b2h_launch_vit2(b2h_ctx *ctx, int C2, const WorkList &wl, const SeqDev &sd, int nitems_hint, StageOut out, cudaStream_t strm)
{
  if (!out.status) return B2H_EINVAL;
  auto go = [&](auto kernel) -> int {
    const size_t smem = (size_t)32 * 32 * (C2 / 2) * 4;
    int occ = 1;
    { const int st = b2h_kernel_occupancy(ctx, (const void *)kernel, 256, smem, &occ); if (st != B2H_OK) return st; }
    static const int occ_cap = getenv("B2H_VIT_OCC") ? atoi(getenv("B2H_VIT_OCC")) : 0;       // experiments: resident CTAs per SM (co-residency with SSV)
    if (occ_cap > 0 && occ > occ_cap) occ = occ_cap;
    int grid = ctx->sm_count * occ;
    if (nitems_hint > 0 && grid > nitems_hint) grid = nitems_hint;
    if (grid < 1) grid = 1;
    B2H_CUDA(cudaMemsetAsync(wl.counter, 0, sizeof(int), strm));
    kernel<<<grid, 256, smem, strm>>>(wl, sd, out);
    ctx->launches++;
    B2H_CUDA(cudaGetLastError());
    return B2H_OK;
  };
  switch (C2) {
    case 2: return go(rvit2_kernel<2>);   case 4: return go(rvit2_kernel<4>);   case 6: return go(rvit2_kernel<6>);   case 8: return go(rvit2_kernel<8>);
    case 10: return go(rvit2_kernel<10>); case 12: return go(rvit2_kernel<12>); case 14: return go(rvit2_kernel<14>); case 16: return go(rvit2_kernel<16>);
  }
  return B2H_EINVAL;
}

// kind: 0 Viterbi, 1 Forward, 2 Backward.  (C, W): nodes per lane, warps per comparison -- see b2h_reg_class().
int b2h_launch_dpreg(b2h_ctx *ctx, int kind, int C, int W, const WorkList &wl, const SeqDev &sd, int nitems_hint, StageOut out, cudaStream_t strm)
{
#define B2H_REG_CASE(CC, WW) \
    case (WW) * 64 + (CC): \
      if (kind == 0) return launch_reg(ctx, rvit_kernel<CC, WW>, CC, WW, wl, sd, nitems_hint, out, strm); \
      if (kind == 1) return launch_reg(ctx, rfwd_kernel<CC, WW>, CC, WW, wl, sd, nitems_hint, out, strm); \
      if (kind == 2) return launch_reg(ctx, rbck_kernel<CC, WW>, CC, WW, wl, sd, nitems_hint, out, strm); \
      break;
  switch (W * 64 + C) {
    B2H_REG_CASE(2, 1) B2H_REG_CASE(3, 1) B2H_REG_CASE(4, 1) B2H_REG_CASE(5, 1) B2H_REG_CASE(6, 1) B2H_REG_CASE(7, 1) B2H_REG_CASE(8, 1)
    B2H_REG_CASE(9, 1) B2H_REG_CASE(10, 1) B2H_REG_CASE(11, 1) B2H_REG_CASE(12, 1) B2H_REG_CASE(14, 1) B2H_REG_CASE(16, 1)
    B2H_REG_CASE(9, 2) B2H_REG_CASE(10, 2) B2H_REG_CASE(11, 2) B2H_REG_CASE(12, 2) B2H_REG_CASE(14, 2) B2H_REG_CASE(16, 2)
    B2H_REG_CASE(10, 4) B2H_REG_CASE(12, 4)
  }
#undef B2H_REG_CASE
  return B2H_EINVAL;
}
