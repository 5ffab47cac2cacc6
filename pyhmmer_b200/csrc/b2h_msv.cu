// b2h_msv.cu -- SSV / MSV uint8 filters as sm_100a kernels.
//
// Replaces p7_SSVFilter (impl_sse/ssvfilter.c:876-926) and p7_MSVFilter (impl_sse/msvfilter.c:74-208).
// This is a new design, not a translation of the SSE code:
//
//  * one GROUP of G lanes (1 .. 32: b2h_ssv_tile picks the tile with the fewest shared-memory wavefronts per cell) per
//    (profile x sequence) comparison, so a warp runs 32/G comparisons side by side; persistent CTAs
//    pull sequences (longest first) from a global work counter;
//  * the profile's emission table is staged ONCE per CTA into shared memory with a TMA bulk copy
//    (cp.async.bulk + mbarrier), pre-swizzled on the host so that every row step is one
//    conflict-free LDS.128 per lane and 4 registers, whatever rows the groups of a warp are reading;
//  * the DP row lives in registers as packed 16x2 cells.  Lane l of a group owns the 2*NR consecutive
//    model nodes l*2NR+1 .. (l+1)*2NR; register j packs nodes (j, j+NR) of that chunk, so the diagonal
//    move M(i,k) <- M(i-1,k-1) is a pure register renaming plus ONE warp shuffle and one PRMT per row.
//    Narrow groups make NR large (M=200: G=8, NR=13), which is what the kernel wants: its ceiling is
//    the shared-memory pipe (2 B of scores per cell), and the one shuffle per row is amortised over
//    64*NR cells of the warp while only 2*G nodes of padding can be wasted per comparison;
//  * sm_100a has no native byte SIMD (vmaxu4/vaddus4 are emulated); cells are fp16x2: every value of
//    the recurrence is an integer of magnitude < 2048 until a comparison has overflowed anyway, so fp16
//    arithmetic is exact and the cell update max(m + e, 0) is ONE HFMA2.RELU for two cells on the FMA
//    pipe; non-negative fp16 values order like their bit patterns, so the running maximum is ONE integer
//    VIMNMX3.S16x2 (DPX) for four cells on the ALU pipe.  The results are bit-identical to the reference
//    (derivation in DESIGN.md, "SSV in wide lanes").
//
// Semantics reproduced exactly: SSV is authoritative when it can prove the J state was not used;
// otherwise (eslENORESULT) the comparison is redone by the full MSV recurrence with saturating
// uint8 arithmetic; overflow returns +inf/eslERANGE.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <cmath>
#include "b2h_internal.h"

namespace {

constexpr int SSV_THREADS = 256;
constexpr uint32_t FULL = 0xffffffffu;

// ---- mbarrier / TMA bulk-copy helpers (PTX ISA 8.x; SASS: SYNCS + UBLKCP) ----
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

// Load the NR packed emission words of this lane for one residue row.  `addr` is the 32-bit shared
// address of (row word 0) + lane*16 for the full groups; `addr_rem` that of the leftover registers.
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr)); return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v;
}
template <int G, int NR>
__device__ __forceinline__ void load_row(uint32_t addr, uint32_t addr_rem, uint32_t (&e)[NR])
{
  constexpr int FULLQ = NR / 4, REM = NR % 4, GS = (G < 8) ? 8 : G;     // 16-byte slots per chunk (b2h_ssv_row_bytes)
#pragma unroll
  for (int g = 0; g < FULLQ; g++) {
    uint4 v = lds128(addr + g * (GS * 16));
    e[4*g+0] = v.x; e[4*g+1] = v.y; e[4*g+2] = v.z; e[4*g+3] = v.w;
  }
#pragma unroll
  for (int r = 0; r < REM; r++) e[4*FULLQ + r] = lds32(addr_rem + r * 128);
}

// esl_gumbel_surv (vendor/easel/esl_gumbel.c:129): same expression, double precision
__device__ __forceinline__ double gumbel_surv(double x, double mu, double lambda)
{
  const double y = lambda * (x - mu);
  const double ey = -exp(-y);
  return (fabs(ey) < 5e-9) ? -ey : 1.0 - exp(ey);
}

__device__ __forceinline__ void surv_append(const SurvList &l, int p, int s, float a, float b)
{
  const int slot = atomicAdd(l.n, 1);            // *n may run past cap: the host checks and re-runs with a smaller batch
  if (slot < l.cap) { l.p[slot] = p; l.s[slot] = s; if (l.a) l.a[slot] = a; if (l.b) l.b[slot] = b; atomicAdd(l.cnt + p, 1); if (l.cnts) atomicAdd(l.cnts + s, 1); }
}

// First-level filter decision of p7_Pipeline (p7_pipeline.c:721-725): P-value of the MSV score against F1.
__device__ __forceinline__ bool msv_passes(float usc, float nullsc, const ProfDev &P, double F1)
{
  const float seq_score = (float)((double)(usc - nullsc) / 0.69314718055994529);
  const double pv = gumbel_surv((double)seq_score, (double)P.evparam[0], (double)P.evparam[1]);
  return !(pv > F1);
}

// p7_SSVFilter's post-processing (ssvfilter.c:881-923) applied to the wide-lane maximum `maxw`
// (cells are kept relative to the constant begin score, so the reference's get_xE() == maxw + 128).
__device__ __forceinline__ void ssv_finish(int maxw, const ProfDev &p, int tjb, float &sc, int &status)
{
  if (tjb + p.tbm + p.tec + p.bias >= 127) { sc = 0.f; status = B2H_ENORESULT; return; }
  if (maxw >= 127 - p.bias) {                      // xE >= 255 - bias_b
    sc = INFINITY;
    status = (p.base - tjb - p.tbm < 128) ? B2H_ENORESULT : B2H_ERANGE;
    return;
  }
  int xE = maxw + p.base - tjb - p.tbm;            // xE += base - tjb - tbm; xE -= 128
  if (xE >= 255 - p.bias) { sc = INFINITY; status = B2H_ERANGE; return; }
  int xJ = xE - p.tec;
  if (xJ > p.base) { sc = 0.f; status = B2H_ENORESULT; return; }
  sc = ((float)(xJ - tjb) - (float)p.base);
  sc /= p.scale_b;
  sc -= 3.0f;
  status = B2H_OK;
}

// ------------------------------------------------------------------------------------------------
// SSV pass over (profile x sequence) for every profile of one register-tile class (G, NR).
// Work item = (profile, chunk of B2H_SSV_CHUNK sequences in length-sorted order), profile-major, so
// a persistent CTA re-stages the emission table only when it crosses a profile boundary.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hfma2_relu_add(uint32_t m, uint32_t e)
{
  uint32_t r;
  asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(m), "r"(0x3c003c00u), "r"(e));
  return r;
}

template <int G, int NR>
__global__ void __launch_bounds__(SSV_THREADS) ssv_kernel(const SsvArgs a)
{
  extern __shared__ __align__(128) uint32_t s_tab[];      // [32 residues][ROWB bytes]
  __shared__ uint64_t s_bar;
  __shared__ int s_item;
  constexpr int FULLQ = NR / 4, REM = NR % 4, NG = 32 / G;
  constexpr int GS = (G < 8) ? 8 : G;                       // 16-byte slots per chunk: groups narrower than a quarter-warp read replicas
  constexpr uint32_t ROWB = (uint32_t)FULLQ * GS * 16 + (uint32_t)REM * 128;
  constexpr uint32_t TAB_BYTES = (uint32_t)B2H_NCODE * ROWB;
  constexpr uint32_t PADW = 0x01010101u * B2H_PAD_CODE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int gl = lane & (G - 1), grp = lane / G;

  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  uint32_t phase = 0;
  int cur_pc = -1;

  const int src_lane = (lane & ~(G - 1)) | ((gl + G - 1) & (G - 1));   // rotate inside the group: its lane 0 reads the last cell of its last lane, which is always padding (= 0)
  const uint32_t tab_lane = smem_u32(s_tab) + (lane & (GS - 1)) * 16;
  const uint32_t tab_rem  = smem_u32(s_tab) + FULLQ * GS * 16 + lane * 4;
  const int nitems = a.ncls * a.chunks;
  int budget = a.items_per_cta > 0 ? a.items_per_cta : 0x7fffffff;

  for (;;) {
    __syncthreads();                                       // every warp is done with the previous item (and table)
    if (budget-- <= 0) break;                              // (uniform) retire: the grid holds enough CTAs for every item
    if (threadIdx.x == 0) s_item = atomicAdd(a.counter, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= nitems) break;
    const int pc = item / a.chunks, chunk = item - pc * a.chunks;
    const int pidx = a.cls[pc];
    const ProfDev &P = a.profs[pidx];
    if (pc != cur_pc) {
      cur_pc = pc;
      if (threadIdx.x == 0) { mbar_expect_tx(&s_bar, TAB_BYTES); tma_load_1d(s_tab, a.wide ? P.ssv_emis_w : P.ssv_emis, TAB_BYTES, &s_bar); }
      mbar_wait(&s_bar, phase); phase ^= 1;
    }
    const int e_end = min(a.sd.n, (chunk + 1) * B2H_SSV_CHUNK);
    for (int e0 = chunk * B2H_SSV_CHUNK + warp * NG; e0 < e_end; e0 += nwarps * NG) {
      const int e = e0 + grp;                              // this group's comparison; groups past the end idle on padding rows
      const bool valid = e < e_end;
      const int s = valid ? a.sd.order[e] : 0;
      const int L = valid ? a.sd.len[s] : 0;
      const uint32_t *seqw = reinterpret_cast<const uint32_t *>(a.sd.res + a.sd.off[s]);
      const int nwords = (L + 3) >> 2;                     // tail rows are B2H_PAD_CODE rows: every score -127, harmless
      const int nwmax = (NG == 1) ? nwords : __reduce_max_sync(FULL, nwords);

      uint32_t m[NR];
#pragma unroll
      for (int j = 0; j < NR; j++) m[j] = 0u;
      uint32_t xe = 0u;

      uint32_t nextw = (gl < nwords) ? __ldg(seqw + gl) : PADW;
      for (int w0 = 0; w0 < nwmax; w0 += G) {
        const uint32_t myw = nextw;
        nextw = (w0 + G + gl < nwords) ? __ldg(seqw + w0 + G + gl) : PADW;   // the next G words are in flight during these 4*G rows
        const int nw = min(G, nwmax - w0);
        for (int wi = 0; wi < nw; wi++) {
          const uint32_t wr = __shfl_sync(FULL, myw, wi, G);
#pragma unroll
          for (int rr = 0; rr < 4; rr++) {
            const uint32_t x = __byte_perm(wr, 0u, 0x4440u + rr);
            uint32_t ev[NR];
            load_row<G, NR>(tab_lane + x * ROWB, tab_rem + x * ROWB, ev);
            const uint32_t t  = __shfl_sync(FULL, m[NR-1], src_lane);
            const uint32_t s0 = __byte_perm(t, m[NR-1], 0x5432u);     // lo <- previous lane's last cell, hi <- own cell NR-1
#pragma unroll
            for (int j = NR - 1; j >= 1; j--) m[j] = hfma2_relu_add(m[j-1], ev[j]);
            m[0] = hfma2_relu_add(s0, ev[0]);
#pragma unroll
            for (int j = 0; j + 1 < NR; j += 2) xe = __vimax3_s16x2(xe, m[j], m[j+1]);
            if (NR & 1) xe = __vimax3_s16x2(xe, m[NR-1], m[NR-1]);
          }
        }
      }
      int v = max((int)(xe & 0xffffu), (int)(xe >> 16));
      if (G == 32) v = __reduce_max_sync(FULL, v);
      else {
#pragma unroll
        for (int o = G / 2; o >= 1; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
      }
      {                                                      // fp16 bit pattern (>= +0) -> integer value; inf -> "overflowed"
        const float f = __half2float(__ushort_as_half((unsigned short)v));
        v = (f > 30000.0f) ? 30000 : (int)f;
      }
      if (a.mode == 3) {                                      // scan orientation: fold this chunk's maximum into its sequence's
        if (gl == 0 && valid) atomicMax(a.raw + (size_t)pidx * a.raw_stride + a.parent[s], v);
      } else
      if (gl == 0 && valid && (a.mode != 2 || L > 0)) {         // p7_Pipeline returns at once for an empty target (p7_pipeline.c:713): it enters no list
        float sc; int status;
        ssv_finish(v, P, (int)a.sd.tjb[s], sc, status);
        if (a.mode == 0) { a.out_sc[s] = sc; a.out_status[s] = status; }
        else if (status == B2H_ENORESULT) surv_append(a.R, pidx, s, 0.f, 0.f);
        else if (a.mode == 1) { a.out_sc[s] = sc; a.out_status[s] = status; }
        else if (msv_passes(sc, a.sd.null1[s], P, a.F1)) surv_append(a.A, pidx, s, sc, 0.f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Full MSV (with the J state, msvfilter.c:106-207) for the comparisons SSV could not decide (~1 %).
// Generic in M: one warp per comparison, lanes interleaved over nodes, the row lives in shared
// memory as bytes; the uint8 saturating arithmetic is written out explicitly.
// ------------------------------------------------------------------------------------------------
struct Item { int p, e_begin, e_end; };
__device__ __forceinline__ bool next_item(const WorkList &wl, int *s_item, Item &it)
{
  __syncthreads();
  if (threadIdx.x == 0) *s_item = atomicAdd(wl.counter, 1);
  __syncthreads();
  const int item = *s_item + wl.itemoff[wl.plo];
  if (item >= wl.itemoff[wl.phi]) return false;
  int lo = wl.plo, hi = wl.phi;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (wl.itemoff[mid] <= item) lo = mid; else hi = mid; }
  it.p = lo;
  it.e_begin = wl.poff[lo] + (item - wl.itemoff[lo]) * B2H_ITEM_ENTRIES;
  it.e_end   = min(wl.poff[lo + 1], it.e_begin + B2H_ITEM_ENTRIES);
  return true;
}

__global__ void __launch_bounds__(256) msv_kernel(const WorkList wl, const SeqDev sd, int max_Mpad, int mode,
                                                   float *out_sc, int32_t *out_status, const SurvList A, double F1)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t s_bar;
  __shared__ int s_item;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int S = max_Mpad + 64;
  uint8_t *s_cost = smem;                                   // [32][Mpad]
  uint8_t *rows = smem + (size_t)32 * max_Mpad + (size_t)warp * 2 * S;

  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  uint32_t phase = 0;
  int cur_p = -1;
  Item it;
  while (next_item(wl, &s_item, it)) {
    const ProfDev &P = wl.profs[it.p];
    const int Mpad = P.Mpad;
    if (it.p != cur_p) {
      cur_p = it.p;
      if (threadIdx.x == 0) { mbar_expect_tx(&s_bar, 32u * Mpad); tma_load_1d(s_cost, P.msv_cost8, 32u * Mpad, &s_bar); }
      mbar_wait(&s_bar, phase); phase ^= 1;
    }
    const int bias = P.bias, base = P.base, tec = P.tec;
    for (int e = it.e_begin + warp; e < it.e_end; e += nwarps) {
      const int s = wl.ent_s[e];
      const int L = sd.len[s];
      const uint8_t *seq = sd.res + sd.off[s];
      const int tjb  = sd.tjb[s];
      const int tjbm = (tjb + P.tbm) & 0xff;               // (int8)tjb + (int8)tbm splatted into bytes (msvfilter.c:116)
      for (int j = lane; j < 2 * S; j += 32) rows[j] = 0;
      __syncwarp();
      int xJ = 0, xB = max(base - tjbm, 0);
      bool overflow = false;
      for (int i = 1; i <= L; i++) {
        const uint8_t *cost = s_cost + (size_t)seq[i - 1] * Mpad;
        uint8_t *cur = rows + (i & 1) * S, *prv = rows + ((i & 1) ^ 1) * S;
        int xEm = 0;
        for (int k = lane + 1; k <= Mpad; k += 32) {
          int sv = max((int)prv[k - 1], xB);               // max_epu8(mpv, xBv)
          sv = min(sv + bias, 255);                        // adds_epu8(sv, biasv)
          sv = max(sv - (int)cost[k - 1], 0);              // subs_epu8(sv, *rsc)
          cur[k] = (uint8_t)sv;
          xEm = max(xEm, sv);
        }
        const int xE0 = __reduce_max_sync(FULL, xEm);
        if (xE0 + bias >= 255) { overflow = true; break; } // adds_epu8(xEv, biasv) == 255
        const int xE = max(xE0 - tec, 0);
        xJ = max(xJ, xE);
        xB = max(max(base, xJ) - tjbm, 0);
        __syncwarp();
      }
      if (lane == 0) {
        float sc; int status = B2H_OK;
        if (overflow) { sc = INFINITY; status = B2H_ERANGE; }
        else { sc = ((float)(xJ - tjb) - (float)base); sc /= P.scale_b; sc -= 3.0f; }
        if (mode == 1) { out_sc[s] = sc; out_status[s] = status; }
        else if (msv_passes(sc, sd.null1[s], P, F1)) surv_append(A, it.p, s, sc, 0.f);
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Full MSV in the SSV kernel's register tiles: the same lane-striped fp16x2 table, G lanes per comparison, but with the
// B/E/J states of msvfilter.c:106-207 -- one group-wide maximum and a handful of scalar updates per row.  The byte
// saturations of the reference never act before its own overflow test fires (every cell and xB stay below 255 - bias
// until then), so the unsaturated recurrence  m' = relu(max(m_diag, xB) + bias - cost)  is exact up to that point and
// the overflow test is the reference's.
// ------------------------------------------------------------------------------------------------
template <int G, int NR>
__global__ void __launch_bounds__(128) rmsv_kernel(const WorkList wl, const SeqDev sd, int mode,
                                                    float *out_sc, int32_t *out_status, const SurvList A, double F1)
{
  extern __shared__ __align__(128) uint32_t s_tab[];
  __shared__ uint64_t s_bar;
  __shared__ int s_item;
  constexpr int FULLQ = NR / 4, REM = NR % 4, NG = 32 / G;
  constexpr int GS = (G < 8) ? 8 : G;                       // 16-byte slots per chunk: groups narrower than a quarter-warp read replicas
  constexpr uint32_t ROWB = (uint32_t)FULLQ * GS * 16 + (uint32_t)REM * 128;
  constexpr uint32_t TAB_BYTES = (uint32_t)B2H_NCODE * ROWB;
  constexpr uint32_t PADW = 0x01010101u * B2H_PAD_CODE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int gl = lane & (G - 1), grp = lane / G;
  const int src_lane = (lane & ~(G - 1)) | ((gl + G - 1) & (G - 1));
  const uint32_t tab_lane = smem_u32(s_tab) + (lane & (GS - 1)) * 16;
  const uint32_t tab_rem  = smem_u32(s_tab) + FULLQ * GS * 16 + lane * 4;

  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  uint32_t phase = 0;
  int cur_p = -1;
  Item it;
  while (next_item(wl, &s_item, it)) {
    const ProfDev &P = wl.profs[it.p];
    if (it.p != cur_p) {
      cur_p = it.p;
      if (threadIdx.x == 0) { mbar_expect_tx(&s_bar, TAB_BYTES); tma_load_1d(s_tab, (mode & 8) ? P.ssv_emis_w : P.ssv_emis, TAB_BYTES, &s_bar); }
      mbar_wait(&s_bar, phase); phase ^= 1;
    }
    const float bias = (float)P.bias, base = (float)P.base, tec = (float)P.tec;
    for (int e0 = it.e_begin + warp * NG; e0 < it.e_end; e0 += nwarps * NG) {
      const int e = e0 + grp;
      const bool valid = e < it.e_end;
      const int s = valid ? wl.ent_s[e] : 0;
      const int L = valid ? sd.len[s] : 0;
      const uint32_t *seqw = reinterpret_cast<const uint32_t *>(sd.res + sd.off[s]);
      const int nwords = (L + 3) >> 2;                     // tail rows are B2H_PAD_CODE rows: every cell 0, xE = 0, specials unchanged
      const int nwmax = (NG == 1) ? nwords : __reduce_max_sync(FULL, nwords);
      const int tjb = valid ? (int)sd.tjb[s] : 0;
      const float tjbm = (float)((tjb + P.tbm) & 0xff);    // (int8)tjb + (int8)tbm splatted into bytes (msvfilter.c:116)

      uint32_t m[NR];
#pragma unroll
      for (int j = 0; j < NR; j++) m[j] = 0u;
      float xJ = 0.f, xB = fmaxf(base - tjbm, 0.f);
      bool overflow = false;

      uint32_t nextw = (gl < nwords) ? __ldg(seqw + gl) : PADW;
      for (int w0 = 0; w0 < nwmax; w0 += G) {
        const uint32_t myw = nextw;
        nextw = (w0 + G + gl < nwords) ? __ldg(seqw + w0 + G + gl) : PADW;
        const int nw = min(G, nwmax - w0);
        for (int wi = 0; wi < nw; wi++) {
          const uint32_t wr = __shfl_sync(FULL, myw, wi, G);
#pragma unroll
          for (int rr = 0; rr < 4; rr++) {
            const uint32_t x = __byte_perm(wr, 0u, 0x4440u + rr);
            uint32_t ev[NR];
            load_row<G, NR>(tab_lane + x * ROWB, tab_rem + x * ROWB, ev);
            const __half2 xb2 = __float2half2_rn(xB);
            const uint32_t xBh = *reinterpret_cast<const uint32_t *>(&xb2);
            const uint32_t t  = __shfl_sync(FULL, m[NR-1], src_lane);
            const uint32_t s0 = __byte_perm(t, m[NR-1], 0x5432u);
            // max(m_diag, xB) on the bit patterns (both non-negative fp16), then + (bias - cost) and the floor at 0
#pragma unroll
            for (int j = NR - 1; j >= 1; j--) m[j] = hfma2_relu_add(__vimax3_s16x2(m[j-1], xBh, xBh), ev[j]);
            m[0] = hfma2_relu_add(__vimax3_s16x2(s0, xBh, xBh), ev[0]);
            uint32_t xe = 0u;
#pragma unroll
            for (int j = 0; j + 1 < NR; j += 2) xe = __vimax3_s16x2(xe, m[j], m[j+1]);
            if (NR & 1) xe = __vimax3_s16x2(xe, m[NR-1], m[NR-1]);
            int v = max((int)(xe & 0xffffu), (int)(xe >> 16));
            if (G == 32) v = __reduce_max_sync(FULL, v);
            else {
#pragma unroll
              for (int o = G / 2; o >= 1; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
            }
            const float xE0 = __half2float(__ushort_as_half((unsigned short)v));
            if (!overflow) {
              if (xE0 + bias >= 255.f) overflow = true;      // adds_epu8(xEv, biasv) saturates (msvfilter.c:167-180)
              else {
                const float xE = fmaxf(xE0 - tec, 0.f);
                xJ = fmaxf(xJ, xE);
                xB = fmaxf(fmaxf(base, xJ) - tjbm, 0.f);
              }
            }
          }
        }
      }
      if (gl == 0 && valid) {
        float sc; int status = B2H_OK;
        if (overflow) { sc = INFINITY; status = B2H_ERANGE; }
        else { sc = ((xJ - (float)tjb) - base); sc /= P.scale_b; sc -= 3.0f; }
        if ((mode & 7) == 1) { out_sc[s] = sc; out_status[s] = status; }
        else if (msv_passes(sc, sd.null1[s], P, F1)) surv_append(A, it.p, s, sc, 0.f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Group a survivor list by profile (counting sort on the per-profile counts the epilogues kept).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) group_scan_kernel(const int *cnt, int P, int32_t *poff, int32_t *itemoff, int *fill)
{
  __shared__ int s_part[1024];
  __shared__ int s_ipart[1024];
  const int t = threadIdx.x;
  const int per = (P + 1023) / 1024;
  const int b = t * per, e = min(P, b + per);
  int sum = 0, isum = 0;
  for (int i = b; i < e; i++) { sum += cnt[i]; isum += (cnt[i] + B2H_ITEM_ENTRIES - 1) / B2H_ITEM_ENTRIES; }
  s_part[t] = sum; s_ipart[t] = isum;
  __syncthreads();
  if (t == 0) {
    int a = 0, ia = 0;
    for (int i = 0; i < 1024; i++) { const int v = s_part[i], iv = s_ipart[i]; s_part[i] = a; s_ipart[i] = ia; a += v; ia += iv; }
    poff[P] = a; itemoff[P] = ia;
  }
  __syncthreads();
  int a = s_part[t], ia = s_ipart[t];
  for (int i = b; i < e; i++) {
    poff[i] = a; itemoff[i] = ia; fill[i] = 0;
    a += cnt[i]; ia += (cnt[i] + B2H_ITEM_ENTRIES - 1) / B2H_ITEM_ENTRIES;
  }
}

__global__ void group_scatter_kernel(const SurvList in, const int32_t *poff, int *fill, Grouped out)
{
  const int n = min(*in.n, in.cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int p = in.p[i];
    const int pos = poff[p] + atomicAdd(fill + p, 1);
    out.p[pos] = p; out.s[pos] = in.s[i];
    if (out.a) out.a[pos] = in.a[i];
    if (out.b) out.b[pos] = in.b[i];
  }
}

// p7_SSVFilter's post-processing + the F1 test of p7_Pipeline for the raw maxima of the scan orientation
__global__ void ssv_finish_kernel(const ProfDev *profs, int P, const SeqDev sd, const int *raw, const SurvList A, const SurvList R, double F1)
{
  const long long total = (long long)P * sd.n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx / sd.n), s = (int)(idx - (long long)p * sd.n);
    if (sd.len[s] <= 0) continue;                          // p7_Pipeline skips empty targets
    const ProfDev &Pf = profs[p];
    float sc; int status;
    ssv_finish(raw[idx], Pf, (int)sd.tjb[s], sc, status);
    if (status == B2H_ENORESULT) surv_append(R, p, s, 0.f, 0.f);
    else if (msv_passes(sc, sd.null1[s], Pf, F1)) surv_append(A, p, s, sc, 0.f);
  }
}

template <int G, int NR>
int launch_ssv_tile(b2h_ctx *ctx, const SsvArgs &a, cudaStream_t strm)
{
  const size_t smem = (size_t)B2H_NCODE * b2h_ssv_row_bytes(G, NR);
  constexpr int NG = 32 / G;                                // comparisons per warp: an item of B2H_SSV_CHUNK sequences keeps at most CHUNK / NG warps busy
  const int fit = 32 * std::max(1, std::min(SSV_THREADS / 32, (B2H_SSV_CHUNK + NG - 1) / NG));
  const int threads = (a.threads > 0 && a.threads <= SSV_THREADS) ? a.threads : fit;
  int occ = 1;
  { const int st = b2h_kernel_occupancy(ctx, (const void *)ssv_kernel<G, NR>, threads, smem, &occ); if (st != B2H_OK) return st; }
  static const int occ_cap = getenv("B2H_SSV_OCC") ? atoi(getenv("B2H_SSV_OCC")) : 0;       // experiments: resident CTAs per SM
  if (occ_cap > 0 && occ > occ_cap) occ = occ_cap;
  int grid = ctx->sm_count * occ;
  const long long nitems = (long long)a.ncls * a.chunks;
  if (grid > nitems) grid = (int)(nitems > 0 ? nitems : 1);
  if (a.items_per_cta > 0) grid = (int)std::max<long long>(grid, (nitems + a.items_per_cta - 1) / a.items_per_cta);
  ssv_kernel<G, NR><<<grid, threads, smem, strm>>>(a);
  ctx->launches++;
  B2H_CUDA(cudaGetLastError());
  return B2H_OK;
}

} // namespace

int b2h_launch_ssv(b2h_ctx *ctx, int G, int NR, const SsvArgs &a, cudaStream_t strm)
{
  if (a.ncls <= 0 || a.sd.n <= 0) return B2H_OK;
  B2H_CUDA(cudaMemsetAsync(a.counter, 0, sizeof(int), strm));
  switch (G * 64 + NR) {
#define CASE(g, n) case (g) * 64 + (n): return launch_ssv_tile<g, n>(ctx, a, strm);
#define CASE8(g, n) CASE(g, n) CASE(g, n + 1) CASE(g, n + 2) CASE(g, n + 3) CASE(g, n + 4) CASE(g, n + 5) CASE(g, n + 6) CASE(g, n + 7)
    CASE8(8, 1) CASE8(8, 9) CASE8(8, 17) CASE8(8, 25)
    CASE(4, 4) CASE(4, 5) CASE(4, 6) CASE(4, 7) CASE(4, 8) CASE8(4, 9) CASE8(4, 17) CASE8(4, 25)
    CASE(2, 4) CASE(2, 5) CASE(2, 6) CASE(2, 7) CASE(2, 8) CASE8(2, 9) CASE8(2, 17) CASE8(2, 25)
    CASE(1, 4) CASE(1, 5) CASE(1, 6) CASE(1, 7) CASE(1, 8) CASE8(1, 9) CASE8(1, 17) CASE8(1, 25)
    CASE8(16, 17) CASE8(16, 25)
    CASE(32, 18) CASE(32, 20) CASE(32, 22) CASE(32, 24) CASE(32, 26) CASE(32, 28) CASE(32, 30) CASE(32, 32) CASE(32, 40) CASE(32, 48)
#undef CASE8
#undef CASE
  }
  ctx->err = "no SSV kernel for this register tile";
  return B2H_EINVAL;
}

int b2h_launch_ssv_finish(b2h_ctx *ctx, const ProfDev *profs, int P, const SeqDev &sd, const int *raw, SurvList A, SurvList R, double F1)
{
  const long long total = (long long)P * sd.n;
  if (total <= 0) return B2H_OK;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 16);
  ssv_finish_kernel<<<grid, 256, 0, ctx->stream>>>(profs, P, sd, raw, A, R, F1);
  ctx->launches++;
  B2H_CUDA(cudaGetLastError());
  return B2H_OK;
}

int b2h_launch_msv(b2h_ctx *ctx, const WorkList &wl, const SeqDev &sd, int max_Mpad, int nitems_hint, int mode,
                   float *out_sc, int32_t *out_status, SurvList A, double F1)
{
  const int nwarps = 8;
  const size_t smem = (size_t)32 * max_Mpad + (size_t)nwarps * 2 * (max_Mpad + 64) + 128;
  if (smem > 220 * 1024) { ctx->err = "model too long for the MSV kernel"; return B2H_EINVAL; }
  int occ = 1;
  { const int st = b2h_kernel_occupancy(ctx, (const void *)msv_kernel, nwarps * 32, smem, &occ); if (st != B2H_OK) return st; }
  int grid = ctx->sm_count * occ;
  if (nitems_hint > 0 && grid > nitems_hint) grid = nitems_hint;
  B2H_CUDA(cudaMemsetAsync(wl.counter, 0, sizeof(int), ctx->stream));
  msv_kernel<<<grid, nwarps * 32, smem, ctx->stream>>>(wl, sd, max_Mpad, mode, out_sc, out_status, A, F1);
  ctx->launches++;
  B2H_CUDA(cudaGetLastError());
  return B2H_OK;
}

namespace {
template <int G, int NR>
int launch_rmsv_tile(b2h_ctx *ctx, const WorkList &wl, const SeqDev &sd, int mode, float *out_sc, int32_t *out_status, const SurvList &A, double F1, cudaStream_t strm)
{
  const size_t smem = (size_t)B2H_NCODE * b2h_ssv_row_bytes(G, NR);
  int occ = 1;
  { const int st = b2h_kernel_occupancy(ctx, (const void *)rmsv_kernel<G, NR>, 128, smem, &occ); if (st != B2H_OK) return st; }
  B2H_CUDA(cudaMemsetAsync(wl.counter, 0, sizeof(int), strm));
  rmsv_kernel<G, NR><<<ctx->sm_count * occ, 128, smem, strm>>>(wl, sd, mode, out_sc, out_status, A, F1);
  ctx->launches++;
  B2H_CUDA(cudaGetLastError());
  return B2H_OK;
}
} // namespace

// Full MSV over a grouped work list whose profiles are sorted by model length: one launch per SSV register tile
// (tiles[p] = G*64 + NR of profile p of the list), the launches side by side on the side streams.  mode | 8: the tiles are
// the profiles' WIDE tiles (ProfDev::ssv_emis_w; scan orientation).
int b2h_launch_msv_tiled(b2h_ctx *ctx, const WorkList &wl_in, const SeqDev &sd, const std::vector<int> &tiles, int mode,
                         float *out_sc, int32_t *out_status, SurvList A, double F1)
{
  const int P = (int)tiles.size();
  ForkJoin fj(ctx);
  int cls = 0;
  for (int plo = 0; plo < P; cls++) {
    int phi = plo;
    while (phi < P && tiles[phi] == tiles[plo]) phi++;
    WorkList wl = wl_in; wl.plo = plo; wl.phi = phi; wl.counter = ctx->d_counters + 32 + (cls % 24);
    cudaStream_t strm = fj.next();
    int st = B2H_EINVAL;
    switch (tiles[plo]) {
#define CASE(g, n) case (g) * 64 + (n): st = launch_rmsv_tile<g, n>(ctx, wl, sd, mode, out_sc, out_status, A, F1, strm); break;
#define CASE8(g, n) CASE(g, n) CASE(g, n + 1) CASE(g, n + 2) CASE(g, n + 3) CASE(g, n + 4) CASE(g, n + 5) CASE(g, n + 6) CASE(g, n + 7)
      CASE8(8, 1) CASE8(8, 9) CASE8(8, 17) CASE8(8, 25)
      CASE(4, 4) CASE(4, 5) CASE(4, 6) CASE(4, 7) CASE(4, 8) CASE8(4, 9) CASE8(4, 17) CASE8(4, 25)
      CASE(2, 4) CASE(2, 5) CASE(2, 6) CASE(2, 7) CASE(2, 8) CASE8(2, 9) CASE8(2, 17) CASE8(2, 25)
      CASE(1, 4) CASE(1, 5) CASE(1, 6) CASE(1, 7) CASE(1, 8) CASE8(1, 9) CASE8(1, 17) CASE8(1, 25)
      CASE8(16, 17) CASE8(16, 25)
      CASE(32, 18) CASE(32, 20) CASE(32, 22) CASE(32, 24) CASE(32, 26) CASE(32, 28) CASE(32, 30) CASE(32, 32) CASE(32, 40) CASE(32, 48)
#undef CASE8
#undef CASE
    }
    if (st != B2H_OK) { if (st == B2H_EINVAL) ctx->err = "no MSV kernel for this register tile"; return st; }
    plo = phi;
  }
  return B2H_OK;
}

int b2h_launch_group(b2h_ctx *ctx, const SurvList &in, int P, Grouped out)
{
  group_scan_kernel<<<1, 1024, 0, ctx->stream>>>(in.cnt, P, out.poff, out.itemoff, out.fill);
  ctx->launches++;
  B2H_CUDA(cudaGetLastError());
  group_scatter_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(in, out.poff, out.fill, out);
  ctx->launches++;
  B2H_CUDA(cudaGetLastError());
  return B2H_OK;
}
