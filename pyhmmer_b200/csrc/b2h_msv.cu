// b2h_msv.cu -- SSV / MSV uint8 filters as sm_100a kernels.
//
// Replaces p7_SSVFilter (impl_sse/ssvfilter.c:876-926) and p7_MSVFilter (impl_sse/msvfilter.c:74-208).
// This is a new design, not a translation of the SSE code:
//
//  * one WARP per (profile x sequence) comparison; persistent CTAs pull sequences (longest first)
//    from a global work counter;
//  * the profile's emission table is staged ONCE per CTA into shared memory with a TMA bulk copy
//    (cp.async.bulk + mbarrier), pre-swizzled on the host so that every row step is one
//    conflict-free LDS.128 per lane;
//  * the DP row lives in registers as packed s16x2 cells.  Lane z owns the 2*NR consecutive model
//    nodes z*2NR+1 .. (z+1)*2NR; register j packs nodes (j, j+NR) of that chunk, so the diagonal
//    move M(i,k) <- M(i-1,k-1) is a pure register renaming plus ONE warp shuffle and one PRMT per row;
//  * sm_100a has no native byte SIMD (vmaxu4/vaddus4 are emulated), but it has the DPX 16x2 ops:
//    the SSV cell update is ONE VIADDMNMX.S16x2 for two cells, the running maximum ONE VIMNMX3.S16x2
//    for four cells.  16-bit lanes make the uint8 saturation explicit (clamps) instead of implicit,
//    and the results are bit-identical to the reference (derivation in DESIGN.md, "SSV in wide lanes").
//
// Semantics reproduced exactly: SSV is authoritative when it can prove the J state was not used;
// otherwise (eslENORESULT) the comparison is redone by the full MSV recurrence with saturating
// uint8 arithmetic; overflow returns +inf/eslERANGE.
#include <cuda_runtime.h>
#include <cmath>
#include "b2h_internal.h"

namespace {

constexpr int SSV_THREADS = 256;

struct MsvProf {
  const uint32_t *emis;     // SSV signed scores, lane-striped
  const uint32_t *cost;     // MSV costs, lane-striped
  int   M;
  int   tbm, tec, base, bias;
  float scale_b;
};

struct MsvArgs {
  MsvProf        prof;
  const uint8_t *res;
  const int64_t *off;
  const int32_t *len;
  const uint8_t *tjb;
  const int32_t *order;
  int            nseq;
  int           *counter;     // [0]: work cursor of the SSV pass  [1]: #entries in redo list  [2]: cursor of the MSV pass
  int32_t       *redo;        // sequences whose SSV result was eslENORESULT
  float         *out_sc;
  int32_t       *out_status;
  int            msv_fallback; // 0: report p7_SSVFilter's own status; 1: queue ENORESULT for the MSV pass
  uint32_t       zero;         // always 0: a register ptxas cannot constant-fold (packed-zero operand of VIADDMNMX)
};

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
template <int NR>
__device__ __forceinline__ void load_row(uint32_t addr, uint32_t addr_rem, uint32_t (&e)[NR])
{
  constexpr int FULL = NR / 4, REM = NR % 4;
#pragma unroll
  for (int g = 0; g < FULL; g++) {
    uint4 v = lds128(addr + g * 512);
    e[4*g+0] = v.x; e[4*g+1] = v.y; e[4*g+2] = v.z; e[4*g+3] = v.w;
  }
  if (REM == 1) e[4*FULL] = lds32(addr_rem);
  if (REM == 2) { uint2 v = lds64(addr_rem); e[4*FULL] = v.x; e[4*FULL+1] = v.y; }
  if (REM == 3) { e[4*FULL] = lds32(addr_rem); e[4*FULL+1] = lds32(addr_rem + 4); e[4*FULL+2] = lds32(addr_rem + 8); }
}
// an opaque zero: keeps ptxas from re-materialising the constant operand of every VIADDMNMX
__device__ __forceinline__ uint32_t opaque_zero() { uint32_t z; asm volatile("mov.u32 %0, 0;" : "=r"(z)); return z; }

// p7_SSVFilter's post-processing (ssvfilter.c:881-923) applied to the wide-lane maximum `maxw`
// (cells are kept relative to the constant begin score, so the reference's get_xE() == maxw + 128).
__device__ __forceinline__ void ssv_finish(int maxw, const MsvProf &p, int tjb, float &sc, int &status)
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
// SSV pass: all sequences.
// ------------------------------------------------------------------------------------------------
template <int NR>
__global__ void __launch_bounds__(SSV_THREADS) ssv_kernel(const MsvArgs a)
{
  extern __shared__ __align__(128) uint32_t s_tab[];      // [32 residues][NR*32 words]
  __shared__ uint64_t s_bar;
  constexpr uint32_t TAB_BYTES = (uint32_t)B2H_NCODE * NR * 128u;
  constexpr int ROW_WORDS = NR * 32;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) { mbar_expect_tx(&s_bar, TAB_BYTES); tma_load_1d(s_tab, a.prof.emis, TAB_BYTES, &s_bar); }
  mbar_wait(&s_bar, 0);

  const int src_lane = (lane + 31) & 31;                   // rotate: lane 0 reads lane 31's last cell, which is always padding (= 0)
  const uint32_t tab_lane = smem_u32(s_tab) + lane * 16;
  const uint32_t tab_rem  = smem_u32(s_tab) + (NR / 4) * 512 + lane * (NR % 4) * 4;
  const uint32_t zero = a.zero;

  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(a.counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= a.nseq) break;
    const int s = a.order[item];
    const int L = a.len[s];
    const uint32_t *seqw = reinterpret_cast<const uint32_t *>(a.res + a.off[s]);
    const int nwords = (L + 3) >> 2;                       // tail rows are B2H_PAD_CODE rows: every score -127, harmless

    uint32_t m[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) m[j] = 0u;
    uint32_t xe = 0u;

    for (int w0 = 0; w0 < nwords; w0 += 32) {
      const uint32_t myw = (w0 + lane < nwords) ? __ldg(seqw + w0 + lane) : 0x1f1f1f1fu;
      const int nw = min(32, nwords - w0);
      for (int wi = 0; wi < nw; wi++) {
        const uint32_t wr = __shfl_sync(0xffffffffu, myw, wi);
#pragma unroll
        for (int rr = 0; rr < 4; rr++) {
          const uint32_t x = __byte_perm(wr, 0u, 0x4440u + rr);
          uint32_t e[NR];
          load_row<NR>(tab_lane + x * (ROW_WORDS * 4), tab_rem + x * (ROW_WORDS * 4), e);
          const uint32_t t  = __shfl_sync(0xffffffffu, m[NR-1], src_lane);
          const uint32_t s0 = __byte_perm(t, m[NR-1], 0x5432u);     // lo <- previous lane's last cell, hi <- own cell NR-1
#pragma unroll
          for (int j = NR - 1; j >= 1; j--) m[j] = __viaddmax_s16x2(m[j-1], e[j], zero);
          m[0] = __viaddmax_s16x2(s0, e[0], zero);
#pragma unroll
          for (int j = 0; j + 1 < NR; j += 2) xe = __vimax3_s16x2(xe, m[j], m[j+1]);
          if (NR & 1) xe = __vimax3_s16x2(xe, m[NR-1], m[NR-1]);
        }
      }
    }
    int v = max((int)(xe & 0xffffu), (int)(xe >> 16));
    v = __reduce_max_sync(0xffffffffu, v);
    if (lane == 0) {
      float sc; int status;
      ssv_finish(v, a.prof, (int)a.tjb[s], sc, status);
      if (a.msv_fallback && status == B2H_ENORESULT) {
        const int slot = atomicAdd(a.counter + 1, 1);
        a.redo[slot] = s;
      } else {
        a.out_sc[s] = sc; a.out_status[s] = status;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Full MSV pass (with the J state): only the comparisons SSV could not decide.
// Saturating uint8 arithmetic of msvfilter.c:132-207 made explicit in s16 lanes.
// ------------------------------------------------------------------------------------------------
template <int NR>
__global__ void __launch_bounds__(SSV_THREADS) msv_kernel(const MsvArgs a)
{
  extern __shared__ __align__(128) uint32_t s_tab[];
  __shared__ uint64_t s_bar;
  constexpr uint32_t TAB_BYTES = (uint32_t)B2H_NCODE * NR * 128u;
  constexpr int ROW_WORDS = NR * 32;
  const int lane = threadIdx.x & 31;
  const int nredo = a.counter[1];
  if (nredo == 0) return;

  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) { mbar_expect_tx(&s_bar, TAB_BYTES); tma_load_1d(s_tab, a.prof.cost, TAB_BYTES, &s_bar); }
  mbar_wait(&s_bar, 0);

  const int src_lane = (lane + 31) & 31;
  const uint32_t tab_lane = smem_u32(s_tab) + lane * 16;
  const uint32_t tab_rem  = smem_u32(s_tab) + (NR / 4) * 512 + lane * (NR % 4) * 4;
  const int bias = a.prof.bias, base = a.prof.base, tec = a.prof.tec;
  const uint32_t biasv = (uint32_t)bias * 0x00010001u;

  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(a.counter + 2, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= nredo) break;
    const int s = a.redo[item];
    const int L = a.len[s];
    const uint8_t *seq = a.res + a.off[s];
    const int tjb  = a.tjb[s];
    const int tjbm = (tjb + a.prof.tbm) & 0xff;              // (int8)tjb + (int8)tbm splatted into bytes (msvfilter.c:116)

    uint32_t m[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) m[j] = 0u;
    int xJ = 0;
    int xB = max(base - tjbm, 0);
    bool overflow = false;

    for (int i = 0; i < L; i++) {
      const uint32_t x = seq[i];
      uint32_t c[NR];
      load_row<NR>(tab_lane + x * (ROW_WORDS * 4), tab_rem + x * (ROW_WORDS * 4), c);
      const uint32_t xBv = (uint32_t)xB * 0x00010001u;
      const uint32_t t  = __shfl_sync(0xffffffffu, m[NR-1], src_lane);
      const uint32_t s0 = __byte_perm(t, m[NR-1], 0x5432u);
      uint32_t xev = 0u;
#pragma unroll
      for (int j = NR - 1; j >= 0; j--) {
        uint32_t sv = (j == 0) ? s0 : m[j-1];
        sv = __vmaxs2(sv, xBv);                               // max_epu8(mpv, xBv)
        sv = __viaddmin_s16x2(sv, biasv, 0x00ff00ffu);        // adds_epu8(sv, biasv)
        sv = __vsub2(sv, c[j]);                               // subs_epu8(sv, cost) ...
        sv = __vimax_s16x2_relu(sv, sv);                      // ... saturating at 0
        m[j] = sv;
        xev = __vmaxs2(xev, sv);
      }
      int xE = max((int)(xev & 0xffffu), (int)(xev >> 16));
      xE = __reduce_max_sync(0xffffffffu, xE);
      if (xE + bias >= 255) { overflow = true; break; }       // adds_epu8(xEv, biasv) == 255
      xE = max(xE - tec, 0);
      xJ = max(xJ, xE);
      xB = max(max(base, xJ) - tjbm, 0);
    }
    if (lane == 0) {
      if (overflow) { a.out_sc[s] = INFINITY; a.out_status[s] = B2H_ERANGE; }
      else {
        float sc = ((float)(xJ - tjb) - (float)base);
        sc /= a.prof.scale_b;
        sc -= 3.0f;
        a.out_sc[s] = sc; a.out_status[s] = B2H_OK;
      }
    }
  }
}

template <int NR>
int launch_nr(b2h_ctx *ctx, const MsvArgs &a, int with_msv)
{
  const size_t smem = (size_t)B2H_NCODE * NR * 128;
  static bool attr_done = false;
  if (!attr_done) {
    B2H_CUDA(cudaFuncSetAttribute(ssv_kernel<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2H_CUDA(cudaFuncSetAttribute(msv_kernel<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  int occ = 1;
  B2H_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ssv_kernel<NR>, SSV_THREADS, smem));
  if (occ < 1) occ = 1;
  const int warps_per_cta = SSV_THREADS / 32;
  int grid = ctx->sm_count * occ;
  int need = (a.nseq + warps_per_cta - 1) / warps_per_cta;
  if (grid > need) grid = need > 0 ? need : 1;
  ssv_kernel<NR><<<grid, SSV_THREADS, smem, ctx->stream>>>(a);
  ctx->launches++;
  B2H_CUDA(cudaGetLastError());
  if (with_msv) {
    msv_kernel<NR><<<grid, SSV_THREADS, smem, ctx->stream>>>(a);
    ctx->launches++;
    B2H_CUDA(cudaGetLastError());
  }
  return B2H_OK;
}

} // namespace

int b2h_launch_ssv_dense(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *db, int with_msv_fallback,
                         float *d_sc, int32_t *d_status)
{
  MsvArgs a;
  a.prof.emis = p->d_ssv_emis; a.prof.cost = p->d_msv_cost; a.prof.M = p->M;
  a.prof.tbm = p->tbm_b; a.prof.tec = p->tec_b; a.prof.base = p->base_b; a.prof.bias = p->bias_b; a.prof.scale_b = p->scale_b;
  a.res = db->d_res; a.off = db->d_off; a.len = db->d_len; a.tjb = db->d_tjb; a.order = db->d_order;
  a.nseq = (int)db->n; a.counter = ctx->d_counters;
  a.out_sc = d_sc; a.out_status = d_status; a.msv_fallback = with_msv_fallback;
  a.redo = nullptr; a.zero = 0u;
  if (db->n == 0) return B2H_OK;
  int32_t *d_redo = nullptr;
  B2H_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 4 * sizeof(int), ctx->stream));
  if (with_msv_fallback) {
    B2H_CUDA(cudaMallocAsync(&d_redo, db->n * sizeof(int32_t), ctx->stream));
    a.redo = d_redo;
  }
  int st;
  switch (p->NR) {
#define CASE(n) case n: st = launch_nr<n>(ctx, a, with_msv_fallback); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(8) CASE(10) CASE(12) CASE(16) CASE(20) CASE(24) CASE(32) CASE(40) CASE(48)
#undef CASE
    default: st = B2H_EINVAL;
  }
  if (d_redo) cudaFreeAsync(d_redo, ctx->stream);
  return st;
}
