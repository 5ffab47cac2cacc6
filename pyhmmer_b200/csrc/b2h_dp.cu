// b2h_dp.cu -- ViterbiFilter (int16), Forward/Backward parsers (fp32) and the bias-composition
// filter as sm_100a kernels.
//
// Replaces p7_ViterbiFilter (impl_sse/vitfilter.c:83-248), forward_engine / backward_engine in
// parser mode (impl_sse/fwdback.c:256-463, 468-733) and p7_bg_FilterScore -> esl_hmm_Forward
// (p7_bg.c:471, vendor/easel/esl_hmm.c:353).
//
// Execution model (new; the SSE code's striping is not used anywhere):
//   * work = a list of (profile, sequence) comparisons that survived the previous stage, grouped by
//     profile; persistent CTAs pull chunks of 16 comparisons of one profile from a global cursor;
//   * the profile's transition (and, if they fit, emission) rows are staged into shared memory by
//     TMA bulk copies once per CTA per profile;
//   * one WARP per comparison walks the sequence row by row; lanes are interleaved over model nodes
//     (node k of a 32-column tile on lane (k-1)%32), the two most recent DP rows live in shared
//     memory, so the (i-1,k-1) dependencies are plain conflict-free shared loads;
//   * the serial D->D chain of a row is a warp-level inclusive scan: max-plus maps
//     d -> max(A, d+T) for Viterbi (exact in integers, so the result is bit-identical to the
//     reference's lazy-F evaluation), affine maps d -> A + d*T for Forward/Backward.
#include <cuda_runtime.h>
#include <cstdlib>
#include <cmath>
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

// Block-wide: fetch the next work item.  Starts with a barrier, so every warp has finished with the
// previous item (and with the shared tables) before anything is overwritten.
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

struct DpCfg { int max_Mpad; int rsc_in_smem; };

// =================================================================================================
// ViterbiFilter
// =================================================================================================
__global__ void __launch_bounds__(256) vit_kernel(const WorkList wl, const SeqDev sd, const DpCfg cfg, const StageOut out)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t s_bar;
  __shared__ int s_item;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int Mp_max = cfg.max_Mpad;
  const int S = Mp_max + 34;                                        // row stride (elements)
  int16_t *s_tsc = reinterpret_cast<int16_t *>(smem);                                    // [8][Mpad]
  int16_t *s_rsc = s_tsc + 8 * Mp_max;                                                   // [32][Mpad] (optional)
  int16_t *s_rows = s_rsc + (cfg.rsc_in_smem ? 32 * Mp_max : 0);
  int16_t *rows = s_rows + (size_t)warp * 6 * S;

  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  uint32_t phase = 0;
  int cur_p = -1;
  Item it;
  while (next_item(wl, &s_item, it)) {
    const ProfDev &P = wl.profs[it.p];
    const int Mpad = P.Mpad;
    if (it.p != cur_p) {
      cur_p = it.p;
      if (threadIdx.x == 0) {
        const uint32_t tb = 8u * Mpad * 2u, rb = cfg.rsc_in_smem ? 32u * Mpad * 2u : 0u;
        mbar_expect_tx(&s_bar, tb + rb);
        tma_load_1d(s_tsc, P.vit_tsc, tb, &s_bar);
        if (rb) tma_load_1d(s_rsc, P.vit_rsc, rb, &s_bar);
      }
      mbar_wait(&s_bar, phase); phase ^= 1;
    }
    const int16_t *tBM = s_tsc, *tMM = s_tsc + Mpad, *tIM = s_tsc + 2 * Mpad, *tDM = s_tsc + 3 * Mpad,
                  *tMD = s_tsc + 4 * Mpad, *tMI = s_tsc + 5 * Mpad, *tII = s_tsc + 6 * Mpad, *tDD = s_tsc + 7 * Mpad;
    const int16_t *rsc_base = cfg.rsc_in_smem ? s_rsc : P.vit_rsc;
    const int xwEm = P.xw_E_move, xwEl = P.xw_E_loop, base_w = P.base_w, ddbound = P.ddbound_w;

    for (int e = it.e_begin + warp; e < it.e_end; e += nwarps) {
      const int s = wl.ent_s[e];
      const int L = sd.len[s];
      const uint8_t *seq = sd.res + sd.off[s];
      const int xw_move = sd.xwmove[s];                 // xw[N|C|J][MOVE] for this target length; LOOPs are 0

      for (int j = lane; j < 6 * S; j += 32) rows[j] = (int16_t)NEG16;
      __syncwarp();
      int xN = base_w, xB = (int16_t)(xN + xw_move), xJ = NEG16, xC = NEG16;
      bool overflow = false;

      for (int i = 1; i <= L; i++) {
        const int x = seq[i - 1];
        int16_t *cur = rows + (i & 1) * 3 * S, *prv = rows + ((i & 1) ^ 1) * 3 * S;
        const int16_t *pM = prv, *pI = prv + S, *pD = prv + 2 * S;
        int16_t *cM = cur, *cI = cur + S, *cD = cur + 2 * S;
        const int16_t *rs = rsc_base + (size_t)x * Mpad;
        int xEm = NEG16, Dm = NEG16;
        for (int k = lane + 1; k <= Mpad; k += 32) {
          const int c = k - 1;
          const int mp = pM[k - 1], ip = pI[k - 1], dp = pD[k - 1], mo = pM[k], io = pI[k];
          int m = max(xB + (int)tBM[c], NEG16);
          m = max(m, mp + (int)tMM[c]);
          m = max(m, ip + (int)tIM[c]);
          m = max(m, dp + (int)tDM[c]);
          m = max(m + (int)rs[c], NEG16);
          xEm = max(xEm, m);
          cM[k] = (int16_t)min(m, 32767);
          const int d = max(m + (int)tMD[c], NEG16);
          Dm = max(Dm, d);
          cD[k + 1] = (int16_t)min(d, 32767);
          cI[k] = (int16_t)max(max(mo + (int)tMI[c], io + (int)tII[c]), NEG16);
        }
        const int xE = __reduce_max_sync(FULL, xEm);
        if (xE >= 32767) { overflow = true; break; }
        // specials: C int expressions stored to int16_t (vitfilter.c:176-180)
        xC = (int16_t)max(xC, xE + xwEm);                 // xw[C][LOOP] = 0
        xJ = (int16_t)max(xJ, xE + xwEl);                 // xw[J][LOOP] = 0
        xB = (int16_t)max(xJ + xw_move, xN + xw_move);    // xw[N][LOOP] = 0 keeps xN = base_w
        const int Dmax = __reduce_max_sync(FULL, Dm);
        __syncwarp();
        if (Dmax + ddbound > xB) {
          // full D->D closure of this row: inclusive max-plus scan over nodes 2..Mpad
          int carry = NEG16;
          for (int k0 = 1; k0 <= Mpad; k0 += 32) {
            const int k = k0 + lane;
            int A = cD[k];
            int T = (k >= 2 && k - 2 < Mpad) ? (int)tDD[k - 2] : NEG16;
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
              const int A2 = __shfl_up_sync(FULL, A, dlt), T2 = __shfl_up_sync(FULL, T, dlt);
              if (lane >= dlt) { A = max(A, A2 + T); T = T2 + T; }
            }
            const int D = max(A, carry + T);
            cD[k] = (int16_t)D;
            carry = __shfl_sync(FULL, D, 31);
          }
          __syncwarp();
        }
      }
      if (lane == 0) {
        float sc; int st = B2H_OK;
        if (overflow) { sc = INFINITY; st = B2H_ERANGE; }
        else if (xC > NEG16) {
          sc = (float)xC + (float)xw_move - (float)base_w;
          sc /= P.scale_w;
          sc -= 3.0f;
        } else sc = -INFINITY;
        out.sc[e] = sc;
        if (out.status) out.status[e] = st;
      }
    }
  }
}

// =================================================================================================
// Forward parser
// =================================================================================================
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

__global__ void __launch_bounds__(256) fwd_kernel(const WorkList wl, const SeqDev sd, const DpCfg cfg, const StageOut out)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t s_bar;
  __shared__ int s_item;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int Mp_max = cfg.max_Mpad;
  const int S = Mp_max + 34;
  float *s_tsc = reinterpret_cast<float *>(smem);
  float *s_rsc = s_tsc + 8 * Mp_max;
  float *s_rows = s_rsc + (cfg.rsc_in_smem ? 32 * Mp_max : 0);
  float *rows = s_rows + (size_t)warp * 6 * S;

  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  uint32_t phase = 0;
  int cur_p = -1;
  Item it;
  while (next_item(wl, &s_item, it)) {
    const ProfDev &P = wl.profs[it.p];
    const int Mpad = P.Mpad;
    if (it.p != cur_p) {
      cur_p = it.p;
      if (threadIdx.x == 0) {
        const uint32_t tb = 8u * Mpad * 4u, rb = cfg.rsc_in_smem ? 32u * Mpad * 4u : 0u;
        mbar_expect_tx(&s_bar, tb + rb);
        tma_load_1d(s_tsc, P.fwd_tsc, tb, &s_bar);
        if (rb) tma_load_1d(s_rsc, P.fwd_rsc, rb, &s_bar);
      }
      mbar_wait(&s_bar, phase); phase ^= 1;
    }
    const float *tBM = s_tsc, *tMM = s_tsc + Mpad, *tIM = s_tsc + 2 * Mpad, *tDM = s_tsc + 3 * Mpad,
                *tMD = s_tsc + 4 * Mpad, *tMI = s_tsc + 5 * Mpad, *tII = s_tsc + 6 * Mpad, *tDD = s_tsc + 7 * Mpad;
    const float *rsc_base = cfg.rsc_in_smem ? s_rsc : P.fwd_rsc;
    const float tEC = P.xf_E_move, tEJ = P.xf_E_loop;

    for (int e = it.e_begin + warp; e < it.e_end; e += nwarps) {
      const int s = wl.ent_s[e];
      const int L = sd.len[s];
      const uint8_t *seq = sd.res + sd.off[s];
      const float pmove = sd.pmove[s], ploop = 1.0f - pmove;
      float *xout = (out.fwd_xmx && out.xoff[e] >= 0) ? out.fwd_xmx + out.xoff[e] * 6 : nullptr;

      for (int j = lane; j < 6 * S; j += 32) rows[j] = 0.0f;
      __syncwarp();
      float xN = 1.0f, xJ = 0.0f, xC = 0.0f, xB = pmove, xE = 0.0f;
      float totscale = 0.0f;
      if (xout && lane == 0) { xout[0] = 0.f; xout[1] = 1.f; xout[2] = 0.f; xout[3] = xB; xout[4] = 0.f; xout[5] = 1.f; }

      for (int i = 1; i <= L; i++) {
        const int x = seq[i - 1];
        float *cur = rows + (i & 1) * 3 * S, *prv = rows + ((i & 1) ^ 1) * 3 * S;
        const float *pM = prv, *pI = prv + S, *pD = prv + 2 * S;
        float *cM = cur, *cI = cur + S, *cD = cur + 2 * S;
        const float *rs = rsc_base + (size_t)x * Mpad;
        float esum = 0.0f;
        float carry = 0.0f;                                // D(i, k0) entering the tile
        for (int k0 = 1; k0 <= Mpad; k0 += 32) {
          const int k = k0 + lane, c = k - 1;
          const float mp = pM[k - 1], ip = pI[k - 1], dp = pD[k - 1], mo = pM[k], io = pI[k];
          float m = xB * tBM[c];
          m += mp * tMM[c];
          m += ip * tIM[c];
          m += dp * tDM[c];
          m *= rs[c];
          cM[k] = m;
          cI[k] = mo * tMI[c] + io * tII[c];
          // D(i,k+1) = m*tMD[k] + D(i,k)*tDD[k]: inclusive affine scan over the tile, then apply the carry
          float A = m * tMD[c], T = tDD[c];
#pragma unroll
          for (int dlt = 1; dlt < 32; dlt <<= 1) {
            const float A2 = __shfl_up_sync(FULL, A, dlt), T2 = __shfl_up_sync(FULL, T, dlt);
            if (lane >= dlt) { A = A + A2 * T; T = T2 * T; }
          }
          const float dnext = A + carry * T;               // D(i, k+1)
          cD[k + 1] = dnext;
          esum += m + dnext;                               // M(i,k) and D(i,k+1) both reach E (D(i,1) = 0)
          carry = __shfl_sync(FULL, dnext, 31);
        }
        xE = warp_sum(esum);
        xN = xN * ploop;
        xC = (xC * ploop) + (xE * tEC);
        xJ = (xJ * ploop) + (xE * tEJ);
        xB = (xJ * pmove) + (xN * pmove);
        float scale = 1.0f;
        __syncwarp();
        if (xE > 1.0e4f) {                                 // sparse rescaling (fwdback.c:418-435)
          xN = xN / xE; xC = xC / xE; xJ = xJ / xE; xB = xB / xE;
          const float inv = 1.0f / xE;
          for (int k = lane + 1; k <= Mpad + 1; k += 32) { cM[k] *= inv; cI[k] *= inv; cD[k] *= inv; }
          scale = xE;
          totscale = (float)((double)totscale + log((double)xE));
          xE = 1.0f;
          __syncwarp();
        }
        if (xout && lane == 0) {
          float *r = xout + (size_t)i * 6;
          r[0] = xE; r[1] = xN; r[2] = xJ; r[3] = xB; r[4] = xC; r[5] = scale;
        }
      }
      if (lane == 0) {
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
// Backward parser (needs the Forward pass's per-row scale factors)
// =================================================================================================
__global__ void __launch_bounds__(256) bck_kernel(const WorkList wl, const SeqDev sd, const DpCfg cfg, const StageOut out)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t s_bar;
  __shared__ int s_item;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int Mp_max = cfg.max_Mpad;
  const int S = Mp_max + 34;
  float *s_tsc = reinterpret_cast<float *>(smem);
  float *s_rsc = s_tsc + 8 * Mp_max;
  float *s_rows = s_rsc + (cfg.rsc_in_smem ? 32 * Mp_max : 0);
  float *rows = s_rows + (size_t)warp * 6 * S;

  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  uint32_t phase = 0;
  int cur_p = -1;
  Item it;
  while (next_item(wl, &s_item, it)) {
    const ProfDev &P = wl.profs[it.p];
    const int Mpad = P.Mpad, M = P.M;
    if (it.p != cur_p) {
      cur_p = it.p;
      if (threadIdx.x == 0) {
        const uint32_t tb = 8u * Mpad * 4u, rb = cfg.rsc_in_smem ? 32u * Mpad * 4u : 0u;
        mbar_expect_tx(&s_bar, tb + rb);
        tma_load_1d(s_tsc, P.fwd_tsc, tb, &s_bar);
        if (rb) tma_load_1d(s_rsc, P.fwd_rsc, rb, &s_bar);
      }
      mbar_wait(&s_bar, phase); phase ^= 1;
    }
    const float *tBM = s_tsc, *tMM = s_tsc + Mpad, *tIM = s_tsc + 2 * Mpad, *tDM = s_tsc + 3 * Mpad,
                *tMD = s_tsc + 4 * Mpad, *tMI = s_tsc + 5 * Mpad, *tII = s_tsc + 6 * Mpad, *tDD = s_tsc + 7 * Mpad;
    const float *rsc_base = cfg.rsc_in_smem ? s_rsc : P.fwd_rsc;
    const float tEC = P.xf_E_move, tEJ = P.xf_E_loop;

    for (int e = it.e_begin + warp; e < it.e_end; e += nwarps) {
      const int s = wl.ent_s[e];
      const int L = sd.len[s];
      const uint8_t *seq = sd.res + sd.off[s];
      const float pmove = sd.pmove[s], ploop = 1.0f - pmove;
      const float *fx = out.fwd_xmx + out.xoff[e] * 6;
      float *bx = out.bck_xmx ? out.bck_xmx + out.xoff[e] * 6 : nullptr;

      // Row L.  Node k lives at index k; index M+1.. are zero so that "k+1" reads past the model give 0.
      float *cur = rows + (L & 1) * 3 * S;
      float xJ = 0.0f, xB = 0.0f, xN = 0.0f, xC = pmove, xE = xC * tEC;
      for (int j = lane; j < 6 * S; j += 32) rows[j] = 0.0f;
      __syncwarp();
      bool own_scales = false;
      float totscale;
      {
        float *cM = cur, *cD = cur + 2 * S;
        // D(L,k) = xE + tDD[k]*D(L,k+1), k = M..1 ; then M(L,k) = xE + tMD[k]*D(L,k+1)
        float carry = 0.0f;
        for (int k0 = ((M - 1) / 32) * 32 + 1; k0 >= 1; k0 -= 32) {
          const int k = k0 + lane, c = k - 1;
          const bool in = (k <= M);
          float A = in ? xE : 0.0f, T = in ? tDD[c] : 0.0f;
#pragma unroll
          for (int dlt = 1; dlt < 32; dlt <<= 1) {
            const float A2 = __shfl_down_sync(FULL, A, dlt), T2 = __shfl_down_sync(FULL, T, dlt);
            if (lane + dlt < 32) { A = A + T * A2; T = T * T2; }
          }
          const float d = A + T * carry;
          if (in) cD[k] = d;
          carry = __shfl_sync(FULL, d, 0);
        }
        __syncwarp();
        for (int k = lane + 1; k <= M; k += 32) cM[k] = xE + tMD[k - 1] * cD[k + 1];
        const float scL = fx[(size_t)L * 6 + 5];
        if (scL > 1.0f) {
          xE = xE / scL; xN = xN / scL; xC = xC / scL; xJ = xJ / scL; xB = xB / scL;
          const float inv = 1.0f / scL;
          __syncwarp();
          for (int k = lane + 1; k <= M; k += 32) { cM[k] *= inv; cD[k] *= inv; }
        }
        totscale = (float)log((double)scL);
        if (bx && lane == 0) { float *r = bx + (size_t)L * 6; r[0] = xE; r[1] = xN; r[2] = xJ; r[3] = xB; r[4] = xC; r[5] = scL; }
        __syncwarp();
      }

      for (int i = L - 1; i >= 1; i--) {
        const int x = seq[i];                               // residue x_{i+1}
        float *cur = rows + (i & 1) * 3 * S, *prv = rows + ((i & 1) ^ 1) * 3 * S;
        const float *pM = prv, *pI = prv + S;
        float *cM = cur, *cI = cur + S, *cD = cur + 2 * S;
        const float *rs = rsc_base + (size_t)x * Mpad;
        // phase 1: I(i,k), partial M/D(i,k), and B(i) = sum_k M(i+1,k) e(k) tBM[k]
        float bsum = 0.0f;
        for (int k = lane + 1; k <= Mpad; k += 32) {
          const int c = k - 1;
          const bool nxt = (k < M);
          const float mpv = nxt ? pM[k + 1] * rs[c + 1] : 0.0f;      // M(i+1,k+1) * e(M_{k+1}, x_{i+1})
          const float ipv = pI[k];
          const float tmm = nxt ? tMM[c + 1] : 0.0f, tim = nxt ? tIM[c + 1] : 0.0f, tdm = nxt ? tDM[c + 1] : 0.0f;
          cI[k] = (ipv * tII[c]) + (mpv * tim);
          cD[k] = mpv * tdm;
          cM[k] = (ipv * tMI[c]) + (mpv * tmm);
          if (k <= M) bsum += (pM[k] * rs[c]) * tBM[c];
        }
        xB = warp_sum(bsum);
        xC = xC * ploop;
        xJ = (xB * pmove) + (xJ * ploop);
        xN = (xB * pmove) + (xN * ploop);
        xE = (xC * tEC) + (xJ * tEJ);
        __syncwarp();
        // phases 3-4: D(i,k) = partial + xE + tDD[k]*D(i,k+1)  (reverse affine scan)
        float carry = 0.0f;
        for (int k0 = ((M - 1) / 32) * 32 + 1; k0 >= 1; k0 -= 32) {
          const int k = k0 + lane, c = k - 1;
          const bool in = (k <= M);
          float A = in ? cD[k] + xE : 0.0f, T = in ? tDD[c] : 0.0f;
#pragma unroll
          for (int dlt = 1; dlt < 32; dlt <<= 1) {
            const float A2 = __shfl_down_sync(FULL, A, dlt), T2 = __shfl_down_sync(FULL, T, dlt);
            if (lane + dlt < 32) { A = A + T * A2; T = T * T2; }
          }
          const float d = A + T * carry;
          if (in) cD[k] = d;
          carry = __shfl_sync(FULL, d, 0);
        }
        __syncwarp();
        // phase 5 (+ the M->E part of phase 3)
        for (int k = lane + 1; k <= M; k += 32) cM[k] = (cM[k] + xE) + tMD[k - 1] * ((k < M) ? cD[k + 1] : 0.0f);
        if (xB > 1.0e16f) own_scales = true;
        float scale = own_scales ? ((xB > 1.0e4f) ? xB : 1.0f) : fx[(size_t)i * 6 + 5];
        __syncwarp();
        if (scale > 1.0f) {
          xE /= scale; xN /= scale; xJ /= scale; xB /= scale; xC /= scale;
          const float inv = 1.0f / scale;
          for (int k = lane + 1; k <= M; k += 32) { cM[k] *= inv; cD[k] *= inv; cI[k] *= inv; }
          totscale = (float)((double)totscale + log((double)scale));
          __syncwarp();
        }
        if (bx && lane == 0) { float *r = bx + (size_t)i * 6; r[0] = xE; r[1] = xN; r[2] = xJ; r[3] = xB; r[4] = xC; r[5] = scale; }
      }
      // termination at i = 0
      {
        const float *pM = rows + (1 & 1) * 3 * S;
        const float *rs = rsc_base + (size_t)seq[0] * Mpad;
        float bsum = 0.0f;
        if (L >= 1) for (int k = lane + 1; k <= M; k += 32) bsum += (pM[k] * rs[k - 1]) * tBM[k - 1];
        xB = warp_sum(bsum);
        xN = (xB * pmove) + (xN * ploop);
        if (lane == 0) {
          if (bx) { bx[0] = 0.f; bx[1] = xN; bx[2] = 0.f; bx[3] = xB; bx[4] = 0.f; bx[5] = 1.0f; }
          int st = B2H_OK; float sc;
          if (isnan(xN) || (L > 0 && xN == 0.0f) || isinf(xN)) { st = B2H_ERANGE; sc = isnan(xN) ? NAN : (xN == 0.0f ? -INFINITY : INFINITY); }
          else sc = (float)((double)totscale + log((double)xN));
          out.sc[e] = sc;
          if (out.status) out.status[e] = own_scales ? (st | 0x100) : st;
        }
      }
      __syncwarp();
    }
  }
}

// =================================================================================================
// Bias-composition filter: Forward of a 2-state HMM, one THREAD per comparison (O(L) scalar work).
// Operation order and operand types follow esl_hmm_Forward (esl_hmm.c:353-412) exactly.
// =================================================================================================
__global__ void __launch_bounds__(128) bias_kernel(const WorkList wl, const SeqDev sd, int nentries_total_hint, float *filtersc)
{
  const int nent = wl.poff[wl.P];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nent; e += gridDim.x * blockDim.x) {
    int lo = 0, hi = wl.P;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (wl.poff[mid] <= e) lo = mid; else hi = mid; }
    const ProfDev &P = wl.profs[lo];
    const float *eo = P.bias_eo;
    const int s = wl.ent_s[e];
    const int L = sd.len[s];
    const uint8_t *seq = sd.res + sd.off[s];
    const float p1 = sd.p1[s];
    const float t00 = p1, t01 = 1.0f - p1, t10 = P.bias_t10, t11 = P.bias_t11;   // p7_bg_SetLength overwrites t[0][0..1] (p7_bg.c:191-194)
    const float pi0 = 0.999f, pi1 = 0.001f;
    float logsc = 0.0f;
    int x = seq[0];
    float d0 = __fmul_rn(eo[x * 2 + 0], pi0), d1 = __fmul_rn(eo[x * 2 + 1], pi1);
    float mx = fmaxf(d1, fmaxf(d0, 0.0f));
    d0 = __fdiv_rn(d0, mx); d1 = __fdiv_rn(d1, mx);
    logsc = __fadd_rn(logsc, (float)log((double)mx));
    for (int i = 2; i <= L; i++) {
      x = seq[i - 1];
      float n0 = __fadd_rn(__fadd_rn(0.0f, __fmul_rn(d0, t00)), __fmul_rn(d1, t10));
      float n1 = __fadd_rn(__fadd_rn(0.0f, __fmul_rn(d0, t01)), __fmul_rn(d1, t11));
      n0 = __fmul_rn(n0, eo[x * 2 + 0]);
      n1 = __fmul_rn(n1, eo[x * 2 + 1]);
      mx = fmaxf(n1, fmaxf(n0, 0.0f));
      d0 = __fdiv_rn(n0, mx); d1 = __fdiv_rn(n1, mx);
      logsc = __fadd_rn(logsc, (float)log((double)mx));
    }
    float last = __fadd_rn(__fadd_rn(0.0f, __fmul_rn(d0, 1.0f)), __fmul_rn(d1, 1.0f));
    logsc = __fadd_rn(logsc, (float)log((double)last));
    filtersc[e] = __fadd_rn(__fadd_rn(logsc, sd.flta[s]), sd.fltb[s]);
  }
}

struct LaunchCfg { int nwarps; int rsc_in_smem; size_t smem; };

LaunchCfg pick_cfg(int max_Mpad, int elem_bytes)
{
  const size_t budget = 200 * 1024;
  const size_t S = (size_t)max_Mpad + 34;
  const size_t tsc = (size_t)8 * max_Mpad * elem_bytes, rsc = (size_t)32 * max_Mpad * elem_bytes, row = 6 * S * elem_bytes;
  LaunchCfg c;
  c.rsc_in_smem = (tsc + rsc + 2 * row <= budget) ? 1 : 0;
  size_t fixed = tsc + (c.rsc_in_smem ? rsc : 0);
  size_t avail = budget > fixed ? budget - fixed : 0;
  int nw = (int)std::min<size_t>(8, avail / row);
  if (nw < 1) nw = 1;
  // several small CTAs per SM beat one large one when the model is short
  c.nwarps = nw;
  c.smem = fixed + (size_t)nw * row + 128;
  return c;
}

// Profiles of a work list are sorted by Mpad.  Models of up to 512 nodes go to the register-resident kernels
// (b2h_dpreg.cu: 2 / 4 / 8 / 12 / 16 nodes per lane); longer ones to the shared-memory kernels of this file, one launch
// per size class so that mid-sized models do not inherit the largest model's shared-memory footprint.
// With <out2>, every size class gets a second launch (kernel2 / kind2) right behind the first ON THE SAME STREAM: the
// Backward pass of a class starts as soon as that class's Forward is done instead of waiting for the slowest class.
template <typename K, typename K2>
int launch_dp(b2h_ctx *ctx, K kernel, int kind, const WorkList &wl_in, const SeqDev &sd, const std::vector<int> &mpads, int elem_bytes, int nitems_hint, const StageOut &out,
              K2 kernel2, int kind2, const StageOut *out2)
{
  static const int bounds[] = {2048, 3072, 1 << 30};
  const int P = (int)mpads.size();
  int plo = 0, cls = 0;
  ForkJoin fj(ctx);
  const b2h_regclass *rcls; const int nrcls = b2h_reg_classes(&rcls);
  for (int rc = 0; rc < nrcls && plo < P; rc++) {
    int phi = plo;
    while (phi < P && mpads[phi] <= rcls[rc].bound) phi++;
    if (phi > plo) {
      WorkList wl = wl_in; wl.plo = plo; wl.phi = phi; wl.counter = wl_in.counter + cls;
      cudaStream_t strm = fj.next();
      StageOut o = out;
      static const bool vit32 = getenv("B2H_VIT32") != nullptr;        // debugging aid: 32-bit Viterbi kernel only
      if (kind == 0 && rcls[rc].W == 1 && out.status && !vit32) {
        // packed 16-bit kernel first; what it cannot decide exactly (strong hits, mostly) is redone in 32-bit lanes
        int st2 = b2h_launch_vit2(ctx, (rcls[rc].C + 1) & ~1, wl, sd, nitems_hint, out, strm);
        if (st2 != B2H_OK) return st2;
        o.redo_only = 1;
        static const bool noredo = getenv("B2H_VIT2_NOREDO") != nullptr; // debugging aid: leave B2H_REDO in the status array
        if (noredo) { plo = phi; cls++; continue; }
      }
      int st = b2h_launch_dpreg(ctx, kind, rcls[rc].C, rcls[rc].W, wl, sd, nitems_hint, o, strm);
      if (st != B2H_OK) return st;
      if (out2 && (st = b2h_launch_dpreg(ctx, kind2, rcls[rc].C, rcls[rc].W, wl, sd, nitems_hint, *out2, strm)) != B2H_OK) return st;
      plo = phi; cls++;
    }
  }
  while (plo < P) {
    int b = 0; while (mpads[plo] > bounds[b]) b++;
    int phi = plo; int mx = 0;
    while (phi < P && mpads[phi] <= bounds[b]) { mx = std::max(mx, mpads[phi]); phi++; }
    LaunchCfg c = pick_cfg(mx, elem_bytes);
    if ((size_t)6 * ((size_t)mx + 34) * elem_bytes + (size_t)8 * mx * elem_bytes > 220 * 1024) {
      ctx->err = "model too long for the shared-memory DP kernels"; return B2H_EINVAL;
    }
    B2H_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    int occ = 1;
    B2H_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, c.nwarps * 32, c.smem));
    if (occ < 1) occ = 1;
    int grid = ctx->sm_count * occ;
    if (nitems_hint > 0 && grid > nitems_hint) grid = nitems_hint;
    if (grid < 1) grid = 1;
    WorkList wl = wl_in; wl.plo = plo; wl.phi = phi; wl.counter = wl_in.counter + cls;
    cudaStream_t strm = fj.next();
    B2H_CUDA(cudaMemsetAsync(wl.counter, 0, sizeof(int), strm));
    DpCfg cfg; cfg.max_Mpad = mx; cfg.rsc_in_smem = c.rsc_in_smem;
    kernel<<<grid, c.nwarps * 32, c.smem, strm>>>(wl, sd, cfg, out);
    ctx->launches++;
    B2H_CUDA(cudaGetLastError());
    if (out2) {
      B2H_CUDA(cudaFuncSetAttribute(kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
      B2H_CUDA(cudaMemsetAsync(wl.counter, 0, sizeof(int), strm));
      kernel2<<<grid, c.nwarps * 32, c.smem, strm>>>(wl, sd, cfg, *out2);
      ctx->launches++;
      B2H_CUDA(cudaGetLastError());
    }
    plo = phi; cls++;
  }
  return B2H_OK;
}

} // namespace

int b2h_launch_viterbi(b2h_ctx *ctx, const WorkList &wl, const SeqDev &sd, const std::vector<int> &mpads, int nitems_hint, StageOut out)
{ return launch_dp(ctx, vit_kernel, 0, wl, sd, mpads, 2, nitems_hint, out, vit_kernel, 0, nullptr); }
int b2h_launch_forward(b2h_ctx *ctx, const WorkList &wl, const SeqDev &sd, const std::vector<int> &mpads, int nitems_hint, StageOut out)
{ return launch_dp(ctx, fwd_kernel, 1, wl, sd, mpads, 4, nitems_hint, out, fwd_kernel, 1, nullptr); }
int b2h_launch_backward(b2h_ctx *ctx, const WorkList &wl, const SeqDev &sd, const std::vector<int> &mpads, int nitems_hint, StageOut out)
{ return launch_dp(ctx, bck_kernel, 2, wl, sd, mpads, 4, nitems_hint, out, bck_kernel, 2, nullptr); }
// Forward (with stored specials) and Backward of the same work list, class by class on one stream each
int b2h_launch_forward_backward(b2h_ctx *ctx, const WorkList &wl, const SeqDev &sd, const std::vector<int> &mpads, int nitems_hint, StageOut fwd, StageOut bck)
{ return launch_dp(ctx, fwd_kernel, 1, wl, sd, mpads, 4, nitems_hint, fwd, bck_kernel, 2, &bck); }

int b2h_launch_bias(b2h_ctx *ctx, const WorkList &wl, const SeqDev &sd, int nentries_hint, float *filtersc)
{
  int grid = (nentries_hint + 127) / 128;
  if (grid < 1) grid = 1;
  if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
  bias_kernel<<<grid, 128, 0, ctx->stream>>>(wl, sd, nentries_hint, filtersc);
  ctx->launches++;
  B2H_CUDA(cudaGetLastError());
  return B2H_OK;
}
