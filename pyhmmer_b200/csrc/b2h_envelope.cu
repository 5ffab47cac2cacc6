// b2h_envelope.cu -- the O(M*Ld) part of rescore_isolated_domain (p7_domaindef.c:814-982) for every envelope of a
// batch of survivors, as sm_100a kernels:
//
//   efwd_kernel   p7_Forward  (fwdback.c:256-463, do_full)   full matrix F + specials, envelope score
//   ebck_kernel   p7_Backward (fwdback.c:468-733) fused with p7_Decoding (decoding.c:76-134): posterior matrix
//                 PP(i,k) = F(i,k)*B(i,k) (row factor kept apart) and the column sums p7_Null2_ByExpectation needs
//   eoa_kernel    p7_OptimalAccuracy (optacc.c:58-176) fill, then p7_OATrace (optacc.c:225-268) as a warp-serial walk
//
// One group of W warps per envelope (persistent CTAs pull envelopes, largest first, from a counter); lane gl of the
// group owns the C consecutive nodes gl*C .. gl*C+C-1, its cells of the previous row and its transition scores live
// in registers (the layout of b2h_dpreg.cu), rows are streamed to / from HBM as three (two) planes of Mp = 32*C*W
// floats.  Emission rows come straight from the node-major table in L2 (prefetched one row ahead): groups of one
// CTA work on different profiles, so there is no per-CTA table to stage.
// Unihit configuration (p7_oprofile_ReconfigUnihit): xf[E][MOVE] = 1, xf[E][LOOP] = 0, pmove = 2/(L+2).
#include <cuda_runtime.h>
#include <cmath>
#include <algorithm>
#include "b2h_internal.h"

namespace {

constexpr uint32_t FULL = 0xffffffffu;
constexpr int MAXGRP = 8;
#define NEGINF (-INFINITY)

template <int W>
__device__ __forceinline__ void group_sync(int grp)
{
  if (W > 1) asm volatile("bar.sync %0, %1;" :: "r"(grp + 1), "r"(W * 32) : "memory");
  else __syncwarp();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// next envelope of this launch for the group (all threads of the group get the same value)
template <int W>
__device__ __forceinline__ int next_env(const EnvDev &ev, int *s_slot, int grp, int wi, int lane)
{
  group_sync<W>(grp);
  if (wi == 0 && lane == 0) s_slot[grp] = atomicAdd(ev.counter, 1);
  group_sync<W>(grp);
  return ev.e_lo + s_slot[grp];
}

// C consecutive floats of a node-major row (k0 .. k0+C-1), zeros beyond Mpad
template <int C>
__device__ __forceinline__ void load_nodes(const float *row, int k0, int Mpad, float (&r)[C])
{
#pragma unroll
  for (int g = 0; g < C / 4; g++) {
    if (k0 + 4 * g < Mpad) {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(row + k0 + 4 * g));
      r[4*g] = v.x; r[4*g+1] = v.y; r[4*g+2] = v.z; r[4*g+3] = v.w;
    } else { r[4*g] = r[4*g+1] = r[4*g+2] = r[4*g+3] = 0.f; }
  }
}
template <int C>
__device__ __forceinline__ void load_plane(const float *p, float (&r)[C])     // plain (coherent) loads of scratch written by an earlier kernel
{
#pragma unroll
  for (int g = 0; g < C / 4; g++) {
    const float4 v = *reinterpret_cast<const float4 *>(p + 4 * g);
    r[4*g] = v.x; r[4*g+1] = v.y; r[4*g+2] = v.z; r[4*g+3] = v.w;
  }
}
template <int C>
__device__ __forceinline__ void store_plane(float *p, const float (&r)[C])
{
#pragma unroll
  for (int g = 0; g < C / 4; g++) *reinterpret_cast<float4 *>(p + 4 * g) = make_float4(r[4*g], r[4*g+1], r[4*g+2], r[4*g+3]);
}

// =================================================================================================
// Forward, full matrix
// =================================================================================================
template <int C, int W>
__global__ void __launch_bounds__(256) efwd_kernel(const EnvDev ev, const SeqDev sd)
{
  __shared__ int s_slot[MAXGRP];
  __shared__ float s_x[2][MAXGRP][8][W];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = warp / W, wi = warp % W;
  const int gl = wi * 32 + lane;
  constexpr int Mp = 32 * C * W;
  const int k0 = gl * C;
  for (;;) {
    const int e = next_env<W>(ev, s_slot, grp, wi, lane);
    if (e >= ev.e_hi) break;
    const ProfDev &P = ev.profs[ev.prof[e]];
    const int Mpad = P.Mpad;
    float tBM[C], tMM[C], tIM[C], tDM[C], tMD[C], tMI[C], tII[C], tDD[C];
    {
      const float *ts = P.fwd_tsc;
#pragma unroll
      for (int c = 0; c < C; c++) {
        const bool in = k0 + c < Mpad;
        tBM[c] = in ? ts[0 * Mpad + k0 + c] : 0.f; tMM[c] = in ? ts[1 * Mpad + k0 + c] : 0.f; tIM[c] = in ? ts[2 * Mpad + k0 + c] : 0.f;
        tDM[c] = in ? ts[3 * Mpad + k0 + c] : 0.f; tMD[c] = in ? ts[4 * Mpad + k0 + c] : 0.f; tMI[c] = in ? ts[5 * Mpad + k0 + c] : 0.f;
        tII[c] = in ? ts[6 * Mpad + k0 + c] : 0.f; tDD[c] = in ? ts[7 * Mpad + k0 + c] : 0.f;
      }
    }
    float tDDin = __shfl_up_sync(FULL, tDD[C - 1], 1);
    if (lane == 0) tDDin = (W > 1 && wi > 0) ? 1.0f : 0.f;     // the value entering warp wi > 0 is handed over complete
    const int s = ev.seq[e], L = ev.Ld[e], i0 = ev.i0[e];
    const uint8_t *seq = sd.res + sd.off[s] + (i0 - 1);        // residue of envelope row i is seq[i-1]
    const float pmove = ev.pmove[e], ploop = 1.0f - pmove;
    const float tEC = 1.0f, tEJ = 0.0f;
    const float *rsc = (ev.rsc_off && ev.rsc_off[e] >= 0) ? ev.rsc_pool + ev.rsc_off[e] : P.fwd_rsc;   // this envelope's own emission odds (long targets)
    float *Fm = ev.F + 3 * ev.moff[e];
    float *xout = ev.fx + ev.xoff[e] * 6;
    float M[C], I[C], D[C];
#pragma unroll
    for (int c = 0; c < C; c++) { M[c] = 0.f; I[c] = 0.f; D[c] = 0.f; }
    float xN = 1.0f, xJ = 0.0f, xC = 0.0f, xB = pmove, xE = 0.0f, totscale = 0.0f;
    float cM = 0.f, cI = 0.f, cD = 0.f;
    if (gl == 0) { xout[0] = 0.f; xout[1] = 1.f; xout[2] = 0.f; xout[3] = xB; xout[4] = 0.f; xout[5] = 1.f; }
    float rn[C];
    if (L >= 1) load_nodes<C>(rsc + (size_t)seq[0] * Mpad, k0, Mpad, rn);

    for (int i = 1; i <= L; i++) {
      float r[C];
#pragma unroll
      for (int c = 0; c < C; c++) r[c] = rn[c];
      if (i < L) load_nodes<C>(rsc + (size_t)seq[i] * Mpad, k0, Mpad, rn);     // prefetch the next row's emissions
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
        float Tp = __shfl_up_sync(FULL, Tt, 1);
        if (lane == 0) Tp = 1.0f;
        float sT = 0.f;
#pragma unroll
        for (int c = 0; c < C; c++) { D[c] = D[c] + din * T[c]; esum += D[c]; sT += T[c]; }
        const float S1 = warp_sum(esum), S2 = warp_sum(Tp * sT);
        float (*X)[W] = s_x[i & 1][grp];
        if (lane == 31) { X[0][wi] = A; X[1][wi] = Tt; X[2][wi] = aout; X[3][wi] = tDD[C - 1]; X[4][wi] = M[C - 1]; X[5][wi] = I[C - 1]; }
        if (lane == 0)  { X[6][wi] = S1; X[7][wi] = S2; }
        group_sync<W>(grp);
        float o = 0.f, o_in = 0.f;
        xE = 0.f;
#pragma unroll
        for (int w = 0; w < W; w++) {
          xE += X[6][w] + o * X[7][w];
          if (w == wi) o_in = o;
          const float Dl = X[0][w] + X[1][w] * o;
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
      float *row = Fm + (size_t)i * 3 * Mp + k0;
      store_plane<C>(row, M); store_plane<C>(row + Mp, D); store_plane<C>(row + 2 * Mp, I);
      if (gl == 0) { float *q = xout + (size_t)i * 6; q[0] = xE; q[1] = xN; q[2] = xJ; q[3] = xB; q[4] = xC; q[5] = scale; }
    }
    if (gl == 0) {
      int st = B2H_OK; float sc;
      if (isnan(xC) || (L > 0 && xC == 0.0f) || isinf(xC)) { st = B2H_ERANGE; sc = INFINITY; }   // the caller ignores p7_Forward's status
      else sc = (float)((double)totscale + log((double)(xC * pmove)));
      ev.envsc[e] = sc;
      ev.status[e] = st;
    }
  }
}

// =================================================================================================
// Backward fused with posterior decoding
// =================================================================================================
template <int C, int W>
__global__ void __launch_bounds__(256) ebck_kernel(const EnvDev ev, const SeqDev sd)
{
  __shared__ int s_slot[MAXGRP];
  __shared__ float s_x[MAXGRP][4][W];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = warp / W, wi = warp % W;
  const int gl = wi * 32 + lane;
  constexpr int Mp = 32 * C * W;
  const int k0 = gl * C;
  float (*X)[W] = s_x[grp];
  for (;;) {
    const int e = next_env<W>(ev, s_slot, grp, wi, lane);
    if (e >= ev.e_hi) break;
    const ProfDev &P = ev.profs[ev.prof[e]];
    const int Mpad = P.Mpad, M = P.M;
    float tBM[C], tMD[C], tMI[C], tII[C], tDD[C], tMMn[C], tIMn[C], tDMn[C];
    {
      const float *ts = P.fwd_tsc;
#pragma unroll
      for (int c = 0; c < C; c++) {
        const int k = k0 + c;
        const bool in = k < M, inn = (k + 1) < M;
        tBM[c] = in ? ts[0 * Mpad + k] : 0.f; tMD[c] = in ? ts[4 * Mpad + k] : 0.f; tMI[c] = in ? ts[5 * Mpad + k] : 0.f;
        tII[c] = in ? ts[6 * Mpad + k] : 0.f; tDD[c] = in ? ts[7 * Mpad + k] : 0.f;
        tMMn[c] = inn ? ts[1 * Mpad + k + 1] : 0.f; tIMn[c] = inn ? ts[2 * Mpad + k + 1] : 0.f; tDMn[c] = inn ? ts[3 * Mpad + k + 1] : 0.f;
      }
    }
    const int s = ev.seq[e], L = ev.Ld[e], i0 = ev.i0[e];
    const uint8_t *seq = sd.res + sd.off[s] + (i0 - 1);
    const float pmove = ev.pmove[e], ploop = 1.0f - pmove;
    const float tEC = 1.0f, tEJ = 0.0f;
    const float *rsc = (ev.rsc_off && ev.rsc_off[e] >= 0) ? ev.rsc_pool + ev.rsc_off[e] : P.fwd_rsc;
    const float *Fm = ev.F + 3 * ev.moff[e];
    float *Pm = ev.PP + 2 * ev.moff[e];
    const float *fx = ev.fx + ev.xoff[e] * 6;
    float *bx = ev.bx + ev.xoff[e] * 6;
    float Mv[C], Iv[C], Dv[C], aM[C], aI[C];
#pragma unroll
    for (int c = 0; c < C; c++) { aM[c] = 0.f; aI[c] = 0.f; }
    float xJ = 0.0f, xB = 0.0f, xN = 0.0f, xC = pmove, xE = xC * tEC;
    bool own_scales = false;
    float dext = 0.f;

    auto close_dd = [&](void) {
      float T[C];
      T[C - 1] = tDD[C - 1];
#pragma unroll
      for (int c = C - 2; c >= 0; c--) { Dv[c] = Dv[c] + tDD[c] * Dv[c + 1]; T[c] = tDD[c] * T[c + 1]; }
      float A = Dv[0], Tt = T[0];
#pragma unroll
      for (int dlt = 1; dlt < 32; dlt <<= 1) {
        const float A2 = __shfl_down_sync(FULL, A, dlt), T2 = __shfl_down_sync(FULL, Tt, dlt);
        if (lane + dlt < 32) { A = A + Tt * A2; Tt = Tt * T2; }
      }
      float din = __shfl_down_sync(FULL, A, 1);
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
    auto group_sum = [&](float v) -> float {
      if (W == 1) return v;
      if (lane == 0) X[0][wi] = v;
      group_sync<W>(grp);
      float t = X[0][0];
#pragma unroll
      for (int w = 1; w < W; w++) t += X[0][w];
      return t;
    };
    // posterior row i: F(i,k) * B(i,k) for M and I; column sums weighted by Forward's scale of the row
    float fM[C], fI[C];                                          // Forward row of the posterior row being produced (loaded a row ahead)
    auto fetch_f = [&](int i) { const float *fr = Fm + (size_t)i * 3 * Mp + k0; load_plane<C>(fr, fM); load_plane<C>(fr + 2 * Mp, fI); };
    auto emit_pp = [&](int i) {
      float q[C];
      const float fsc = fx[(size_t)i * 6 + 5];
      float *pr = Pm + (size_t)i * 2 * Mp + k0;
#pragma unroll
      for (int c = 0; c < C; c++) { q[c] = fM[c] * Mv[c]; aM[c] += q[c] * fsc; }
      store_plane<C>(pr, q);
#pragma unroll
      for (int c = 0; c < C; c++) { q[c] = fI[c] * Iv[c]; aI[c] += q[c] * fsc; }
      store_plane<C>(pr + Mp, q);
    };
    group_sync<W>(grp);

    // row L
    if (L >= 1) fetch_f(L);
#pragma unroll
    for (int c = 0; c < C; c++) { const bool in = (k0 + c) < M; Dv[c] = in ? xE : 0.f; Iv[c] = 0.f; }
    close_dd();
    float totscale;
    {
      float dnext = __shfl_down_sync(FULL, Dv[0], 1); if (lane == 31) dnext = (W > 1) ? dext : 0.f;
#pragma unroll
      for (int c = 0; c < C; c++) { const bool in = (k0 + c) < M; const float dn = (c == C - 1) ? dnext : Dv[c + 1]; Mv[c] = in ? xE + tMD[c] * dn : 0.f; }
      const float scL = fx[(size_t)L * 6 + 5];
      if (scL > 1.0f) {
        xE = xE / scL; xN = xN / scL; xC = xC / scL; xJ = xJ / scL; xB = xB / scL;
        const float inv = 1.0f / scL;
#pragma unroll
        for (int c = 0; c < C; c++) { Mv[c] *= inv; Dv[c] *= inv; }
      }
      totscale = (float)log((double)scL);
      if (gl == 0) { float *q = bx + (size_t)L * 6; q[0] = xE; q[1] = xN; q[2] = xJ; q[3] = xB; q[4] = xC; q[5] = scL; }
      if (W > 1 && lane == 0) X[3][wi] = Mv[0];
      if (L >= 1) emit_pp(L);
    }

    for (int i = L - 1; i >= 1; i--) {
      const int x = seq[i];                               // x_{i+1}
      fetch_f(i);
      float r[C];
      load_nodes<C>(rsc + (size_t)x * Mpad, k0, Mpad, r);
      float me[C];
#pragma unroll
      for (int c = 0; c < C; c++) me[c] = Mv[c] * r[c];
      float bsum = 0.f;
#pragma unroll
      for (int c = 0; c < C; c++) bsum += me[c] * tBM[c];
      xB = group_sum(warp_sum(bsum));
      float menext = __shfl_down_sync(FULL, me[0], 1);
      if (lane == 31) {
        menext = 0.f;
        if (W > 1 && wi < W - 1) { const int kn = (wi + 1) * 32 * C; menext = X[3][wi + 1] * ((kn < Mpad) ? __ldg(rsc + (size_t)x * Mpad + kn) : 0.f); }
      }
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
      for (int c = 0; c < C; c++) { const bool in = (k0 + c) < M; Dv[c] = in ? Dv[c] + xE : 0.f; }
      close_dd();
      float dnext = __shfl_down_sync(FULL, Dv[0], 1); if (lane == 31) dnext = (W > 1) ? dext : 0.f;
#pragma unroll
      for (int c = 0; c < C; c++) { const bool in = (k0 + c) < M; const float dn = (c == C - 1) ? dnext : Dv[c + 1]; Mv[c] = in ? (Mv[c] + xE) + tMD[c] * dn : 0.f; }
      if (xB > 1.0e16f) own_scales = true;
      const float scale = own_scales ? ((xB > 1.0e4f) ? xB : 1.0f) : fx[(size_t)i * 6 + 5];
      if (scale > 1.0f) {
        xE /= scale; xN /= scale; xJ /= scale; xB /= scale; xC /= scale;
        const float inv = 1.0f / scale;
#pragma unroll
        for (int c = 0; c < C; c++) { Mv[c] *= inv; Dv[c] *= inv; Iv[c] *= inv; }
        totscale = (float)((double)totscale + log((double)scale));
      }
      if (gl == 0) { float *q = bx + (size_t)i * 6; q[0] = xE; q[1] = xN; q[2] = xJ; q[3] = xB; q[4] = xC; q[5] = scale; }
      if (W > 1 && lane == 0) X[3][wi] = Mv[0];
      emit_pp(i);
    }
    {
      float r[C];
      load_nodes<C>(rsc + (size_t)seq[0] * Mpad, k0, Mpad, r);
      float bsum = 0.f;
      if (L >= 1) {
#pragma unroll
        for (int c = 0; c < C; c++) bsum += (Mv[c] * r[c]) * tBM[c];
      }
      xB = group_sum(warp_sum(bsum));
      xN = (xB * pmove) + (xN * ploop);
      if (gl == 0) { bx[0] = 0.f; bx[1] = xN; bx[2] = 0.f; bx[3] = xB; bx[4] = 0.f; bx[5] = 1.0f; }
      // column sums of the posterior matrix: sum_i pp(i,k) * rs[i], rs[i] = (1/xN(0)) * fwd scale(i)
      const float sp = 1.0f / xN;
      float *em = ev.em + ev.moff_n[e] + k0, *ei = ev.ei + ev.moff_n[e] + k0;
#pragma unroll
      for (int c = 0; c < C; c++) { aM[c] *= sp; aI[c] *= sp; }
      store_plane<C>(em, aM); store_plane<C>(ei, aI);
      if (gl == 0) {
        int st = ev.status[e];
        if (own_scales) st |= 0x100;                                  // the row factors are not a constant: rescored on the host
        if (isinf(sp) || isnan(sp)) st |= 0x200;                      // eslERANGE from p7_Decoding
        ev.status[e] = st;
      }
    }
  }
}

// =================================================================================================
// Optimal accuracy: fill, then traceback
// =================================================================================================
enum { G_BM = 1, G_MM = 2, G_IM = 4, G_DM = 8, G_MD = 16, G_MI = 32, G_II = 64, G_DD = 128 };
enum { ST_M = 1, ST_D, ST_I, ST_S, ST_N, ST_B, ST_E, ST_C, ST_T, ST_J };      // same numbering as b2h_domaindef.cpp

template <int C, int W>
__global__ void __launch_bounds__(256) eoa_kernel(const EnvDev ev, const SeqDev sd)
{
  __shared__ int s_slot[MAXGRP];
  __shared__ float s_x[2][MAXGRP][8][W];
  __shared__ float s_y[MAXGRP][2][W];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = warp / W, wi = warp % W;
  const int gl = wi * 32 + lane;
  constexpr int Mp = 32 * C * W;
  const int k0 = gl * C;
  for (;;) {
    const int e = next_env<W>(ev, s_slot, grp, wi, lane);
    if (e >= ev.e_hi) break;
    const ProfDev &P = ev.profs[ev.prof[e]];
    const int Mpad = P.Mpad, M = P.M;
    const float *ts = P.fwd_tsc;
    int gm[C];                                                   // which transitions of node k0+c are possible (t > 0)
#pragma unroll
    for (int c = 0; c < C; c++) {
      const int k = k0 + c; int g = 0;
      if (k < M) {
        if (ts[0 * Mpad + k] > 0.f) g |= G_BM; if (ts[1 * Mpad + k] > 0.f) g |= G_MM; if (ts[2 * Mpad + k] > 0.f) g |= G_IM; if (ts[3 * Mpad + k] > 0.f) g |= G_DM;
        if (ts[4 * Mpad + k] > 0.f) g |= G_MD; if (ts[5 * Mpad + k] > 0.f) g |= G_MI; if (ts[6 * Mpad + k] > 0.f) g |= G_II; if (ts[7 * Mpad + k] > 0.f) g |= G_DD;
      }
      gm[c] = g;
    }
    int gprev = __shfl_up_sync(FULL, gm[C - 1], 1);             // gates of the node left of this lane's first node
    if (lane == 0) gprev = (W > 1 && wi > 0 && k0 - 1 < M) ? ((ts[4 * Mpad + k0 - 1] > 0.f ? G_MD : 0) | (ts[7 * Mpad + k0 - 1] > 0.f ? G_DD : 0)) : 0;
    const int L = ev.Ld[e];
    const float pmove = ev.pmove[e], ploop = 1.0f - pmove;
    const float eM = 1.0f, eL = 0.0f;
    const float *Pm = ev.PP + 2 * ev.moff[e];
    float *Om = ev.OA + 3 * ev.moff[e];
    const float *fx = ev.fx + ev.xoff[e] * 6, *bx = ev.bx + ev.xoff[e] * 6;
    float *ox = ev.ox + ev.xoff[e] * 6;
    const float sp = 1.0f / bx[1];
    float Mv[C], Iv[C], Dv[C];
#pragma unroll
    for (int c = 0; c < C; c++) { Mv[c] = NEGINF; Iv[c] = NEGINF; Dv[c] = NEGINF; }
    {   // row 0: every cell -inf
      float *row = Om + k0;
      store_plane<C>(row, Mv); store_plane<C>(row + Mp, Dv); store_plane<C>(row + 2 * Mp, Iv);
    }
    float xE = NEGINF, xN = 0.f, xJ = NEGINF, xB = 0.f, xC = NEGINF;
    float sum_n = 0.f, sum_j = 0.f, sum_c = 0.f;                 // null2: sums of the special-state posteriors
    if (gl == 0) { ox[0] = xE; ox[1] = xN; ox[2] = xJ; ox[3] = xB; ox[4] = xC; ox[5] = 0.f; }
    float cM = NEGINF, cI = NEGINF, cD = NEGINF;                 // W > 1: previous row's cells of the left warp's last node
    group_sync<W>(grp);

    uint8_t *Bp = ev.BP + ev.moff[e];                            // back-pointers: one byte per cell, (Ld+1) rows of Mp
    float qMn[C], qIn[C];
    if (L >= 1) { const float *pr = Pm + (size_t)1 * 2 * Mp + k0; load_plane<C>(pr, qMn); load_plane<C>(pr + Mp, qIn); }
    for (int i = 1; i <= L; i++) {
      const float rs = sp * fx[(size_t)i * 6 + 5];
      float qM[C], qI[C];
#pragma unroll
      for (int c = 0; c < C; c++) { qM[c] = qMn[c]; qI[c] = qIn[c]; }
      if (i < L) { const float *pr = Pm + (size_t)(i + 1) * 2 * Mp + k0; load_plane<C>(pr, qMn); load_plane<C>(pr + Mp, qIn); }   // next row, one row ahead
      float mp = __shfl_up_sync(FULL, Mv[C - 1], 1), ip = __shfl_up_sync(FULL, Iv[C - 1], 1), dp = __shfl_up_sync(FULL, Dv[C - 1], 1);
      if (lane == 0) { mp = cM; ip = cI; dp = cD; }
      float xEm = NEGINF;
      uint32_t bp[C];                                            // bits 0-1: M came from M/I/D/B; bit 2: I came from I; bit 3: D came from D
#pragma unroll
      for (int c = C - 1; c >= 0; c--) {
        float pm = (c == 0) ? mp : Mv[c - 1], pi = (c == 0) ? ip : Iv[c - 1], pd = (c == 0) ? dp : Dv[c - 1];
        const int g = gm[c];
        float iv = (g & G_MI) ? Mv[c] : 0.0f;
        iv = fmaxf(iv, (g & G_II) ? Iv[c] : 0.0f);
        // p7_OATrace's choices at this cell, made now while the candidates are in registers (select_i / select_m, optacc.c:300-371)
        const float i0p = (g & G_MI) ? Mv[c] : NEGINF, i1p = (g & G_II) ? Iv[c] : NEGINF;
        uint32_t b = (i0p >= i1p) ? 0u : 4u;
        float sv = (g & G_BM) ? xB : 0.0f;
        sv = fmaxf(sv, (g & G_MM) ? pm : 0.0f);
        sv = fmaxf(sv, (g & G_IM) ? pi : 0.0f);
        sv = fmaxf(sv, (g & G_DM) ? pd : 0.0f);
        if (k0 + c == 0) { pm = 0.0f; pi = 0.0f; pd = 0.0f; }     // node 1: zeros shift in (rightshiftz)
        const float p0 = (g & G_MM) ? pm : NEGINF, p1 = (g & G_IM) ? pi : NEGINF, p2 = (g & G_DM) ? pd : NEGINF, p3 = (g & G_BM) ? xB : NEGINF;
        { uint32_t best = 0u; float bv = p0;
          if (p1 > bv) { bv = p1; best = 1u; }
          if (p2 > bv) { bv = p2; best = 2u; }
          if (p3 > bv) { bv = p3; best = 3u; }
          b |= best; }
        bp[c] = b;
        const bool in = (k0 + c) < M;
        sv = in ? sv + qM[c] * rs : NEGINF;
        xEm = fmaxf(xEm, sv);
        Mv[c] = sv;
        Iv[c] = in ? iv + qI[c] * rs : NEGINF;
      }
      // D(k) = max( gate(tMD[k-1], M(k-1)), gate(tDD[k-1], D(k-1)) ): partials from M first
      const float mdl = (gm[C - 1] & G_MD) ? Mv[C - 1] : 0.0f;
      float dleft = __shfl_up_sync(FULL, mdl, 1);
      if (lane == 0) dleft = NEGINF;                           // node 1 (and, for wi > 0, fixed after the exchange)
      if (W > 1) {
        float (*X)[W] = s_x[i & 1][grp];
        if (lane == 31) { X[0][wi] = Mv[C - 1]; X[1][wi] = Iv[C - 1]; X[3][wi] = mdl; }
        group_sync<W>(grp);
        if (wi > 0) { cM = X[0][wi - 1]; cI = X[1][wi - 1]; if (lane == 0) dleft = X[3][wi - 1]; }
      }
      Dv[0] = dleft;
#pragma unroll
      for (int c = 1; c < C; c++) Dv[c] = (gm[c - 1] & G_MD) ? Mv[c - 1] : 0.0f;
      // close the D->D chain: f_k(v) = max(D_k, gate_k ? v : 0), serial in the lane, then a scan of the lane composites
      bool pm[C];
      {
        bool ok = (gprev & G_DD) != 0;                          // D(first-1) -> D(first)
        pm[0] = ok;
        if (!ok && !(lane == 0 && wi == 0)) Dv[0] = fmaxf(Dv[0], 0.0f);
#pragma unroll
        for (int c = 1; c < C; c++) {
          const bool g = (gm[c - 1] & G_DD) != 0;
          Dv[c] = fmaxf(Dv[c], g ? Dv[c - 1] : 0.0f);
          ok = ok && g; pm[c] = ok;
        }
      }
      float A = Dv[C - 1]; int Pk = pm[C - 1] ? 1 : 0;          // lane composite: v -> max(A, Pk ? v : -inf)   (floors are inside A)
#pragma unroll
      for (int dlt = 1; dlt < 32; dlt <<= 1) {
        const float A2 = __shfl_up_sync(FULL, A, dlt); const int P2 = __shfl_up_sync(FULL, Pk, dlt);
        if (lane >= dlt) { A = fmaxf(A, Pk ? A2 : NEGINF); Pk = Pk & P2; }
      }
      float din = __shfl_up_sync(FULL, A, 1);                    // closed D of the previous lane's last node
      if (lane == 0) din = NEGINF;
      if (W > 1) {
        float (*Y)[W] = s_y[grp];
        const int Pp = __shfl_up_sync(FULL, Pk, 1);
        if (lane == 31) { Y[0][wi] = A; Y[1][wi] = __int_as_float(Pk); }
        group_sync<W>(grp);
        float d = NEGINF;
#pragma unroll
        for (int w = 0; w < W - 1; w++) if (w < wi) d = fmaxf(Y[0][w], __float_as_int(Y[1][w]) ? d : NEGINF);
        if (wi > 0) { cD = d; din = (lane == 0) ? d : fmaxf(din, Pp ? d : NEGINF); }
      }
#pragma unroll
      for (int c = 0; c < C; c++) { const bool in = (k0 + c) < M; Dv[c] = in ? fmaxf(Dv[c], pm[c] ? din : NEGINF) : NEGINF; xEm = fmaxf(xEm, Dv[c]); }
      {   // select_d (optacc.c:340-352): D(k) from M(k-1) unless D(k-1) is strictly better; <din> is the closed D left of this lane
        float mleft = __shfl_up_sync(FULL, Mv[C - 1], 1);
        if (lane == 0) mleft = (W > 1 && wi > 0) ? cM : NEGINF;
#pragma unroll
        for (int c = 0; c < C; c++) {
          const int gp = (c == 0) ? gprev : gm[c - 1];
          const float p0 = (gp & G_MD) ? ((c == 0) ? mleft : Mv[c - 1]) : NEGINF;
          const float p1 = (gp & G_DD) ? ((c == 0) ? din : Dv[c - 1]) : NEGINF;
          if (!(p0 >= p1)) bp[c] |= 8u;
        }
        uint32_t *brow = reinterpret_cast<uint32_t *>(Bp + (size_t)i * Mp + k0);
#pragma unroll
        for (int g = 0; g < C / 4; g++) brow[g] = bp[4*g] | (bp[4*g+1] << 8) | (bp[4*g+2] << 16) | (bp[4*g+3] << 24);
      }
      xE = warp_max(xEm);
      if (W > 1) {
        float (*X)[W] = s_x[i & 1][grp];
        if (lane == 0) X[4][wi] = xE;
        if (lane == 31) X[2][wi] = Dv[C - 1];
        group_sync<W>(grp);
        xE = X[4][0];
#pragma unroll
        for (int w = 1; w < W; w++) xE = fmaxf(xE, X[4][w]);
        if (wi > 0) cD = X[2][wi - 1];
      }
      const float ppN = fx[(size_t)(i - 1) * 6 + 1] * bx[(size_t)i * 6 + 1] * ploop * sp;
      const float ppJ = fx[(size_t)(i - 1) * 6 + 2] * bx[(size_t)i * 6 + 2] * ploop * sp;
      const float ppC = fx[(size_t)(i - 1) * 6 + 4] * bx[(size_t)i * 6 + 4] * ploop * sp;
      sum_n += ppN; sum_j += ppJ; sum_c += ppC;
      float t1 = (ploop == 0.0f) ? 0.0f : xJ + ppJ;
      float t2 = (eL == 0.0f) ? 0.0f : xE;
      xJ = fmaxf(t1, t2);
      t1 = (ploop == 0.0f) ? 0.0f : xC + ppC;
      t2 = (eM == 0.0f) ? 0.0f : xE;
      xC = fmaxf(t1, t2);
      xN = (ploop == 0.0f) ? 0.0f : xN + ppN;
      t1 = (pmove == 0.0f) ? 0.0f : xN;
      t2 = (pmove == 0.0f) ? 0.0f : xJ;
      xB = fmaxf(t1, t2);
      float *row = Om + (size_t)i * 3 * Mp + k0;
      store_plane<C>(row, Mv); store_plane<C>(row + Mp, Dv); store_plane<C>(row + 2 * Mp, Iv);
      if (gl == 0) { float *q = ox + (size_t)i * 6; q[0] = xE; q[1] = xN; q[2] = xJ; q[3] = xB; q[4] = xC; q[5] = 0.f; }
    }
    if (gl == 0) { ev.oasc[e] = xC; float *q = ev.xnull + (size_t)e * 4; q[0] = sum_n; q[1] = sum_c; q[2] = sum_j; q[3] = sp; }
    __threadfence_block();
    group_sync<W>(grp);                                           // the matrix is complete for the walk below

    // ---- p7_OATrace: the first warp of the group walks back from (L, C) ----
    if (wi == 0) {
      const int Q = max(2, (M - 1) / 4 + 1);
      int4 *rec = ev.trace + ev.toff[e];
      const int cap = ev.tcap[e];
      int nrec = 0, i = L, k = 0, s0 = ST_C;
      bool bad = false;
      auto O = [&](int r, int kk, int pl) -> float { return Om[(size_t)r * 3 * Mp + (size_t)pl * Mp + (kk - 1)]; };   // node kk (1-based) lives at column kk-1
      auto OX = [&](int r, int q) -> float { return ox[(size_t)r * 6 + q]; };
      int guard = (L + M + 8) * 4 + 64;
      // the record of a step is stored one step later, so that the loads of its posterior probability are not on the
      // critical path of the walk (the next back-pointer load is issued first)
      int ps = -1, pk = 0, pi_ = 0; float pa = 0.f, pb = 0.f;     // pending record: state, k, i, postprob = pa * pb
      while (s0 != ST_S) {
        if (guard-- <= 0) { bad = true; break; }
        int s1 = -1;
        uint32_t b = 0;
        if (s0 == ST_M || s0 == ST_D || s0 == ST_I) b = Bp[(size_t)i * Mp + (k - 1)];
        if (ps >= 0) { if (lane == 0) rec[nrec - 1] = make_int4(ps, pk, pi_, __float_as_int(pa * pb)); ps = -1; }
        switch (s0) {
          case ST_M: { const uint32_t ch = b & 3u; s1 = (ch == 0u) ? ST_M : (ch == 1u) ? ST_I : (ch == 2u) ? ST_D : ST_B; k--; i--; break; }
          case ST_D: s1 = (b & 8u) ? ST_D : ST_M; k--; break;
          case ST_I: s1 = (b & 4u) ? ST_I : ST_M; i--; break;
          case ST_N: s1 = (i == 0) ? ST_S : ST_N; break;
          case ST_C: {
            const float ppC = fx[(size_t)(i - 1) * 6 + 4] * bx[(size_t)i * 6 + 4] * ploop * sp;
            const float p0 = (ploop == 0.0f) ? NEGINF : OX(i - 1, 4) + ppC;
            const float p1 = (eM == 0.0f) ? NEGINF : OX(i, 0);
            s1 = (p0 > p1) ? ST_C : ST_E; break; }
          case ST_J: {
            const float ppJ = fx[(size_t)(i - 1) * 6 + 2] * bx[(size_t)i * 6 + 2] * ploop * sp;
            const float p0 = (ploop == 0.0f) ? NEGINF : OX(i - 1, 2) + ppJ;
            const float p1 = (eL == 0.0f) ? NEGINF : OX(i, 0);
            s1 = (p0 > p1) ? ST_J : ST_E; break; }
          case ST_E: {
            // the reference scans its striped row (q outer, lane r inner): M cells take ties (>=), D cells need >  (optacc.c:404-421).
            // Equivalent: the row maximum v*; if any M cell equals v*, the LAST such M in scan order, else the FIRST such D.
            float vmax = NEGINF;
            for (int kk = lane + 1; kk <= M; kk += 32) vmax = fmaxf(vmax, fmaxf(O(i, kk, 0), O(i, kk, 1)));
            vmax = warp_max(vmax);
            int lastM = -1, firstD = 0x7fffffff;
            for (int kk = lane + 1; kk <= M; kk += 32) {
              const int q = (kk - 1) % Q, r = (kk - 1) / Q, pos = q * 8 + r;
              if (O(i, kk, 0) == vmax) lastM = max(lastM, pos);
              if (O(i, kk, 1) == vmax) firstD = min(firstD, pos + 4);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { lastM = max(lastM, __shfl_xor_sync(FULL, lastM, o)); firstD = min(firstD, __shfl_xor_sync(FULL, firstD, o)); }
            if (lastM >= 0) { const int q = lastM / 8, r = lastM % 8; k = r * Q + q + 1; s1 = ST_M; }
            else if (firstD != 0x7fffffff) { const int q = firstD / 8, r = (firstD % 8) - 4; k = r * Q + q + 1; s1 = ST_D; }
            else { bad = true; }
            break; }
          case ST_B: {
            const float p0 = (pmove == 0.0f) ? NEGINF : OX(i, 1);
            const float p1 = (pmove == 0.0f) ? NEGINF : OX(i, 2);
            s1 = (p0 > p1) ? ST_N : ST_J; break; }
          default: bad = true; break;
        }
        if (bad || s1 == -1) { bad = true; break; }
        if (nrec >= cap) { bad = true; break; }
        // posterior probability of the new state: loads issued now, consumed when the record is stored
        pa = 0.f; pb = 0.f;
        if (s1 == ST_M || s1 == ST_I) {
          pb = sp * fx[(size_t)i * 6 + 5];
          pa = Pm[(size_t)i * 2 * Mp + (size_t)(s1 == ST_M ? 0 : 1) * Mp + (k - 1)];
        } else if (s1 == s0 && (s1 == ST_N || s1 == ST_C || s1 == ST_J)) {
          const int q = (s1 == ST_N) ? 1 : (s1 == ST_J) ? 2 : 4;
          pa = fx[(size_t)(i - 1) * 6 + q] * bx[(size_t)i * 6 + q] * ploop; pb = sp;
        }
        ps = s1; pk = k; pi_ = i;
        nrec++;
        if ((s1 == ST_N || s1 == ST_J || s1 == ST_C) && s1 == s0) i--;
        s0 = s1;
      }
      if (ps >= 0 && lane == 0) rec[nrec - 1] = make_int4(ps, pk, pi_, __float_as_int(pa * pb));
      if (lane == 0) { ev.tlen[e] = bad ? -1 : nrec; }
    }
  }
}

template <typename K>
int launch_env(b2h_ctx *ctx, K kernel, const EnvDev &ev, const SeqDev &sd, cudaStream_t strm)
{
  int occ = 1;
  { const int st = b2h_kernel_occupancy(ctx, (const void *)kernel, 256, 0, &occ); if (st != B2H_OK) return st; }
  const int n = ev.e_hi - ev.e_lo;
  int grid = std::min(ctx->sm_count * occ, std::max(1, n));
  B2H_CUDA(cudaMemsetAsync(ev.counter, 0, sizeof(int), strm));
  kernel<<<grid, 256, 0, strm>>>(ev, sd);
  ctx->launches++;
  B2H_CUDA(cudaGetLastError());
  return B2H_OK;
}

} // namespace

// kind: 0 Forward, 1 Backward + decoding, 2 optimal accuracy + trace.  (C, W) from B2H_ENV_CLASSES.
int b2h_launch_envelope(b2h_ctx *ctx, int kind, int C, int W, const EnvDev &ev, const SeqDev &sd, cudaStream_t strm)
{
#define B2H_ENV_CASE(CC, WW) \
    case (WW) * 64 + (CC): \
      if (kind == 0) return launch_env(ctx, efwd_kernel<CC, WW>, ev, sd, strm); \
      if (kind == 1) return launch_env(ctx, ebck_kernel<CC, WW>, ev, sd, strm); \
      if (kind == 2) return launch_env(ctx, eoa_kernel<CC, WW>, ev, sd, strm); \
      break;
  switch (W * 64 + C) {
    B2H_ENV_CASE(4, 1) B2H_ENV_CASE(8, 1) B2H_ENV_CASE(12, 1) B2H_ENV_CASE(12, 2) B2H_ENV_CASE(12, 4) B2H_ENV_CASE(12, 8)
  }
#undef B2H_ENV_CASE
  return B2H_EINVAL;
}
