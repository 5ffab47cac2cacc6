// b2h_search.cu -- the fused filter cascade of p7_Pipeline (p7_pipeline.c:697-936) for
// P profiles x N sequences, with on-device survivor compaction between stages.
//
//   SSV (all P*N comparisons) --eslENORESULT--> full MSV
//        | P(msv) <= F1
//   bias filter (2-state HMM) | P <= F1 --P <= F2 (skip Viterbi, p7_pipeline.c:748)-----.
//        | P > F2                                                                       |
//   ViterbiFilter | P(vit) <= F2 -------------------------------------------------------+
//   ForwardParser | P(fwd) <= F3
//   list D  --(device->host: a few bytes per survivor)-->  Forward/Backward parsers with stored
//   specials for D only, then host-side domain definition (b2h_domaindef.cpp).
//
// Every list is a struct-of-arrays in HBM appended to with atomics by the stage epilogues and
// regrouped by profile (counting sort) before the next stage, so that the next kernel can stage one
// profile's tables per CTA.  Nothing but list D and its O(L) special-state rows crosses PCIe.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <thread>
#include <vector>
#include "b2h_internal.h"
#include "b2h_domaindef.h"

namespace {

__device__ __forceinline__ double gumbel_surv(double x, double mu, double lambda)
{
  const double y = lambda * (x - mu);
  const double ey = -exp(-y);
  return (fabs(ey) < 5e-9) ? -ey : 1.0 - exp(ey);
}
__device__ __forceinline__ double exp_surv(double x, double mu, double lambda)
{
  return (x < mu) ? 1.0 : exp(-lambda * (x - mu));
}
__device__ __forceinline__ void surv_append(const SurvList &l, int p, int s, float a, float b)
{
  const int slot = atomicAdd(l.n, 1);
  if (slot < l.cap) { l.p[slot] = p; l.s[slot] = s; if (l.a) l.a[slot] = a; if (l.b) l.b[slot] = b; atomicAdd(l.cnt + p, 1); if (l.cnts) atomicAdd(l.cnts + s, 1); }
}
__device__ __forceinline__ float bits(float sc, float null) { return (float)((double)(sc - null) / 0.69314718055994529); }

// after the bias filter (p7_pipeline.c:728-754): entries of the grouped MSV-survivor list
__global__ void bias_post_kernel(const ProfDev *profs, const SeqDev sd, const Grouped g, const int32_t *nent,
                                 const float *filtersc, int do_bias, double F1, double F2, int *cnt_bias, int *cnts_bias, SurvList V, SurvList F)
{
  const int n = *nent;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int p = g.p[e], s = g.s[e];
    const ProfDev &P = profs[p];
    const float usc = g.a[e];
    const float fsc = do_bias ? filtersc[e] : sd.null1[s];
    const double pv = gumbel_surv((double)bits(usc, fsc), (double)P.evparam[0], (double)P.evparam[1]);
    if (do_bias && pv > F1) continue;
    atomicAdd(cnt_bias + p, 1);
    if (cnts_bias) atomicAdd(cnts_bias + s, 1);
    if (pv > F2) surv_append(V, p, s, fsc, 0.f); else surv_append(F, p, s, fsc, 0.f);
  }
}

// Bias filter and ViterbiFilter computed side by side for every MSV survivor; the two tests of p7_pipeline.c:728-762 are
// then applied in the reference's order: bias P-value > F1 drops the comparison; P <= F2 skips the Viterbi test.
__global__ void bias_vit_post_kernel(const ProfDev *profs, const SeqDev sd, const Grouped g, const int32_t *nent,
                                     const float *filtersc, const float *vfsc, int do_bias, double F1, double F2, int *cnt_bias, int *cnts_bias, SurvList F)
{
  const int n = *nent;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int p = g.p[e], s = g.s[e];
    const ProfDev &P = profs[p];
    const float usc = g.a[e];
    const float fsc = do_bias ? filtersc[e] : sd.null1[s];
    const double pv = gumbel_surv((double)bits(usc, fsc), (double)P.evparam[0], (double)P.evparam[1]);
    if (do_bias && pv > F1) continue;
    atomicAdd(cnt_bias + p, 1);
    if (cnts_bias) atomicAdd(cnts_bias + s, 1);
    if (pv > F2) {
      const double pvv = gumbel_surv((double)bits(vfsc[e], fsc), (double)P.evparam[2], (double)P.evparam[3]);
      if (pvv > F2) continue;
    }
    surv_append(F, p, s, fsc, 0.f);
  }
}

__global__ void vit_post_kernel(const ProfDev *profs, const Grouped g, const int32_t *nent, const float *vfsc, double F2, SurvList F)
{
  const int n = *nent;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int p = g.p[e];
    const ProfDev &P = profs[p];
    const double pv = gumbel_surv((double)bits(vfsc[e], g.a[e]), (double)P.evparam[2], (double)P.evparam[3]);
    if (pv > F2) continue;
    surv_append(F, p, g.s[e], g.a[e], 0.f);
  }
}

__global__ void fwd_post_kernel(const ProfDev *profs, const Grouped g, const int32_t *nent, const float *fwdsc, const int32_t *fst,
                                double F3, SurvList D, int *nerr, const int64_t *xoff, int64_t *dxoff)
{
  const int n = *nent;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int p = g.p[e], s = g.s[e];
    const ProfDev &P = profs[p];
    if (fst[e] != B2H_OK) { atomicAdd(nerr, 1); continue; }
    const double pv = exp_surv((double)bits(fwdsc[e], g.a[e]), (double)P.evparam[4], (double)P.evparam[5]);
    if (pv > F3) continue;
    const int slot = atomicAdd(D.n, 1);                  // surv_append, plus where this comparison's Forward specials were stored
    if (slot < D.cap) {
      D.p[slot] = p; D.s[slot] = s; D.a[slot] = fwdsc[e]; D.b[slot] = g.a[e]; dxoff[slot] = xoff ? xoff[e] : -1;
      atomicAdd(D.cnt + p, 1); if (D.cnts) atomicAdd(D.cnts + s, 1);
    }
  }
}

// Row offsets of the Forward special-state rows of a grouped list: exclusive prefix sum of (L + 1) over the entries, -1 for the
// entries that would not fit in <cap_rows> (their survivors take the Forward pass again on the survivor lane).  One CTA.
__global__ void __launch_bounds__(1024) xoff_scan_kernel(const int32_t *ent_s, const int32_t *nent, const int32_t *len, int64_t cap_rows, int64_t *xoff)
{
  __shared__ long long s_part[1024];
  const int n = *nent, t = threadIdx.x;
  const int per = (n + 1023) / 1024, b = min(n, t * per), e = min(n, b + per);
  long long sum = 0;
  for (int i = b; i < e; i++) sum += len[ent_s[i]] + 1;
  s_part[t] = sum;
  __syncthreads();
  if (t == 0) { long long a = 0; for (int i = 0; i < 1024; i++) { const long long v = s_part[i]; s_part[i] = a; a += v; } }
  __syncthreads();
  long long a = s_part[t];
  for (int i = b; i < e; i++) { const long long r = len[ent_s[i]] + 1; xoff[i] = (a + r <= cap_rows) ? a : -1; a += r; }
}

// Copy the special-state rows of the F3 survivors out of the cascade's Forward buffer into the compact layout the Backward
// pass and the host expect: rows [src[e], src[e] + n[e]) -> [dst[e], ...), 6 floats per row.  One CTA per entry (grid-stride).
__global__ void gather_rows_kernel(const float *src, float *dst, const int64_t *src_off, const int64_t *dst_off, const int32_t *ent_s,
                                   const int32_t *len, int n)
{
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const int64_t words = (int64_t)(len[ent_s[e]] + 1) * 6;
    const float *a = src + src_off[e] * 6; float *b = dst + dst_off[e] * 6;
    for (int64_t i = threadIdx.x; i < words; i += blockDim.x) b[i] = a[i];
  }
}

// A pool of device buffers for one batch of the cascade
struct Pool {                                        // allocations and frees are ordered on the lane current at construction
  b2h_ctx *ctx; cudaStream_t stream; std::vector<void *> ptrs;
  explicit Pool(b2h_ctx *c) : ctx(c), stream(c->stream) {}
  template <typename T> int get(T **out, size_t n) {
    void *p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T), stream);
    if (e != cudaSuccess) { ctx->err = std::string("cudaMallocAsync: ") + cudaGetErrorString(e); return B2H_EMEM; }
    ptrs.push_back(p); *out = (T *)p; return B2H_OK;
  }
  ~Pool() { for (void *p : ptrs) cudaFreeAsync(p, stream); }
};
// page-locked host array from the context's pool
template <typename T> struct Pinned {
  b2h_ctx *ctx = nullptr; T *p = nullptr; size_t n = 0;
  Pinned() {}
  Pinned(const Pinned &) = delete; Pinned &operator=(const Pinned &) = delete;
  int alloc(b2h_ctx *c, size_t count) { release(); ctx = c; n = count; p = (T *)b2h_pin_get(c, std::max<size_t>(count, 1) * sizeof(T)); if (!p) { c->err = "cudaHostAlloc failed"; return B2H_EMEM; } return B2H_OK; }
  void release() { if (p) b2h_pin_put(ctx, p); p = nullptr; n = 0; }
  ~Pinned() { release(); }
  T *data() { return p; } const T *data() const { return p; }
  T &operator[](size_t i) { return p[i]; } const T &operator[](size_t i) const { return p[i]; }
};

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

#define TRY(x) do { int st_ = (x); if (st_ != B2H_OK) return st_; } while (0)

int make_list(Pool &pool, SurvList &l, size_t cap, int P, int *ctr, int *cnt, bool with_b)
{
  TRY(pool.get(&l.p, cap)); TRY(pool.get(&l.s, cap)); TRY(pool.get(&l.a, cap));
  l.b = nullptr;
  if (with_b) TRY(pool.get(&l.b, cap));
  l.n = ctr; l.cnt = cnt; l.cap = (int)cap;
  return B2H_OK;
}

} // namespace

// One batch (wave) of profiles [p0, p1) against the whole database.  cascade_enqueue queues the whole cascade on the
// current lane and returns at once; cascade_collect waits for that wave only (an event), reads list D and the
// counters.  Between the two the caller queues the NEXT wave, so the GPU never idles while the host digests a wave.
struct CascadeWave {
  std::unique_ptr<Pool> pool;
  int p0 = 0, P = 0;
  std::vector<int> perm;
  Pinned<int> hctr; Pinned<ProfDev> hprof; Pinned<int32_t> hcls;
  SurvList D; int64_t *D_xoff = nullptr;
  b2h_ctx *ctx = nullptr;
  float *d_fx = nullptr; cudaStream_t fx_stream = nullptr;     // Forward special-state rows of every entry of list F (a context buffer: the survivor lane reads them after this wave is gone)
  cudaEvent_t done = nullptr, ssv_done = nullptr, bias_fork = nullptr, bias_join = nullptr;
  ~CascadeWave() { for (cudaEvent_t e : {done, ssv_done, bias_fork, bias_join}) if (e) cudaEventDestroy(e); if (d_fx) b2h_bigbuf_put(ctx, d_fx, fx_stream); }
};

static int cascade_enqueue(b2h_ctx *ctx, const b2h_profile *const *profiles, int p0, int p1, const b2h_seqdb *db,
                           const b2h_search_params *prm, CascadeWave &cw, int *seqcnt /* device [4][N] per-sequence pass counts, or NULL */)
{
  const int P = p1 - p0, N = (int)db->n;
  const size_t cap = (size_t)P * N;
  const SeqDev sd = b2h_seqdev(db);
  B2H_CUDA(cudaSetDevice(ctx->device));
  cw.pool.reset(new Pool(ctx));
  Pool &pool = *cw.pool;
  cw.p0 = p0; cw.P = P;
  // profile descriptors + register-tile classes (staged page-locked and kept with the wave: the uploads must be truly
  // asynchronous, this wave is queued while the previous one is still running)
  TRY(cw.hprof.alloc(ctx, P)); TRY(cw.hcls.alloc(ctx, P));
  ProfDev *hprof = cw.hprof.data();
  std::map<int, std::vector<int32_t>> classes;
  std::vector<int> &perm = cw.perm; perm.resize(P); std::vector<int> mpads(P);                  // batch-local profile order: ascending model size
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return profiles[p0 + a]->M < profiles[p0 + b]->M; });   // (Mpad and the SSV tile are monotone in M)
  int max_Mpad = 0;
  // Scan orientation (few, possibly long sequences: hmmscan): every comparison is a latency chain, so the SSV / MSV kernels
  // take the profiles' wide tiles (fewest registers per lane = shortest row); the search orientation takes the tiles with
  // the fewest shared-memory wavefronts per cell.
  static const int scan_max = getenv("B2H_SCAN_MAXSEQ") ? atoi(getenv("B2H_SCAN_MAXSEQ")) : 2048;
  const bool scan_mode = (int)db->n <= scan_max;
  auto tile_of = [&](const ProfDev &pd) { return scan_mode ? pd.Gw * 64 + pd.NRw : pd.G * 64 + pd.NR; };
  for (int i = 0; i < P; i++) {
    hprof[i] = b2h_profdev(profiles[p0 + perm[i]]);
    mpads[i] = hprof[i].Mpad;
    classes[tile_of(hprof[i])].push_back(i);
    max_Mpad = std::max(max_Mpad, hprof[i].Mpad);
  }
  ProfDev *d_prof; TRY(pool.get(&d_prof, P));
  B2H_CUDA(cudaMemcpyAsync(d_prof, hprof, P * sizeof(ProfDev), cudaMemcpyHostToDevice, ctx->stream));
  int32_t *hcls = cw.hcls.data(); int ncls_tot = 0;
  std::vector<std::pair<int, std::pair<int, int>>> cls_ranges;   // SSV tile G*64+NR -> (offset, count)
  for (auto &kv : classes) { cls_ranges.push_back({kv.first, {ncls_tot, (int)kv.second.size()}}); for (int32_t v : kv.second) hcls[ncls_tot++] = v; }
  int32_t *d_cls; TRY(pool.get(&d_cls, P));
  B2H_CUDA(cudaMemcpyAsync(d_cls, hcls, P * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));

  // counters: [0..5] list sizes n(A,R,V,F,D,err), then 6 per-profile count arrays (A,R,V,F,D,bias) + fill
  int *d_ctr; TRY(pool.get(&d_ctr, 8 + (size_t)7 * P));
  B2H_CUDA(cudaMemsetAsync(d_ctr, 0, (8 + (size_t)7 * P) * sizeof(int), ctx->stream));
  int *cntA = d_ctr + 8, *cntR = cntA + P, *cntV = cntR + P, *cntF = cntV + P, *cntD = cntF + P, *cntB = cntD + P, *fill = cntB + P;
  SurvList A, R, V, F, D;
  TRY(make_list(pool, A, cap, P, d_ctr + 0, cntA, false));
  TRY(make_list(pool, R, cap, P, d_ctr + 1, cntR, false));
  TRY(make_list(pool, V, cap, P, d_ctr + 2, cntV, false));
  TRY(make_list(pool, F, cap, P, d_ctr + 3, cntF, false));
  TRY(make_list(pool, D, cap, P, d_ctr + 4, cntD, true));
  if (seqcnt) { A.cnts = seqcnt; F.cnts = seqcnt + 2 * (size_t)N; D.cnts = seqcnt + 3 * (size_t)N; }
  Grouped G;
  TRY(pool.get(&G.p, cap)); TRY(pool.get(&G.s, cap)); TRY(pool.get(&G.a, cap)); G.b = nullptr;
  TRY(pool.get(&G.poff, (size_t)P + 1)); TRY(pool.get(&G.itemoff, (size_t)P + 1)); G.fill = fill;
  float *stage_sc; int32_t *stage_st; TRY(pool.get(&stage_sc, cap)); TRY(pool.get(&stage_st, cap));
  WorkList wl; wl.profs = d_prof; wl.ent_s = G.s; wl.poff = G.poff; wl.itemoff = G.itemoff; wl.P = P; wl.counter = ctx->d_counters + 8; wl.plo = 0; wl.phi = P;
  const int32_t *nent = G.poff + P;                  // number of entries of the current grouped list (device)
  const int pgrid = ctx->sm_count * 8;

  // 1. SSV over every comparison
  // Scan orientation (few, possibly long sequences: hmmscan): one comparison per profile would keep one lane group of a
  // CTA busy and walk the whole sequence serially.  SSV is exactly decomposable along the sequence -- an ungapped diagonal
  // spans at most M rows -- so every sequence is taken as overlapping chunks (b2h_chunkview), the kernel folds the chunk
  // maxima into raw[profile][sequence] and a finish kernel applies p7_SSVFilter's post-processing and the F1 test.
  if (scan_mode) {
    StageTimer tm(ctx, 0);
    int *raw; TRY(pool.get(&raw, cap));
    B2H_CUDA(cudaMemsetAsync(raw, 0, cap * sizeof(int), ctx->stream));
    {
      ForkJoin fj(ctx);
      int ssv_cls = 0;
      for (auto &cr : cls_ranges) {
        const int Gc = cr.first / 64, NRc = cr.first % 64, NG = 32 / Gc;
        const int O = ((2 * Gc * NRc + 127) / 128) * 128;
        auto nchunks = [&](int S) { long long c = 0; for (int s = 0; s < N; s++) c += (db->h_len[s] + S - 1) / S; return c; };
        int S = 4 * O;                                        // 1.25x the cells; finer cuts (2x, 1x the overlap) when the job is small
        if ((long long)cr.second.second * nchunks(S) < 16384) S = 2 * O;
        if ((long long)cr.second.second * nchunks(S) < 16384) S = O;
        const b2h_chunkview *vp = b2h_seqdb_chunk_view(db, O, S, ctx->stream);
        if (!vp) return B2H_EMEM;
        const b2h_chunkview v = *vp;
        if (v.n == 0) continue;
        SsvArgs a;
        a.profs = d_prof; a.cls = d_cls + cr.second.first; a.ncls = cr.second.second;
        a.sd = sd; a.sd.off = v.d_off; a.sd.len = v.d_len; a.sd.order = v.d_order; a.sd.n = v.n;
        a.chunks = (v.n + B2H_SSV_CHUNK - 1) / B2H_SSV_CHUNK; a.counter = ctx->d_counters + 32 + (ssv_cls++ % 24); a.mode = 3;
        a.items_per_cta = 2;
        a.out_sc = nullptr; a.out_status = nullptr; a.A = A; a.R = R; a.F1 = prm->F1;
        a.parent = v.d_parent; a.raw = raw; a.raw_stride = N; a.wide = 1;
        a.threads = 32 * std::max(1, std::min(8, (std::min(v.n, B2H_SSV_CHUNK) + NG - 1) / NG));
        TRY(b2h_launch_ssv(ctx, Gc, NRc, a, fj.next()));
      }
    }
    TRY(b2h_launch_ssv_finish(ctx, d_prof, P, sd, raw, A, R, prm->F1));
  } else
  { StageTimer tm(ctx, 0);
  ForkJoin fj(ctx);
  int ssv_cls = 0;
  for (auto &cr : cls_ranges) {
    SsvArgs a;
    a.profs = d_prof; a.cls = d_cls + cr.second.first; a.ncls = cr.second.second; a.sd = sd;
    a.chunks = (N + B2H_SSV_CHUNK - 1) / B2H_SSV_CHUNK; a.counter = ctx->d_counters + 32 + (ssv_cls++ % 24); a.mode = 2;
    static const int ipc = getenv("B2H_SSV_ITEMS_PER_CTA") ? atoi(getenv("B2H_SSV_ITEMS_PER_CTA")) : 2;   // measured (ms/step): persistent 42.0, 8 -> 43.2, 4 -> 41.3, 2 -> 40.4, 1 -> 40.4
    a.items_per_cta = ipc;
    a.out_sc = nullptr; a.out_status = nullptr; a.A = A; a.R = R; a.F1 = prm->F1;
    TRY(b2h_launch_ssv(ctx, cr.first / 64, cr.first % 64, a, fj.next()));
  } }
  // The tail of the cascade can move to the high-priority POST lane so that the SSV launches of the next wave (queued on
  // the main lane right behind) run next to it.  B2H_OVERLAP = 1: everything after SSV, 2: Viterbi + Forward, 3: Forward
  // only, 0: nothing.  Measured on B200 (ms/step), round 1: 0 -> 44.9, 1 -> 48.0, 2 -> 46.4, 3 -> 43.4: the ALU-bound stages
  // lose more from sharing the SMs with SSV than the overlap wins; the latency-bound Forward pass hides for free.  Round 2,
  // with the leaner packed Viterbi row (ALU pipe 61-74 % instead of 90 %): 100 x 50 000: 3 -> 38.8, 2 -> 37.9 (the stages still
  // trade time, SSV 24.5 -> 31 ms, but the sum comes out a millisecond ahead); 20 000 x 100 000: 3 -> 10.4 s, 2 -> 12.6 s.
  // 3 stays the default: it never loses.
  static const int overlap = getenv("B2H_OVERLAP") ? atoi(getenv("B2H_OVERLAP")) : 3;
  std::unique_ptr<b2h_lane_switch> post;
  auto to_post_lane = [&]() -> int {
    B2H_CUDA(cudaEventCreateWithFlags(&cw.ssv_done, cudaEventDisableTiming));
    B2H_CUDA(cudaEventRecord(cw.ssv_done, ctx->stream));
    post.reset(new b2h_lane_switch(ctx, B2H_LANE_POST));
    B2H_CUDA(cudaStreamWaitEvent(ctx->stream, cw.ssv_done, 0));
    wl.counter = ctx->d_counters + 8;
    return B2H_OK;
  };
  if (overlap == 1) TRY(to_post_lane());
  wl.counter = ctx->d_counters + 8;
  // 2. full MSV for the comparisons SSV could not decide
  {
    Grouped GR = G; GR.a = nullptr;
    { StageTimer tg(ctx, 6); TRY(b2h_launch_group(ctx, R, P, GR)); }
    StageTimer tm(ctx, 1);
    static const bool smem_msv = getenv("B2H_MSV_SMEM") != nullptr;      // debugging aid: the shared-memory MSV kernel
    if (smem_msv) TRY(b2h_launch_msv(ctx, wl, sd, max_Mpad, 0, 2, nullptr, nullptr, A, prm->F1));
    else { std::vector<int> tiles(P); for (int i = 0; i < P; i++) tiles[i] = tile_of(hprof[i]);
           TRY(b2h_launch_msv_tiled(ctx, wl, sd, tiles, scan_mode ? (2 | 8) : 2, nullptr, nullptr, A, prm->F1)); }
  }
  // 3 + 4. bias filter and ViterbiFilter on the MSV survivors.  The bias filter is a serial chain of divides per
  // residue, one thread per comparison: latency bound, a few hundred microseconds whatever the list size.  Instead of
  // waiting for it, the Viterbi launches take the whole MSV-survivor list (6 % more comparisons than the bias-filtered
  // one) and run next to it; both P-value tests are applied afterwards, in the reference's order.  B2H_SERIAL_BIAS=1
  // restores the serial form (bias -> regroup -> Viterbi).
  static const bool serial_bias = getenv("B2H_SERIAL_BIAS") != nullptr;
  { StageTimer tg(ctx, 6); TRY(b2h_launch_group(ctx, A, P, G)); }
  if (serial_bias) {
    { StageTimer tm(ctx, 2);
      if (prm->do_biasfilter) TRY(b2h_launch_bias(ctx, wl, sd, ctx->sm_count * 128 * 8, stage_sc));
      bias_post_kernel<<<pgrid, 256, 0, ctx->stream>>>(d_prof, sd, G, nent, stage_sc, prm->do_biasfilter, prm->F1, prm->F2, cntB, seqcnt ? seqcnt + N : nullptr, V, F);
      ctx->launches++; }
    if (overlap == 2) TRY(to_post_lane());
    { StageTimer tg(ctx, 6); TRY(b2h_launch_group(ctx, V, P, G)); }
    { StageTimer tm(ctx, 3);
      StageOut so; so.sc = stage_sc; so.status = stage_st; so.fwd_xmx = so.bck_xmx = nullptr; so.xoff = nullptr;
      TRY(b2h_launch_viterbi(ctx, wl, sd, mpads, 0, so));
      vit_post_kernel<<<pgrid, 256, 0, ctx->stream>>>(d_prof, G, nent, stage_sc, prm->F2, F);
      ctx->launches++; }
  } else {
    float *bias_sc; TRY(pool.get(&bias_sc, cap));
    if (overlap == 2) TRY(to_post_lane());
    StageTimer tm(ctx, 3);
    if (prm->do_biasfilter) {
      B2H_CUDA(cudaEventCreateWithFlags(&cw.bias_fork, cudaEventDisableTiming));
      B2H_CUDA(cudaEventCreateWithFlags(&cw.bias_join, cudaEventDisableTiming));
      B2H_CUDA(cudaEventRecord(cw.bias_fork, ctx->stream));
      cudaStream_t keep = ctx->stream;
      ctx->stream = ctx->bias_stream;                              // (b2h_launch_bias launches on the current stream)
      cudaStreamWaitEvent(ctx->stream, cw.bias_fork, 0);
      const int rc = b2h_launch_bias(ctx, wl, sd, ctx->sm_count * 128 * 8, bias_sc);
      cudaEventRecord(cw.bias_join, ctx->stream);
      ctx->stream = keep;
      if (rc != B2H_OK) return rc;
    }
    StageOut so; so.sc = stage_sc; so.status = stage_st; so.fwd_xmx = so.bck_xmx = nullptr; so.xoff = nullptr;
    TRY(b2h_launch_viterbi(ctx, wl, sd, mpads, 0, so));
    if (prm->do_biasfilter) B2H_CUDA(cudaStreamWaitEvent(ctx->stream, cw.bias_join, 0));
    bias_vit_post_kernel<<<pgrid, 256, 0, ctx->stream>>>(d_prof, sd, G, nent, bias_sc, stage_sc, prm->do_biasfilter, prm->F1, prm->F2, cntB, seqcnt ? seqcnt + N : nullptr, F);
    ctx->launches++;
  }
  // 5. Forward parser
  if (overlap == 3) TRY(to_post_lane());
  { StageTimer tg(ctx, 6); TRY(b2h_launch_group(ctx, F, P, G)); }
  { StageTimer tm(ctx, 4);
    // The Forward stage keeps the special-state rows of every comparison it scores (a prefix sum over the grouped list gives
    // the row offsets), so that the F3 survivors need only the Backward pass afterwards (p7_pipeline.c:764-773 runs Forward
    // once, too).  The buffer is sized on the host for the common case; entries beyond it are marked and redone later.
    static const long long fx_budget = getenv("B2H_FX_ROWS") ? atoll(getenv("B2H_FX_ROWS")) : ((long long)16 << 20);   // rows of 6 floats (384 MB)
    const long long cap_rows = std::min<long long>(fx_budget, (long long)P * ((long long)db->nres + N));
    int64_t *d_xoffF = nullptr, *d_Dx = nullptr;
    TRY(pool.get(&d_xoffF, cap)); TRY(pool.get(&d_Dx, cap));
    StageOut so; so.sc = stage_sc; so.status = stage_st; so.fwd_xmx = so.bck_xmx = nullptr; so.xoff = nullptr;
    if (cap_rows > 0 && !getenv("B2H_FWD_TWICE")) {
      cw.ctx = ctx;
      cw.d_fx = (float *)b2h_bigbuf_get(ctx, (size_t)cap_rows * 6 * sizeof(float), ctx->stream);
      if (cw.d_fx) {
        cw.fx_stream = ctx->stream;
        xoff_scan_kernel<<<1, 1024, 0, ctx->stream>>>(G.s, nent, sd.len, (int64_t)cap_rows, d_xoffF);
        ctx->launches++;
        so.fwd_xmx = cw.d_fx; so.xoff = d_xoffF;
      }
    }
    TRY(b2h_launch_forward(ctx, wl, sd, mpads, 0, so));
    fwd_post_kernel<<<pgrid, 256, 0, ctx->stream>>>(d_prof, G, nent, stage_sc, stage_st, prm->F3, D, d_ctr + 5, so.xoff, d_Dx);
    ctx->launches++;
    cw.D_xoff = d_Dx; }
  B2H_CUDA(cudaGetLastError());

  // 6. the counters come home (page-locked: the copy must not block the host, the next wave is queued behind it)
  TRY(cw.hctr.alloc(ctx, 8 + (size_t)7 * P));
  B2H_CUDA(cudaMemcpyAsync(cw.hctr.data(), d_ctr, (8 + (size_t)7 * P) * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  B2H_CUDA(cudaEventCreateWithFlags(&cw.done, cudaEventDisableTiming));
  B2H_CUDA(cudaEventRecord(cw.done, ctx->stream));
  cw.D = D;
  return B2H_OK;
}

// Wait for the wave, then fetch list D over <copy_stream> (the wave's own lane may already hold the next wave's kernels).
static int cascade_collect(b2h_ctx *ctx, CascadeWave &cw, cudaStream_t copy_stream, std::vector<b2h_survivor> &outD, int64_t *counters /*[P][4], global index*/)
{
  B2H_CUDA(cudaEventSynchronize(cw.done));
  b2h_resolve_timers(ctx);
  const int P = cw.P, p0 = cw.p0;
  const int *hctr = cw.hctr.data();
  if (hctr[5] > 0) { ctx->err = "numerical overflow in the Forward parser"; return B2H_ERANGE; }
  const int nD = hctr[4];
  for (int i = 0; i < P; i++) {
    int64_t *c = counters + (size_t)(p0 + cw.perm[i]) * 4;
    c[0] = hctr[8 + i]; c[1] = hctr[8 + 5 * P + i]; c[2] = hctr[8 + 3 * P + i]; c[3] = hctr[8 + 4 * P + i];
  }
  if (nD) {
    Pinned<int32_t> h; TRY(h.alloc(ctx, (size_t)4 * nD));
    Pinned<int64_t> hx; TRY(hx.alloc(ctx, (size_t)nD));
    const SurvList &D = cw.D;
    B2H_CUDA(cudaMemcpyAsync(hx.data(), cw.D_xoff, nD * sizeof(int64_t), cudaMemcpyDeviceToHost, copy_stream));
    B2H_CUDA(cudaMemcpyAsync(h.data(), D.p, nD * sizeof(int32_t), cudaMemcpyDeviceToHost, copy_stream));
    B2H_CUDA(cudaMemcpyAsync(h.data() + nD, D.s, nD * sizeof(int32_t), cudaMemcpyDeviceToHost, copy_stream));
    B2H_CUDA(cudaMemcpyAsync(h.data() + 2 * (size_t)nD, D.a, nD * sizeof(float), cudaMemcpyDeviceToHost, copy_stream));
    B2H_CUDA(cudaMemcpyAsync(h.data() + 3 * (size_t)nD, D.b, nD * sizeof(float), cudaMemcpyDeviceToHost, copy_stream));
    B2H_CUDA(cudaStreamSynchronize(copy_stream));
    outD.reserve(outD.size() + nD);
    const float *fa = reinterpret_cast<const float *>(h.data() + 2 * (size_t)nD), *fb = reinterpret_cast<const float *>(h.data() + 3 * (size_t)nD);
    for (int i = 0; i < nD; i++) { b2h_survivor v; v.profile = p0 + cw.perm[h[i]]; v.seq = h[(size_t)nD + i]; v.fwdsc = fa[i]; v.filtersc = fb[i]; v.fxoff = cw.d_fx ? hx[i] : -1; outD.push_back(v); }
  }
  cw.pool.reset();                                    // stream-ordered frees on the wave's lane, after everything queued there so far
  return B2H_OK;
}

// Forward + Backward parsers with stored special-state rows for the F3 survivors (sorted by profile),
// chunked so that the specials of one chunk stay within a fixed budget; then host domain definition.
// The GPU backend of the domain definition's phase B: every envelope of a chunk of survivors through
// efwd/ebck/eoa kernels (b2h_envelope.cu) on the context's envelope stream, results back as b2h_env_job fields.
// Called from the domain-definition thread while the main thread feeds the cascade of the next wave.
struct EnvGpu : b2h_env_backend {
  b2h_ctx *ctx; const b2h_seqdb *db; double ms = 0.0; size_t nenv = 0;
  EnvGpu(b2h_ctx *c, const b2h_seqdb *d) : ctx(c), db(d) {}
  template <typename T> int dalloc(std::vector<void *> &keep, T **out, size_t n) {
    void *p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T), ctx->env_stream);
    if (e != cudaSuccess) { ctx->err = std::string("cudaMallocAsync (envelopes): ") + cudaGetErrorString(e); return B2H_EMEM; }
    keep.push_back(p); *out = (T *)p; return B2H_OK;
  }
  int run(const std::vector<b2h_ddef_task> &tasks, std::vector<b2h_env_job> &jobs) override {
    const double t0 = now_ms();
    B2H_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->env_stream;
    const SeqDev sd = b2h_seqdev(db);
    // order: by size class, inside a class as given (largest first)
    const size_t n = jobs.size();
    std::vector<int> cls(n), perm;
    for (size_t q = 0; q < n; q++) {
      const int M = tasks[jobs[q].task].prof->M; int c = 0;
      while (c < B2H_N_ENV_CLASSES && M > B2H_ENV_CLASSES[c].bound) c++;
      cls[q] = c;                                            // == B2H_N_ENV_CLASSES: longer than any class (cannot happen: SSV tiles stop at 3071)
    }
    // bytes of matrix scratch per pass: 6 GB, or -- when a pass would not hold the job (long-target envelopes of kilobase
    // models are ~50 MB each) -- up to a third of what the device has free, so that few, well-filled passes are launched
    size_t BUDGET = (size_t)6 << 30;
    {
      size_t want = 0;
      for (size_t q = 0; q < n; q++) {
        const int c = std::min(cls[q], B2H_N_ENV_CLASSES - 1);
        want += (size_t)(jobs[q].j - jobs[q].i + 2) * (size_t)32 * B2H_ENV_CLASSES[c].C * B2H_ENV_CLASSES[c].W * (8 * sizeof(float) + 1);
      }
      size_t free_b = 0, total_b = 0;
      if (want > BUDGET && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) BUDGET = std::max(BUDGET, std::min(want, free_b / 3));
    }
    bool fwd_only = n > 0;
    for (size_t q = 0; q < n; q++) fwd_only = fwd_only && jobs[q].fwd_only;
    std::vector<char> done(n, 0);
    for (;;) {
      // take envelopes class by class until the scratch budget is full
      perm.clear();
      std::vector<int> cls_lo(B2H_N_ENV_CLASSES + 1, 0);
      size_t bytes = 0; bool full = false;
      for (int c = 0; c < B2H_N_ENV_CLASSES && !full; c++) {
        cls_lo[c] = (int)perm.size();
        const size_t Mp = (size_t)32 * B2H_ENV_CLASSES[c].C * B2H_ENV_CLASSES[c].W;
        for (size_t q = 0; q < n; q++) {
          if (done[q] || cls[q] != c) continue;
          const size_t need = (size_t)(jobs[q].j - jobs[q].i + 2) * Mp * (8 * sizeof(float) + 1);
          if (!perm.empty() && bytes + need > BUDGET) { full = true; break; }
          bytes += need; perm.push_back((int)q); done[q] = 1;
        }
        cls_lo[c + 1] = (int)perm.size();
      }
      for (int c = 0; c < B2H_N_ENV_CLASSES; c++) cls_lo[c + 1] = std::max(cls_lo[c + 1], cls_lo[c]);
      const int m = (int)perm.size();
      if (m == 0) break;
      // per-envelope arrays
      std::vector<ProfDev> hprof; std::vector<const b2h_profile *> seen;
      std::vector<int32_t> h_prof(m), h_seq(m), h_i0(m), h_Ld(m), h_tcap(m);
      std::vector<float> h_pmove(m);
      std::vector<int64_t> h_moff(m), h_moffn(m), h_xoff(m), h_toff(m), h_rscoff(m, -1);
      int64_t moff = 0, moffn = 0, xoff = 0, toff = 0, rscoff = 0;
      bool any_rsc = false;
      for (int z = 0; z < m; z++) {
        const b2h_env_job &jb = jobs[perm[z]];
        const b2h_ddef_task &t = tasks[jb.task];
        size_t pi = 0; while (pi < seen.size() && seen[pi] != t.prof) pi++;
        if (pi == seen.size()) { seen.push_back(t.prof); hprof.push_back(b2h_profdev(t.prof)); }
        const int c = cls[perm[z]];
        const int64_t Mp = (int64_t)32 * B2H_ENV_CLASSES[c].C * B2H_ENV_CLASSES[c].W;
        const int Ld = jb.j - jb.i + 1;
        h_prof[z] = (int32_t)pi; h_seq[z] = t.surv.seq; h_i0[z] = jb.i; h_Ld[z] = Ld;
        h_pmove[z] = (2.0f + 0.0f) / ((float)(jb.cfg_len > 0 ? jb.cfg_len : t.L) + 2.0f + 0.0f);     // configure(m, false, L)
        if (jb.rsc) { h_rscoff[z] = rscoff; rscoff += (int64_t)t.prof->Kp * t.prof->Mpad; any_rsc = true; }
        h_tcap[z] = Ld + t.prof->M + 16;
        h_moff[z] = moff; h_moffn[z] = moffn; h_xoff[z] = xoff; h_toff[z] = toff;
        moff += (int64_t)(Ld + 1) * Mp; moffn += Mp; xoff += Ld + 1; toff += h_tcap[z];
      }
      std::vector<void *> keep;
      EnvDev ev;
      ProfDev *d_prof = nullptr; int32_t *d_pi = nullptr, *d_seq = nullptr, *d_i0 = nullptr, *d_Ld = nullptr, *d_tcap = nullptr; float *d_pmove = nullptr;
      int64_t *d_moff = nullptr, *d_moffn = nullptr, *d_xoff = nullptr, *d_toff = nullptr;
      int rc = B2H_OK;
      auto A = [&](int r) { if (rc == B2H_OK) rc = r; };
      A(dalloc(keep, &d_prof, hprof.size())); A(dalloc(keep, &d_pi, m)); A(dalloc(keep, &d_seq, m)); A(dalloc(keep, &d_i0, m)); A(dalloc(keep, &d_Ld, m));
      A(dalloc(keep, &d_tcap, m)); A(dalloc(keep, &d_pmove, m)); A(dalloc(keep, &d_moff, m)); A(dalloc(keep, &d_moffn, m)); A(dalloc(keep, &d_xoff, m)); A(dalloc(keep, &d_toff, m));
      A(dalloc(keep, &ev.F, (size_t)moff * 3)); A(dalloc(keep, &ev.PP, (size_t)moff * 2)); A(dalloc(keep, &ev.OA, (size_t)moff * 3)); A(dalloc(keep, &ev.BP, (size_t)moff));
      A(dalloc(keep, &ev.fx, (size_t)xoff * 6)); A(dalloc(keep, &ev.bx, (size_t)xoff * 6)); A(dalloc(keep, &ev.ox, (size_t)xoff * 6));
      A(dalloc(keep, &ev.envsc, m)); A(dalloc(keep, &ev.oasc, m)); A(dalloc(keep, &ev.em, (size_t)moffn)); A(dalloc(keep, &ev.ei, (size_t)moffn));
      A(dalloc(keep, &ev.xnull, (size_t)m * 4)); A(dalloc(keep, &ev.status, m)); A(dalloc(keep, &ev.tlen, m)); A(dalloc(keep, &ev.trace, (size_t)toff));
      float *d_rsc = nullptr; int64_t *d_rscoff = nullptr;
      std::vector<float> h_rsc;
      if (any_rsc) {                                           // the envelopes' own emission odds, padded to the device row length
        A(dalloc(keep, &d_rsc, (size_t)rscoff)); A(dalloc(keep, &d_rscoff, m));
        h_rsc.assign((size_t)rscoff, 0.0f);
        for (int z = 0; z < m; z++) {
          const b2h_env_job &jb = jobs[perm[z]];
          if (!jb.rsc) continue;
          const b2h_profile *pf = tasks[jb.task].prof;
          for (int x = 0; x < pf->Kp; x++) memcpy(h_rsc.data() + h_rscoff[z] + (size_t)x * pf->Mpad, jb.rsc + (size_t)x * pf->M, (size_t)pf->M * sizeof(float));
        }
      }
      auto release = [&]() { for (void *p : keep) cudaFreeAsync(p, st); };
      if (rc != B2H_OK) { release(); return rc; }
#define H2D(dst, src) cudaMemcpyAsync(dst, (src).data(), (src).size() * sizeof((src)[0]), cudaMemcpyHostToDevice, st)
      H2D(d_prof, hprof); H2D(d_pi, h_prof); H2D(d_seq, h_seq); H2D(d_i0, h_i0); H2D(d_Ld, h_Ld); H2D(d_tcap, h_tcap); H2D(d_pmove, h_pmove);
      H2D(d_moff, h_moff); H2D(d_moffn, h_moffn); H2D(d_xoff, h_xoff); H2D(d_toff, h_toff);
      if (any_rsc) { H2D(d_rsc, h_rsc); H2D(d_rscoff, h_rscoff); ev.rsc_pool = d_rsc; ev.rsc_off = d_rscoff; }
#undef H2D
      cudaMemsetAsync(ev.tlen, 0xff, (size_t)m * sizeof(int32_t), st);
      ev.profs = d_prof; ev.prof = d_pi; ev.seq = d_seq; ev.i0 = d_i0; ev.Ld = d_Ld; ev.pmove = d_pmove;
      ev.moff = d_moff; ev.moff_n = d_moffn; ev.xoff = d_xoff; ev.toff = d_toff; ev.tcap = d_tcap; ev.counter = ctx->d_env_counter;
      // every size class on its own stream (forked from / joined to the envelope stream): a class launch is as long as
      // its longest envelope and fills only a few SMs, so the classes overlap
      cudaEventRecord(ctx->env_fork, st);
      for (int c = 0; c < B2H_N_ENV_CLASSES && rc == B2H_OK; c++) {
        if (cls_lo[c + 1] <= cls_lo[c]) continue;
        cudaStream_t sc = ctx->env_side[c];
        cudaStreamWaitEvent(sc, ctx->env_fork, 0);
        ev.e_lo = cls_lo[c]; ev.e_hi = cls_lo[c + 1]; ev.counter = ctx->d_env_counter + c;
        for (int kind = 0; kind < (fwd_only ? 1 : 3) && rc == B2H_OK; kind++) rc = b2h_launch_envelope(ctx, kind, B2H_ENV_CLASSES[c].C, B2H_ENV_CLASSES[c].W, ev, sd, sc);
        cudaEventRecord(ctx->env_join[c], sc);
        cudaStreamWaitEvent(st, ctx->env_join[c], 0);
      }
      if (rc != B2H_OK) { cudaStreamSynchronize(st); release(); return rc; }
      std::vector<float> r_envsc(m), r_oasc(m), r_xnull((size_t)m * 4), r_em((size_t)moffn), r_ei((size_t)moffn);
      std::vector<int32_t> r_status(m), r_tlen(m), r_trace((size_t)toff * 4);
#define D2H(dst, src) cudaMemcpyAsync((dst).data(), src, (dst).size() * sizeof((dst)[0]), cudaMemcpyDeviceToHost, st)
      D2H(r_envsc, ev.envsc); D2H(r_oasc, ev.oasc); D2H(r_xnull, ev.xnull); D2H(r_em, ev.em); D2H(r_ei, ev.ei);
      D2H(r_status, ev.status); D2H(r_tlen, ev.tlen); D2H(r_trace, ev.trace);
#undef D2H
      cudaError_t ce = cudaStreamSynchronize(st);
      release();
      if (ce != cudaSuccess) { ctx->err = std::string("envelope kernels: ") + cudaGetErrorString(ce); return B2H_ECUDA; }
      for (int z = 0; z < m; z++) {
        b2h_env_job &jb = jobs[perm[z]];
        const int M = tasks[jb.task].prof->M;
        const int base = r_status[z] & 0xff;                                             // a Forward range error only makes envsc = inf
        if (fwd_only) { jb.status = (base == B2H_OK) ? 0 : 1; jb.envsc = r_envsc[z]; continue; }
        if ((base != B2H_OK && base != B2H_ERANGE) || (r_status[z] & 0x300) || r_tlen[z] < 0) { jb.status = 1; continue; }
        jb.status = 0; jb.envsc = r_envsc[z]; jb.oasc = r_oasc[z];
        jb.xn = r_xnull[(size_t)z * 4 + 0]; jb.xc = r_xnull[(size_t)z * 4 + 1]; jb.xj = r_xnull[(size_t)z * 4 + 2];
        jb.em.assign(r_em.begin() + h_moffn[z], r_em.begin() + h_moffn[z] + M);
        jb.ei.assign(r_ei.begin() + h_moffn[z], r_ei.begin() + h_moffn[z] + M);
        jb.trace.assign(r_trace.begin() + h_toff[z] * 4, r_trace.begin() + (h_toff[z] + r_tlen[z]) * 4);
      }
      nenv += (size_t)m;
    }
    ms += now_ms() - t0;
    return B2H_OK;
  }
};

// One chunk of survivors whose parser specials are on the host.  The domain definition of the chunks runs on the host
// thread pool, driven by ONE background thread that takes the chunks in order, while the calling thread goes on feeding
// the GPU (it never waits for a domain definition before the very end).
struct DdefJob {
  Pinned<float> fx, bx; Pinned<int32_t> bst; std::vector<b2h_ddef_task> tasks;   // page-locked: the D2H copies run at PCIe speed
  b2h_results *res = nullptr;          // where this chunk's hits go: the result object of its wave
  int wave = -1; bool last = false;    // last chunk of its wave (a wave without survivors sends an empty marker job)
};
struct DdefQueue {
  b2h_ddef_pool &pool; const b2h_search_params *prm; b2h_env_backend *backend;
  std::function<void(int)> wave_done;  // called on the domain-definition thread when a wave's last chunk is through
  std::thread th; std::mutex mu; std::condition_variable cv; std::deque<std::unique_ptr<DdefJob>> q;
  bool closing = false; int status = B2H_OK; double total_ms = 0.0;
  DdefQueue(b2h_ddef_pool &p, const b2h_search_params *pr, b2h_env_backend *be, std::function<void(int)> done)
    : pool(p), prm(pr), backend(be), wave_done(std::move(done)) {
    th = std::thread([this]() {
      for (;;) {
        std::unique_ptr<DdefJob> job;
        { std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return closing || !q.empty(); });
          if (q.empty()) return;
          job = std::move(q.front()); q.pop_front(); }
        const double t0 = now_ms();
        int st = status;                                                                        // after a failure the rest is dropped
        if (st == B2H_OK && !job->tasks.empty()) st = pool.run(job->tasks, prm, job->res, backend);
        const double ms = now_ms() - t0;
        { std::lock_guard<std::mutex> lk(mu); if (status == B2H_OK) status = st; total_ms += ms; }
        if (getenv("B2H_TRACE") && !job->tasks.empty()) fprintf(stderr, "[b2h_search]   chunk of %zu survivors: host domain definition %.1f ms on %d threads\n", job->tasks.size(), ms, pool.nthreads);
        if (job->last && st == B2H_OK && wave_done) wave_done(job->wave);
      }
    });
  }
  void push(std::unique_ptr<DdefJob> job) { { std::lock_guard<std::mutex> lk(mu); q.push_back(std::move(job)); } cv.notify_one(); }
  int finish() {                                               // drain the queue, stop the thread
    if (th.joinable()) { { std::lock_guard<std::mutex> lk(mu); closing = true; } cv.notify_one(); th.join(); }
    return status;
  }
  ~DdefQueue() { finish(); }
};

// The survivors of one wave on their way through Forward (with stored special-state rows) and Backward: sorted by
// profile, cut into chunks whose specials stay within a fixed budget, every chunk queued on the current (survivor) lane
// with its D2H copies behind it; <done> fires when everything has arrived in the page-locked job buffers.
struct SurvChunk {
  std::unique_ptr<Pool> pool; std::unique_ptr<DdefJob> job;
  std::vector<ProfDev> hprof; std::vector<int32_t> poff, itemoff, ent_s; std::vector<int64_t> xoff, srcoff;   // H2D sources: alive until <done>
  size_t i0 = 0; int n = 0;
};
struct SurvPending {
  std::vector<b2h_survivor> surv; std::vector<std::unique_ptr<SurvChunk>> chunks; cudaEvent_t done = nullptr; size_t wave = 0;
  b2h_ctx *ctx = nullptr;
  float *fx_src = nullptr; cudaStream_t fx_stream = nullptr;    // the cascade's Forward special-state rows (list F) of this wave
  ~SurvPending() { if (done) cudaEventDestroy(done); if (fx_src) b2h_bigbuf_put(ctx, fx_src, fx_stream); }
};

static int survivors_enqueue(b2h_ctx *ctx, const b2h_profile *const *profiles, const b2h_seqdb *db, SurvPending &sp)
{
  std::vector<b2h_survivor> &surv = sp.surv;
  std::sort(surv.begin(), surv.end(), [&](const b2h_survivor &x, const b2h_survivor &y) {
    const int mx = profiles[x.profile]->Mpad, my = profiles[y.profile]->Mpad;
    if (mx != my) return mx < my;
    return x.profile != y.profile ? x.profile < y.profile : x.seq < y.seq; });
  const SeqDev sd = b2h_seqdev(db);
  const size_t ROW_BUDGET = (size_t)32 << 20;        // rows of 6 floats per chunk (x2 matrices = 1.5 GB)
  size_t i0 = 0;
  while (i0 < surv.size()) {
    size_t i1 = i0, rows = 0;
    while (i1 < surv.size() && (i1 == i0 || rows + db->h_len[surv[i1].seq] + 1 <= ROW_BUDGET)) { rows += db->h_len[surv[i1].seq] + 1; i1++; }
    const int n = (int)(i1 - i0);
    sp.chunks.emplace_back(new SurvChunk());
    SurvChunk &ck = *sp.chunks.back();
    ck.i0 = i0; ck.n = n;
    // work list of this chunk: profiles present, in order
    std::vector<ProfDev> &hprof = ck.hprof; std::vector<int32_t> &poff = ck.poff, &itemoff = ck.itemoff, &ent_s = ck.ent_s; std::vector<int64_t> &xoff = ck.xoff;
    ent_s.resize(n); xoff.resize(n); ck.srcoff.resize(n);
    std::vector<int> mpads;
    int64_t acc = 0; int items = 0;
    bool have_fwd = sp.fx_src != nullptr;                  // every entry's Forward rows were stored by the cascade?
    for (int e = 0; e < n; e++) {
      const b2h_survivor &v = surv[i0 + e];
      ck.srcoff[e] = v.fxoff; if (v.fxoff < 0) have_fwd = false;
      if (e == 0 || v.profile != surv[i0 + e - 1].profile) {
        if (e) items += ((e - poff.back()) + B2H_ITEM_ENTRIES - 1) / B2H_ITEM_ENTRIES;
        hprof.push_back(b2h_profdev(profiles[v.profile])); poff.push_back(e); itemoff.push_back(items); mpads.push_back(hprof.back().Mpad);
      }
      ent_s[e] = v.seq; xoff[e] = acc; acc += db->h_len[v.seq] + 1;
    }
    items += ((n - poff.back()) + B2H_ITEM_ENTRIES - 1) / B2H_ITEM_ENTRIES;
    poff.push_back(n); itemoff.push_back(items);
    const int Pc = (int)hprof.size();
    ck.pool.reset(new Pool(ctx));
    Pool &pool = *ck.pool;
    ProfDev *d_prof; int32_t *d_poff, *d_itemoff, *d_ent; int64_t *d_xoff; float *d_fx, *d_bx, *d_fsc, *d_bsc; int32_t *d_fst, *d_bst;
    TRY(pool.get(&d_prof, Pc)); TRY(pool.get(&d_poff, Pc + 1)); TRY(pool.get(&d_itemoff, Pc + 1)); TRY(pool.get(&d_ent, n));
    TRY(pool.get(&d_xoff, n)); TRY(pool.get(&d_fx, (size_t)acc * 6)); TRY(pool.get(&d_bx, (size_t)acc * 6));
    TRY(pool.get(&d_fsc, n)); TRY(pool.get(&d_bsc, n)); TRY(pool.get(&d_fst, n)); TRY(pool.get(&d_bst, n));
    B2H_CUDA(cudaMemcpyAsync(d_prof, hprof.data(), Pc * sizeof(ProfDev), cudaMemcpyHostToDevice, ctx->stream));
    B2H_CUDA(cudaMemcpyAsync(d_poff, poff.data(), (Pc + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    B2H_CUDA(cudaMemcpyAsync(d_itemoff, itemoff.data(), (Pc + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    B2H_CUDA(cudaMemcpyAsync(d_ent, ent_s.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    B2H_CUDA(cudaMemcpyAsync(d_xoff, xoff.data(), n * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    WorkList wl; wl.profs = d_prof; wl.ent_s = d_ent; wl.poff = d_poff; wl.itemoff = d_itemoff; wl.P = Pc; wl.counter = ctx->d_counters + 8; wl.plo = 0; wl.phi = Pc;
    StageOut sf; sf.sc = d_fsc; sf.status = d_fst; sf.fwd_xmx = d_fx; sf.bck_xmx = nullptr; sf.xoff = d_xoff;
    StageOut sb; sb.sc = d_bsc; sb.status = d_bst; sb.fwd_xmx = d_fx; sb.bck_xmx = d_bx; sb.xoff = d_xoff;
    if (have_fwd) {                                        // the rows are there: compact them, then Backward only
      StageTimer tm(ctx, 5);
      int64_t *d_srcoff; TRY(pool.get(&d_srcoff, n));
      B2H_CUDA(cudaMemcpyAsync(d_srcoff, ck.srcoff.data(), n * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
      gather_rows_kernel<<<std::min(n, ctx->sm_count * 8), 256, 0, ctx->stream>>>(sp.fx_src, d_fx, d_srcoff, d_xoff, d_ent, sd.len, n);
      ctx->launches++;
      B2H_CUDA(cudaGetLastError());
      B2H_CUDA(cudaMemsetAsync(d_bst, 0, n * sizeof(int32_t), ctx->stream));
      TRY(b2h_launch_backward(ctx, wl, sd, mpads, items, sb));
    } else { StageTimer tm(ctx, 5); TRY(b2h_launch_forward_backward(ctx, wl, sd, mpads, items, sf, sb)); }
    ck.job.reset(new DdefJob());
    Pinned<float> &fx = ck.job->fx, &bx = ck.job->bx; Pinned<int32_t> &bst = ck.job->bst;
    TRY(fx.alloc(ctx, (size_t)acc * 6)); TRY(bx.alloc(ctx, (size_t)acc * 6)); TRY(bst.alloc(ctx, n));
    B2H_CUDA(cudaMemcpyAsync(fx.data(), d_fx, (size_t)acc * 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    B2H_CUDA(cudaMemcpyAsync(bx.data(), d_bx, (size_t)acc * 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    B2H_CUDA(cudaMemcpyAsync(bst.data(), d_bst, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    i0 = i1;
  }
  if (sp.fx_src) { b2h_bigbuf_put(ctx, sp.fx_src, ctx->stream); sp.fx_src = nullptr; }   // (its last users are the gathers queued above)
  B2H_CUDA(cudaEventCreateWithFlags(&sp.done, cudaEventDisableTiming));
  B2H_CUDA(cudaEventRecord(sp.done, ctx->stream));
  return B2H_OK;
}

// <done> has fired: hand the chunks to the domain-definition thread, one task per survivor
static int survivors_complete(b2h_ctx *ctx, const b2h_profile *const *profiles, const b2h_seqdb *db, SurvPending &sp, DdefQueue &ddef, b2h_results *wres)
{
  b2h_resolve_timers(ctx);
  if (sp.chunks.empty()) {                                       // nothing survived: the wave is complete as soon as its turn comes
    std::unique_ptr<DdefJob> marker(new DdefJob());
    marker->res = wres; marker->wave = (int)sp.wave; marker->last = true;
    ddef.push(std::move(marker));
    return B2H_OK;
  }
  for (auto &ckp : sp.chunks) {
    SurvChunk &ck = *ckp;
    DdefJob &job = *ck.job;
    job.res = wres; job.wave = (int)sp.wave; job.last = (&ckp == &sp.chunks.back());
    job.tasks.resize(ck.n);
    for (int e = 0; e < ck.n; e++) {
      b2h_ddef_task &t = job.tasks[e];
      const b2h_survivor &v = sp.surv[ck.i0 + e];
      t.surv = v; t.prof = profiles[v.profile];
      t.dsq = db->h_res + db->h_off[v.seq]; t.L = db->h_len[v.seq];
      t.fx = job.fx.data() + (size_t)ck.xoff[e] * 6; t.bx = job.bx.data() + (size_t)ck.xoff[e] * 6;
      t.bck_own_scales = (job.bst[e] & 0x100) != 0;
    }
    ck.pool.reset();                                           // stream-ordered frees on the chunk's lane
    ddef.push(std::move(ck.job));
  }
  return B2H_OK;
}

// The plan of one search: profiles longest first, cut into waves (see search_impl).
struct WavePlan { std::vector<int> order; std::vector<size_t> bounds; };

static WavePlan plan_waves(const b2h_profile *const *profiles, size_t P, const b2h_seqdb *db)
{
  // Profiles are processed longest first, in waves of about equal DP volume: while the host threads define the
  // domains of one wave's survivors the GPU already runs the cascade of the next wave, and the wave whose host
  // work cannot be hidden (the last) holds the shortest models.
  WavePlan pl;
  const size_t N = db->n;
  pl.order.resize(P);
  std::iota(pl.order.begin(), pl.order.end(), 0);
  pl.bounds.push_back(0);
  if (N == 0 || P == 0) return pl;
  std::stable_sort(pl.order.begin(), pl.order.end(), [&](int a, int b) { return profiles[a]->M > profiles[b]->M; });
  double cells = 0.0;
  for (size_t i = 0; i < P; i++) cells += profiles[i]->M;
  const size_t CAP = (size_t)1 << 25;                 // comparisons per batch: every list is sized for the worst case
  const size_t pb = std::max<size_t>(1, CAP / N);
  // Waves shrink geometrically: the host work of every wave but the last hides behind the next wave's cascade, so
  // the last wave -- whose survivor passes, envelope kernels and host domain definition are exposed -- is the smallest.
  // Waves only pay when a wave's cascade is long enough to hide the previous wave's host work behind it; a small job
  // (hmmscan of one query: latency-bound launches of a few ms whatever their size) runs as one wave.
  const double job_cells = cells * (double)db->nres;
  int nwaves = (P >= 16 && job_cells >= 1.5e11) ? 4 : (P >= 6 && job_cells >= 4e10) ? 2 : 1;   // measured on B200 (100 profiles x 50k sequences, 3.7e11 cells, ms/step): 3 waves 44.1, 4 waves 42.6, 5 waves 45.4
  double ratio = 0.6;
  if (const char *ev = getenv("B2H_WAVES")) nwaves = std::max(1, atoi(ev));
  if (const char *ev = getenv("B2H_WAVE_RATIO")) ratio = std::min(1.0, std::max(0.05, atof(ev)));
  std::vector<double> cum(nwaves + 1, 0.0);            // cumulative share of the DP cells after each wave
  { double wsum = 0.0, wgt = 1.0; for (int i = 0; i < nwaves; i++) { wsum += wgt; cum[i + 1] = wsum; wgt *= ratio; }
    for (int i = 0; i <= nwaves; i++) cum[i] /= wsum; }
  double acc = 0.0; size_t start = 0; int slot = 1;
  for (size_t i = 0; i < P; i++) {
    acc += profiles[pl.order[i]]->M;
    if (i + 1 == P || i + 1 - start >= pb || acc >= cells * cum[slot]) {
      pl.bounds.push_back(i + 1); start = i + 1;
      while (slot < nwaves && acc >= cells * cum[slot]) slot++;
    }
  }
  return pl;
}

// Receives the result object of every wave, in wave order, as soon as the wave's last survivor is through domain
// definition (on the domain-definition thread).  Hits carry the caller's profile indices, sorted by (profile, target);
// counters are [P][4] with the rows of the wave's profiles filled; profiles = the wave's profile indices.
struct WaveSink { virtual void wave(std::unique_ptr<b2h_results> r) = 0; virtual ~WaveSink() {} };

static int search_impl(b2h_ctx *ctx, const b2h_profile *const *profiles, size_t P, const b2h_seqdb *db,
                       const b2h_search_params *prm, const WavePlan &plan, WaveSink &sink, std::vector<int64_t> *seq_counters)
{
  struct ExitTrace { double t; ~ExitTrace() { if (getenv("B2H_TRACE")) fprintf(stderr, "[b2h_search] returned after %.1f ms\n", now_ms() - t); } } exit_trace{now_ms()};
  const size_t N = db->n;
  if (N == 0 || P == 0) return B2H_OK;
  {
    const double t0 = now_ms();
    const std::vector<int> &order = plan.order;
    const std::vector<size_t> &bounds = plan.bounds;
    std::vector<const b2h_profile *> sp(P);
    for (size_t i = 0; i < P; i++) sp[i] = profiles[order[i]];
    std::vector<int64_t> scnt(P * 4, 0);
    b2h_ddef_pool ddpool(prm->host_threads);
    EnvGpu envgpu(ctx, db);
    const bool host_env = getenv("B2H_ENVELOPES_ON_HOST") != nullptr;     // debugging aid: rescore envelopes with the host code
    const size_t nw = bounds.size() - 1;
    std::vector<std::unique_ptr<b2h_results>> wres(nw);
    for (auto &r : wres) r.reset(new b2h_results());
    DdefQueue ddef(ddpool, prm, host_env ? nullptr : &envgpu, [&](int w) {
      // the wave is final: hits in the caller's profile numbering, ordered by (profile, target), with its counters
      std::unique_ptr<b2h_results> r = std::move(wres[w]);
      for (b2h_hit &h : r->hits) h.profile = order[h.profile];
      std::stable_sort(r->hits.begin(), r->hits.end(), [](const b2h_hit &x, const b2h_hit &y) {
        return x.profile != y.profile ? x.profile < y.profile : x.seq < y.seq; });
      r->counters.assign(P * 4, 0);
      r->profiles.reserve(bounds[w + 1] - bounds[w]);
      for (size_t i = bounds[w]; i < bounds[w + 1]; i++) {
        r->profiles.push_back(order[i]);
        for (int c = 0; c < 4; c++) r->counters[(size_t)order[i] * 4 + c] = scnt[i * 4 + c];
      }
      sink.wave(std::move(r));
    });
    size_t nsurv = 0;
    int *d_seqcnt = nullptr;                             // per-sequence pass counters (hmmscan of several queries)
    struct SeqCntGuard { int *&p; ~SeqCntGuard() { if (p) cudaFree(p); } } seqcnt_guard{d_seqcnt};
    if (prm->seq_counters) {
      B2H_CUDA(cudaSetDevice(ctx->device));
      B2H_CUDA(cudaMalloc(&d_seqcnt, 4 * N * sizeof(int)));
      B2H_CUDA(cudaMemsetAsync(d_seqcnt, 0, 4 * N * sizeof(int), ctx->stream));
    }
    // Software pipeline over the waves, driven by events.  The cascade of waves w+1, w+2 is queued on the main lane before
    // the host looks at wave w.  When a wave's cascade has finished, its survivor list is fetched and its Forward/Backward
    // passes are queued on one of two high-priority survivor lanes (their kernels slip in between the following wave's SSV
    // launches, and the passes of two consecutive waves may run side by side); when those have arrived on the host, the
    // chunk goes to the domain-definition thread.  This thread never blocks on either while the other can make progress.
    std::vector<std::unique_ptr<CascadeWave>> waves(nw);
    int st = B2H_OK;
    size_t ahead = 2;                                    // waves queued beyond the one being collected
    if (const char *ev = getenv("B2H_AHEAD")) ahead = (size_t)std::max(1, atoi(ev));
    size_t queued = 0;
    auto top_up = [&](size_t upto) {                     // keep the lanes fed: queue waves [queued, upto]
      for (; queued <= upto && queued < nw && st == B2H_OK; queued++) {
        waves[queued].reset(new CascadeWave());
        st = cascade_enqueue(ctx, sp.data(), (int)bounds[queued], (int)bounds[queued + 1], db, prm, *waves[queued], d_seqcnt);
      }
    };
    auto fired = [](cudaEvent_t e) { const cudaError_t q = cudaEventQuery(e); if (q != cudaSuccess) (void)cudaGetLastError(); return q != cudaErrorNotReady; };
    const bool trace = getenv("B2H_TRACE") != nullptr;
    std::deque<std::unique_ptr<SurvPending>> pend;
    size_t next_collect = 0;
    while (st == B2H_OK && (next_collect < nw || !pend.empty())) {
      if (next_collect < nw) top_up(next_collect + ahead);
      if (st != B2H_OK) break;
      if (!pend.empty() && fired(pend.front()->done)) {
        SurvPending &p = *pend.front();
        if (trace) fprintf(stderr, "[b2h_search]   wave %zu: %zu survivors through Forward/Backward at +%.1f ms\n", p.wave, p.surv.size(), now_ms() - t0);
        st = survivors_complete(ctx, sp.data(), db, p, ddef, wres[p.wave].get());
        nsurv += p.surv.size();
        pend.pop_front();
        continue;
      }
      if (next_collect < nw && pend.size() < 2 && fired(waves[next_collect]->done)) {
        const size_t w = next_collect++;
        std::unique_ptr<SurvPending> p(new SurvPending());
        p->wave = w;
        const int lane_id = (w & 1) ? B2H_LANE_SURV2 : B2H_LANE_SURV;
        st = cascade_collect(ctx, *waves[w], ctx->lanes[lane_id].stream, p->surv, scnt.data());
        p->ctx = ctx; p->fx_src = waves[w]->d_fx; p->fx_stream = ctx->lanes[lane_id].stream; waves[w]->d_fx = nullptr;
        waves[w].reset();
        if (trace) fprintf(stderr, "[b2h_search]   wave %zu (%zu profiles): cascade collected at +%.1f ms (queued up to wave %zu)\n", w, bounds[w + 1] - bounds[w], now_ms() - t0, queued - 1);
        if (st == B2H_OK) { b2h_lane_switch lane(ctx, lane_id); st = survivors_enqueue(ctx, sp.data(), db, *p); }
        pend.push_back(std::move(p));
        continue;
      }
      // nothing is ready: block on the only thing in flight, or yield briefly when there are two
      if (pend.empty()) { if (cudaEventSynchronize(waves[next_collect]->done) != cudaSuccess) { ctx->err = "cascade failed"; st = B2H_ECUDA; } }
      else if (next_collect >= nw || pend.size() >= 2) { if (cudaEventSynchronize(pend.front()->done) != cudaSuccess) { ctx->err = "survivor passes failed"; st = B2H_ECUDA; } }
      else std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    const double tg = now_ms() - t0;
    if (st != B2H_OK) { cudaStreamSynchronize(ctx->stream); for (auto &l : ctx->lanes) cudaStreamSynchronize(l.stream); ddef.finish(); pend.clear(); waves.clear(); return st; }
    const double t1 = now_ms();
    st = ddef.finish();
    if (st != B2H_OK) { ctx->err = "domain definition failed"; return st; }
    if (d_seqcnt && seq_counters) {                     // every wave has been collected: the counts are final
      std::vector<int> h(4 * N);
      if (cudaMemcpy(h.data(), d_seqcnt, 4 * N * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) { ctx->err = "per-sequence counters"; return B2H_ECUDA; }
      seq_counters->resize(4 * N);
      for (size_t s = 0; s < N; s++) for (int c = 0; c < 4; c++) (*seq_counters)[s * 4 + c] = h[(size_t)c * N + s];
    }
    if (getenv("B2H_TRACE")) fprintf(stderr, "[b2h_search] %zu waves: GPU cascade + survivor parsers %.1f ms, %zu survivors, host domain definition %.1f ms in total (%.1f ms not hidden; %zu envelopes on the GPU, %.1f ms), all %.1f ms\n",
                                     bounds.size() - 1, tg, nsurv, ddef.total_ms, now_ms() - t1, envgpu.nenv, envgpu.ms, now_ms() - t0);
  }
  return B2H_OK;
}

// A search whose waves are handed out as they complete (b2h_search_begin / _next / _end): the pipeline above runs on a
// driver thread of its own, the caller -- a Python thread assembling `TopHits`, an all-gather per wave -- works on wave w
// while the GPU is busy with the waves behind it.  Only the last (smallest) wave's post-processing stays exposed.
struct b2h_search_job : WaveSink {
  b2h_ctx *ctx = nullptr; std::vector<const b2h_profile *> profiles; const b2h_seqdb *db = nullptr; b2h_search_params prm;
  WavePlan plan; std::thread th; std::mutex mu; std::condition_variable cv;
  std::deque<std::unique_ptr<b2h_results>> ready; bool finished = false; int status = B2H_OK; size_t handed = 0;
  void wave(std::unique_ptr<b2h_results> r) override { { std::lock_guard<std::mutex> lk(mu); ready.push_back(std::move(r)); } cv.notify_all(); }
};

extern "C" {

int b2h_search(b2h_ctx *ctx, const b2h_profile *const *profiles, size_t P, const b2h_seqdb *db,
               const b2h_search_params *prm, b2h_results **out)
{
  if (!ctx || !profiles || !db || !prm || !out || db->ctx != ctx) return B2H_EINVAL;
  for (size_t i = 0; i < P; i++) if (!profiles[i] || profiles[i]->ctx != ctx) return B2H_EINVAL;
  *out = nullptr;
  struct Collect : WaveSink {
    std::vector<std::unique_ptr<b2h_results>> parts;
    void wave(std::unique_ptr<b2h_results> r) override { parts.push_back(std::move(r)); }
  } sink;
  std::unique_ptr<b2h_results> res(new b2h_results());
  res->counters.assign(P * 4, 0);
  const WavePlan plan = plan_waves(profiles, P, db);
  const int st = search_impl(ctx, profiles, P, db, prm, plan, sink, &res->seq_counters);
  if (st != B2H_OK) return st;
  // one result: the waves' records back to back (every profile belongs to one wave), ordered by (profile, target)
  for (auto &part : sink.parts) {
    const int64_t dbase = (int64_t)res->doms.size(), tbase = (int64_t)res->text.size();
    for (b2h_hit &h : part->hits) { h.dom_offset += dbase; res->hits.push_back(h); }
    for (b2h_domain &d : part->doms) { d.text_offset += tbase; res->doms.push_back(d); }
    res->text.insert(res->text.end(), part->text.begin(), part->text.end());
    for (int32_t p : part->profiles) for (int c = 0; c < 4; c++) res->counters[(size_t)p * 4 + c] = part->counters[(size_t)p * 4 + c];
  }
  std::stable_sort(res->hits.begin(), res->hits.end(), [](const b2h_hit &x, const b2h_hit &y) {
    return x.profile != y.profile ? x.profile < y.profile : x.seq < y.seq; });
  *out = res.release();
  return B2H_OK;
}

int b2h_search_begin(b2h_ctx *ctx, const b2h_profile *const *profiles, size_t P, const b2h_seqdb *db,
                     const b2h_search_params *prm, b2h_search_job **out, size_t *nwaves)
{
  if (!ctx || !profiles || !db || !prm || !out || db->ctx != ctx || prm->seq_counters) return B2H_EINVAL;
  for (size_t i = 0; i < P; i++) if (!profiles[i] || profiles[i]->ctx != ctx) return B2H_EINVAL;
  b2h_search_job *job = new b2h_search_job();
  job->ctx = ctx; job->profiles.assign(profiles, profiles + P); job->db = db; job->prm = *prm;
  job->plan = plan_waves(job->profiles.data(), P, db);
  if (nwaves) *nwaves = job->plan.bounds.size() - 1;
  job->th = std::thread([job]() {
    cudaSetDevice(job->ctx->device);
    int st;
    try { st = search_impl(job->ctx, job->profiles.data(), job->profiles.size(), job->db, &job->prm, job->plan, *job, nullptr); }
    catch (const std::bad_alloc &) { job->ctx->err = "out of host memory"; st = B2H_EMEM; }      // (nothing may leave a thread of the library)
    catch (const std::exception &e) { job->ctx->err = e.what(); st = B2H_EINVAL; }
    { std::lock_guard<std::mutex> lk(job->mu); job->status = st; job->finished = true; }
    job->cv.notify_all();
  });
  *out = job;
  return B2H_OK;
}

// Blocks until the next wave is complete: *out = its results (the caller destroys them), or NULL after the last wave.
int b2h_search_next(b2h_search_job *job, b2h_results **out)
{
  if (!job || !out) return B2H_EINVAL;
  *out = nullptr;
  std::unique_lock<std::mutex> lk(job->mu);
  job->cv.wait(lk, [&] { return !job->ready.empty() || job->finished; });
  if (!job->ready.empty()) { *out = job->ready.front().release(); job->ready.pop_front(); job->handed++; return B2H_OK; }
  return job->status;
}

// Waits for the driver thread and frees the job (results not fetched are dropped).  Returns the status of the search.
int b2h_search_end(b2h_search_job *job)
{
  if (!job) return B2H_EINVAL;
  if (job->th.joinable()) job->th.join();
  const int st = job->status;
  delete job;
  return st;
}

const int32_t *b2h_results_profiles(const b2h_results *r, size_t *n) { if (n) *n = r ? r->profiles.size() : 0; return r ? r->profiles.data() : nullptr; }

int b2h_debug_domaindef(const b2h_profile *p, const uint8_t *dsq, int L, const float *fwd_xmx, const float *bck_xmx,
                        float fwdsc, const b2h_search_params *prm, b2h_results **out)
{
  if (!p || !dsq || !fwd_xmx || !bck_xmx || !prm || !out) return B2H_EINVAL;
  b2h_results *res = new b2h_results();
  res->counters.assign(4, 0);
  std::vector<b2h_ddef_task> tasks(1);
  tasks[0].surv.profile = 0; tasks[0].surv.seq = 0; tasks[0].surv.fwdsc = fwdsc; tasks[0].surv.filtersc = 0.f;
  tasks[0].prof = p; tasks[0].dsq = dsq; tasks[0].L = L; tasks[0].fx = fwd_xmx; tasks[0].bx = bck_xmx; tasks[0].bck_own_scales = false;
  b2h_ddef_pool pool(1);
  int st = pool.run(tasks, prm, res);
  if (st != B2H_OK) { delete res; return st; }
  *out = res;
  return B2H_OK;
}

int b2h_longtarget_domains(const b2h_profile *p, const b2h_lt_window *windows, size_t n, const b2h_search_params *prm, b2h_results **out)
{
  if (!p || (!windows && n) || !prm || !out || p->max_length <= 0) return B2H_EINVAL;
  for (size_t i = 0; i < n; i++) if (!windows[i].dsq || windows[i].L < 1 || !windows[i].fwd_xmx || !windows[i].bck_xmx) return B2H_EINVAL;
  b2h_results *res = new b2h_results();
  res->counters.assign(4, 0);
  b2h_ddef_pool pool(prm->host_threads);
  // the batched driver the GPU path uses, without a backend: every envelope is rescored by the host code
  std::vector<int32_t> idx(n);
  for (size_t i = 0; i < n; i++) idx[i] = (int32_t)i;
  const int st = getenv("B2H_LT_WINDOW_BY_WINDOW") ? b2h_longtarget_domains_host(p, windows, n, prm, pool.nthreads, res)
                                                   : b2h_longtarget_domains_backend(p, windows, idx.data(), n, prm, pool.nthreads, nullptr, res);
  if (st != B2H_OK) { delete res; return st; }
  *out = res;
  return B2H_OK;
}

int b2h_longtarget_hits(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *windows, const int64_t *window_start,
                        const int64_t *seq_start, const int32_t *complement, const int32_t *seq,
                        const b2h_search_params *prm, b2h_results **out)
{
  if (!ctx || !p || !windows || p->ctx != ctx || windows->ctx != ctx || !prm || !out || p->max_length <= 0) return B2H_EINVAL;
  const size_t n = windows->n;
  if (n && (!window_start || !seq_start || !complement || !seq)) return B2H_EINVAL;
  *out = nullptr;
  b2h_results *res = new b2h_results();
  res->counters.assign(4, 0);
  struct Guard { b2h_results *r; ~Guard() { delete r; } } guard{res};
  B2H_CUDA(cudaSetDevice(ctx->device));
  b2h_ddef_pool pool(prm->host_threads);
  const SeqDev sd = b2h_seqdev(windows);
  const ProfDev hprof = b2h_profdev(p);
  const std::vector<int> mpads(1, p->Mpad);
  const size_t ROW_BUDGET = (size_t)8 << 20;           // rows of 6 floats per batch, two matrices
  for (size_t i0 = 0; i0 < n;) {
    size_t i1 = i0, rows = 0;
    while (i1 < n && (i1 == i0 || rows + windows->h_len[i1] + 1 <= ROW_BUDGET)) { rows += windows->h_len[i1] + 1; i1++; }
    const int nb = (int)(i1 - i0);
    std::vector<int32_t> ent(nb); std::vector<int64_t> xoff(nb);
    int64_t acc = 0;
    for (int e = 0; e < nb; e++) { ent[e] = (int32_t)(i0 + e); xoff[e] = acc; acc += windows->h_len[i0 + e] + 1; }
    const int items = (nb + B2H_ITEM_ENTRIES - 1) / B2H_ITEM_ENTRIES;
    const int32_t poff[2] = {0, nb}, itemoff[2] = {0, items};
    std::vector<float> fx((size_t)acc * 6), bx((size_t)acc * 6);
    std::vector<int32_t> bst(nb);
    {
      Pool dev(ctx);
      ProfDev *d_prof; int32_t *d_poff, *d_itemoff, *d_ent; int64_t *d_xoff; float *d_fx, *d_bx, *d_fsc, *d_bsc; int32_t *d_fst, *d_bst;
      TRY(dev.get(&d_prof, 1)); TRY(dev.get(&d_poff, 2)); TRY(dev.get(&d_itemoff, 2)); TRY(dev.get(&d_ent, nb)); TRY(dev.get(&d_xoff, nb));
      TRY(dev.get(&d_fx, (size_t)acc * 6)); TRY(dev.get(&d_bx, (size_t)acc * 6));
      TRY(dev.get(&d_fsc, nb)); TRY(dev.get(&d_bsc, nb)); TRY(dev.get(&d_fst, nb)); TRY(dev.get(&d_bst, nb));
      B2H_CUDA(cudaMemcpyAsync(d_prof, &hprof, sizeof(ProfDev), cudaMemcpyHostToDevice, ctx->stream));
      B2H_CUDA(cudaMemcpyAsync(d_poff, poff, sizeof poff, cudaMemcpyHostToDevice, ctx->stream));
      B2H_CUDA(cudaMemcpyAsync(d_itemoff, itemoff, sizeof itemoff, cudaMemcpyHostToDevice, ctx->stream));
      B2H_CUDA(cudaMemcpyAsync(d_ent, ent.data(), nb * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
      B2H_CUDA(cudaMemcpyAsync(d_xoff, xoff.data(), nb * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
      WorkList wl; wl.profs = d_prof; wl.ent_s = d_ent; wl.poff = d_poff; wl.itemoff = d_itemoff; wl.P = 1; wl.counter = ctx->d_counters + 8; wl.plo = 0; wl.phi = 1;
      StageOut sf; sf.sc = d_fsc; sf.status = d_fst; sf.fwd_xmx = d_fx; sf.bck_xmx = nullptr; sf.xoff = d_xoff;
      StageOut sb; sb.sc = d_bsc; sb.status = d_bst; sb.fwd_xmx = d_fx; sb.bck_xmx = d_bx; sb.xoff = d_xoff;
      const double t_fb = now_ms();
      TRY(b2h_launch_forward_backward(ctx, wl, sd, mpads, items, sf, sb));
      B2H_CUDA(cudaMemcpyAsync(fx.data(), d_fx, (size_t)acc * 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
      B2H_CUDA(cudaMemcpyAsync(bx.data(), d_bx, (size_t)acc * 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
      B2H_CUDA(cudaMemcpyAsync(bst.data(), d_bst, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      B2H_CUDA(cudaStreamSynchronize(ctx->stream));
      if (getenv("B2H_TRACE")) fprintf(stderr, "[b2h_longtarget_hits] %d windows, %lld rows: Forward / Backward parsers + specials to the host %.2f ms\n", nb, (long long)acc, now_ms() - t_fb);
    }
    std::vector<b2h_lt_window> lw(nb);
    for (int e = 0; e < nb; e++) {
      const size_t w = i0 + e;
      b2h_lt_window &x = lw[e];
      x.dsq = windows->h_res + windows->h_off[w]; x.L = windows->h_len[w];
      x.fwd_xmx = fx.data() + (size_t)xoff[e] * 6; x.bck_xmx = bx.data() + (size_t)xoff[e] * 6;
      x.window_start = window_start[w]; x.seq_start = seq_start[w]; x.complement = complement[w]; x.seq = seq[w];
      x.bck_own_scales = (bst[e] & 0x100) != 0; x.reserved = 0;
    }
    const size_t h0 = res->hits.size();
    static const bool lt_host_env = getenv("B2H_ENVELOPES_ON_HOST") != nullptr;     // debugging aid: rescore envelopes with the host code
    if (lt_host_env) TRY(b2h_longtarget_domains_host(p, lw.data(), (size_t)nb, prm, pool.nthreads, res));
    else { EnvGpu envgpu(ctx, windows); TRY(b2h_longtarget_domains_backend(p, lw.data(), ent.data(), (size_t)nb, prm, pool.nthreads, &envgpu, res)); }
    for (size_t h = h0; h < res->hits.size(); h++) res->hits[h].profile += (int32_t)i0;      // window index in the whole list
    i0 = i1;
  }
  guard.r = nullptr;
  *out = res;
  return B2H_OK;
}

size_t            b2h_results_nhits   (const b2h_results *r) { return r ? r->hits.size() : 0; }
const b2h_hit    *b2h_results_hits    (const b2h_results *r) { return r ? r->hits.data() : nullptr; }
size_t            b2h_results_ndomains(const b2h_results *r) { return r ? r->doms.size() : 0; }
const b2h_domain *b2h_results_domains (const b2h_results *r) { return r ? r->doms.data() : nullptr; }
const char       *b2h_results_text    (const b2h_results *r, size_t *n) { if (n) *n = r ? r->text.size() : 0; return r ? r->text.data() : nullptr; }
const int64_t    *b2h_results_counters(const b2h_results *r) { return r ? r->counters.data() : nullptr; }
const int64_t    *b2h_results_seq_counters(const b2h_results *r) { return (r && !r->seq_counters.empty()) ? r->seq_counters.data() : nullptr; }
void              b2h_results_destroy (b2h_results *r) { delete r; }

} // extern "C"
