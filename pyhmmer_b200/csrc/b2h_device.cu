// b2h_device.cu -- context, sequence arena and profile upload for libb2h.so.
#include <cuda_fp16.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <tuple>
#include <numeric>
#include <thread>
#include <chrono>
#include "b2h_internal.h"

extern "C" {

int b2h_ctx_create(int device, b2h_ctx **out)
{
  if (!out) return B2H_EINVAL;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
    return B2H_ECUDA;                       // no CPU fallback, by design
  b2h_ctx *ctx = new b2h_ctx();
  ctx->device = device;
  B2H_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B2H_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) { delete ctx; return B2H_ECUDA; }   // sm_100a cubins only
  ctx->sm_count = prop.multiProcessorCount;
  B2H_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  B2H_CUDA(cudaMalloc(&ctx->d_counters, 64 * sizeof(int)));
  B2H_CUDA(cudaMalloc(&ctx->d_env_counter, 16 * sizeof(int)));
  // the survivor lane and the envelope streams outrank the cascade: their kernels are short, latency-bound and on the
  // critical path of the host-side domain definition, the cascade's persistent CTAs would otherwise starve them
  { int lo = 0, hi = 0; B2H_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi)); ctx->prio_hi = hi; }
  for (auto &l : ctx->lanes) {
    B2H_CUDA(cudaStreamCreateWithPriority(&l.stream, cudaStreamNonBlocking, ctx->prio_hi));
    B2H_CUDA(cudaMalloc(&l.counters, 64 * sizeof(int)));
  }
  B2H_CUDA(cudaStreamCreateWithPriority(&ctx->env_stream, cudaStreamNonBlocking, ctx->prio_hi));
  B2H_CUDA(cudaStreamCreateWithPriority(&ctx->bias_stream, cudaStreamNonBlocking, ctx->prio_hi));
  B2H_CUDA(cudaEventCreateWithFlags(&ctx->env_fork, cudaEventDisableTiming));
  for (int i = 0; i < 8; i++) { B2H_CUDA(cudaStreamCreateWithPriority(&ctx->env_side[i], cudaStreamNonBlocking, ctx->prio_hi)); B2H_CUDA(cudaEventCreateWithFlags(&ctx->env_join[i], cudaEventDisableTiming)); }
  {  // keep stream-ordered allocations cached across searches instead of returning them to the driver at every sync
    cudaMemPool_t pool; uint64_t keep = ~0ull;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  *out = ctx;
  return B2H_OK;
}

void b2h_ctx_destroy(b2h_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->d_counters) cudaFree(ctx->d_counters);
  if (ctx->d_env_counter) cudaFree(ctx->d_env_counter);
  if (ctx->env_stream) cudaStreamDestroy(ctx->env_stream);
  if (ctx->bias_stream) cudaStreamDestroy(ctx->bias_stream);
  if (ctx->env_fork) cudaEventDestroy(ctx->env_fork);
  for (int i = 0; i < 8; i++) { if (ctx->env_side[i]) cudaStreamDestroy(ctx->env_side[i]); if (ctx->env_join[i]) cudaEventDestroy(ctx->env_join[i]); }
  for (auto &pf : ctx->pinned_free) cudaFreeHost(pf.first);
  for (auto &pf : ctx->pin_pool) cudaFreeHost(pf.first);
  for (auto &b : ctx->bigbufs) { if (b.p) cudaFree(b.p); if (b.free_ev) cudaEventDestroy(b.free_ev); }
  for (auto &l : ctx->lanes) {
    if (l.counters) cudaFree(l.counters);
    if (l.stream) cudaStreamDestroy(l.stream);
    for (cudaStream_t q : l.side) cudaStreamDestroy(q);
    for (cudaEvent_t e : l.side_done) cudaEventDestroy(e);
    if (l.fork_ev) cudaEventDestroy(l.fork_ev);
  }
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  for (cudaStream_t q : ctx->side) cudaStreamDestroy(q);
  for (cudaEvent_t e : ctx->side_done) cudaEventDestroy(e);
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

int b2h_ctx_set_stream(b2h_ctx *ctx, void *s)
{
  if (!ctx) return B2H_EINVAL;
  ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
  return B2H_OK;
}

int b2h_ctx_synchronize(b2h_ctx *ctx)
{
  if (!ctx) return B2H_EINVAL;
  B2H_CUDA(cudaStreamSynchronize(ctx->stream));
  return B2H_OK;
}

int b2h_ctx_set_profiling(b2h_ctx *ctx, int on) { if (!ctx) return B2H_EINVAL; ctx->profiling = on; return B2H_OK; }
int b2h_ctx_stage_ms(b2h_ctx *ctx, double *ms8, int reset)
{
  if (!ctx || !ms8) return B2H_EINVAL;
  for (int i = 0; i < 8; i++) { ms8[i] = ctx->stage_ms[i]; if (reset) ctx->stage_ms[i] = 0; }
  return B2H_OK;
}
const char *b2h_ctx_last_error(const b2h_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t    b2h_ctx_launch_count(const b2h_ctx *ctx) { return ctx ? ctx->launches.load() : (uint64_t)0; }

// ------------------------------------------------------------------------------------------
// sequence arena
// ------------------------------------------------------------------------------------------
static void *pinned_get(b2h_ctx *ctx, size_t bytes, size_t *cap)
{
  size_t best = (size_t)-1;
  for (size_t i = 0; i < ctx->pinned_free.size(); i++)
    if (ctx->pinned_free[i].second >= bytes && (best == (size_t)-1 || ctx->pinned_free[i].second < ctx->pinned_free[best].second)) best = i;
  if (best != (size_t)-1) {
    void *p = ctx->pinned_free[best].first; *cap = ctx->pinned_free[best].second;
    ctx->pinned_free.erase(ctx->pinned_free.begin() + best);
    return p;
  }
  for (auto &pf : ctx->pinned_free) cudaFreeHost(pf.first);      // too small for this database: do not hoard them
  ctx->pinned_free.clear();
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  *cap = bytes;
  return p;
}

static int seqdb_build(b2h_ctx *ctx, size_t n, const int64_t *len,
                       const uint8_t *const *dsq, const uint8_t *packed, const int64_t *poff, b2h_seqdb **out)
{
  if (!ctx || !out || (n && !len && !poff)) return B2H_EINVAL;
  *out = nullptr;
  b2h_seqdb *db = new b2h_seqdb();
  db->ctx = ctx; db->n = n;
  db->h_len.resize(n); db->h_off.resize(n + 1);
  size_t total = 0;
  for (size_t s = 0; s < n; s++) {
    int64_t L = poff ? poff[s+1] - poff[s] : len[s];
    if (L < 0 || L > 100000000) { delete db; return B2H_EINVAL; }
    db->h_len[s] = (int32_t)L; db->h_off[s] = (int64_t)total;
    db->nres += L; db->maxL = std::max(db->maxL, (int)L);
    total += ((size_t)L + 15) & ~(size_t)15;
    if (L % 16 == 0) total += 16;           // always at least one pad byte after the last residue
  }
  db->h_off[n] = (int64_t)total;
  db->arena_bytes = total;

  // block layout
  const size_t n1 = std::max<size_t>(n, 1);
  size_t o = 0;
  auto sect = [&](size_t bytes) { const size_t at = o; o = (o + bytes + 255) & ~(size_t)255; return at; };
  const size_t o_res = sect(std::max<size_t>(total, 16)), o_off = sect((n + 1) * sizeof(int64_t)), o_len = sect(n1 * 4), o_ord = sect(n1 * 4),
               o_tjb = sect(n1), o_xwm = sect(n1 * 2), o_pmv = sect(n1 * 4), o_nl1 = sect(n1 * 4), o_p1 = sect(n1 * 4),
               o_fa = sect(n1 * 4), o_fb = sect(n1 * 4);
  db->block_bytes = o;
  B2H_CUDA(cudaSetDevice(ctx->device));
  db->h_block = (uint8_t *)pinned_get(ctx, o, &db->h_block_cap);
  if (!db->h_block) { delete db; ctx->err = "cudaHostAlloc failed"; return B2H_EMEM; }
  uint8_t *hb = db->h_block;
  db->h_res = hb + o_res;

  // residues: clamp codes to the 32-row tables, pad each sequence to its 16-byte boundary
  {
    auto fill = [&](size_t s0, size_t s1) {
      for (size_t s = s0; s < s1; s++) {
        const uint8_t *src = poff ? packed + poff[s] : dsq[s] + 1;      // Easel dsq is 1-based
        uint8_t *dst = hb + o_res + db->h_off[s];
        const int32_t L = db->h_len[s];
        for (int32_t i = 0; i < L; i++) dst[i] = src[i] < B2H_NCODE ? src[i] : (uint8_t)B2H_PAD_CODE;
        const int64_t end = db->h_off[s + 1] - db->h_off[s];
        for (int64_t i = L; i < end; i++) dst[i] = (uint8_t)B2H_PAD_CODE;
      }
    };
    const int T = (total > ((size_t)4 << 20)) ? b2h_rank_threads(8) : 1;
    if (T <= 1) fill(0, n);
    else {
      std::vector<std::thread> th;
      size_t s0 = 0;
      for (int t = 0; t < T; t++) {                                     // contiguous ranges of about equal bytes
        size_t s1 = s0;
        const int64_t want = (int64_t)(total * (size_t)(t + 1) / T);
        if (t == T - 1) s1 = n; else { s1 = std::upper_bound(db->h_off.begin() + s0, db->h_off.begin() + n, want) - db->h_off.begin(); s1 = std::min(std::max(s1, s0), n); }
        th.emplace_back(fill, s0, s1);
        s0 = s1;
      }
      for (auto &t : th) t.join();
    }
    if (total < 16) memset(hb + o_res + total, B2H_PAD_CODE, 16 - total);
  }
  memcpy(hb + o_off, db->h_off.data(), (n + 1) * sizeof(int64_t));
  if (n) memcpy(hb + o_len, db->h_len.data(), n * 4);

  // order[]: sequence indices by decreasing length (stable) -- counting sort when the length range is small
  {
    int32_t *order = (int32_t *)(hb + o_ord);
    if ((size_t)db->maxL <= 8 * n + 4096) {
      std::vector<int32_t> start((size_t)db->maxL + 2, 0);
      for (size_t s = 0; s < n; s++) start[db->maxL - db->h_len[s] + 1]++;
      for (size_t i = 1; i < start.size(); i++) start[i] += start[i - 1];
      for (size_t s = 0; s < n; s++) order[start[db->maxL - db->h_len[s]]++] = (int32_t)s;
    } else {
      std::iota(order, order + n, 0);
      std::stable_sort(order, order + n, [&](int a, int b) { return db->h_len[a] > db->h_len[b]; });
    }
  }
  // L-dependent scalars depend on L only: memoise (databases have few distinct lengths)
  {
    uint8_t *tjb = hb + o_tjb; int16_t *xwm = (int16_t *)(hb + o_xwm);
    float *pmove = (float *)(hb + o_pmv), *null1 = (float *)(hb + o_nl1), *p1 = (float *)(hb + o_p1), *flta = (float *)(hb + o_fa), *fltb = (float *)(hb + o_fb);
    std::vector<int32_t> lens(db->h_len);
    std::vector<int32_t> slot;
    std::vector<b2h_len_params> cacheP;
    const bool direct = (size_t)db->maxL <= 8 * n + 4096;
    if (direct) slot.assign((size_t)db->maxL + 1, -1);
    else { std::sort(lens.begin(), lens.end()); lens.erase(std::unique(lens.begin(), lens.end()), lens.end()); cacheP.resize(lens.size()); for (size_t i = 0; i < lens.size(); i++) b2h_length_params(lens[i], 1.0f, &cacheP[i]); }
    for (size_t s = 0; s < n; s++) {
      const int32_t L = db->h_len[s];
      size_t i;
      if (direct) { if (slot[L] < 0) { slot[L] = (int32_t)cacheP.size(); cacheP.emplace_back(); b2h_length_params(L, 1.0f, &cacheP.back()); } i = slot[L]; }
      else i = std::lower_bound(lens.begin(), lens.end(), L) - lens.begin();
      const b2h_len_params &q = cacheP[i];
      tjb[s] = q.tjb_b; xwm[s] = q.xw_move; pmove[s] = q.pmove; null1[s] = q.null1; p1[s] = q.p1;
      flta[s] = q.flt_len_a; fltb[s] = q.flt_len_b;
    }
  }

  // one stream-ordered allocation, one H2D copy (the pinned source lives as long as the database)
  cudaError_t e;
  if ((e = cudaMallocAsync((void **)&db->d_block, db->block_bytes, ctx->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(db->d_block, hb, db->block_bytes, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) {
    ctx->err = std::string("sequence database upload: ") + cudaGetErrorString(e);
    b2h_seqdb_destroy(db);
    return B2H_ECUDA;
  }
  db->h2d_bytes = db->block_bytes;
  uint8_t *b = db->d_block;
  db->d_res = b + o_res; db->d_off = (int64_t *)(b + o_off); db->d_len = (int32_t *)(b + o_len); db->d_order = (int32_t *)(b + o_ord);
  db->d_tjb = b + o_tjb; db->d_xwmove = (int16_t *)(b + o_xwm); db->d_pmove = (float *)(b + o_pmv); db->d_null1 = (float *)(b + o_nl1);
  db->d_p1 = (float *)(b + o_p1); db->d_flta = (float *)(b + o_fa); db->d_fltb = (float *)(b + o_fb);
  *out = db;
  return B2H_OK;
}

int b2h_seqdb_create(b2h_ctx *ctx, const uint8_t *const *dsq, const int64_t *len, size_t n, b2h_seqdb **out)
{ return seqdb_build(ctx, n, len, dsq, nullptr, nullptr, out); }

int b2h_seqdb_create_packed(b2h_ctx *ctx, const uint8_t *residues, const int64_t *offsets, size_t n, b2h_seqdb **out)
{ return seqdb_build(ctx, n, nullptr, nullptr, residues, offsets, out); }

} // extern "C"

const b2h_chunkview *b2h_seqdb_chunk_view(const b2h_seqdb *cdb, int O, int S, cudaStream_t strm)
{
  b2h_seqdb *db = const_cast<b2h_seqdb *>(cdb);
  for (const b2h_chunkview &v : db->views) if (v.O == O && v.S == S) return &v;
  b2h_ctx *ctx = db->ctx;
  std::vector<int64_t> off; std::vector<int32_t> len, parent;
  for (size_t s = 0; s < db->n; s++)
    for (int64_t a = 0; a < db->h_len[s]; a += S) {
      off.push_back(db->h_off[s] + a);
      len.push_back((int32_t)std::min<int64_t>(db->h_len[s] - a, (int64_t)S + O));
      parent.push_back((int32_t)s);
    }
  const size_t n = off.size(), n1 = std::max<size_t>(n, 1);
  std::vector<int32_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return len[a] > len[b]; });
  b2h_chunkview v; v.O = O; v.S = S; v.n = (int)n;
  const size_t o_off = 0, o_len = (n1 * 8 + 255) & ~(size_t)255, o_ord = o_len + ((n1 * 4 + 255) & ~(size_t)255), o_par = o_ord + ((n1 * 4 + 255) & ~(size_t)255),
               bytes = o_par + n1 * 4;
  if (cudaSetDevice(ctx->device) != cudaSuccess || cudaMalloc(&v.d_block, bytes) != cudaSuccess) { ctx->err = "chunk view allocation failed"; return nullptr; }
  uint8_t *b = (uint8_t *)v.d_block;
  v.d_off = (int64_t *)(b + o_off); v.d_len = (int32_t *)(b + o_len); v.d_order = (int32_t *)(b + o_ord); v.d_parent = (int32_t *)(b + o_par);
  if (n) {                                                  // pageable sources: the copies are staged before these calls return
    cudaMemcpyAsync(v.d_off, off.data(), n * 8, cudaMemcpyHostToDevice, strm);
    cudaMemcpyAsync(v.d_len, len.data(), n * 4, cudaMemcpyHostToDevice, strm);
    cudaMemcpyAsync(v.d_order, order.data(), n * 4, cudaMemcpyHostToDevice, strm);
    cudaMemcpyAsync(v.d_parent, parent.data(), n * 4, cudaMemcpyHostToDevice, strm);
    cudaStreamSynchronize(strm);
  }
  if (db->views.capacity() < 64) db->views.reserve(64);     // (callers copy the view at once; keep the storage stable anyway)
  db->views.push_back(v);
  return &db->views.back();
}

extern "C" {

void b2h_seqdb_destroy(b2h_seqdb *db)
{
  if (!db) return;
  b2h_ctx *ctx = db->ctx;
  if (ctx) cudaSetDevice(ctx->device);
  for (b2h_chunkview &v : db->views) if (v.d_block) cudaFree(v.d_block);
  if (db->d_block) { if (ctx) cudaFreeAsync(db->d_block, ctx->stream); else cudaFree(db->d_block); }
  if (db->h_block) {
    if (ctx) {
      cudaStreamSynchronize(ctx->stream);             // the H2D copy of this block may still be in flight
      ctx->pinned_free.emplace_back((void *)db->h_block, db->h_block_cap);
    } else cudaFreeHost(db->h_block);
  }
  delete db;
}
size_t  b2h_seqdb_nseq(const b2h_seqdb *db) { return db ? db->n : 0; }
int64_t b2h_seqdb_nres(const b2h_seqdb *db) { return db ? db->nres : 0; }

// ------------------------------------------------------------------------------------------
// profile upload
// ------------------------------------------------------------------------------------------
// Lane-striped fp16x2 table of the SSV kernel for a (G, NR) register tile (b2h_ssv_tile): lane l of a group of G
// lanes owns model nodes l*2*NR+1 .. (l+1)*2*NR, word j of that lane packs the scores of cells c=j (low half) and
// c=j+NR (high half).  One residue row is
//     [NR/4 chunks][G lanes][4 words]   one LDS.128 per lane and chunk; a group reads G*16 contiguous bytes
//     [NR%4 words][32 physical lanes]   leftover words replicated for every group of the warp, so that the 32 lanes
//                                       of a warp hit 32 different banks whatever rows their groups are reading
// and every row is a multiple of 128 bytes.
static size_t ssv_word_index(int G, int NR, int x, int j, int lane /* physical lane 0..31 */)
{
  const int full = NR / 4, GS = G < 8 ? 8 : G, slot = lane % GS;     // G < 8: slot = lane & 7, i.e. group-local lane + G * (group & (8/G - 1))
  const size_t base = (size_t)x * (b2h_ssv_row_bytes(G, NR) / 4);
  if (j < full * 4) return base + ((size_t)(j / 4) * GS + slot) * 4 + (j % 4);
  return base + (size_t)full * GS * 4 + (size_t)(j - full * 4) * 32 + lane;
}

int b2h_profile_create_host(const b2h_oprofile_desc *d, b2h_profile **out)
{ return b2h_profile_upload(nullptr, d, out); }

// Host half of an upload: the profile object and, when a context is given, the staged image of its device block
// (sections 256-byte aligned; offsets in <offs>).  Pure CPU work, safe to run for many profiles in parallel.
struct ProfStage { std::vector<uint8_t> bytes; uint8_t *ext = nullptr; size_t size = 0; size_t offs[11] = {0}; };   // image in <bytes>, or written straight to <ext>
// Size of the staged device image of a profile (sections 256-byte aligned, in the order profile_build adds them).
static void profile_classes(int M, int K, int *G, int *NR, int *regC, int *regW)
{
  *G = *NR = *regC = *regW = 0;
  b2h_ssv_tile(M, G, NR, K == 20);
  const b2h_regclass *rcls; const int nrcls = b2h_reg_classes(&rcls);
  for (int rc = 0; rc < nrcls; rc++) if (M <= rcls[rc].bound) { *regC = rcls[rc].C; *regW = rcls[rc].W; break; }
}
static size_t profile_stage_bytes(int M, int G, int NR, int regC, int regW)
{
  const size_t Mp = (size_t)((M + 31) & ~31);
  size_t used = 0;
  auto add = [&](size_t bytes) { used = ((used + 255) & ~(size_t)255) + bytes; };
  int Gw = 0, NRw = 0; b2h_ssv_tile(M, &Gw, &NRw, false);
  add((size_t)B2H_NCODE * b2h_ssv_row_bytes(G, NR)); add(B2H_NCODE * Mp);
  if (Gw != G || NRw != NR) add((size_t)B2H_NCODE * b2h_ssv_row_bytes(Gw, NRw));     // the wide tile's table (last section)
  add(B2H_NCODE * Mp * 2); add(8 * Mp * 2); add(B2H_NCODE * Mp * 4); add(8 * Mp * 4); add((size_t)B2H_NCODE * 2 * 4);
  if (regC) { add((size_t)B2H_NCODE * 32 * regC * regW * 4); add((size_t)B2H_NCODE * 32 * regC * regW * 4); }
  if (regC && regW == 1) add((size_t)B2H_NCODE * 32 * ((regC + 1) / 2) * 4);       // packed Viterbi table
  return used;
}

static int profile_build(b2h_ctx *ctx, const b2h_oprofile_desc *d, b2h_profile **out, ProfStage *stg)
{
  if (!d || !out || d->M < 1 || d->Kp > B2H_NCODE - 1 || d->K > B2H_MAXABET) return B2H_EINVAL;
  *out = nullptr;
  const int M = d->M, Kp = d->Kp;
  int G = 0, NR = 0;
  if (!b2h_ssv_tile(M, &G, &NR, d->K == 20)) { if (ctx) ctx->err = "model too long for the register-tiled SSV kernel (M > 3071)"; return B2H_EINVAL; }
  b2h_profile *p = new b2h_profile();
  p->ctx = ctx; p->M = M; p->K = d->K; p->Kp = Kp; p->max_length = d->max_length; p->multihit = d->mode_multihit;
  p->NR = NR; p->G = G; p->tbm_b = d->tbm_b; p->tec_b = d->tec_b; p->base_b = d->base_b; p->bias_b = d->bias_b; p->scale_b = d->scale_b;
  memcpy(p->xw, d->xw, sizeof p->xw); p->base_w = d->base_w; p->ddbound_w = d->ddbound_w; p->scale_w = d->scale_w;
  memcpy(p->xf, d->xf, sizeof p->xf);
  memcpy(p->evparam, d->evparam, sizeof p->evparam); memcpy(p->cutoff, d->cutoff, sizeof p->cutoff);
  memcpy(p->compo, d->compo, sizeof p->compo); memcpy(p->bgf, d->bgf, sizeof p->bgf);
  p->Mpad = (M + 31) & ~31;
  p->h_fwd_rsc.assign(d->fwd_rsc, d->fwd_rsc + (size_t)Kp * M);
  p->h_fwd_tsc.assign(d->fwd_tsc, d->fwd_tsc + (size_t)8 * M);
  if (d->degen) p->h_degen.assign(d->degen, d->degen + (size_t)Kp * d->K);
  p->symbols = (d->K == 20) ? "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~" : "ACGT-RYMKSWHBVDN*~";

  // --- SSV signed scores / MSV costs, lane-striped ---
  auto cost_of = [&](int x, int k) -> int {          // k is 1-based; anything off the model is the -inf cost
    return (x < Kp && k <= M) ? (int)d->msv_cost[(size_t)x * M + (k-1)] : 255;
  };
  uint16_t half_of[512];                                 // fp16 bit patterns of the integers -256 .. 255 (exact)
  for (int v = -256; v < 256; v++) half_of[v + 256] = __half_as_ushort(__float2half_rn((float)v));
  auto ssv_table = [&](int G, int NR, std::vector<uint32_t> &ssv) {
  ssv.assign((size_t)B2H_NCODE * (b2h_ssv_row_bytes(G, NR) / 4), 0u);
  for (int x = 0; x < B2H_NCODE; x++)
    for (int gl = 0; gl < G; gl++)
      for (int j = 0; j < NR; j++) {
        const int klo = gl * 2 * NR + j + 1, khi = klo + NR;
        const int clo = cost_of(x, klo), chi = cost_of(x, khi);
        // score = bias - cost, stored as fp16 (exact: |score| <= 255).  MSV adds the bias and subtracts the cost as
        // unsigned bytes (msvfilter.c:155-158); SSV subtracts sbv = clamp(cost - bias, .., 127) as a signed byte
        // (p7_oprofile.c:721-761), and the clamp is immaterial there: an SSV cell is < 127 - bias until the comparison
        // has overflowed, so any decrement >= 127 takes it to the floor either way.  One table serves both kernels.
        const int slo = (int)d->bias_b - clo, shi = (int)d->bias_b - chi;
        const uint32_t w = ((uint32_t)half_of[shi + 256] << 16) | (uint32_t)half_of[slo + 256];
        if (j < (NR / 4) * 4) for (int lane = gl; lane < (G < 8 ? 8 : G); lane += G) ssv[ssv_word_index(G, NR, x, j, lane)] = w;   // (G < 8: one copy per group of a quarter-warp)
        else for (int lane = gl; lane < 32; lane += G) ssv[ssv_word_index(G, NR, x, j, lane)] = w;     // leftover words: one copy per group of the warp
      }
  };
  std::vector<uint32_t> ssv, ssv_w;
  ssv_table(G, NR, ssv);
  p->Gw = G; p->NRw = NR;
  b2h_ssv_tile(M, &p->Gw, &p->NRw, false);
  if (p->Gw != G || p->NRw != NR) ssv_table(p->Gw, p->NRw, ssv_w);

  // --- Viterbi / Forward tables, padded ---
  const int Mp = p->Mpad;
  std::vector<int16_t> vr((size_t)B2H_NCODE * Mp, (int16_t)-32768), vt((size_t)8 * Mp, (int16_t)-32768);
  std::vector<float>   fr((size_t)B2H_NCODE * Mp, 0.0f),            ft((size_t)8 * Mp, 0.0f);
  std::vector<uint8_t> mc((size_t)B2H_NCODE * Mp, (uint8_t)255);
  for (int x = 0; x < Kp; x++)
    for (int k = 0; k < M; k++) {
      vr[(size_t)x * Mp + k] = d->vit_rsc[(size_t)x * M + k];
      mc[(size_t)x * Mp + k] = d->msv_cost[(size_t)x * M + k];
      fr[(size_t)x * Mp + k] = d->fwd_rsc[(size_t)x * M + k];
    }
  for (int t = 0; t < 8; t++)
    for (int k = 0; k < M; k++) {
      vt[(size_t)t * Mp + k] = d->vit_tsc[(size_t)t * M + k];
      ft[(size_t)t * Mp + k] = d->fwd_tsc[(size_t)t * M + k];
    }

  // --- lane-grouped emission tables of the register-resident DP kernels (b2h_dpreg.cu), models up to 512 nodes ---
  p->regC = 0; p->regW = 0;
  const b2h_regclass *rcls; const int nrcls = b2h_reg_classes(&rcls);
  for (int rc = 0; rc < nrcls; rc++)
    if (M <= rcls[rc].bound) { p->regC = rcls[rc].C; p->regW = rcls[rc].W; break; }
  std::vector<int32_t> vr32; std::vector<float> frr;
  if (p->regC) {
    const int C = p->regC, W = p->regW, G = (C % 4 == 0) ? 4 : (C % 2 == 0) ? 2 : 1, stride = 32 * C * W;
    vr32.assign((size_t)B2H_NCODE * stride, -32768); frr.assign((size_t)B2H_NCODE * stride, 0.0f);
    for (int x = 0; x < Kp; x++)
      for (int gl = 0; gl < 32 * W; gl++)
        for (int c = 0; c < C; c++) {
          const int k0 = gl * C + c;
          if (k0 >= M) continue;
          const int wi = gl / 32, lane = gl % 32;
          const size_t idx = (size_t)x * stride + (size_t)wi * 32 * C + (size_t)(c / G) * 32 * G + (size_t)lane * G + (c % G);
          vr32[idx] = d->vit_rsc[(size_t)x * M + k0];
          frr[idx] = d->fwd_rsc[(size_t)x * M + k0];
        }
  }

  // --- packed s16x2 emission table of the 16-bit ViterbiFilter kernel (single-warp classes) ---
  std::vector<uint32_t> vr2;
  p->v2C = 0; p->v2_ok = 0; p->tbm_min = 0;
  if (p->regC && p->regW == 1) {
    const int C2 = (p->regC + 1) & ~1, H = C2 / 2, g = (H % 4 == 0) ? 4 : (H % 2 == 0) ? 2 : 1;
    int rmax = -32768, tbm_min = 0;
    for (int x = 0; x < Kp; x++) for (int k = 0; k < M; k++) rmax = std::max(rmax, (int)d->vit_rsc[(size_t)x * M + k]);
    for (int k = 0; k < M; k++) tbm_min = std::min(tbm_min, (int)d->vit_tsc[k]);
    p->v2C = C2; p->tbm_min = tbm_min; p->v2_ok = (rmax <= B2H_V2_RMAX);
    auto val = [&](int x, int k0) -> uint32_t {           // clamped from below at V2_TF; off the model = -inf
      const int v = (x < Kp && k0 < M) ? (int)d->vit_rsc[(size_t)x * M + k0] : -32768;
      return (uint32_t)(uint16_t)(int16_t)std::max(v, B2H_V2_TF);
    };
    vr2.resize((size_t)B2H_NCODE * 32 * H);
    for (int x = 0; x < B2H_NCODE; x++)
      for (int lane = 0; lane < 32; lane++)
        for (int j = 0; j < H; j++)
          vr2[(size_t)x * 32 * H + (size_t)(j / g) * 32 * g + (size_t)lane * g + (j % g)] = val(x, lane * C2 + j) | (val(x, lane * C2 + j + H) << 16);
  }

  // --- bias-filter 2-state HMM (p7_bg_SetFilter p7_bg.c:429, esl_hmm_Configure esl_hmm.c:118) ---
  std::vector<float> eo((size_t)B2H_NCODE * 2, 1.0f);
  {
    const int K = d->K;
    for (int x = 0; x < K; x++) { eo[x*2+0] = d->bgf[x] / d->bgf[x]; eo[x*2+1] = d->compo[x] / d->bgf[x]; }
    for (int x = K + 1; x <= Kp - 3; x++)
      for (int s = 0; s < 2; s++) {
        float num = 0.0f, denom = 0.0f;
        for (int y = 0; y < K; y++)
          if (d->degen && d->degen[(size_t)x * K + y]) { num += (s == 0 ? d->bgf[y] : d->compo[y]); denom += d->bgf[y]; }
        eo[x*2+s] = (denom > 0.0f) ? num / denom : 0.0f;
      }
    float L1 = (float)((double)(float)M / 8.0);
    p->bias_t10 = 1.0f / (L1 + 1.0f);
    p->bias_t11 = L1 / (L1 + 1.0f);
  }

  if (ctx && stg) {
    std::vector<uint8_t> &stage = stg->bytes;
    if (!stg->ext) stage.resize(profile_stage_bytes(M, G, NR, p->regC, p->regW));
    uint8_t *dst = stg->ext ? stg->ext : stage.data();
    size_t used = 0;
    auto add = [&](const void *src, size_t bytes) -> size_t {
      const size_t off = (used + 255) & ~(size_t)255;
      memcpy(dst + off, src, bytes);
      used = off + bytes;
      return off;
    };
    stg->offs[0] = add(ssv.data(), ssv.size() * 4); stg->offs[1] = add(mc.data(), mc.size());
    stg->offs[2] = add(vr.data(), vr.size() * 2);   stg->offs[3] = add(vt.data(), vt.size() * 2);
    stg->offs[4] = add(fr.data(), fr.size() * 4);   stg->offs[5] = add(ft.data(), ft.size() * 4);
    stg->offs[6] = add(eo.data(), eo.size() * 4);
    if (p->regC) { stg->offs[7] = add(vr32.data(), vr32.size() * 4); stg->offs[8] = add(frr.data(), frr.size() * 4); }
    if (p->v2C) stg->offs[9] = add(vr2.data(), vr2.size() * 4);
    stg->offs[10] = ssv_w.empty() ? stg->offs[0] : add(ssv_w.data(), ssv_w.size() * 4);
    if (used > profile_stage_bytes(M, G, NR, p->regC, p->regW)) { delete p; return B2H_EINVAL; }   // (cannot happen: same arithmetic)
    stg->size = used;
    p->h2d_bytes = used;
  }
  *out = p;
  return B2H_OK;
}

// Device half: one stream-ordered allocation and one H2D copy for all tables of the profile.
static int profile_commit(b2h_ctx *ctx, b2h_profile *p, const ProfStage &stg)
{
  cudaError_t e;
  cudaSetDevice(ctx->device);
  if ((e = cudaMallocAsync((void **)&p->d_block, stg.size, ctx->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(p->d_block, stg.bytes.data(), stg.size, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) {
    ctx->err = cudaGetErrorString(e); return B2H_ECUDA;        // (pageable source: the copy is staged before the call returns)
  }
  uint8_t *b = (uint8_t *)p->d_block;
  p->d_ssv_emis = (uint32_t *)(b + stg.offs[0]); p->d_msv_cost8 = b + stg.offs[1];
  p->d_vit_rsc = (int16_t *)(b + stg.offs[2]); p->d_vit_tsc = (int16_t *)(b + stg.offs[3]);
  p->d_fwd_rsc = (float *)(b + stg.offs[4]); p->d_fwd_tsc = (float *)(b + stg.offs[5]); p->d_bias_eo = (float *)(b + stg.offs[6]);
  if (p->regC) { p->d_vit_rsc32 = (int32_t *)(b + stg.offs[7]); p->d_fwd_rscr = (float *)(b + stg.offs[8]); }
  if (p->v2C) p->d_vit_rsc2 = (uint32_t *)(b + stg.offs[9]);
  p->d_ssv_emis_w = (uint32_t *)(b + stg.offs[10]);
  return B2H_OK;
}

int b2h_profile_upload(b2h_ctx *ctx, const b2h_oprofile_desc *d, b2h_profile **out)
{
  ProfStage stg;
  b2h_profile *p = nullptr;
  if (out) *out = nullptr;
  int st = profile_build(ctx, d, &p, ctx ? &stg : nullptr);
  if (st != B2H_OK) return st;
  if (ctx && (st = profile_commit(ctx, p, stg)) != B2H_OK) { b2h_profile_destroy(p); return st; }
  *out = p;
  return B2H_OK;
}

// Batched upload: the host halves are built in parallel, the staged images are packed into ONE page-locked buffer and go
// to ONE device allocation with ONE copy (100 Pfam-sized profiles: ~9 MB, 0.4 ms instead of 100 pageable copies).
int b2h_profile_upload_many(b2h_ctx *ctx, const b2h_oprofile_desc *const *descs, size_t n, b2h_profile **out)
{
  if (!ctx || !descs || !out) return B2H_EINVAL;
  for (size_t i = 0; i < n; i++) out[i] = nullptr;
  if (n == 0) return B2H_OK;
  std::vector<ProfStage> stg(n);
  std::vector<int> status(n, B2H_OK);
  const int T = (int)std::min<size_t>((size_t)b2h_rank_threads(16), n);
  auto run_parallel = [&](const std::function<void(size_t)> &fn) {
    auto work = [&](int t) { for (size_t i = t; i < n; i += T) fn(i); };
    if (T <= 1) work(0);
    else { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(work, t); for (auto &x : th) x.join(); }
  };
  std::vector<size_t> base(n + 1, 0);
  for (size_t i = 0; i < n; i++) {
    if (!descs[i] || descs[i]->M < 1) return B2H_EINVAL;
    int G, NR, rC, rW; profile_classes(descs[i]->M, descs[i]->K, &G, &NR, &rC, &rW);
    if (!G) { ctx->err = "model too long for the register-tiled SSV kernel (M > 3071)"; return B2H_EINVAL; }
    base[i + 1] = base[i] + ((profile_stage_bytes(descs[i]->M, G, NR, rC, rW) + 255) & ~(size_t)255);
  }
  cudaSetDevice(ctx->device);
  static const bool trace = getenv("B2H_TRACE") != nullptr;
  const auto tr0 = std::chrono::steady_clock::now();
  uint8_t *h = (uint8_t *)b2h_pin_get(ctx, base[n]);
  if (!h) { ctx->err = "cudaHostAlloc failed"; return B2H_EMEM; }
  const auto tr1 = std::chrono::steady_clock::now();
  run_parallel([&](size_t i) { stg[i].ext = h + base[i]; status[i] = profile_build(ctx, descs[i], &out[i], &stg[i]); });
  const auto tr2 = std::chrono::steady_clock::now();
  int st = B2H_OK;
  for (size_t i = 0; i < n && st == B2H_OK; i++) st = status[i];
  b2h_devblock *blk = nullptr;
  if (st == B2H_OK) {
    blk = new b2h_devblock();
    cudaError_t e;
    if ((e = cudaMallocAsync(&blk->d, base[n], ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(blk->d, h, base[n], cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {       // the staging buffer goes back to the pool below
      ctx->err = cudaGetErrorString(e); st = B2H_ECUDA;
      if (blk->d) cudaFreeAsync(blk->d, ctx->stream);
      delete blk; blk = nullptr;
    }
  }
  if (h) b2h_pin_put(ctx, h);
  if (trace) {
    const auto tr3 = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "[b2h_profile_upload_many] %zu profiles, %.1f MB on %d threads: staging %.2f ms, tables %.2f ms, alloc + copy %.2f ms\n",
            n, base[n] / 1e6, T, ms(tr0, tr1), ms(tr1, tr2), ms(tr2, tr3));
  }
  if (st == B2H_OK) {
    blk->refs = (int)n;
    for (size_t i = 0; i < n; i++) {
      b2h_profile *p = out[i]; const ProfStage &sg = stg[i];
      uint8_t *b = (uint8_t *)blk->d + base[i];
      p->shared = blk;
      p->d_ssv_emis = (uint32_t *)(b + sg.offs[0]); p->d_msv_cost8 = b + sg.offs[1];
      p->d_vit_rsc = (int16_t *)(b + sg.offs[2]); p->d_vit_tsc = (int16_t *)(b + sg.offs[3]);
      p->d_fwd_rsc = (float *)(b + sg.offs[4]); p->d_fwd_tsc = (float *)(b + sg.offs[5]); p->d_bias_eo = (float *)(b + sg.offs[6]);
      if (p->regC) { p->d_vit_rsc32 = (int32_t *)(b + sg.offs[7]); p->d_fwd_rscr = (float *)(b + sg.offs[8]); }
      if (p->v2C) p->d_vit_rsc2 = (uint32_t *)(b + sg.offs[9]);
      p->d_ssv_emis_w = (uint32_t *)(b + sg.offs[10]);
    }
  } else {
    for (size_t i = 0; i < n; i++) { b2h_profile_destroy(out[i]); out[i] = nullptr; }
  }
  return st;
}

void b2h_profile_destroy(b2h_profile *p)
{
  if (!p) return;
  if (p->ctx) cudaSetDevice(p->ctx->device);
  if (p->d_block) { if (p->ctx) cudaFreeAsync(p->d_block, p->ctx->stream); else cudaFree(p->d_block); }
  if (p->shared && --p->shared->refs == 0) {
    if (p->ctx) cudaFreeAsync(p->shared->d, p->ctx->stream); else cudaFree(p->shared->d);
    delete p->shared;
  }
  delete p;
}

int b2h_profile_set_annotation(b2h_profile *p, const char *consensus, const char *rf, const char *cs, const char *symbols)
{
  if (!p) return B2H_EINVAL;
  p->consensus = consensus ? consensus : "";
  p->rf = rf ? rf : "";
  p->cs = cs ? cs : "";
  if (symbols && *symbols) p->symbols = symbols;
  return B2H_OK;
}

int b2h_profile_set_model_mask(b2h_profile *p, const char *mm)
{
  if (!p) return B2H_EINVAL;
  p->mm = mm ? mm : "";
  return B2H_OK;
}

} // extern "C"

extern "C" size_t b2h_seqdb_h2d_bytes(const b2h_seqdb *db) { return db ? db->h2d_bytes : 0; }
extern "C" size_t b2h_profile_h2d_bytes(const b2h_profile *p) { return p ? p->h2d_bytes : 0; }

extern "C" int b2h_ssv_tile_info(int M, int *G, int *NR, double *wavefronts_per_row)
{
  int g = 0, nr = 0;
  if (M < 1 || !b2h_ssv_tile(M, &g, &nr, true)) return B2H_EINVAL;                   // (the tile of a PROTEIN profile)
  if (G) *G = g;
  if (NR) *NR = nr;
  if (wavefronts_per_row) *wavefronts_per_row = b2h_ssv_tile_cost(g, nr) * 32.0 / g;  // per WARP and row: LDS.128 x NR/4, LDS.32 x NR%4, diagonal SHFL, residue SHFL / 4 rows
  return B2H_OK;
}

// Page-locked result buffers.  Every buffer carries its capacity in a 64-byte header in front of the pointer handed
// out; buffers go back to the pool (never to the driver) until the context is destroyed.
void *b2h_pin_get(b2h_ctx *ctx, size_t bytes)
{
  const size_t need = bytes + 64;
  {
    std::lock_guard<std::mutex> lk(ctx->pin_mu);
    size_t best = (size_t)-1;
    for (size_t i = 0; i < ctx->pin_pool.size(); i++)
      if (ctx->pin_pool[i].second >= need && (best == (size_t)-1 || ctx->pin_pool[i].second < ctx->pin_pool[best].second)) best = i;
    if (best != (size_t)-1) {
      void *p = ctx->pin_pool[best].first;
      ctx->pin_pool.erase(ctx->pin_pool.begin() + best);
      return (uint8_t *)p + 64;
    }
  }
  size_t cap = (size_t)1 << 16;
  while (cap < need) cap += cap / 2 + ((size_t)1 << 16);       // geometric size classes: a slightly larger search reuses the buffer
  void *p = nullptr;
  cudaSetDevice(ctx->device);
  if (cudaHostAlloc(&p, cap, cudaHostAllocDefault) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
  *(size_t *)p = cap;
  return (uint8_t *)p + 64;
}
void b2h_pin_put(b2h_ctx *ctx, void *q)
{
  if (!q) return;
  void *p = (uint8_t *)q - 64;
  std::lock_guard<std::mutex> lk(ctx->pin_mu);
  ctx->pin_pool.emplace_back(p, *(size_t *)p);
}


void *b2h_bigbuf_get(b2h_ctx *ctx, size_t bytes, cudaStream_t strm)
{
  b2h_ctx::BigBuf *best = nullptr;
  for (auto &b : ctx->bigbufs) if (!b.in_use && b.bytes >= bytes && (!best || b.bytes < best->bytes)) best = &b;
  if (!best) {
    for (auto &b : ctx->bigbufs) if (!b.in_use && (!best || b.bytes > best->bytes)) best = &b;       // grow the largest idle one
    if (best) { cudaEventSynchronize(best->free_ev); cudaFree(best->p); best->p = nullptr; best->bytes = 0; }
    else { ctx->bigbufs.emplace_back(); best = &ctx->bigbufs.back(); cudaEventCreateWithFlags(&best->free_ev, cudaEventDisableTiming); }
    if (cudaMalloc(&best->p, bytes) != cudaSuccess) { (void)cudaGetLastError(); best->p = nullptr; best->bytes = 0; return nullptr; }
    best->bytes = bytes;
    cudaEventRecord(best->free_ev, strm);
  }
  best->in_use = true;
  cudaStreamWaitEvent(strm, best->free_ev, 0);             // its previous user (on whatever lane) is done
  return best->p;
}
void b2h_bigbuf_put(b2h_ctx *ctx, void *p, cudaStream_t strm)
{
  for (auto &b : ctx->bigbufs) if (b.p == p) { cudaEventRecord(b.free_ev, strm); b.in_use = false; return; }
}

int b2h_kernel_occupancy(b2h_ctx *ctx, const void *kernel, int threads, size_t smem, int *occ)
{
  static std::mutex mu;
  static std::map<std::tuple<int, const void *, size_t, int>, int> cache;
  const auto key = std::make_tuple(ctx->device, kernel, smem, threads);
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *occ = it->second; return B2H_OK; }
  }
  {                                                        // the limit only ever grows (launches with less still fit)
    static std::map<std::pair<int, const void *>, size_t> limit;
    std::lock_guard<std::mutex> lk(mu);
    size_t &cur = limit[std::make_pair(ctx->device, kernel)];
    // (static + dynamic may cross 48 KB even when the dynamic part alone does not)
    if (smem > cur) { B2H_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cur = smem; }
  }
  int o = 1;
  B2H_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kernel, threads, smem));
  if (o < 1) o = 1;
  std::lock_guard<std::mutex> lk(mu);
  cache[key] = o; *occ = o;
  return B2H_OK;
}
