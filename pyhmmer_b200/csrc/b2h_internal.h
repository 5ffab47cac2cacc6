// b2h_internal.h -- shared declarations for libb2h.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <atomic>
#include <cstdio>
#include <mutex>
#include <utility>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>
#include "b2h.h"

// ---------------------------------------------------------------------------------------------
// Data layout in HBM
//
//  sequence arena   one uint8 buffer; sequence s occupies res[off[s] .. off[s]+len[s]) with
//                   off[s] % 16 == 0 and the tail up to the next multiple of 16 filled with
//                   B2H_PAD_CODE, so kernels may read whole 4/16-byte words.  Per-sequence
//                   side arrays (len, L-dependent scalars) are SoA.  `order[]` lists sequence
//                   indices by decreasing length: persistent CTAs pull work in that order so
//                   the longest comparisons start first.
//  profile tables   node-major (k-1 indexed) copies of the three score systems, plus a
//                   pre-swizzled 16-bit "lane-striped" emission table per integer filter laid
//                   out exactly as the kernels' LDS.128 reads want it (see b2h_msv.cu).
// ---------------------------------------------------------------------------------------------

#define B2H_PAD_CODE   31     // residue code used for arena padding; tables have 32 residue rows
#define B2H_NCODE      32
#define B2H_MAX_NR     48     // SSV register tiles: a group of G lanes holds 2*G*NR cells; G=32, NR=48 => M <= 3071

struct b2h_ctx {
  int           device = 0;
  int           sm_count = 0;
  cudaStream_t  own_stream = nullptr;
  cudaStream_t  stream = nullptr;
  std::string   err;
  std::atomic<uint64_t> launches{0};   // (the envelope kernels are launched from the domain-definition thread)
  int          *d_counters = nullptr;   // small pool of work counters
  int           profiling = 0;
  // side streams: independent launches of one stage (size classes) run concurrently, forked from / joined to <stream>
  std::vector<cudaStream_t> side; std::vector<cudaEvent_t> side_done; cudaEvent_t fork_ev = nullptr; int side_used = 0;
  double        stage_ms[8] = {0};
  std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev_open;   // (stage, begin, end) awaiting a sync
  cudaStream_t  env_stream = nullptr;  // stream of the envelope kernels (they overlap the next wave's cascade)
  int          *d_env_counter = nullptr;
  cudaStream_t  env_side[8] = {nullptr}; cudaEvent_t env_fork = nullptr, env_join[8] = {nullptr};   // one side stream per envelope size class
  // page-locked host buffers of destroyed sequence databases, kept for the next one (pinning costs ~0.3 ms/MB)
  std::vector<std::pair<void *, size_t>> pinned_free;
  // Extra launch lanes (a stream, its side streams and its own work counters; all high priority).  b2h_lane_switch
  // swaps one in as the current lane:
  //   B2H_LANE_SURV  survivor Forward/Backward passes of wave w, issued while later waves are already queued
  //   B2H_LANE_POST  the cascade stages after SSV (MSV, bias, Viterbi, Forward) of wave w, which run NEXT TO the SSV
  //                  launches of wave w+1 on the main lane: SSV saturates the shared-memory pipe, these the ALU pipe
  struct Lane { cudaStream_t stream = nullptr; std::vector<cudaStream_t> side; std::vector<cudaEvent_t> side_done;
                cudaEvent_t fork_ev = nullptr; int *counters = nullptr; };
  Lane lanes[3]; int prio_hi = 0;
  cudaStream_t bias_stream = nullptr;  // the bias filter of a wave runs here, next to that wave's Viterbi launches
  // page-locked result buffers (survivor lists, parser special rows) recycled between searches
  std::mutex    pin_mu; std::vector<std::pair<void *, size_t>> pin_pool;
  // Large device buffers that travel between lanes (the Forward special-state rows a cascade wave leaves for the survivor
  // lane): kept for the life of the context and handed over with an event, because the stream-ordered allocator cannot
  // reuse a block freed on one stream for an allocation on another without new physical memory.
  struct BigBuf { void *p = nullptr; size_t bytes = 0; cudaEvent_t free_ev = nullptr; bool in_use = false; };
  std::vector<BigBuf> bigbufs;
};
// a device buffer of at least <bytes> for work queued on <strm> (nullptr: out of memory); b2h_bigbuf_put: the work queued on
// <strm> so far is its last user
void *b2h_bigbuf_get(b2h_ctx *ctx, size_t bytes, cudaStream_t strm);
void  b2h_bigbuf_put(b2h_ctx *ctx, void *p, cudaStream_t strm);

// RAII: make one of the extra lanes the current one (only the thread that drives b2h_search does this)
#define B2H_LANE_SURV 0
#define B2H_LANE_POST 1
#define B2H_LANE_SURV2 2      /* second survivor lane: the survivor passes of two consecutive waves may be in flight together */
struct b2h_lane_switch {
  b2h_ctx *c; int i;
  b2h_lane_switch(b2h_ctx *ctx, int lane) : c(ctx), i(lane) { swap(); }
  ~b2h_lane_switch() { swap(); }
  void swap() { b2h_ctx::Lane &l = c->lanes[i]; std::swap(c->stream, l.stream); std::swap(c->side, l.side); std::swap(c->side_done, l.side_done);
                std::swap(c->fork_ev, l.fork_ev); std::swap(c->d_counters, l.counters); }
};
// occupancy (resident CTAs per SM) of a kernel at <threads> / <smem> dynamic bytes, raising its dynamic shared-memory
// limit on first use; cached per (device, kernel, smem): the runtime queries cost ~10 us each and a search launches hundreds of kernels
int b2h_kernel_occupancy(b2h_ctx *ctx, const void *kernel, int threads, size_t smem, int *occ);
void *b2h_pin_get(b2h_ctx *ctx, size_t bytes);          // page-locked buffer of at least <bytes> (nullptr: out of memory)
void  b2h_pin_put(b2h_ctx *ctx, void *p);

// device block shared by the profiles of one batched upload; freed when the last of them is destroyed
struct b2h_devblock { void *d = nullptr; std::atomic<int> refs{0}; };

// A view of the arena as overlapping chunks (scan orientation: few long sequences): chunk c of sequence s covers residues
// [c*S, min(L, c*S + S + O)) of s.  An ungapped diagonal of a model of M <= O + 1 nodes spans at most M rows, so it lies
// wholly inside one chunk and the maximum over the chunks of the chunk-local SSV maxima IS the SSV maximum of the sequence.
struct b2h_chunkview {
  int O = 0, S = 0, n = 0;
  void *d_block = nullptr;
  int64_t *d_off = nullptr; int32_t *d_len = nullptr, *d_order = nullptr, *d_parent = nullptr;
};

struct b2h_seqdb {
  b2h_ctx  *ctx = nullptr;
  size_t    n = 0;
  int64_t   nres = 0;
  int       maxL = 0;
  size_t    arena_bytes = 0;
  size_t    h2d_bytes = 0;         // bytes copied host->device when the database was made resident
  std::vector<int32_t> h_len;
  std::vector<int64_t> h_off;
  // One page-locked host block and one device block with the same layout: arena | off | len | order | tjb | xwmove |
  // pmove | null1 | p1 | flta | fltb (sections 256-byte aligned), moved by a single H2D copy.
  uint8_t  *h_block = nullptr; size_t h_block_cap = 0;
  uint8_t  *d_block = nullptr; size_t block_bytes = 0;
  const uint8_t *h_res = nullptr;  // host copy of the arena inside h_block (domain definition reads residues)
  uint8_t  *d_res = nullptr;
  int64_t  *d_off = nullptr;
  int32_t  *d_len = nullptr;
  int32_t  *d_order = nullptr;
  // L-dependent scalars, multihit (nj=1) configuration -- what the search loop uses
  uint8_t  *d_tjb = nullptr;
  int16_t  *d_xwmove = nullptr;
  float    *d_pmove = nullptr;     // ploop = 1 - pmove is recomputed (same float op)
  float    *d_null1 = nullptr;
  float    *d_p1 = nullptr;
  float    *d_flta = nullptr, *d_fltb = nullptr;
  std::vector<b2h_chunkview> views;   // built on demand by b2h_seqdb_chunk_view (driving thread of b2h_search only)
};
// the (O, S) chunk view of a database, built and uploaded (stream-ordered on <strm>) at first use
const b2h_chunkview *b2h_seqdb_chunk_view(const b2h_seqdb *db, int O, int S, cudaStream_t strm);

struct b2h_profile {
  b2h_ctx *ctx = nullptr;
  void *d_block = nullptr;         // the single device allocation all d_* table pointers below point into (own upload), or
  struct b2h_devblock *shared = nullptr;   // the allocation shared by every profile of one b2h_profile_upload_many call
  int M = 0, K = 0, Kp = 0, max_length = 0, multihit = 1;
  // MSV
  int NR = 0, G = 0;              // SSV register tile: G lanes per comparison, NR packed cell registers (2*NR nodes) per lane
  int NRw = 0, Gw = 0;            // the 8 / 16 / 32-lane tile of the same model: fewest registers per lane = shortest row.  The scan
                                  // orientation (one long query: a comparison is a latency chain, not a throughput problem) uses it.
  uint32_t *d_ssv_emis_w = nullptr;   // its table (== d_ssv_emis when the two tiles coincide)
  uint8_t tbm_b = 0, tec_b = 0, base_b = 0, bias_b = 0;
  float scale_b = 0;
  uint32_t *d_ssv_emis = nullptr; // [32 residues][b2h_ssv_row_bytes(G, NR)] packed fp16x2 signed scores (SSV), lane-striped
  uint8_t  *d_msv_cost8 = nullptr; // [32][Mpad] node-major u8 costs (255 = -inf) for the full MSV kernel
  // Viterbi
  int16_t *d_vit_rsc = nullptr;   // [32][Mpad]
  int16_t *d_vit_tsc = nullptr;   // [8][Mpad]
  int16_t xw[4][2] = {{0}};
  int16_t base_w = 0, ddbound_w = 0;
  float scale_w = 0;
  // Forward/Backward
  float *d_fwd_rsc = nullptr;     // [32][Mpad]
  float *d_fwd_tsc = nullptr;     // [8][Mpad]
  float xf[4][2] = {{0}};
  int Mpad = 0;                   // M rounded up to a multiple of 32
  float evparam[B2H_NEVPARAM];
  float cutoff[B2H_NCUTOFFS];
  float compo[B2H_MAXABET];
  float bgf[B2H_MAXABET];
  float *d_bias_eo = nullptr;     // [32][2] bias-filter emission odds (esl_hmm_Configure)
  size_t h2d_bytes = 0;           // size of the single device block (= bytes uploaded)
  int regC = 0, regW = 0;         // nodes per lane / warps per comparison of the register-resident DP kernels (0 = model too long)
  uint32_t *d_vit_rsc2 = nullptr; int v2C = 0, v2_ok = 0, tbm_min = 0;   // packed s16x2 emission table [32][H/g][32 lanes][g], H = v2C/2 (b2h_dpreg.cu)
  int32_t *d_vit_rsc32 = nullptr; // [32][regW][regC/G][32][G] int32 emission scores, lane-grouped (b2h_dpreg.cu)
  float   *d_fwd_rscr = nullptr;  // same layout, fp32 odds ratios
  // host copies for the domain-definition stage
  std::vector<float> h_fwd_rsc, h_fwd_tsc;   // [Kp][M], [8][M] node-major odds ratios
  std::vector<uint8_t> h_degen;              // [Kp][K]
  std::string consensus, rf, cs, symbols;    // 1..M annotation (index k-1), alphabet symbols
  std::string mm;                            // model mask (P7_OPROFILE.mm, 'm' = masked node), index k-1; empty = none
  float bias_t10 = 0, bias_t11 = 0;   // fhmm->t[1][0], t[1][1]
};

#define B2H_CUDA(call)                                                                   \
  do { cudaError_t e_ = (call);                                                          \
       if (e_ != cudaSuccess) {                                                          \
         if (ctx) { char b_[256]; snprintf(b_, sizeof b_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); ctx->err = b_; } \
         return B2H_ECUDA; } } while (0)

// stage timing (only when ctx->profiling): events are resolved at the next host sync of the search
struct StageTimer {
  b2h_ctx *ctx; int stage; cudaEvent_t e0 = nullptr, e1 = nullptr;
  static cudaEvent_t get(b2h_ctx *c) { if (c->ev_used == c->ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); c->ev_pool.push_back(e); } return c->ev_pool[c->ev_used++]; }
  StageTimer(b2h_ctx *c, int st) : ctx(c), stage(st) { if (c->profiling) { e0 = get(c); e1 = get(c); cudaEventRecord(e0, c->stream); } }
  ~StageTimer() { if (e0) { cudaEventRecord(e1, ctx->stream); ctx->ev_open.push_back({stage, {e0, e1}}); } }
};
static inline void b2h_resolve_timers(b2h_ctx *c) {       // stages of a wave still running on the other lane stay open
  size_t keep = 0;
  for (auto &o : c->ev_open) {
    float ms = 0.f;
    const cudaError_t e = cudaEventElapsedTime(&ms, o.second.first, o.second.second);
    if (e == cudaSuccess) c->stage_ms[o.first] += ms;
    else if (e == cudaErrorNotReady) c->ev_open[keep++] = o;
  }
  (void)cudaGetLastError();
  c->ev_open.resize(keep);
  if (keep == 0) c->ev_used = 0;
}

// fork/join of the side streams around a group of independent launches (one per size class: a class launch of a few
// hundred comparisons is as long as its longest comparison and fills a few SMs, so the classes must overlap)
#define B2H_NSIDE 12
struct ForkJoin {
  b2h_ctx *ctx; int n = 0;
  explicit ForkJoin(b2h_ctx *c) : ctx(c) {
    if (!c->fork_ev) cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming);
    cudaEventRecord(c->fork_ev, c->stream);
  }
  cudaStream_t next() {                                   // stream for the next independent launch
    const int i = n++ % B2H_NSIDE;
    if ((int)ctx->side.size() <= i) {
      int prio = 0; cudaStreamGetPriority(ctx->stream, &prio);     // side streams inherit the priority of their lane
      cudaStream_t s; cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, prio); ctx->side.push_back(s);
      cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); ctx->side_done.push_back(e);
    }
    if (n <= B2H_NSIDE) cudaStreamWaitEvent(ctx->side[i], ctx->fork_ev, 0);
    return ctx->side[i];
  }
  ~ForkJoin() {
    const int used = n < B2H_NSIDE ? n : B2H_NSIDE;
    for (int i = 0; i < used; i++) { cudaEventRecord(ctx->side_done[i], ctx->side[i]); cudaStreamWaitEvent(ctx->stream, ctx->side_done[i], 0); }
  }
};

// SSV register tile of a model of M nodes: a group of G lanes (8, 16 or 32: 4, 2 or 1 comparisons per warp) in which
// every lane owns 2*NR consecutive nodes.  Needs 2*G*NR >= M+1 so that the last cell of the group's last lane is always
// padding (see b2h_msv.cu).  The narrowest group that fits is taken: it wastes the fewest padded cells (granularity
// 2*G nodes) and spreads the per-row shuffle over the most cells.
// Host threads a helper loop of this process may use: <want>, but no more than this rank's fair share of the node's hardware
// threads in a one-process-per-GPU job (torchrun exports LOCAL_WORLD_SIZE) -- eight ranks that each spawn sixteen packing
// threads on a 32-thread host only slow each other down.
static inline int b2h_rank_threads(int want) {
  int hw = (int)std::thread::hardware_concurrency(), ranks = 1;
  if (const char *ev = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(ev));
  return std::max(1, std::min(want, std::max(1, hw / ranks)));
}

// Shared-memory wavefronts one DP row of one COMPARISON costs with a (G, NR) tile: a warp holds 32/G comparisons and moves
// 4 wavefronts per LDS.128, 1 per leftover LDS.32, 1 for the diagonal shuffle (none when a single lane owns the model) and
// 1/4 for the residue-word shuffle.
static inline double b2h_ssv_tile_cost(int G, int NR) {
  return (4.0 * (NR / 4) + (NR % 4) + (G > 1 ? 1.0 : 0.0) + 0.25) * G / 32.0;
}
// The register tile of a model of M nodes: G lanes per comparison, NR packed registers (2 NR cells) per lane, 2 G NR >= M + 1
// (one spare cell: the value that wraps around the group must be a padding cell).  The fewer lanes share a model, the
// fewer shuffles a cell costs and the better the lanes are used, so short models go to groups of 4, 2 or 1 lanes
// (<small_ok>: protein profiles; the long-target kernels keep the 8 / 16 / 32-lane tiles) -- the cheapest tile by the
// wavefront count above among the instantiated ones: G = 1, 2, 4: NR 4 .. 32; G = 8: NR 1 .. 32; G = 16: NR 17 .. 32;
// G = 32: the list below.
static inline bool b2h_ssv_tile(int M, int *G, int *NR, bool small_ok = false) {
  const int need = M + 1;
  int bg = 0, bnr = 0;
  if (need <= 16 * 32) { bg = 8;  bnr = (need + 15) / 16; }
  else if (need <= 32 * 32) { bg = 16; bnr = (need + 31) / 32; }
  else {
    const int nr = (need + 63) / 64;
    static const int allowed[] = {18, 20, 22, 24, 26, 28, 30, 32, 40, 48};
    for (int a : allowed) if (a >= nr) { bg = 32; bnr = a; break; }
    if (!bg) return false;
  }
  if (small_ok) {
    double best = b2h_ssv_tile_cost(bg, bnr);
    for (int g = 4; g >= 1; g >>= 1) {                     // (ties keep the wider group: fewer registers per lane)
      const int nr = std::max(4, (need + 2 * g - 1) / (2 * g));
      if (nr > 32) continue;
      const double c = b2h_ssv_tile_cost(g, nr);
      if (c < best - 1e-9) { best = c; bg = g; bnr = nr; }
    }
  }
  *G = bg; *NR = bnr;
  return true;
}
// bytes of one residue row of the lane-striped SSV table of a (G, NR) tile.  An LDS.128 wavefront serves a quarter-warp
// (8 lanes): with fewer than 8 lanes per group the 8/G groups of a quarter-warp read different residue rows at the same
// in-row offset, so every 16-byte chunk is stored 8/G times side by side -- slot (lane & 7) -- and the eight lanes always
// fall into eight different bank quartets, whatever rows they read.  The row geometry of G < 8 is that of G = 8.
static inline size_t b2h_ssv_row_bytes(int G, int NR) { return (size_t)(NR / 4) * (G < 8 ? 8 : G) * 16 + (size_t)(NR % 4) * 128; }


// ---------------------------------------------------------------------------------------------
// Device-side views (passed to kernels by value or through small device arrays)
// ---------------------------------------------------------------------------------------------
struct SeqDev {
  const uint8_t *res; const int64_t *off; const int32_t *len;
  const uint8_t *tjb; const int16_t *xwmove; const float *pmove, *null1, *p1, *flta, *fltb;
  const int32_t *order; int n;
};
static inline SeqDev b2h_seqdev(const b2h_seqdb *db) {
  SeqDev s; s.res = db->d_res; s.off = db->d_off; s.len = db->d_len; s.tjb = db->d_tjb; s.xwmove = db->d_xwmove;
  s.pmove = db->d_pmove; s.null1 = db->d_null1; s.p1 = db->d_p1; s.flta = db->d_flta; s.fltb = db->d_fltb;
  s.order = db->d_order; s.n = (int)db->n; return s;
}

struct ProfDev {
  const uint32_t *ssv_emis; const uint8_t *msv_cost8;
  const int16_t *vit_rsc, *vit_tsc;
  const float *fwd_rsc, *fwd_tsc, *bias_eo;
  const int32_t *vit_rsc32; const float *fwd_rscr;
  const uint32_t *vit_rsc2; int v2C, v2_ok, tbm_min;     // packed ViterbiFilter: table, nodes per lane (even), usable, min tBM
  int M, Mpad, NR, G;
  const uint32_t *ssv_emis_w; int NRw, Gw;               // the wide tile and its table (scan orientation)
  int tbm, tec, base, bias; float scale_b;
  int xw_E_move, xw_E_loop, base_w, ddbound_w; float scale_w;
  float xf_E_move, xf_E_loop;
  float evparam[B2H_NEVPARAM];
  float bias_t10, bias_t11;
};
static inline ProfDev b2h_profdev(const b2h_profile *p) {
  ProfDev d; d.ssv_emis = p->d_ssv_emis; d.msv_cost8 = p->d_msv_cost8; d.vit_rsc = p->d_vit_rsc; d.vit_tsc = p->d_vit_tsc;
  d.fwd_rsc = p->d_fwd_rsc; d.fwd_tsc = p->d_fwd_tsc; d.bias_eo = p->d_bias_eo; d.vit_rsc32 = p->d_vit_rsc32; d.fwd_rscr = p->d_fwd_rscr;
  d.vit_rsc2 = p->d_vit_rsc2; d.v2C = p->v2C; d.v2_ok = p->v2_ok; d.tbm_min = p->tbm_min;
  d.ssv_emis_w = p->d_ssv_emis_w; d.NRw = p->NRw; d.Gw = p->Gw;
  d.M = p->M; d.Mpad = p->Mpad; d.NR = p->NR; d.G = p->G; d.tbm = p->tbm_b; d.tec = p->tec_b; d.base = p->base_b; d.bias = p->bias_b; d.scale_b = p->scale_b;
  d.xw_E_move = p->xw[0][0]; d.xw_E_loop = p->xw[0][1]; d.base_w = p->base_w; d.ddbound_w = p->ddbound_w; d.scale_w = p->scale_w;
  d.xf_E_move = p->xf[0][0]; d.xf_E_loop = p->xf[0][1];
  for (int i = 0; i < B2H_NEVPARAM; i++) d.evparam[i] = p->evparam[i];
  d.bias_t10 = p->bias_t10; d.bias_t11 = p->bias_t11; return d;
}

// A stage's work: comparisons (entries) grouped by profile.  Entry e of profile p lives at
// ent_s[poff[p] .. poff[p+1]); CTAs pull "items" = chunks of B2H_ITEM_ENTRIES consecutive entries of
// one profile from *counter; itemoff[p] is the first item id of profile p, itemoff[P] the item count.
#define B2H_ITEM_ENTRIES 16
struct WorkList {
  const ProfDev *profs;      // [P]
  const int32_t *ent_s;      // sequence index of each entry
  const int32_t *poff;       // [P+1]
  const int32_t *itemoff;    // [P+1]
  int            P;
  int           *counter;
  int            plo, phi;   // this launch covers profiles [plo, phi) only (profiles are sorted by size; one launch per size class)
};

// A batch of envelopes for the kernels of b2h_envelope.cu: entries [e_lo, e_hi) of one (C, W) size class, largest first.
// Matrices of envelope e: F (3 planes), PP (2 planes), OA (3 planes) of (Ld+1) rows x Mp floats, at plane offset moff[e];
// special-state rows (6 floats) at row offset xoff[e]; trace records at toff[e] (capacity tcap[e]).
struct EnvDev {
  const ProfDev *profs; const int32_t *prof, *seq, *i0, *Ld; const float *pmove;
  const int64_t *moff, *moff_n, *xoff, *toff; const int32_t *tcap;
  float *F, *PP, *OA, *fx, *bx, *ox; uint8_t *BP;
  float *envsc, *oasc, *em, *ei, *xnull; int32_t *status, *tlen; int4 *trace;
  int *counter; int e_lo, e_hi;
  const float *rsc_pool = nullptr; const int64_t *rsc_off = nullptr;   // per-envelope emission odds [Kp][Mpad] (offset in floats, -1: the profile's)
};
int b2h_launch_envelope(b2h_ctx *ctx, int kind, int C, int W, const EnvDev &ev, const SeqDev &sd, cudaStream_t strm);
struct b2h_envclass { int bound, C, W; };
static const b2h_envclass B2H_ENV_CLASSES[] = {{128, 4, 1}, {256, 8, 1}, {384, 12, 1}, {768, 12, 2}, {1536, 12, 4}, {3072, 12, 8}};
static const int B2H_N_ENV_CLASSES = 6;

// per-entry outputs of a DP stage
struct StageOut { float *sc; int32_t *status; float *fwd_xmx, *bck_xmx; const int64_t *xoff;
                  int redo_only = 0; };    // Viterbi: only entries whose status is B2H_REDO (left by the packed 16x2 kernel)
#define B2H_REDO 0x7e00d0     // internal status: "decide this comparison with the exact 32-bit kernel"

// Packed (s16x2) ViterbiFilter, b2h_dpreg.cu: state cells are stored as true value + V2_SIG in 16-bit halves and clamped from
// below at the stored floor V2_FLOOR; table values are clamped from below at V2_TF.  With these constants no 16-bit sum can
// wrap and every value >= V2_LO is exact ("Viterbi in packed 16-bit lanes"); comparisons that leave the safe
// range (a cell >= V2_HI, begin floor below V2_LO, final score below V2_LO) are flagged B2H_REDO.
#define B2H_V2_SIG    2400
#define B2H_V2_LO    (-4401)
#define B2H_V2_FLOOR (B2H_V2_LO + B2H_V2_SIG)      /* -2001, stored coordinates */
#define B2H_V2_TF    (-30767)
#define B2H_V2_HI     26366                        /* = V2_LO - V2_TF */
#define B2H_V2_RMAX   3900                         /* largest emission score a profile may have: V2_HI + V2_SIG + RMAX <= 32767 */

// mpads[p] = Mpad of profile p of the work list, ascending (one launch per size class)
int b2h_launch_viterbi(b2h_ctx *ctx, const WorkList &wl, const SeqDev &sd, const std::vector<int> &mpads, int nitems_hint, StageOut out);
int b2h_launch_forward(b2h_ctx *ctx, const WorkList &wl, const SeqDev &sd, const std::vector<int> &mpads, int nitems_hint, StageOut out);
int b2h_launch_backward(b2h_ctx *ctx, const WorkList &wl, const SeqDev &sd, const std::vector<int> &mpads, int nitems_hint, StageOut out);
int b2h_launch_forward_backward(b2h_ctx *ctx, const WorkList &wl, const SeqDev &sd, const std::vector<int> &mpads, int nitems_hint, StageOut fwd, StageOut bck);
int b2h_launch_dpreg(b2h_ctx *ctx, int kind, int C, int W, const WorkList &wl, const SeqDev &sd, int nitems_hint, StageOut out, cudaStream_t strm);
int b2h_launch_vit2(b2h_ctx *ctx, int C2, const WorkList &wl, const SeqDev &sd, int nitems_hint, StageOut out, cudaStream_t strm);
// register-resident DP size classes: nodes per lane C and warps per comparison W for a model of M nodes (0,0: too long)
struct b2h_regclass { int bound, C, W; };
static const b2h_regclass B2H_REG_CLASSES_FINE[] = {
  {64, 2, 1}, {96, 3, 1}, {128, 4, 1}, {160, 5, 1}, {192, 6, 1}, {224, 7, 1}, {256, 8, 1}, {288, 9, 1}, {320, 10, 1}, {352, 11, 1},
  {384, 12, 1}, {448, 14, 1}, {512, 16, 1},
  {576, 9, 2}, {640, 10, 2}, {704, 11, 2}, {768, 12, 2}, {896, 14, 2}, {1024, 16, 2}, {1280, 10, 4}, {1536, 12, 4}};
static const b2h_regclass B2H_REG_CLASSES_EVEN[] = {
  {64, 2, 1}, {128, 4, 1}, {192, 6, 1}, {256, 8, 1}, {320, 10, 1}, {384, 12, 1}, {448, 14, 1}, {512, 16, 1},
  {640, 10, 2}, {768, 12, 2}, {896, 14, 2}, {1024, 16, 2}, {1280, 10, 4}, {1536, 12, 4}};
static const b2h_regclass B2H_REG_CLASSES_COARSE[] = {{64, 2, 1}, {128, 4, 1}, {256, 8, 1}, {384, 12, 1}, {512, 16, 1},
                                                      {640, 10, 2}, {768, 12, 2}, {1024, 16, 2}, {1536, 12, 4}};
// B2H_REG_CLASSSET = fine | even | coarse (experiments; default below)
static inline int b2h_reg_classes(const b2h_regclass **tab) {
  static int which = -1;
  if (which < 0) { const char *ev = getenv("B2H_REG_CLASSSET"); which = !ev ? 0 : (ev[0] == 'e' ? 1 : ev[0] == 'c' ? 2 : 0); }
  if (which == 1) { *tab = B2H_REG_CLASSES_EVEN; return 14; }
  if (which == 2) { *tab = B2H_REG_CLASSES_COARSE; return 9; }
  *tab = B2H_REG_CLASSES_FINE; return 21;
}
int b2h_launch_bias(b2h_ctx *ctx, const WorkList &wl, const SeqDev &sd, int nentries_hint, float *filtersc);

// A compacted list of comparisons produced by a stage's epilogue (device memory).  Appends are
// atomic (*n, and cnt[p] for the later grouping by profile); a,b carry the stage's scores.
struct SurvList { int32_t *p, *s; float *a, *b; int *n; int *cnt; int cap; int *cnts = nullptr; };   // cnts: optional per-sequence counts

#define B2H_SSV_CHUNK 128      // sequences per SSV work item
struct SsvArgs {
  const ProfDev *profs;        // all profiles of the batch
  const int32_t *cls;          // indices of the profiles of this register-tile class (device)
  int            ncls;
  SeqDev         sd;
  int            chunks;       // ceil(nseq / B2H_SSV_CHUNK)
  int            items_per_cta; // a CTA retires after this many work items (0: persistent): SM slots turn over every ~100 us, so the
                               // short high-priority kernels of the survivor / envelope lanes are not kept waiting by a long SSV launch
  int           *counter;
  int            mode;         // 0: dense p7_SSVFilter  1: dense, queue eslENORESULT in R  2: cascade (P-value test, A and R)
                               // 3: scan orientation: <sd> is a chunk view, the chunk maxima are folded into raw[profile][parent]
  const int32_t *parent = nullptr; int *raw = nullptr; int raw_stride = 0;   // mode 3
  int            threads = 0;  // CTA size (0 = the full 256): a chunk view of one long query holds a handful of chunks per profile
  int            wide = 0;     // 1: (G, NR) is the profiles' WIDE tile, read ProfDev::ssv_emis_w (scan orientation)
  float         *out_sc; int32_t *out_status;
  SurvList       A, R;
  double         F1;
};
int b2h_launch_ssv(b2h_ctx *ctx, int G, int NR, const SsvArgs &a, cudaStream_t strm);
// scan orientation, after the mode-3 launches: p7_SSVFilter's post-processing and the F1 test for every (profile, sequence)
// from raw[P][n] -- appends to A (passed) and R (eslENORESULT: needs the full MSV filter)
int b2h_launch_ssv_finish(b2h_ctx *ctx, const ProfDev *profs, int P, const SeqDev &sd, const int *raw, SurvList A, SurvList R, double F1);
// full MSV (with J) over a grouped work list; mode 1: dense outputs indexed by sequence, 2: cascade append to A
int b2h_launch_msv(b2h_ctx *ctx, const WorkList &wl, const SeqDev &sd, int max_Mpad, int nitems_hint, int mode,
                   float *out_sc, int32_t *out_status, SurvList A, double F1);
int b2h_launch_msv_tiled(b2h_ctx *ctx, const WorkList &wl, const SeqDev &sd, const std::vector<int> &tiles, int mode,
                         float *out_sc, int32_t *out_status, SurvList A, double F1);
// group a SurvList by profile: poff/itemoff[P+1] and the grouped arrays (device)
struct Grouped { int32_t *p, *s; float *a, *b; int32_t *poff, *itemoff; int *fill; };
int b2h_launch_group(b2h_ctx *ctx, const SurvList &in, int P, Grouped out);// p7_pli_ExtendAndMergeWindows on a host list, in place (b2h_longtarget.cu); returns the number of windows left
size_t b2h_extend_merge(const b2h_profile *p, b2h_window *w, size_t n, const int64_t *target_len, float pct_overlap);

// one lane of the reference's 4-lane Cephes expf (esl_sse.c:182-246), as used to build Forward odds ratios (b2h_host.cpp)
float b2h_cephes_expf(float x);


