// b2h_api.cu -- per-stage C-ABI entry points with dense host outputs (parity / diagnostic surface).
#include <vector>
#include "b2h_internal.h"

namespace {

struct DenseOut {
  b2h_ctx *ctx; size_t n;
  float *d_sc = nullptr; int32_t *d_status = nullptr;
  DenseOut(b2h_ctx *c, size_t n_) : ctx(c), n(n_) {}
  int alloc() {
    B2H_CUDA(cudaSetDevice(ctx->device));
    B2H_CUDA(cudaMallocAsync(&d_sc, (n ? n : 1) * sizeof(float), ctx->stream));
    B2H_CUDA(cudaMallocAsync(&d_status, (n ? n : 1) * sizeof(int32_t), ctx->stream));
    return B2H_OK;
  }
  int fetch(float *sc, int32_t *status) {
    if (n && sc)     B2H_CUDA(cudaMemcpyAsync(sc, d_sc, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (n && status) B2H_CUDA(cudaMemcpyAsync(status, d_status, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    B2H_CUDA(cudaStreamSynchronize(ctx->stream));
    return B2H_OK;
  }
  ~DenseOut() { if (d_sc) cudaFreeAsync(d_sc, ctx->stream); if (d_status) cudaFreeAsync(d_status, ctx->stream); }
};

bool args_ok(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *db) {
  return ctx && p && db && p->ctx == ctx && db->ctx == ctx;
}

// One profile against every sequence: the trivial work list (entries in the length-sorted order).
struct DenseList {
  b2h_ctx *ctx; ProfDev *d_prof = nullptr; int32_t *d_poff = nullptr, *d_itemoff = nullptr, *d_cls = nullptr;
  WorkList wl; int nitems = 0;
  DenseList(b2h_ctx *c) : ctx(c) {}
  int build(const b2h_profile *p, const b2h_seqdb *db) {
    ProfDev h = b2h_profdev(p);
    const int n = (int)db->n;
    nitems = (n + B2H_ITEM_ENTRIES - 1) / B2H_ITEM_ENTRIES;
    int32_t poff[2] = {0, n}, itemoff[2] = {0, nitems}, cls0 = 0;
    B2H_CUDA(cudaMallocAsync(&d_prof, sizeof(ProfDev), ctx->stream));
    B2H_CUDA(cudaMallocAsync(&d_poff, sizeof poff, ctx->stream));
    B2H_CUDA(cudaMallocAsync(&d_itemoff, sizeof itemoff, ctx->stream));
    B2H_CUDA(cudaMallocAsync(&d_cls, sizeof cls0, ctx->stream));
    B2H_CUDA(cudaMemcpyAsync(d_cls, &cls0, sizeof cls0, cudaMemcpyHostToDevice, ctx->stream));
    B2H_CUDA(cudaMemcpyAsync(d_prof, &h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
    B2H_CUDA(cudaMemcpyAsync(d_poff, poff, sizeof poff, cudaMemcpyHostToDevice, ctx->stream));
    B2H_CUDA(cudaMemcpyAsync(d_itemoff, itemoff, sizeof itemoff, cudaMemcpyHostToDevice, ctx->stream));
    B2H_CUDA(cudaStreamSynchronize(ctx->stream));           // the host staging variables above are on the stack
    wl.profs = d_prof; wl.ent_s = db->d_order; wl.poff = d_poff; wl.itemoff = d_itemoff; wl.P = 1; wl.counter = ctx->d_counters + 8; wl.plo = 0; wl.phi = 1;
    return B2H_OK;
  }
  ~DenseList() { if (d_prof) cudaFreeAsync(d_prof, ctx->stream); if (d_poff) cudaFreeAsync(d_poff, ctx->stream); if (d_itemoff) cudaFreeAsync(d_itemoff, ctx->stream); if (d_cls) cudaFreeAsync(d_cls, ctx->stream); }
};

// scatter entry-ordered results back to database order
__global__ void unpermute_kernel(const int32_t *order, int n, const float *sc_e, const int32_t *st_e, float *sc, int32_t *st)
{
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) { const int s = order[e]; if (sc) sc[s] = sc_e[e]; if (st) st[s] = st_e ? st_e[e] : 0; }
}

typedef int (*dp_launcher)(b2h_ctx *, const WorkList &, const SeqDev &, const std::vector<int> &, int, StageOut);

int dense_dp(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *db, dp_launcher launch, bool backward, float *sc, int32_t *status)
{
  if (!args_ok(ctx, p, db)) return B2H_EINVAL;
  const size_t n = db->n;
  if (n == 0) return B2H_OK;
  DenseOut o(ctx, n), oe(ctx, n), of(ctx, n);
  int st;
  if ((st = o.alloc()) != B2H_OK || (st = oe.alloc()) != B2H_OK) return st;
  DenseList dl(ctx);
  if ((st = dl.build(p, db)) != B2H_OK) return st;
  SeqDev sd = b2h_seqdev(db);
  const std::vector<int> mp(1, p->Mpad);
  StageOut so; so.sc = oe.d_sc; so.status = oe.d_status; so.fwd_xmx = nullptr; so.bck_xmx = nullptr; so.xoff = nullptr;
  float *d_fx = nullptr; int64_t *d_xoff = nullptr;
  if (backward) {
    // Forward with stored specials first: the Backward parser reuses its per-row scale factors
    if ((st = of.alloc()) != B2H_OK) return st;
    std::vector<int64_t> xoff(n);
    int64_t tot = 0;
    std::vector<int32_t> order(n);
    B2H_CUDA(cudaMemcpyAsync(order.data(), db->d_order, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    B2H_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t e = 0; e < n; e++) { xoff[e] = tot; tot += db->h_len[order[e]] + 1; }
    B2H_CUDA(cudaMallocAsync(&d_fx, (size_t)tot * 6 * sizeof(float), ctx->stream));
    B2H_CUDA(cudaMallocAsync(&d_xoff, n * sizeof(int64_t), ctx->stream));
    B2H_CUDA(cudaMemcpyAsync(d_xoff, xoff.data(), n * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    B2H_CUDA(cudaStreamSynchronize(ctx->stream));
    StageOut sf = so; sf.sc = of.d_sc; sf.status = of.d_status; sf.fwd_xmx = d_fx; sf.xoff = d_xoff;
    st = b2h_launch_forward(ctx, dl.wl, sd, mp, dl.nitems, sf);
    if (st == B2H_OK) { so.fwd_xmx = d_fx; so.xoff = d_xoff; st = launch(ctx, dl.wl, sd, mp, dl.nitems, so); }
  } else {
    st = launch(ctx, dl.wl, sd, mp, dl.nitems, so);
  }
  if (st == B2H_OK) {
    unpermute_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(db->d_order, (int)n, oe.d_sc, oe.d_status, o.d_sc, o.d_status);
    ctx->launches++;
    st = o.fetch(sc, status);
  }
  if (d_fx) cudaFreeAsync(d_fx, ctx->stream);
  if (d_xoff) cudaFreeAsync(d_xoff, ctx->stream);
  return st;
}

} // namespace

extern "C" {

// SSV (mode 0) or MSV = SSV + full-MSV fallback (mode 1) for one profile against the whole database
static int dense_msv(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *db, int mode, float *sc, int32_t *status)
{
  if (!args_ok(ctx, p, db)) return B2H_EINVAL;
  const size_t n = db->n;
  if (n == 0) return B2H_OK;
  DenseOut o(ctx, n);
  int st = o.alloc();                                       if (st != B2H_OK) return st;
  DenseList dl(ctx);
  if ((st = dl.build(p, db)) != B2H_OK) return st;
  // redo list R + its grouping
  int32_t *buf = nullptr; int *ctr = nullptr;
  B2H_CUDA(cudaMallocAsync(&buf, (4 * n + 8) * sizeof(int32_t), ctx->stream));
  B2H_CUDA(cudaMallocAsync(&ctr, 8 * sizeof(int), ctx->stream));
  B2H_CUDA(cudaMemsetAsync(ctr, 0, 8 * sizeof(int), ctx->stream));
  SurvList R; R.p = buf; R.s = buf + n; R.a = nullptr; R.b = nullptr; R.n = ctr; R.cnt = ctr + 1; R.cap = (int)n;
  Grouped G; G.p = buf + 2 * n; G.s = buf + 3 * n; G.a = nullptr; G.b = nullptr; G.poff = buf + 4 * n; G.itemoff = buf + 4 * n + 2; G.fill = ctr + 2;
  SsvArgs a;
  a.profs = dl.d_prof; a.cls = dl.d_cls; a.ncls = 1; a.sd = b2h_seqdev(db);
  a.chunks = (int)((n + B2H_SSV_CHUNK - 1) / B2H_SSV_CHUNK); a.counter = ctx->d_counters; a.items_per_cta = 0;
  a.mode = mode; a.out_sc = o.d_sc; a.out_status = o.d_status; a.R = R; a.A = R; a.F1 = 1.0;
  st = b2h_launch_ssv(ctx, p->G, p->NR, a, ctx->stream);
  if (st == B2H_OK && mode == 1) {
    st = b2h_launch_group(ctx, R, 1, G);
    if (st == B2H_OK) {
      WorkList wl = dl.wl; wl.ent_s = G.s; wl.poff = G.poff; wl.itemoff = G.itemoff;
      static const bool smem_msv = getenv("B2H_MSV_SMEM") != nullptr;
      if (smem_msv) st = b2h_launch_msv(ctx, wl, a.sd, p->Mpad, 0, 1, o.d_sc, o.d_status, R, 1.0);
      else st = b2h_launch_msv_tiled(ctx, wl, a.sd, std::vector<int>(1, p->G * 64 + p->NR), 1, o.d_sc, o.d_status, R, 1.0);
    }
  }
  if (st == B2H_OK) st = o.fetch(sc, status);
  cudaFreeAsync(buf, ctx->stream); cudaFreeAsync(ctr, ctx->stream);
  return st;
}

int b2h_ssv_filter(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *db, float *sc, int32_t *status)
{ return dense_msv(ctx, p, db, 0, sc, status); }
int b2h_msv_filter(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *db, float *sc, int32_t *status)
{ return dense_msv(ctx, p, db, 1, sc, status); }

int b2h_viterbi_filter(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *db, float *sc, int32_t *status)
{ return dense_dp(ctx, p, db, b2h_launch_viterbi, false, sc, status); }
int b2h_forward_parser(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *db, float *sc, int32_t *status)
{ return dense_dp(ctx, p, db, b2h_launch_forward, false, sc, status); }
int b2h_backward_parser(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *db, float *sc, int32_t *status)
{ return dense_dp(ctx, p, db, b2h_launch_backward, true, sc, status); }

int b2h_null_scores(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *db, float *null1, float *filtersc)
{
  if (!args_ok(ctx, p, db)) return B2H_EINVAL;
  const size_t n = db->n;
  if (n == 0) return B2H_OK;
  if (null1) B2H_CUDA(cudaMemcpyAsync(null1, db->d_null1, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (filtersc) {
    DenseOut o(ctx, n), oe(ctx, n);
    int st;
    if ((st = o.alloc()) != B2H_OK || (st = oe.alloc()) != B2H_OK) return st;
    DenseList dl(ctx);
    if ((st = dl.build(p, db)) != B2H_OK) return st;
    if ((st = b2h_launch_bias(ctx, dl.wl, b2h_seqdev(db), (int)n, oe.d_sc)) != B2H_OK) return st;
    unpermute_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(db->d_order, (int)n, oe.d_sc, nullptr, o.d_sc, nullptr);
    ctx->launches++;
    return o.fetch(filtersc, nullptr);
  }
  B2H_CUDA(cudaStreamSynchronize(ctx->stream));
  return B2H_OK;
}

} // extern "C"
