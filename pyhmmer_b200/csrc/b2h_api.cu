// b2h_api.cu -- per-stage C-ABI entry points with dense host outputs (parity / diagnostic surface).
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

} // namespace

extern "C" {

int b2h_ssv_filter(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *db, float *sc, int32_t *status)
{
  if (!args_ok(ctx, p, db)) return B2H_EINVAL;
  DenseOut o(ctx, db->n);
  int st = o.alloc();                                       if (st != B2H_OK) return st;
  st = b2h_launch_ssv_dense(ctx, p, db, 0, o.d_sc, o.d_status);   if (st != B2H_OK) return st;
  return o.fetch(sc, status);
}

int b2h_msv_filter(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *db, float *sc, int32_t *status)
{
  if (!args_ok(ctx, p, db)) return B2H_EINVAL;
  DenseOut o(ctx, db->n);
  int st = o.alloc();                                       if (st != B2H_OK) return st;
  st = b2h_launch_ssv_dense(ctx, p, db, 1, o.d_sc, o.d_status);   if (st != B2H_OK) return st;
  return o.fetch(sc, status);
}

} // extern "C"
