/* b2h_pyhmmer_glue.h -- see b2h_pyhmmer_glue.c */
#ifndef B2H_PYHMMER_GLUE_H
#define B2H_PYHMMER_GLUE_H
#include "hmmer.h"
#include "impl_sse/impl_sse.h"
#include "b2h.h"
int b2h_glue_upload_oprofile(b2h_ctx *ctx, const P7_OPROFILE *om, const P7_BG *bg, b2h_profile **out);
int b2h_glue_seqdb(b2h_ctx *ctx, ESL_SQ *const *sq, size_t n, b2h_seqdb **out);
int b2h_glue_search_loop(b2h_ctx *ctx, const b2h_seqdb *db, P7_PIPELINE *pli, P7_OPROFILE *om, P7_BG *bg,
                         ESL_SQ *const *sq, size_t n_targets, P7_TOPHITS *th, unsigned seed, int host_threads);
int b2h_glue_scan_loop(b2h_ctx *ctx, const b2h_profile *const *profs, P7_PIPELINE *pli, const ESL_SQ *sq, P7_BG *bg,
                       P7_OPROFILE *const *om, size_t n_targets, P7_TOPHITS *th, unsigned seed, int host_threads);
int b2h_glue_longtarget_fill(P7_PIPELINE *pli, P7_TOPHITS *th, const void *hits, size_t nh, const void *doms, const char *text,
                             const unsigned char *dup, ESL_SQ *const *sq, P7_OPROFILE *om, P7_BG *bg,
                             int64_t nseqs, int64_t nres, const int64_t *pos_past);
void b2h_glue_seqdb_destroy(b2h_seqdb *db);
void b2h_glue_profile_destroy(b2h_profile *p);
#endif
