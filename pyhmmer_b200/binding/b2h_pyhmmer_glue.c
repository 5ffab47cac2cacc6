/* b2h_pyhmmer_glue.c -- the reference-side half of the binding: pyhmmer's own structures in, pyhmmer's own structures out.
 *
 * Compiled against the headers an installed pyhmmer ships (pyhmmer.libs/include: hmmer.h, easel.h, impl_sse.h) and linked
 * with its liblibhmmer / liblibeasel, plus libb2h.so.  Two entry points stand in for the static loops of pyhmmer's
 * Pipeline (src/pyhmmer/plan7.pyx):
 *
 *   b2h_glue_search_loop   Pipeline._search_loop  (plan7.pyx:6394-6453): one P7_OPROFILE against ESL_SQ*[n]
 *   b2h_glue_scan_loop     Pipeline._scan_loop    (plan7.pyx:6625-6677): one ESL_SQ against P7_OPROFILE*[n]
 *
 * Contract: every reportable hit is appended to <th> exactly as p7_Pipeline would (p7_tophits_CreateNextHit,
 * fields of p7_pipeline.c:840-931), the accounting of <pli> (nseqs, nres, nmodels, nnodes, n_past_*, Z) is left as the
 * sequential loop leaves it, the return value is an Easel status.  All DP runs on the GPU behind b2h_search().
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "p7_config.h"
#include "easel.h"
#include "esl_alphabet.h"
#include "esl_sq.h"
#include "hmmer.h"
#include "impl_sse/impl_sse.h"
#include "b2h.h"
#include <pthread.h>

/* pyhmmer's hmmsearch / hmmscan run one Pipeline per worker THREAD with the GIL released; a b2h context drives one GPU with
 * its own streams and lanes and takes one call at a time, so the workers of a process queue up here (each call fills the
 * whole device anyway). */
static pthread_mutex_t g_engine_mu = PTHREAD_MUTEX_INITIALIZER;
#define ENGINE_LOCK()   pthread_mutex_lock(&g_engine_mu)
#define ENGINE_UNLOCK() pthread_mutex_unlock(&g_engine_mu)
void b2h_glue_seqdb_destroy(b2h_seqdb *db)     { ENGINE_LOCK(); b2h_seqdb_destroy(db);  ENGINE_UNLOCK(); }
void b2h_glue_profile_destroy(b2h_profile *p)  { ENGINE_LOCK(); b2h_profile_destroy(p); ENGINE_UNLOCK(); }

/* ---- P7_OPROFILE -> device profile: de-stripe the SSE tables, copy the scalars (include/b2h.h: b2h_oprofile_desc) ---- */
int b2h_glue_upload_oprofile(b2h_ctx *ctx, const P7_OPROFILE *om, const P7_BG *bg, b2h_profile **out)
{
  const ESL_ALPHABET *abc = om->abc;
  const int M = om->M, K = abc->K, Kp = abc->Kp;
  b2h_oprofile_desc d;
  uint8_t *msv = NULL, *degen = NULL; int16_t *vr = NULL, *vt = NULL; float *fr = NULL, *ft = NULL;
  int x, y, status = B2H_EMEM;
  memset(&d, 0, sizeof d);
  msv = malloc((size_t)Kp * M); vr = malloc(sizeof(int16_t) * (size_t)Kp * M); vt = malloc(sizeof(int16_t) * 8 * (size_t)M);
  fr = malloc(sizeof(float) * (size_t)Kp * M); ft = malloc(sizeof(float) * 8 * (size_t)M); degen = malloc((size_t)Kp * K);
  if (!msv || !vr || !vt || !fr || !ft || !degen) goto DONE;
  status = b2h_destripe_oprofile(M, Kp, (const uint8_t *)om->rbv[0], (const int16_t *)om->rwv[0], (const int16_t *)om->twv,
                                 (const float *)om->rfv[0], (const float *)om->tfv, msv, vr, vt, fr, ft);
  if (status != B2H_OK) goto DONE;
  for (x = 0; x < Kp; x++) for (y = 0; y < K; y++) degen[x * K + y] = (uint8_t)abc->degen[x][y];
  d.M = M; d.K = K; d.Kp = Kp; d.L = om->L; d.mode_multihit = (om->mode == p7_LOCAL || om->mode == p7_GLOCAL) ? 1 : 0; d.max_length = om->max_length;
  d.msv_cost = msv; d.tbm_b = om->tbm_b; d.tec_b = om->tec_b; d.tjb_b = om->tjb_b; d.base_b = om->base_b; d.bias_b = om->bias_b; d.scale_b = om->scale_b;
  d.vit_rsc = vr; d.vit_tsc = vt; memcpy(d.xw, om->xw, sizeof d.xw); d.base_w = om->base_w; d.ddbound_w = om->ddbound_w; d.scale_w = om->scale_w;
  d.fwd_rsc = fr; d.fwd_tsc = ft; memcpy(d.xf, om->xf, sizeof d.xf);
  memcpy(d.evparam, om->evparam, sizeof d.evparam); memcpy(d.cutoff, om->cutoff, sizeof d.cutoff);
  for (x = 0; x < B2H_MAXABET; x++) d.compo[x] = (x < p7_MAXABET) ? om->compo[x] : 0.f;
  for (x = 0; x < K && x < B2H_MAXABET; x++) d.bgf[x] = bg->f[x];
  d.degen = degen;
  ENGINE_LOCK();
  status = b2h_profile_upload(ctx, &d, out);
  ENGINE_UNLOCK();
  if (status == B2H_OK)
    b2h_profile_set_annotation(*out, om->consensus + 1, (om->rf && om->rf[0]) ? om->rf + 1 : NULL, (om->cs && om->cs[0]) ? om->cs + 1 : NULL, abc->sym);
DONE:
  free(msv); free(vr); free(vt); free(fr); free(ft); free(degen);
  return status;
}

int b2h_glue_seqdb(b2h_ctx *ctx, ESL_SQ *const *sq, size_t n, b2h_seqdb **out)
{
  const uint8_t **dsq = malloc(sizeof(uint8_t *) * (n + 1));
  int64_t *len = malloc(sizeof(int64_t) * (n + 1));
  size_t i; int status;
  if (!dsq || !len) { free(dsq); free(len); return B2H_EMEM; }
  for (i = 0; i < n; i++) { dsq[i] = sq[i]->dsq; len[i] = sq[i]->n; }
  ENGINE_LOCK();
  status = b2h_seqdb_create(ctx, dsq, len, n, out);
  ENGINE_UNLOCK();
  free(dsq); free(len);
  return status;
}

/* ---- one hit record -> P7_HIT with its P7_DOMAINs and P7_ALIDISPLAYs (p7_pipeline.c:840-931, p7_alidisplay.c:92-273) ---- */
static char *dupz(const char *s) { char *r = NULL; esl_strdup(s ? s : "", -1, &r); return r; }

static P7_ALIDISPLAY *make_alidisplay(const b2h_domain *d, const char *text, const P7_OPROFILE *om, const ESL_SQ *sq)
{
  P7_ALIDISPLAY *ad = p7_alidisplay_Create_empty();
  const char *t = text + d->text_offset;
  const size_t w = (size_t)d->N + 1;
  if (!ad) return NULL;
  ad->N = d->N;                                  /* "deserialized" form: every string its own allocation, mem == NULL */
  ad->model  = dupz(t); ad->mline = dupz(t + w); ad->aseq = dupz(t + 2 * w); ad->ppline = dupz(t + 3 * w);
  ad->rfline = d->has_rf ? dupz(t + 4 * w) : NULL;
  ad->csline = d->has_cs ? dupz(t + (4 + (d->has_rf ? 1 : 0)) * w) : NULL;
  ad->mmline = NULL; ad->ntseq = NULL;
  ad->hmmname = dupz(om->name); ad->hmmacc = dupz(om->acc); ad->hmmdesc = dupz(om->desc);
  ad->hmmfrom = d->hmmfrom; ad->hmmto = d->hmmto; ad->M = om->M;
  ad->sqname = dupz(sq->name); ad->sqacc = dupz(sq->acc); ad->sqdesc = dupz(sq->desc);
  ad->sqfrom = d->sqfrom; ad->sqto = d->sqto; ad->L = sq->n;
  ad->memsize = 0; ad->mem = NULL;
  return ad;
}

static int fill_hit(P7_PIPELINE *pli, P7_TOPHITS *th, const b2h_hit *h, const b2h_domain *doms, const char *text,
                    const ESL_SQ *sq, const P7_OPROFILE *om)
{
  P7_HIT *hit = NULL;
  int d, status;
  if ((status = p7_tophits_CreateNextHit(th, &hit)) != eslOK) return status;
  if (pli->mode == p7_SEARCH_SEQS) {
    hit->name = dupz(sq->name);
    if (sq->acc[0]  != '\0') hit->acc  = dupz(sq->acc);
    if (sq->desc[0] != '\0') hit->desc = dupz(sq->desc);
  } else {
    hit->name = dupz(om->name); hit->acc = dupz(om->acc); hit->desc = dupz(om->desc);
  }
  hit->ndom = h->ndom; hit->nexpected = h->nexpected; hit->nregions = h->nregions; hit->nclustered = h->nclustered;
  hit->noverlaps = h->noverlaps; hit->nenvelopes = h->nenvelopes;
  hit->pre_score = h->pre_score; hit->pre_lnP = h->pre_lnP;
  hit->score = h->score; hit->lnP = h->lnP;
  hit->sortkey = pli->inc_by_E ? -h->lnP : h->score;
  hit->sum_score = h->sum_score; hit->sum_lnP = h->sum_lnP;
  hit->best_domain = h->best_domain;
  hit->dcl = calloc((size_t)(h->ndom > 0 ? h->ndom : 1), sizeof(P7_DOMAIN));
  if (!hit->dcl) return eslEMEM;
  for (d = 0; d < h->ndom; d++) {
    const b2h_domain *s = doms + h->dom_offset + d;
    P7_DOMAIN *q = &hit->dcl[d];
    q->ienv = s->ienv; q->jenv = s->jenv; q->iali = s->iali; q->jali = s->jali; q->iorf = 0; q->jorf = 0;
    q->envsc = s->envsc; q->domcorrection = s->domcorrection; q->dombias = s->dombias; q->oasc = s->oasc;
    q->bitscore = s->bitscore; q->lnP = s->lnP; q->is_reported = FALSE; q->is_included = FALSE; q->scores_per_pos = NULL;
    if ((q->ad = make_alidisplay(s, text, om, sq)) == NULL) return eslEMEM;
  }
  if (pli->use_bit_cutoffs) {
    if (p7_pli_TargetReportable(pli, hit->score, hit->lnP)) {
      hit->flags |= p7_IS_REPORTED;
      if (p7_pli_TargetIncludable(pli, hit->score, hit->lnP)) hit->flags |= p7_IS_INCLUDED;
    }
    for (d = 0; d < hit->ndom; d++)
      if (p7_pli_DomainReportable(pli, hit->dcl[d].bitscore, hit->dcl[d].lnP)) {
        hit->dcl[d].is_reported = TRUE;
        if (p7_pli_DomainIncludable(pli, hit->dcl[d].bitscore, hit->dcl[d].lnP)) hit->dcl[d].is_included = TRUE;
      }
  }
  return eslOK;
}

static void params_of(const P7_PIPELINE *pli, unsigned seed, int host_threads, int seq_counters, b2h_search_params *prm)
{
  memset(prm, 0, sizeof *prm);
  prm->F1 = pli->F1; prm->F2 = pli->F2; prm->F3 = pli->F3;
  prm->do_biasfilter = pli->do_biasfilter; prm->do_null2 = pli->do_null2;
  prm->seed = pli->do_reseeding ? seed : 0; prm->host_threads = host_threads; prm->seq_counters = seq_counters;
}

/* Pipeline._search_loop.  <db> = the targets, resident (b2h_glue_seqdb of the same ESL_SQ array). */
int b2h_glue_search_loop(b2h_ctx *ctx, const b2h_seqdb *db, P7_PIPELINE *pli, P7_OPROFILE *om, P7_BG *bg,
                         ESL_SQ *const *sq, size_t n_targets, P7_TOPHITS *th, unsigned seed, int host_threads)
{
  b2h_profile *prof = NULL; b2h_results *res = NULL; b2h_search_params prm;
  const b2h_profile *plist[1];
  const b2h_hit *hits; const b2h_domain *doms; const char *text; const int64_t *ctr;
  size_t i, nh; uint64_t nseqs0 = pli->nseqs; int status;
  if ((status = p7_pli_NewModel(pli, om, bg)) != eslOK) return status;          /* thresholds of this model, nmodels, nnodes */
  if (n_targets == 0) return eslOK;
  if ((status = b2h_glue_upload_oprofile(ctx, om, bg, &prof)) != B2H_OK) return status;
  params_of(pli, seed, host_threads, 0, &prm);
  plist[0] = prof;
  ENGINE_LOCK();
  status = b2h_search(ctx, plist, 1, db, &prm, &res);
  ENGINE_UNLOCK();
  if (status != B2H_OK) { ENGINE_LOCK(); b2h_profile_destroy(prof); ENGINE_UNLOCK(); return status; }
  nh = b2h_results_nhits(res); hits = b2h_results_hits(res); doms = b2h_results_domains(res); text = b2h_results_text(res, NULL);
  ctr = b2h_results_counters(res);
  /* accounting: what p7_pli_NewSeq does for every target (p7_pipeline.c:576), what p7_Pipeline counts (:725-770) */
  for (i = 0; i < n_targets; i++) pli->nres += sq[i]->n;
  pli->n_past_msv += ctr[0]; pli->n_past_bias += ctr[1]; pli->n_past_vit += ctr[2]; pli->n_past_fwd += ctr[3];
  for (i = 0; i < nh && status == eslOK; i++) {                                  /* in target order, with the running Z */
    const b2h_hit *h = &hits[i];
    pli->nseqs = nseqs0 + (uint64_t)h->seq + 1;
    if (pli->Z_setby == p7_ZSETBY_NTARGETS) pli->Z = (double)pli->nseqs;
    if (p7_pli_TargetReportable(pli, h->score, h->lnP)) status = fill_hit(pli, th, h, doms, text, sq[h->seq], om);
  }
  pli->nseqs = nseqs0 + n_targets;
  if (pli->Z_setby == p7_ZSETBY_NTARGETS) pli->Z = (double)pli->nseqs;
  b2h_results_destroy(res);
  ENGINE_LOCK(); b2h_profile_destroy(prof); ENGINE_UNLOCK();
  return status;
}

/* Pipeline._scan_loop.  <profs> = the device profiles of om[0..n) (b2h_glue_upload_oprofile each, cached by the caller). */
int b2h_glue_scan_loop(b2h_ctx *ctx, const b2h_profile *const *profs, P7_PIPELINE *pli, const ESL_SQ *sq, P7_BG *bg,
                       P7_OPROFILE *const *om, size_t n_targets, P7_TOPHITS *th, unsigned seed, int host_threads)
{
  b2h_seqdb *db = NULL; b2h_results *res = NULL; b2h_search_params prm;
  ESL_SQ *one[1];
  const b2h_hit *hits; const b2h_domain *doms; const char *text; const int64_t *ctr;
  size_t i, nh, t = 0; int status;
  p7_pli_NewSeq(pli, sq);
  if (n_targets == 0) return eslOK;
  one[0] = (ESL_SQ *)sq;
  if ((status = b2h_glue_seqdb(ctx, one, 1, &db)) != B2H_OK) return status;
  params_of(pli, seed, host_threads, 1, &prm);
  ENGINE_LOCK();
  status = b2h_search(ctx, profs, n_targets, db, &prm, &res);
  ENGINE_UNLOCK();
  if (status != B2H_OK) { ENGINE_LOCK(); b2h_seqdb_destroy(db); ENGINE_UNLOCK(); return status; }
  nh = b2h_results_nhits(res); hits = b2h_results_hits(res); doms = b2h_results_domains(res); text = b2h_results_text(res, NULL);
  ctr = b2h_results_seq_counters(res);
  if (ctr) { pli->n_past_msv += ctr[0]; pli->n_past_bias += ctr[1]; pli->n_past_vit += ctr[2]; pli->n_past_fwd += ctr[3]; }
  /* every model in order: p7_pli_NewModel sets its thresholds and the running Z (p7_pipeline.c:497-530), then its hit if any */
  for (i = 0; i < nh && status == eslOK; i++) {
    const b2h_hit *h = &hits[i];                                                  /* sorted by (profile, seq) */
    for (; t <= (size_t)h->profile && status == eslOK; t++) status = p7_pli_NewModel(pli, om[t], bg);
    if (status == eslOK && p7_pli_TargetReportable(pli, h->score, h->lnP)) status = fill_hit(pli, th, h, doms, text, sq, om[h->profile]);
  }
  for (; t < n_targets && status == eslOK; t++) status = p7_pli_NewModel(pli, om[t], bg);
  b2h_results_destroy(res);
  ENGINE_LOCK(); b2h_seqdb_destroy(db); ENGINE_UNLOCK();
  return status;
}


/* LongTargetsPipeline.search_hmm (plan7.pyx:7258-7412) behind its window loop: the engine's long-target search has produced
 * one record per hit -- a single domain in target coordinates (start > end on the reverse strand), its lnP already carrying
 * the search-space term of p7_tophits_ComputeNhmmerEvalues (p7_tophits.c:796), duplicates of overlapping windows already
 * marked as p7_tophits_RemoveDuplicates (:823) marks them, hits ordered by target and alignment position -- and this
 * function turns them into the P7_HITs p7_pli_postDomainDef / postViterbi_LongTarget would have appended
 * (p7_pipeline.c:1195-1262), with the accounting of the window loop.  The caller sorts by key and thresholds, as the
 * reference does. */
int b2h_glue_longtarget_fill(P7_PIPELINE *pli, P7_TOPHITS *th, const void *hits_v, size_t nh, const void *doms_v, const char *text,
                             const unsigned char *dup, ESL_SQ *const *sq, P7_OPROFILE *om, P7_BG *bg,
                             int64_t nseqs, int64_t nres, const int64_t *pos_past)
{
  const b2h_hit *hits = (const b2h_hit *)hits_v; const b2h_domain *doms = (const b2h_domain *)doms_v;
  size_t i; int status;
  if ((status = p7_pli_NewModel(pli, om, bg)) != eslOK) return status;          /* thresholds of this model, nmodels, nnodes */
  pli->nseqs += (uint64_t)nseqs; pli->nres += (uint64_t)nres;
  pli->pos_past_msv += pos_past[0]; pli->pos_past_bias += pos_past[1]; pli->pos_past_vit += pos_past[2]; pli->pos_past_fwd += pos_past[3];
  for (i = 0; i < nh; i++) {
    const b2h_hit *h = &hits[i];
    P7_HIT *hit;
    const size_t before = th->N;
    if ((status = fill_hit(pli, th, h, doms, text, sq[h->seq], om)) != eslOK) return status;
    if (th->N != before + 1) return eslEINCONCEIVABLE;
    hit = &th->unsrt[th->N - 1];
    hit->sortkey = -h->lnP;                                                      /* p7_tophits_ComputeNhmmerEvalues: always by E-value */
    hit->window_length = om->max_length;
    hit->seqidx = h->seq;
    hit->subseq_start = 1;
    if (dup && dup[i]) { hit->flags |= p7_IS_DUPLICATE; hit->flags &= ~(p7_IS_REPORTED | p7_IS_INCLUDED); }
  }
  return eslOK;
}
