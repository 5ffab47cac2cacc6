/* oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A flat C interface (ctypes-friendly: plain pointers and sizes) over the
 * UNMODIFIED reference C library (HMMER 3.4 + Easel, compiled from
 * /root/reference/vendor by oracle/Makefile).  Nothing in pyhmmer_b200/ may
 * link or load this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * The call sequences mirror what pyhmmer itself does around each reference
 * function (cited inline), so that the values returned here are the values a
 * pyhmmer user would observe.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "p7_config.h"
#include "easel.h"
#include "esl_alphabet.h"
#include "esl_sq.h"
#include "esl_hmm.h"
#include "esl_random.h"
#include "esl_randomseq.h"
#include "esl_gumbel.h"
#include "esl_exponential.h"
#include "esl_vectorops.h"
#include "hmmer.h"

typedef struct {
  ESL_ALPHABET *abc;
  P7_BG        *bg;
  P7_HMM       *hmm;
  P7_PROFILE   *gm;
  P7_OPROFILE  *om;
  P7_OMX       *ox;    /* scratch for single-comparison calls */
  P7_OMX       *oxb;
} REFM;

static int g_inited = 0;
static void ref_init(void) {
  if (!g_inited) { p7_FLogsumInit(); g_inited = 1; }   /* as `import pyhmmer.plan7` does (plan7.pyx:9968-9986); impl_Init() is NOT called */
}

/* Read the idx-th (0-based) HMM of an HMMER3 ASCII/binary file and configure it the way
 * Pipeline.search_hmm does for an HMM query: Profile.configure(hmm, bg, L) =
 * p7_ProfileConfig(hmm,bg,gm,L,p7_LOCAL) (plan7.pyx:8082), then OptimizedProfile.convert =
 * p7_oprofile_Convert (plan7.pyx:4961). */
REFM *refm_read(const char *path, int idx, int L)
{
  P7_HMMFILE *hfp = NULL;
  REFM *m = calloc(1, sizeof(REFM));
  int i, status;
  ref_init();
  if (p7_hmmfile_Open(path, NULL, &hfp, NULL) != eslOK) { free(m); return NULL; }
  for (i = 0; i <= idx; i++) {
    if (m->hmm) { p7_hmm_Destroy(m->hmm); m->hmm = NULL; }
    status = p7_hmmfile_Read(hfp, &m->abc, &m->hmm);
    if (status != eslOK) { p7_hmmfile_Close(hfp); free(m); return NULL; }
  }
  p7_hmmfile_Close(hfp);
  m->bg = p7_bg_Create(m->abc);
  m->gm = p7_profile_Create(m->hmm->M, m->abc);
  m->om = p7_oprofile_Create(m->hmm->M, m->abc);
  p7_ProfileConfig(m->hmm, m->bg, m->gm, L, p7_LOCAL);
  p7_oprofile_Convert(m->gm, m->om);
  m->ox  = p7_omx_Create(m->hmm->M, 0, 400);
  m->oxb = p7_omx_Create(m->hmm->M, 0, 400);
  return m;
}

void refm_free(REFM *m)
{
  if (!m) return;
  p7_omx_Destroy(m->ox); p7_omx_Destroy(m->oxb);
  p7_oprofile_Destroy(m->om); p7_profile_Destroy(m->gm);
  p7_bg_Destroy(m->bg); p7_hmm_Destroy(m->hmm); esl_alphabet_Destroy(m->abc);
  free(m);
}

int  refm_M(const REFM *m)   { return m->om->M; }
int  refm_K(const REFM *m)   { return m->abc->K; }
int  refm_Kp(const REFM *m)  { return m->abc->Kp; }
int  refm_abc_type(const REFM *m) { return m->abc->type; }
const char *refm_name(const REFM *m) { return m->hmm->name; }
const char *refm_acc(const REFM *m)  { return m->hmm->acc; }
const char *refm_desc(const REFM *m) { return m->hmm->desc; }
int  refm_max_length(const REFM *m) { return m->om->max_length; }
void refm_evparam(const REFM *m, float *out6) { memcpy(out6, m->om->evparam, 6*sizeof(float)); }
void refm_cutoff(const REFM *m, float *out6)  { memcpy(out6, m->om->cutoff, 6*sizeof(float)); }
void refm_compo(const REFM *m, float *out20)  { memcpy(out20, m->om->compo, p7_MAXABET*sizeof(float)); }
void refm_bg_f(const REFM *m, float *out)     { memcpy(out, m->bg->f, m->abc->K*sizeof(float)); }

/* raw HMM parameters: t[(M+1)*7], mat[(M+1)*K], ins[(M+1)*K] */
void refm_hmm_params(const REFM *m, float *t, float *mat, float *ins)
{
  int M = m->hmm->M, K = m->abc->K, k;
  for (k = 0; k <= M; k++) {
    memcpy(t   + k*7, m->hmm->t[k],   7*sizeof(float));
    memcpy(mat + k*K, m->hmm->mat[k], K*sizeof(float));
    memcpy(ins + k*K, m->hmm->ins[k], K*sizeof(float));
  }
}
/* generic profile: tsc[(M+1)*8], rsc[Kp*(M+1)*2], xsc[4*2] */
void refm_gm_params(const REFM *m, float *tsc, float *rsc, float *xsc)
{
  int M = m->gm->M, Kp = m->abc->Kp, x;
  memcpy(tsc, m->gm->tsc, (size_t)(M+1)*p7P_NTRANS*sizeof(float));
  for (x = 0; x < Kp; x++) memcpy(rsc + (size_t)x*(M+1)*2, m->gm->rsc[x], (size_t)(M+1)*2*sizeof(float));
  memcpy(xsc, m->gm->xsc, 8*sizeof(float));
}

/* optimized-profile scalars: out[0..] = tbm_b tec_b tjb_b base_b bias_b | base_w ddbound_w | xw[4][2] */
void refm_om_ints(const REFM *m, int *out)
{
  const P7_OPROFILE *om = m->om; int i, j, n = 0;
  out[n++] = om->tbm_b; out[n++] = om->tec_b; out[n++] = om->tjb_b; out[n++] = om->base_b; out[n++] = om->bias_b;
  out[n++] = om->base_w; out[n++] = om->ddbound_w;
  for (i = 0; i < 4; i++) for (j = 0; j < 2; j++) out[n++] = om->xw[i][j];
  out[n++] = om->L; out[n++] = om->mode;
}
/* out = scale_b scale_w ncj_roundoff nj | xf[4][2] */
void refm_om_floats(const REFM *m, float *out)
{
  const P7_OPROFILE *om = m->om; int i, j, n = 0;
  out[n++] = om->scale_b; out[n++] = om->scale_w; out[n++] = om->ncj_roundoff; out[n++] = om->nj;
  for (i = 0; i < 4; i++) for (j = 0; j < 2; j++) out[n++] = om->xf[i][j];
}
/* Striped tables, copied out as stored by the reference (impl_sse.h:75-100).
 * sizes: rbv Kp*Q16*16 u8 ; sbv Kp*(Q16+p7O_EXTRA_SB)*16 u8 ; rwv Kp*Q8*8 i16 ; twv 8*Q8*8 i16 ;
 *        rfv Kp*Q4*4 f32 ; tfv 8*Q4*4 f32 */
int refm_Q16(const REFM *m) { return p7O_NQB(m->om->M); }
int refm_Q8 (const REFM *m) { return p7O_NQW(m->om->M); }
int refm_Q4 (const REFM *m) { return p7O_NQF(m->om->M); }
int refm_extra_sb(void)     { return p7O_EXTRA_SB; }
void refm_om_tables(const REFM *m, uint8_t *rbv, uint8_t *sbv, int16_t *rwv, int16_t *twv, float *rfv, float *tfv)
{
  const P7_OPROFILE *om = m->om;
  int Kp = m->abc->Kp, x;
  int Q16 = p7O_NQB(om->M), Q8 = p7O_NQW(om->M), Q4 = p7O_NQF(om->M);
  for (x = 0; x < Kp; x++) {
    if (rbv) memcpy(rbv + (size_t)x*Q16*16, om->rbv[x], (size_t)Q16*16);
    if (sbv) memcpy(sbv + (size_t)x*(Q16+p7O_EXTRA_SB)*16, om->sbv[x], (size_t)(Q16+p7O_EXTRA_SB)*16);
    if (rwv) memcpy(rwv + (size_t)x*Q8*8, om->rwv[x], (size_t)Q8*16);
    if (rfv) memcpy(rfv + (size_t)x*Q4*4, om->rfv[x], (size_t)Q4*16);
  }
  if (twv) memcpy(twv, om->twv, (size_t)8*Q8*16);
  if (tfv) memcpy(tfv, om->tfv, (size_t)8*Q4*16);
}

/* length / mode reconfiguration, exactly the calls the search loop makes (plan7.pyx:6431-6436) */
void refm_set_length(REFM *m, int L) { p7_bg_SetLength(m->bg, L); p7_oprofile_ReconfigLength(m->om, L); }
void refm_set_multihit(REFM *m, int L, int multihit)
{ if (multihit) p7_oprofile_ReconfigMultihit(m->om, L); else p7_oprofile_ReconfigUnihit(m->om, L); }

/* ---- single-comparison filters; dsq is 1..L with sentinels at 0 and L+1 (esl_sq.h:100-102) ---- */
static void grow(REFM *m, int L) { p7_omx_GrowTo(m->ox, m->om->M, 0, L); p7_omx_GrowTo(m->oxb, m->om->M, 0, L); }

int ref_ssv(REFM *m, const uint8_t *dsq, int L, float *sc) { refm_set_length(m, L); return p7_SSVFilter(dsq, L, m->om, sc); }
int ref_msv(REFM *m, const uint8_t *dsq, int L, float *sc) { grow(m, L); refm_set_length(m, L); return p7_MSVFilter(dsq, L, m->om, m->ox, sc); }
int ref_vit(REFM *m, const uint8_t *dsq, int L, float *sc) { grow(m, L); refm_set_length(m, L); return p7_ViterbiFilter(dsq, L, m->om, m->ox, sc); }
int ref_fwd(REFM *m, const uint8_t *dsq, int L, float *sc) { grow(m, L); refm_set_length(m, L); return p7_ForwardParser(dsq, L, m->om, m->ox, sc); }
/* Forward then Backward parser; optionally returns the xmx specials (each (L+1)*p7X_NXCELLS floats) */
int ref_fwdbck(REFM *m, const uint8_t *dsq, int L, float *fsc, float *bsc, float *fx, float *bx)
{
  int st;
  grow(m, L); refm_set_length(m, L);
  st = p7_ForwardParser(dsq, L, m->om, m->ox, fsc);            if (st != eslOK) return st;
  st = p7_BackwardParser(dsq, L, m->om, m->ox, m->oxb, bsc);   if (st != eslOK) return st;
  if (fx) memcpy(fx, m->ox->xmx,  (size_t)(L+1)*p7X_NXCELLS*sizeof(float));
  if (bx) memcpy(bx, m->oxb->xmx, (size_t)(L+1)*p7X_NXCELLS*sizeof(float));
  return eslOK;
}
int ref_nxcells(void) { return p7X_NXCELLS; }

/* null1 (p7_bg.c:357) and bias-filter (p7_bg.c:471) scores, set up as p7_pli_NewModel + the search loop do */
float ref_null1(REFM *m, const uint8_t *dsq, int L)
{ float sc; p7_bg_SetLength(m->bg, L); p7_bg_NullOne(m->bg, dsq, L, &sc); return sc; }
float ref_bias(REFM *m, const uint8_t *dsq, int L)
{
  float sc;
  p7_bg_SetFilter(m->bg, m->om->M, m->om->compo);   /* p7_pli_NewModel, p7_pipeline.c:504 */
  p7_bg_SetLength(m->bg, L);                        /* per sequence, plan7.pyx:6431 */
  p7_bg_FilterScore(m->bg, dsq, L, &sc);
  return sc;
}

/* generic (unstriped, log-space) reference DP on the P7_PROFILE; needs length config on gm */
int ref_generic(REFM *m, const uint8_t *dsq, int L, float *gmsv, float *gvit, float *gfwd, float *gbck)
{
  P7_GMX *gx = p7_gmx_Create(m->gm->M, L);
  p7_ReconfigLength(m->gm, L);
  if (gmsv) p7_GMSV    (dsq, L, m->gm, gx, 2.0, gmsv);
  if (gvit) p7_GViterbi(dsq, L, m->gm, gx, gvit);
  if (gfwd) p7_GForward(dsq, L, m->gm, gx, gfwd);
  if (gbck) p7_GBackward(dsq, L, m->gm, gx, gbck);
  p7_gmx_Destroy(gx);
  return eslOK;
}

/* p7_GDecoding (generic_decoding.c:77) after p7_GForward / p7_GBackward: pp_dp [(L+1)][(M+1)][3] (M,I,D), pp_xmx [(L+1)][5] (E,N,J,B,C) */
/* dom = NULL or [3][L+1]: btot, etot, mocc of p7_GDomainDecoding (generic_decoding.c:207) */
int ref_gdecoding(REFM *m, const uint8_t *dsq, int L, float *pp_dp, float *pp_xmx, float *fsc, float *bsc, float *dom)
{
  int M = m->gm->M, i;
  P7_GMX *fwd = p7_gmx_Create(M, L), *bck = p7_gmx_Create(M, L), *pp = p7_gmx_Create(M, L);
  p7_ReconfigLength(m->gm, L);
  p7_GForward (dsq, L, m->gm, fwd, fsc);
  p7_GBackward(dsq, L, m->gm, bck, bsc);
  p7_GDecoding(m->gm, fwd, bck, pp);
  if (dom) {
    P7_DOMAINDEF *dd = p7_domaindef_Create(NULL);
    p7_domaindef_GrowTo(dd, L);
    dd->btot[0] = dd->etot[0] = dd->mocc[0] = 0.;
    p7_GDomainDecoding(m->gm, fwd, bck, dd);
    memcpy(dom, dd->btot, sizeof(float) * (L + 1)); memcpy(dom + (L + 1), dd->etot, sizeof(float) * (L + 1)); memcpy(dom + 2 * (L + 1), dd->mocc, sizeof(float) * (L + 1));
    p7_domaindef_Destroy(dd);
  }
  for (i = 0; i <= L; i++) {
    memcpy(pp_dp  + (size_t)i * (M + 1) * p7G_NSCELLS, pp->dp[i], sizeof(float) * (M + 1) * p7G_NSCELLS);
    memcpy(pp_xmx + (size_t)i * p7G_NXCELLS, pp->xmx + (size_t)i * p7G_NXCELLS, sizeof(float) * p7G_NXCELLS);
  }
  p7_gmx_Destroy(fwd); p7_gmx_Destroy(bck); p7_gmx_Destroy(pp);
  return eslOK;
}

/* nhmmer's first stage for one target chunk (p7_Pipeline_LongTarget, p7_pipeline.c:1535-1565): p7_oprofile_ReconfigMSVLength
 * to the model's max_length, p7_SSVFilter_longtarget (impl_sse/msvfilter.c:256), then p7_hmm_ScoreDataComputeRest +
 * p7_pli_ExtendAndMergeWindows(.., 0).  raw [cap][3] = n, k, length of every SSV diagonal (+ raw_sc), merged [cap][2] = n,
 * length of the merged windows; also the prefix / suffix length tables [M+1] when asked for.  Returns counts via n_raw/n_merged. */
int ref_longtarget_windows(REFM *m, const uint8_t *dsq, int L, double F1, int cap, int *n_raw, int64_t *raw, float *raw_sc,
                           int *n_merged, int64_t *merged, float *prefix, float *suffix)
{
  P7_SCOREDATA *data = p7_hmm_ScoreDataCreate(m->om, NULL);
  P7_HMM_WINDOWLIST wl;
  int i, M = m->om->M;
  wl.windows = NULL;
  p7_hmmwindow_init(&wl);
  p7_omx_GrowTo(m->ox, M, 0, m->om->max_length);
  p7_oprofile_ReconfigMSVLength(m->om, m->om->max_length);
  p7_SSVFilter_longtarget(dsq, L, m->om, m->ox, data, m->bg, F1, &wl);
  *n_raw = wl.count;
  for (i = 0; i < wl.count && i < cap; i++) {
    raw[i*3+0] = wl.windows[i].n; raw[i*3+1] = wl.windows[i].k; raw[i*3+2] = wl.windows[i].length; raw_sc[i] = wl.windows[i].score;
  }
  p7_hmm_ScoreDataComputeRest(m->om, data);
  if (prefix) memcpy(prefix, data->prefix_lengths, sizeof(float) * (M + 1));
  if (suffix) memcpy(suffix, data->suffix_lengths, sizeof(float) * (M + 1));
  p7_pli_ExtendAndMergeWindows(m->om, data, &wl, 0);
  *n_merged = wl.count;
  for (i = 0; i < wl.count && i < cap; i++) { merged[i*2+0] = wl.windows[i].n; merged[i*2+1] = wl.windows[i].length; }
  free(wl.windows);
  p7_hmm_ScoreDataDestroy(data);
  return eslOK;
}

/* The window-level stages of nhmmer for one target chunk, restated from the reference's own (static) functions with the
 * reference's public calls in the same order, so that every intermediate is observable:
 *   p7_Pipeline_LongTarget (p7_pipeline.c:1496-1706): SSV scan, extend/merge, then per window null1, bias FilterScore,
 *     p7_MSVFilter at the window's length, F1 gate;
 *   p7_pli_postSSV_LongTarget (:1331-1436): B1-scaled bias gate, null1 at min(window, max_length), B2-scaled filter score,
 *     p7_oprofile_ReconfigRestLength, p7_ViterbiFilter_longtarget, p7_pli_ExtendAndMergeWindows(.., 0.5), 80 kb splitting;
 *   p7_pli_postViterbi_LongTarget (:1065-1112): null1, bias, ReconfigRestLength(window), p7_ForwardParser, B3-scaled F3 gate.
 * ref_longtarget_pipeline (below) runs the real p7_Pipeline_LongTarget; tests check that both give the same pos_past_* counters.
 *   msvwin  [cap][2] = n, length           msvsc [cap][3] = null1, FilterScore, MSV score     msvflag: bit0 passed F1, bit1 passed bias gate
 *   vithit  [cap][3] = msv window, i, k    (every landmark p7_ViterbiFilter_longtarget records, in its order)
 *   vitwin  [cap][3] = msv window, n (in the msv window's coordinates), length      vitsc [cap][3] = null1, FilterScore, Forward score
 *   vitpass [cap]    = passed F3          counters [4] = pos_past_msv, pos_past_bias, pos_past_vit, pos_past_fwd */
int ref_longtarget_stages(REFM *m, const uint8_t *dsq, int L, double F1, double F2, double F3, int do_bias, int B1, int B2, int B3, int cap,
                          int *n_msvwin, int64_t *msvwin, float *msvsc, int *msvflag,
                          int *n_vithit, int64_t *vithit, int *n_vitwin, int64_t *vitwin, float *vitsc, int *vitpass, long *counters)
{
  P7_OPROFILE *om = m->om;
  P7_BG *bg = m->bg;
  P7_SCOREDATA *data = p7_hmm_ScoreDataCreate(om, NULL);
  P7_HMM_WINDOWLIST wl, vl;
  int w, i, nh = 0, nv = 0;
  counters[0] = counters[1] = counters[2] = counters[3] = 0;
  wl.windows = NULL; vl.windows = NULL;
  p7_hmmwindow_init(&wl);
  p7_bg_SetFilter(bg, om->M, om->compo);                 /* p7_pli_NewModel (p7_pipeline.c:514) */
  p7_omx_GrowTo(m->ox, om->M, 0, om->max_length);
  p7_oprofile_ReconfigMSVLength(om, om->max_length);
  p7_SSVFilter_longtarget(dsq, L, om, m->ox, data, bg, F1, &wl);
  *n_msvwin = 0; *n_vithit = 0; *n_vitwin = 0;
  if (wl.count > 0) {
    p7_hmm_ScoreDataComputeRest(om, data);
    p7_pli_ExtendAndMergeWindows(om, data, &wl, 0);
    p7_hmmwindow_init(&vl);
    *n_msvwin = wl.count;
    for (w = 0; w < wl.count; w++) {
      P7_HMM_WINDOW *win = wl.windows + w;
      const ESL_DSQ *subseq = dsq + win->n - 1;
      int window_len = win->length;
      float nullsc, bias_filtersc, usc, filtersc, seq_score;
      double P;
      if (w < cap) { msvwin[w*2] = win->n; msvwin[w*2+1] = win->length; msvflag[w] = 0; }
      p7_bg_SetLength(bg, window_len);
      p7_bg_NullOne(bg, subseq, window_len, &nullsc);
      p7_bg_FilterScore(bg, subseq, window_len, &bias_filtersc);
      p7_oprofile_ReconfigMSVLength(om, window_len);
      p7_omx_GrowTo(m->ox, om->M, 0, window_len);
      p7_MSVFilter(subseq, window_len, om, m->ox, &usc);
      if (w < cap) { msvsc[w*3] = nullsc; msvsc[w*3+1] = bias_filtersc; msvsc[w*3+2] = usc; }
      P = esl_gumbel_surv((usc - nullsc) / eslCONST_LOG2, om->evparam[p7_MMU], om->evparam[p7_MLAMBDA]);
      if (P > F1) continue;
      counters[0] += window_len;
      if (w < cap) msvflag[w] |= 1;
      {  /* p7_pli_postSSV_LongTarget */
        int max_window_len = 80000;
        int overlap_len = ESL_MIN(40000, om->max_length);
        int F1_L = ESL_MIN(window_len, B1), F2_L = ESL_MIN(window_len, B2);
        int loc_window_len, overlap;
        if (do_bias) {
          p7_bg_SetLength(bg, window_len);
          p7_bg_FilterScore(bg, subseq, window_len, &bias_filtersc);
          bias_filtersc -= nullsc;
          filtersc = nullsc + (bias_filtersc * (float)((F1_L > window_len ? 1.0 : (float)F1_L / window_len)));
          seq_score = (usc - filtersc) / eslCONST_LOG2;
          P = esl_gumbel_surv(seq_score, om->evparam[p7_MMU], om->evparam[p7_MLAMBDA]);
          if (P > F1) continue;
        } else bias_filtersc = 0;
        counters[1] += window_len;
        if (w < cap) msvflag[w] |= 2;
        loc_window_len = ESL_MIN(window_len, om->max_length);
        p7_bg_SetLength(bg, loc_window_len);
        p7_bg_NullOne(bg, subseq, loc_window_len, &nullsc);
        filtersc = nullsc + (bias_filtersc * (F2_L > window_len ? 1.0 : (float)F2_L / window_len));
        p7_oprofile_ReconfigRestLength(om, loc_window_len);
        p7_omx_GrowTo(m->ox, om->M, 0, window_len);
        p7_ViterbiFilter_longtarget((ESL_DSQ *)subseq, window_len, om, m->ox, filtersc, F2, &vl);
        for (i = 0; i < vl.count; i++, nh++)
          if (nh < cap) { vithit[nh*3] = w; vithit[nh*3+1] = vl.windows[i].n; vithit[nh*3+2] = vl.windows[i].k; }
        p7_pli_ExtendAndMergeWindows(om, data, &vl, 0.5);
        for (i = 0; i < vl.count; i++) {
          if (vl.windows[i].length > max_window_len) {
            uint64_t new_n = vl.windows[i].n; uint32_t new_len = vl.windows[i].length;
            vl.windows[i].length = max_window_len;
            do {
              int shift = max_window_len - overlap_len;
              new_n += shift; new_len -= shift;
              p7_hmmwindow_new(&vl, 0, new_n, 0, 0, ESL_MIN(max_window_len, new_len), 0.0, p7_NOCOMPLEMENT, new_len);
            } while (new_len > max_window_len);
          }
        }
        overlap = 0;
        for (i = 0; i < vl.count; i++, nv++) {
          int vlen = vl.windows[i].length, passed = 0;
          const ESL_DSQ *vsub = subseq + vl.windows[i].n - 1;
          float vnull, vbias, fwdsc;
          int F3_L = ESL_MIN(vlen, B3);
          counters[2] += vlen;
          if (i > 0) counters[2] -= ESL_MAX(0, vl.windows[i-1].n + vl.windows[i-1].length - vl.windows[i].n);
          /* p7_pli_postViterbi_LongTarget up to the F3 gate */
          p7_bg_SetLength(bg, vlen);
          p7_bg_NullOne(bg, vsub, vlen, &vnull);
          if (do_bias) { p7_bg_FilterScore(bg, vsub, vlen, &vbias); vbias -= vnull; } else vbias = 0;
          p7_oprofile_ReconfigRestLength(om, vlen);
          p7_omx_GrowTo(m->ox, om->M, 0, vlen);
          p7_ForwardParser(vsub, vlen, om, m->ox, &fwdsc);
          filtersc = vnull + (vbias * (F3_L > vlen ? 1.0 : (float)F3_L / vlen));
          seq_score = (fwdsc - filtersc) / eslCONST_LOG2;
          P = esl_exp_surv(seq_score, om->evparam[p7_FTAU], om->evparam[p7_FLAMBDA]);
          if (P <= F3) { passed = 1; counters[3] += vlen - overlap; }
          if (nv < cap) {
            vitwin[nv*3] = w; vitwin[nv*3+1] = vl.windows[i].n; vitwin[nv*3+2] = vlen;
            vitsc[nv*3] = vnull; vitsc[nv*3+1] = vbias + vnull; vitsc[nv*3+2] = fwdsc; vitpass[nv] = passed;
          }
          if (passed && i < vl.count - 1) overlap = ESL_MAX(0, vl.windows[i].n + vl.windows[i].length - vl.windows[i+1].n);
          else overlap = 0;
        }
      }
    }
    free(vl.windows);
  }
  *n_vithit = nh; *n_vitwin = nv;
  free(wl.windows);
  p7_hmm_ScoreDataDestroy(data);
  return eslOK;
}

/* p7_ViterbiFilter_longtarget (impl_sse/vitfilter.c:292) on one window, the profile's length model set for <cfg_len>
 * (p7_oprofile_ReconfigRestLength; the pipeline passes min(window, max_length), p7_pipeline.c:1369-1385).
 * hit [cap][2] = i, k of every landmark in the order the reference records them; *thresh = the int16 score threshold it derived. */
int ref_vit_longtarget(REFM *m, const uint8_t *dsq, int L, int cfg_len, float filtersc, double P, int cap, int64_t *hit)
{
  P7_HMM_WINDOWLIST vl;
  int i, n;
  vl.windows = NULL;
  p7_hmmwindow_init(&vl);
  p7_oprofile_ReconfigRestLength(m->om, cfg_len);
  p7_omx_GrowTo(m->ox, m->om->M, 0, L);
  p7_ViterbiFilter_longtarget((ESL_DSQ *)dsq, L, m->om, m->ox, filtersc, P, &vl);
  n = vl.count;
  for (i = 0; i < n && i < cap; i++) { hit[i*2] = vl.windows[i].n; hit[i*2+1] = vl.windows[i].k; }
  free(vl.windows);
  return n;
}

/* The real thing: p7_Pipeline_LongTarget on one chunk with a long-target P7_PIPELINE, as LongTargetsPipeline's search loop
 * drives it (plan7.pyx:7568-7643; top strand).  counters [5] = pos_past_msv, pos_past_bias, pos_past_vit, pos_past_fwd, hits;
 * hits [cap][12] = ienv, jenv, iali, jali, score (bits), bias (dombias), pre_score, lnP, envsc, oasc, hmmfrom, hmmto of every
 * hit appended (unsorted), text = their alignment displays (model | mline | aseq | ppline, NUL-terminated, back to back)
 * when asked for; sq->start = <start> (so that the coordinate arithmetic is exercised), complement as given (the
 * caller passes the reverse-complemented residues, as esl_sq_ReverseComplement leaves them, and start = the chunk's LAST
 * coordinate). */
int ref_longtarget_pipeline(REFM *m, const uint8_t *dsq, int L, double F1, double F2, double F3, int do_bias, int do_null2,
                            long start, int complement, long *counters, int cap, double *hits, char *text, long textcap)
{
  P7_PIPELINE *pli = p7_pipeline_Create(NULL, m->om->M, 100, TRUE, p7_SEARCH_SEQS);
  P7_TOPHITS *th = p7_tophits_Create();
  P7_SCOREDATA *data = p7_hmm_ScoreDataCreate(m->om, NULL);
  ESL_SQ *sq = esl_sq_CreateDigitalFrom(m->abc, "chunk", dsq, L, "", "", NULL);
  int status;
  long h;
  pli->F1 = F1; pli->F2 = F2; pli->F3 = F3;
  pli->do_biasfilter = do_bias; pli->do_null2 = do_null2;
  pli->E = 1e300; pli->domE = 1e300; pli->incE = 1e300; pli->incdomE = 1e300;
  esl_sq_SetSource(sq, "chunk");
  sq->start = start; sq->end = complement ? start - L + 1 : start + L - 1; sq->C = 0; sq->W = L; sq->L = -1;
  p7_pli_NewModel(pli, m->om, m->bg);
  p7_pli_NewSeq(pli, sq);
  status = p7_Pipeline_LongTarget(pli, m->om, data, m->bg, th, 0, sq, complement ? p7_COMPLEMENT : p7_NOCOMPLEMENT, NULL, NULL, NULL);
  counters[0] = pli->pos_past_msv; counters[1] = pli->pos_past_bias; counters[2] = pli->pos_past_vit; counters[3] = pli->pos_past_fwd;
  counters[4] = th->N;
  long tpos = 0;
  for (h = 0; h < (long)th->N && h < cap; h++) {
    P7_HIT *hit = th->unsrt + h;
    if (text) {           /* model | mline | aseq | ppline, each N+1 bytes */
      P7_ALIDISPLAY *ad = hit->dcl[0].ad;
      const char *ln[4] = { ad->model, ad->mline, ad->aseq, ad->ppline };
      int q;
      for (q = 0; q < 4; q++) if (tpos + ad->N + 1 <= textcap) { memcpy(text + tpos, ln[q], ad->N + 1); tpos += ad->N + 1; }
    }
    hits[h*12+0] = hit->dcl[0].ienv; hits[h*12+1] = hit->dcl[0].jenv; hits[h*12+2] = hit->dcl[0].iali; hits[h*12+3] = hit->dcl[0].jali;
    hits[h*12+4] = hit->score; hits[h*12+5] = hit->dcl[0].dombias; hits[h*12+6] = hit->pre_score; hits[h*12+7] = hit->lnP;
    hits[h*12+8] = hit->dcl[0].envsc; hits[h*12+9] = hit->dcl[0].oasc; hits[h*12+10] = hit->dcl[0].ad->hmmfrom; hits[h*12+11] = hit->dcl[0].ad->hmmto;
  }
  esl_sq_Destroy(sq);
  p7_hmm_ScoreDataDestroy(data);
  p7_tophits_Destroy(th);
  p7_pipeline_Destroy(pli);
  return status;
}

/* nhmmer as pyhmmer runs it: LongTargetsPipeline.search_hmm (src/pyhmmer/plan7.pyx:7258-7412) with its window loop
 * _search_loop_longtargets (:7541-7663) restated in C over the reference's own functions: every target is cut into windows
 * of <block_length> residues that keep max_length residues of context, each window goes through p7_Pipeline_LongTarget on
 * the strands asked for, then p7_tophits_ComputeNhmmerEvalues, SortBySeqidxAndAlipos, RemoveDuplicates, SortBySortkey,
 * Threshold.  (idlen_list_assign only fills the display's sequence length and is skipped.)
 * strands: 0 both, 1 top only, 2 bottom only.  out [cap] in final (sorted) order; stats [6] = nres, nseqs, pos_past_msv,
 * pos_past_bias, pos_past_vit, pos_past_fwd. */
typedef struct {
  long   seqidx, ienv, jenv, iali, jali, hmmfrom, hmmto;
  float  score, bias, pre_score, envsc, oasc;
  double lnP;
  int    flags, dom_reported, dom_included, pad;
} REF_LTHIT;

long ref_nhmmer(REFM *m, int nseq, const uint8_t **dsq, const long *len, long block_length, int strands,
                double F1, double F2, double F3, int do_bias, int do_null2, double E, double incE, long evalue_window,
                long cap, REF_LTHIT *out, long *stats, const char *const *names, const char *table_prefix)
{
  P7_OPROFILE *om = m->om;
  P7_PIPELINE *pli = p7_pipeline_Create(NULL, om->M, 100, TRUE, p7_SEARCH_SEQS);
  P7_TOPHITS *th = p7_tophits_Create();
  P7_SCOREDATA *data = p7_hmm_ScoreDataCreate(om, NULL);
  ESL_SQ *tmpsq = esl_sq_CreateDigital(m->abc);
  long C = om->max_length, W = block_length, i, h, nout;
  int t;
  pli->F1 = F1; pli->F2 = F2; pli->F3 = F3;
  pli->do_biasfilter = do_bias; pli->do_null2 = do_null2;
  pli->E = E; pli->incE = incE;
  pli->strands = (strands == 1) ? p7_STRAND_TOPONLY : (strands == 2) ? p7_STRAND_BOTTOMONLY : p7_STRAND_BOTH;
  pli->block_length = (int)W;
  pli->nseqs = 0;
  if (C <= 0 || W <= C) return -1;
  p7_pli_NewModel(pli, om, m->bg);
  for (t = 0; t < nseq; t++) {
    char name[256];
    if (names && names[t]) snprintf(name, sizeof name, "%s", names[t]); else snprintf(name, sizeof name, "seq%d", t);
    tmpsq->idx = t; tmpsq->L = -1;
    esl_sq_SetAccession(tmpsq, ""); esl_sq_SetName(tmpsq, name); esl_sq_SetDesc(tmpsq, ""); esl_sq_SetSource(tmpsq, name);
    esl_sq_GrowTo(tmpsq, ESL_MIN(W + C, len[t]));
    for (i = 0; i < len[t]; i += W - C) {
      tmpsq->C = (i == 0) ? 0 : ESL_MIN(C, len[t] - i);
      tmpsq->W = ESL_MIN(W, len[t] - i - tmpsq->C);
      tmpsq->n = tmpsq->C + tmpsq->W;
      tmpsq->start = i + 1;
      tmpsq->end = i + tmpsq->n;
      memcpy(tmpsq->dsq + 1, dsq[t] + i + 1, tmpsq->n);
      tmpsq->dsq[0] = tmpsq->dsq[tmpsq->n + 1] = eslDSQ_SENTINEL;
      p7_pli_NewSeq(pli, tmpsq);
      if (pli->strands != p7_STRAND_BOTTOMONLY) {
        pli->nres -= tmpsq->C;
        p7_Pipeline_LongTarget(pli, om, data, m->bg, th, pli->nseqs, tmpsq, p7_NOCOMPLEMENT, NULL, NULL, NULL);
        p7_pipeline_Reuse(pli);
      } else pli->nres -= tmpsq->n;
      if (pli->strands != p7_STRAND_TOPONLY) {
        esl_sq_ReverseComplement(tmpsq);
        p7_Pipeline_LongTarget(pli, om, data, m->bg, th, pli->nseqs, tmpsq, p7_COMPLEMENT, NULL, NULL, NULL);
        p7_pipeline_Reuse(pli);
        pli->nres += tmpsq->W;
      }
    }
    esl_sq_Reuse(tmpsq);
    pli->nseqs++;
  }
  /* the window the E-values count in: om->max_length for profile queries, p7_Builder_MaxLength(hmm, beta) for HMM queries
   * (plan7.pyx:7345-7354) -- the caller says which */
  p7_tophits_ComputeNhmmerEvalues(th, (double)pli->nres, evalue_window > 0 ? (int)evalue_window : om->max_length);
  p7_tophits_SortBySeqidxAndAlipos(th);
  p7_tophits_RemoveDuplicates(th, TRUE);
  p7_tophits_SortBySortkey(th);
  p7_tophits_Threshold(th, pli);
  if (table_prefix) {     /* the tables pyhmmer's TopHits.write gives for these hits; idlen_list_assign (nhmmer.c) = the target's length */
    char path[1024]; FILE *fp; int q;
    for (h = 0; h < (long)th->N; h++) th->hit[h]->dcl[0].ad->L = len[th->hit[h]->seqidx];
    for (q = 0; q < 2; q++) {
      snprintf(path, sizeof path, "%s%s", table_prefix, q ? ".pfam" : ".tbl");
      if ((fp = fopen(path, "w")) == NULL) return -2;
      if (q == 0) p7_tophits_TabularTargets(fp, om->name, om->acc, th, pli, TRUE);
      else        p7_tophits_TabularXfam(fp, om->name, om->acc, th, pli);
      fclose(fp);
    }
  }
  stats[0] = pli->nres; stats[1] = pli->nseqs; stats[2] = pli->pos_past_msv; stats[3] = pli->pos_past_bias;
  stats[4] = pli->pos_past_vit; stats[5] = pli->pos_past_fwd;
  nout = th->N;
  for (h = 0; h < nout && h < cap; h++) {
    P7_HIT *hit = th->hit[h];
    REF_LTHIT *o = out + h;
    o->seqidx = hit->seqidx; o->ienv = hit->dcl[0].ienv; o->jenv = hit->dcl[0].jenv; o->iali = hit->dcl[0].iali; o->jali = hit->dcl[0].jali;
    o->hmmfrom = hit->dcl[0].ad->hmmfrom; o->hmmto = hit->dcl[0].ad->hmmto;
    o->score = hit->score; o->bias = hit->dcl[0].dombias; o->pre_score = hit->pre_score; o->envsc = hit->dcl[0].envsc; o->oasc = hit->dcl[0].oasc;
    o->lnP = hit->lnP; o->flags = hit->flags; o->dom_reported = hit->dcl[0].is_reported; o->dom_included = hit->dcl[0].is_included; o->pad = 0;
  }
  esl_sq_Destroy(tmpsq);
  p7_hmm_ScoreDataDestroy(data);
  p7_tophits_Destroy(th);
  p7_pipeline_Destroy(pli);
  return nout;
}

/* p7_Builder_MaxLength on the model's HMM (p7_builder.c:651); the HMM's own max_length is restored. */
int ref_max_length(REFM *m, double emit_thresh)
{
  int saved = m->hmm->max_length, r;
  p7_Builder_MaxLength(m->hmm, emit_thresh);
  r = m->hmm->max_length;
  m->hmm->max_length = saved;
  return r;
}

/* hmmpress's two profile files for every model of an HMM file, written by the reference's own p7_oprofile_Write
 * (impl_sse/io.c:87) after the conversion hmmpress does (hmmpress.c:150-170: p7_ProfileConfig(hmm, bg, gm, 400, p7_LOCAL),
 * p7_oprofile_Convert): <outbase>.h3f + <outbase>.h3p.  Returns the number of models, or -1. */
int ref_press(const char *hmmpath, const char *outbase)
{
  P7_HMMFILE *hfp = NULL;
  ESL_ALPHABET *abc = NULL;
  P7_HMM *hmm = NULL;
  P7_BG *bg = NULL;
  FILE *ffp, *pfp;
  char path[1024];
  int n = 0;
  ref_init();
  if (p7_hmmfile_Open(hmmpath, NULL, &hfp, NULL) != eslOK) return -1;
  snprintf(path, sizeof path, "%s.h3f", outbase); ffp = fopen(path, "wb");
  snprintf(path, sizeof path, "%s.h3p", outbase); pfp = fopen(path, "wb");
  if (!ffp || !pfp) return -1;
  while (p7_hmmfile_Read(hfp, &abc, &hmm) == eslOK) {
    P7_PROFILE *gm = p7_profile_Create(hmm->M, abc);
    P7_OPROFILE *om = p7_oprofile_Create(hmm->M, abc);
    if (!bg) bg = p7_bg_Create(abc);
    p7_ProfileConfig(hmm, bg, gm, 400, p7_LOCAL);
    p7_oprofile_Convert(gm, om);
    om->offs[p7_MOFFSET] = 0; om->offs[p7_FOFFSET] = ftello(ffp); om->offs[p7_POFFSET] = ftello(pfp);
    p7_oprofile_Write(ffp, pfp, om);
    p7_oprofile_Destroy(om); p7_profile_Destroy(gm); p7_hmm_Destroy(hmm); hmm = NULL;
    n++;
  }
  fclose(ffp); fclose(pfp);
  p7_hmmfile_Close(hfp);
  if (bg) p7_bg_Destroy(bg);
  if (abc) esl_alphabet_Destroy(abc);
  return n;
}

/* Builder.build for a single query sequence (plan7.pyx:1150-1260 -> p7_SingleBuilder, p7_builder.c:440): a builder with the
 * given score matrix and gap probabilities, p7_SingleBuilder (p7_Seqmodel, composition, consensus, calibration with the
 * Mersenne Twister seeded <seed>), the resulting HMM written in ASCII to <path>.  alphabet_type: eslAMINO 3, eslDNA 2, eslRNA 1. */
int ref_single_builder(int alphabet_type, const uint8_t *dsq, int L, const char *name, const char *matrix, double popen, double pextend,
                       unsigned seed, const char *path)
{
  ESL_ALPHABET *abc = esl_alphabet_Create(alphabet_type);
  P7_BG *bg = p7_bg_Create(abc);
  P7_BUILDER *bld = p7_builder_Create(NULL, abc);
  ESL_SQ *sq = esl_sq_CreateDigitalFrom(abc, name, dsq, L, NULL, NULL, NULL);
  P7_HMM *hmm = NULL;
  FILE *fp;
  int status;
  ref_init();
  if (seed != 42) { esl_randomness_Destroy(bld->r); bld->r = esl_randomness_CreateFast(seed); }   /* as p7_builder_Create makes it */
  bld->do_reseeding = (seed != 0);
  bld->w_len = -1; bld->w_beta = p7_DEFAULT_WINDOW_BETA;     /* what pyhmmer's Builder sets (plan7.pyx:840-849); Create(NULL, ..) leaves them unset */
  if ((status = p7_builder_LoadScoreSystem(bld, matrix, popen, pextend, bg)) != eslOK) return status;
  if ((status = p7_SingleBuilder(bld, sq, bg, &hmm, NULL, NULL, NULL)) != eslOK) return status;
  if ((fp = fopen(path, "w")) == NULL) return -1;
  p7_hmmfile_WriteASCII(fp, -1, hmm);
  fclose(fp);
  p7_hmm_Destroy(hmm); esl_sq_Destroy(sq); p7_builder_Destroy(bld); p7_bg_Destroy(bg); esl_alphabet_Destroy(abc);
  return eslOK;
}

/* Builder.build_msa (plan7.pyx:1018-1119 -> p7_Builder, p7_builder.c:415) on a digital alignment given as <nseq> rows of <alen>
 * residue codes (no sentinels): default builder (PB weights, entropy-weighted effective sequence number, the alphabet's mixture
 * Dirichlet priors), architecture 0 = fast (symfrac) / 1 = hand (needs <rf>), effn < 0: entropy weighting, 0: none, > 0: set.
 * The HMM is written in ASCII to <path>; the weights p7_Builder left in the alignment come back in <wgt_out> (may be NULL). */
#include "esl_msa.h"
int ref_msa_builder(int alphabet_type, const uint8_t *rows, int nseq, int alen, const char *const *names, const char *msaname,
                    const char *rf, int architecture, double symfrac, double fragthresh, double effn, int laplace, unsigned seed,
                    const char *path, double *wgt_out)
{
  ESL_ALPHABET *abc = esl_alphabet_Create(alphabet_type);
  P7_BG *bg = p7_bg_Create(abc);
  P7_BUILDER *bld = p7_builder_Create(NULL, abc);
  ESL_MSA *msa = esl_msa_CreateDigital(abc, nseq, alen);
  P7_HMM *hmm = NULL;
  FILE *fp;
  int i, j, status;
  ref_init();
  for (i = 0; i < nseq; i++) {
    esl_msa_SetSeqName(msa, i, names[i], -1);
    msa->ax[i][0] = msa->ax[i][alen + 1] = eslDSQ_SENTINEL;
    for (j = 0; j < alen; j++) msa->ax[i][j + 1] = rows[(size_t)i * alen + j];
  }
  esl_msa_SetName(msa, msaname, -1);
  if (rf) { msa->rf = malloc(alen + 1); memcpy(msa->rf, rf, alen); msa->rf[alen] = 0; }
  if (seed != 42) { esl_randomness_Destroy(bld->r); bld->r = esl_randomness_CreateFast(seed); }
  bld->do_reseeding = (seed != 0);
  bld->w_len = -1; bld->w_beta = p7_DEFAULT_WINDOW_BETA;
  bld->arch_strategy = architecture ? p7_ARCH_HAND : p7_ARCH_FAST;
  bld->symfrac = symfrac; bld->fragthresh = fragthresh;
  if (effn == 0.0) bld->effn_strategy = p7_EFFN_NONE;
  else if (effn > 0.0) { bld->effn_strategy = p7_EFFN_SET; bld->eset = effn; }
  if (laplace) { p7_prior_Destroy(bld->prior); bld->prior = p7_prior_CreateLaplace(abc); }
  if ((status = p7_Builder(bld, msa, bg, &hmm, NULL, NULL, NULL, NULL)) != eslOK) return status;
  if (wgt_out) for (i = 0; i < nseq; i++) wgt_out[i] = msa->wgt[i];
  if ((fp = fopen(path, "w")) == NULL) return -1;
  p7_hmmfile_WriteASCII(fp, -1, hmm);
  fclose(fp);
  p7_hmm_Destroy(hmm); esl_msa_Destroy(msa); p7_builder_Destroy(bld); p7_bg_Destroy(bg); esl_alphabet_Destroy(abc);
  return eslOK;
}

/* esl_random stream of the Mersenne Twister seeded <seed> (esl_random.c), and iid digital sequences drawn with it */
void ref_mt_stream(unsigned seed, int n, double *out)
{
  ESL_RANDOMNESS *r = esl_randomness_Create(seed);
  int i;
  for (i = 0; i < n; i++) out[i] = esl_random(r);
  esl_randomness_Destroy(r);
}

double ref_gumbel_surv(double x, double mu, double lambda) { return esl_gumbel_surv(x, mu, lambda); }
double ref_exp_surv(double x, double mu, double lambda)    { return esl_exp_surv(x, mu, lambda); }

/* =====================================================================================
 * The search loop, exactly as pyhmmer runs it (Pipeline._search_loop, plan7.pyx:6394-6453):
 *   p7_pli_NewModel; for each target: p7_pli_NewSeq, p7_bg_SetLength, p7_oprofile_ReconfigLength,
 *   p7_Pipeline, p7_pipeline_Reuse.
 * Reporting thresholds are opened wide (E = 1e300) so that every comparison p7_Pipeline scores to
 * completion comes back; the caller applies thresholds itself.
 * ===================================================================================== */
#include <pthread.h>
#include "esl_getopts.h"

typedef struct {
  int   seq;
  float score, pre_score, sum_score, nexpected;
  double lnP, pre_lnP, sum_lnP;
  int   nregions, nclustered, noverlaps, nenvelopes, ndom, best_domain;
  long  dom_offset;
} REF_HIT;

typedef struct {
  int   ienv, jenv, iali, jali;
  float envsc, domcorrection, dombias, oasc, bitscore;
  double lnP;
  int   hmmfrom, hmmto, sqfrom, sqto, N;
  long  text_offset;          /* model | mline | aseq | ppline, each N+1 bytes */
} REF_DOM;

typedef struct {
  long nhits, ndoms, ntext;
  REF_HIT *hits; REF_DOM *doms; char *text;
  long counters[4];           /* n_past_msv, n_past_bias, n_past_vit, n_past_fwd */
  double seconds;
} REF_RESULT;

static P7_PIPELINE *make_pipeline(const REFM *m, double F1, double F2, double F3, int do_bias, int do_null2, unsigned seed)
{
  P7_PIPELINE *pli = p7_pipeline_Create(NULL, m->om->M, 400, FALSE, p7_SEARCH_SEQS);
  pli->F1 = F1; pli->F2 = F2; pli->F3 = F3;
  pli->do_biasfilter = do_bias; pli->do_null2 = do_null2;
  pli->E = 1e300; pli->domE = 1e300; pli->incE = 1e300; pli->incdomE = 1e300;
  if (seed != 42) { esl_randomness_Init(pli->r, seed); pli->do_reseeding = pli->ddef->do_reseeding = (seed != 0); }
  return pli;
}

static void collect(P7_TOPHITS *th, const int *seqidx_of_hit, REF_RESULT *r)
{
  long h, d, ndoms = 0, ntext = 0;
  for (h = 0; h < (long)th->N; h++) for (d = 0; d < th->unsrt[h].ndom; d++) { ndoms++; ntext += 4 * (th->unsrt[h].dcl[d].ad->N + 1); }
  r->nhits = th->N; r->ndoms = ndoms; r->ntext = ntext;
  r->hits = calloc(th->N + 1, sizeof(REF_HIT)); r->doms = calloc(ndoms + 1, sizeof(REF_DOM)); r->text = calloc(ntext + 1, 1);
  ndoms = 0; ntext = 0;
  for (h = 0; h < (long)th->N; h++) {
    P7_HIT *hit = &th->unsrt[h]; REF_HIT *o = &r->hits[h];
    o->seq = seqidx_of_hit[h];
    o->score = hit->score; o->pre_score = hit->pre_score; o->sum_score = hit->sum_score; o->nexpected = hit->nexpected;
    o->lnP = hit->lnP; o->pre_lnP = hit->pre_lnP; o->sum_lnP = hit->sum_lnP;
    o->nregions = hit->nregions; o->nclustered = hit->nclustered; o->noverlaps = hit->noverlaps; o->nenvelopes = hit->nenvelopes;
    o->ndom = hit->ndom; o->best_domain = hit->best_domain; o->dom_offset = ndoms;
    for (d = 0; d < hit->ndom; d++) {
      P7_DOMAIN *dom = &hit->dcl[d]; REF_DOM *q = &r->doms[ndoms++]; P7_ALIDISPLAY *ad = dom->ad;
      q->ienv = dom->ienv; q->jenv = dom->jenv; q->iali = dom->iali; q->jali = dom->jali;
      q->envsc = dom->envsc; q->domcorrection = dom->domcorrection; q->dombias = dom->dombias; q->oasc = dom->oasc;
      q->bitscore = dom->bitscore; q->lnP = dom->lnP;
      q->hmmfrom = ad->hmmfrom; q->hmmto = ad->hmmto; q->sqfrom = ad->sqfrom; q->sqto = ad->sqto; q->N = ad->N;
      q->text_offset = ntext;
      memcpy(r->text + ntext, ad->model, ad->N + 1);  ntext += ad->N + 1;
      memcpy(r->text + ntext, ad->mline, ad->N + 1);  ntext += ad->N + 1;
      memcpy(r->text + ntext, ad->aseq, ad->N + 1);   ntext += ad->N + 1;
      memcpy(r->text + ntext, ad->ppline, ad->N + 1); ntext += ad->N + 1;
    }
  }
}

/* single-threaded, full results */
REF_RESULT *ref_search(REFM *m, const uint8_t *const *dsq, const int64_t *len, int n,
                       double F1, double F2, double F3, int do_bias, int do_null2, unsigned seed)
{
  REF_RESULT  *r   = calloc(1, sizeof(REF_RESULT));
  P7_PIPELINE *pli = make_pipeline(m, F1, F2, F3, do_bias, do_null2, seed);
  P7_TOPHITS  *th  = p7_tophits_Create();
  int *seqidx = malloc(sizeof(int) * (n + 1));
  int t; char name[32];
  p7_oprofile_ReconfigMultihit(m->om, 400);
  p7_pli_NewModel(pli, m->om, m->bg);
  for (t = 0; t < n; t++) {
    ESL_SQ *sq;
    uint64_t before = th->N;
    snprintf(name, sizeof name, "seq%d", t);
    sq = esl_sq_CreateDigitalFrom(m->abc, name, dsq[t], len[t], NULL, NULL, NULL);
    p7_pli_NewSeq(pli, sq);
    p7_bg_SetLength(m->bg, sq->n);
    p7_oprofile_ReconfigLength(m->om, sq->n);
    p7_Pipeline(pli, m->om, m->bg, sq, NULL, th);
    p7_pipeline_Reuse(pli);
    if (th->N > before) seqidx[before] = t;
    esl_sq_Destroy(sq);
  }
  collect(th, seqidx, r);
  r->counters[0] = pli->n_past_msv; r->counters[1] = pli->n_past_bias; r->counters[2] = pli->n_past_vit; r->counters[3] = pli->n_past_fwd;
  free(seqidx); p7_tophits_Destroy(th); p7_pipeline_Destroy(pli);
  return r;
}
/* The reference's own tabular writers on a search with DEFAULT reporting thresholds, as TopHits.write does (plan7.pyx:9096:
 * p7_tophits_TabularTargets / TabularDomains / TabularXfam after SortBySortkey + Threshold).  names / accs / descs: per target
 * (acc / desc may hold NULL or "").  Files: <prefix>.tbl, <prefix>.domtbl, <prefix>.pfam.  Returns the number of hits. */
long ref_search_tables(REFM *m, const uint8_t *const *dsq, const int64_t *len, int n, const char *const *names,
                       const char *const *accs, const char *const *descs, const char *prefix)
{
  P7_PIPELINE *pli = p7_pipeline_Create(NULL, m->om->M, 400, FALSE, p7_SEARCH_SEQS);
  P7_TOPHITS  *th  = p7_tophits_Create();
  char path[1024];
  const char *ext[3] = { ".tbl", ".domtbl", ".pfam" };
  int t, q;
  long nh;
  p7_oprofile_ReconfigMultihit(m->om, 400);
  p7_pli_NewModel(pli, m->om, m->bg);
  for (t = 0; t < n; t++) {
    ESL_SQ *sq = esl_sq_CreateDigitalFrom(m->abc, names[t], dsq[t], len[t], (descs && descs[t]) ? descs[t] : NULL, (accs && accs[t]) ? accs[t] : NULL, NULL);
    p7_pli_NewSeq(pli, sq);
    p7_bg_SetLength(m->bg, sq->n);
    p7_oprofile_ReconfigLength(m->om, sq->n);
    p7_Pipeline(pli, m->om, m->bg, sq, NULL, th);
    p7_pipeline_Reuse(pli);
    esl_sq_Destroy(sq);
  }
  p7_tophits_SortBySortkey(th);
  p7_tophits_Threshold(th, pli);
  for (q = 0; q < 3; q++) {
    FILE *fp;
    snprintf(path, sizeof path, "%s%s", prefix, ext[q]);
    if ((fp = fopen(path, "w")) == NULL) return -1;
    if (q == 0) p7_tophits_TabularTargets(fp, m->om->name, m->om->acc, th, pli, TRUE);
    if (q == 1) p7_tophits_TabularDomains(fp, m->om->name, m->om->acc, th, pli, TRUE);
    if (q == 2) p7_tophits_TabularXfam(fp, m->om->name, m->om->acc, th, pli);
    fclose(fp);
  }
  {   /* every domain's alignment block as Alignment.__str__ prints it (plan7.pyx:249-270), in hit / domain order */
    FILE *fp; long h; int d;
    snprintf(path, sizeof path, "%s.ali", prefix);
    if ((fp = fopen(path, "w")) == NULL) return -1;
    for (h = 0; h < (long)th->N; h++)
      for (d = 0; d < th->hit[h]->ndom; d++) {
        fprintf(fp, ">> %s %d\n", th->hit[h]->name, d);
        p7_nontranslated_alidisplay_Print(fp, th->hit[h]->dcl[d].ad, 0, -1, FALSE);
      }
    fclose(fp);
  }
  nh = th->N;
  p7_tophits_Destroy(th); p7_pipeline_Destroy(pli);
  return nh;
}

/* TopHits.to_msa (plan7.pyx:8960-9080): the reference search with default thresholds, then p7_tophits_Alignment of the
 * included domains, written in Pfam (one block) Stockholm format to <path>.  flags: 1 = p7_ALL_CONSENSUS_COLS, 2 = p7_TRIM.
 * Returns the number of aligned sequences, 0 when nothing is included. */
#include "esl_msa.h"
#include "esl_msafile.h"
long ref_search_msa(REFM *m, const uint8_t *const *dsq, const int64_t *len, int n, const char *const *names,
                    const char *const *accs, const char *const *descs, int flags, const char *path)
{
  P7_PIPELINE *pli = p7_pipeline_Create(NULL, m->om->M, 400, FALSE, p7_SEARCH_SEQS);
  P7_TOPHITS  *th  = p7_tophits_Create();
  ESL_MSA *msa = NULL;
  FILE *fp;
  int t, status, opt = p7_DEFAULT;
  long nseq = 0;
  p7_oprofile_ReconfigMultihit(m->om, 400);
  p7_pli_NewModel(pli, m->om, m->bg);
  for (t = 0; t < n; t++) {
    ESL_SQ *sq = esl_sq_CreateDigitalFrom(m->abc, names[t], dsq[t], len[t], (descs && descs[t]) ? descs[t] : NULL, (accs && accs[t]) ? accs[t] : NULL, NULL);
    p7_pli_NewSeq(pli, sq);
    p7_bg_SetLength(m->bg, sq->n);
    p7_oprofile_ReconfigLength(m->om, sq->n);
    p7_Pipeline(pli, m->om, m->bg, sq, NULL, th);
    p7_pipeline_Reuse(pli);
    esl_sq_Destroy(sq);
  }
  p7_tophits_SortBySortkey(th);
  p7_tophits_Threshold(th, pli);
  if (flags & 1) opt |= p7_ALL_CONSENSUS_COLS;
  if (flags & 2) opt |= p7_TRIM;
  status = p7_tophits_Alignment(th, m->abc, NULL, NULL, 0, opt, &msa);
  if (status == eslOK) {
    if ((fp = fopen(path, "w")) == NULL) return -1;
    esl_msafile_Write(fp, msa, eslMSAFILE_PFAM);
    fclose(fp);
    nseq = msa->nseq;
    esl_msa_Destroy(msa);
  }
  p7_tophits_Destroy(th); p7_pipeline_Destroy(pli);
  return nseq;
}

/* hmmscan of ONE sequence against a list of models as pyhmmer runs it (Pipeline._scan_loop, plan7.pyx:6625-6677), default
 * thresholds, hits written by the reference's tabular writers to <prefix>.tbl / .domtbl / .pfam.  Returns the hit count. */
long ref_scan_tables(REFM **models, int nmodels, const uint8_t *dsq, long len, const char *qname, const char *qacc, const char *qdesc,
                     const char *prefix)
{
  P7_PIPELINE *pli = p7_pipeline_Create(NULL, 100, (int)len, FALSE, p7_SCAN_MODELS);
  P7_TOPHITS  *th  = p7_tophits_Create();
  ESL_SQ *sq = esl_sq_CreateDigitalFrom(models[0]->abc, qname, dsq, len, qdesc, qacc, NULL);
  char path[1024];
  const char *ext[3] = { ".tbl", ".domtbl", ".pfam" };
  int t, q;
  long nh;
  p7_pli_NewSeq(pli, sq);
  for (t = 0; t < nmodels; t++) {
    REFM *m = models[t];
    p7_oprofile_ReconfigMultihit(m->om, 400);
    p7_pli_NewModel(pli, m->om, m->bg);
    p7_bg_SetLength(m->bg, sq->n);
    p7_oprofile_ReconfigLength(m->om, sq->n);
    p7_Pipeline(pli, m->om, m->bg, sq, NULL, th);
    p7_pipeline_Reuse(pli);
  }
  p7_tophits_SortBySortkey(th);
  p7_tophits_Threshold(th, pli);
  for (q = 0; q < 3; q++) {
    FILE *fp;
    snprintf(path, sizeof path, "%s%s", prefix, ext[q]);
    if ((fp = fopen(path, "w")) == NULL) return -1;
    if (q == 0) p7_tophits_TabularTargets(fp, sq->name, sq->acc, th, pli, TRUE);
    if (q == 1) p7_tophits_TabularDomains(fp, sq->name, sq->acc, th, pli, TRUE);
    if (q == 2) p7_tophits_TabularXfam(fp, sq->name, sq->acc, th, pli);
    fclose(fp);
  }
  nh = th->N;
  esl_sq_Destroy(sq); p7_tophits_Destroy(th); p7_pipeline_Destroy(pli);
  return nh;
}

void ref_result_free(REF_RESULT *r) { if (r) { free(r->hits); free(r->doms); free(r->text); free(r); } }
long ref_result_nhits(const REF_RESULT *r) { return r->nhits; }
long ref_result_ndoms(const REF_RESULT *r) { return r->ndoms; }
const REF_HIT *ref_result_hits(const REF_RESULT *r) { return r->hits; }
const REF_DOM *ref_result_doms(const REF_RESULT *r) { return r->doms; }
const char *ref_result_text(const REF_RESULT *r) { return r->text; }
const long *ref_result_counters(const REF_RESULT *r) { return r->counters; }

/* ---- multi-threaded timing run (bench.py --impl reference / cpu_baseline): every thread owns a pipeline, a
 * background and a clone of the current query profile, and pulls blocks of MT_BLOCK target sequences from a
 * shared counter (the work-queue scheme of HMMER's own threaded hmmsearch; pyhmmer's dispatcher splits the
 * targets the same way, _hmmsearch.py:153-171, but statically).  Dynamic blocks keep the 128 host threads
 * busy when a few targets (true homologs) cost 1000x the others. */
#define MT_BLOCK 128
typedef struct {
  REFM **models; int nmodels;
  const uint8_t *const *dsq; const int64_t *len; int n;
  int *next;                      /* next[q] = next unclaimed target of query q (atomic) */
  double F1, F2, F3; int do_bias, do_null2;
  long nhits; long counters[4];
} MT_JOB;

static void *mt_worker(void *arg)
{
  MT_JOB *job = (MT_JOB *)arg;
  int q, t, c;
  job->nhits = 0; for (c = 0; c < 4; c++) job->counters[c] = 0;
  ESL_SQ *sq = NULL;
  for (q = 0; q < job->nmodels; q++) {
    REFM *m = job->models[q];
    int t0 = __atomic_fetch_add(&job->next[q], MT_BLOCK, __ATOMIC_RELAXED);
    if (t0 >= job->n) continue;
    P7_OPROFILE *om = p7_oprofile_Clone(m->om);
    P7_BG *bg = p7_bg_Clone(m->bg);
    P7_PIPELINE *pli = p7_pipeline_Create(NULL, om->M, 400, FALSE, p7_SEARCH_SEQS);
    P7_TOPHITS *th = p7_tophits_Create();
    if (!sq) sq = esl_sq_CreateDigital(m->abc);
    pli->F1 = job->F1; pli->F2 = job->F2; pli->F3 = job->F3; pli->do_biasfilter = job->do_bias; pli->do_null2 = job->do_null2;
    p7_pli_NewModel(pli, om, bg);
    for (; t0 < job->n; t0 = __atomic_fetch_add(&job->next[q], MT_BLOCK, __ATOMIC_RELAXED)) {
      int t1 = t0 + MT_BLOCK < job->n ? t0 + MT_BLOCK : job->n;
      for (t = t0; t < t1; t++) {
        esl_sq_GrowTo(sq, job->len[t]);
        memcpy(sq->dsq, job->dsq[t], job->len[t] + 2);
        sq->n = job->len[t];
        esl_sq_SetName(sq, "s");
        p7_pli_NewSeq(pli, sq);
        p7_bg_SetLength(bg, sq->n);
        p7_oprofile_ReconfigLength(om, sq->n);
        p7_Pipeline(pli, om, bg, sq, NULL, th);
        p7_pipeline_Reuse(pli);
        esl_sq_Reuse(sq);
      }
    }
    job->nhits += th->N;
    job->counters[0] += pli->n_past_msv; job->counters[1] += pli->n_past_bias; job->counters[2] += pli->n_past_vit; job->counters[3] += pli->n_past_fwd;
    p7_tophits_Destroy(th); p7_pipeline_Destroy(pli); p7_bg_Destroy(bg); p7_oprofile_Destroy(om);
  }
  if (sq) esl_sq_Destroy(sq);
  return NULL;
}

/* returns the number of hits; counters4 receives the summed pipeline counters */
long ref_search_mt(REFM **models, int nmodels, const uint8_t *const *dsq, const int64_t *len, int n, int nthreads,
                   double F1, double F2, double F3, int do_bias, int do_null2, long *counters4)
{
  pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
  MT_JOB *jobs = calloc(nthreads, sizeof(MT_JOB));
  int *next = calloc(nmodels > 0 ? nmodels : 1, sizeof(int));
  int t, c; long nhits = 0;
  for (t = 0; t < nthreads; t++) {
    jobs[t].models = models; jobs[t].nmodels = nmodels; jobs[t].dsq = dsq; jobs[t].len = len; jobs[t].n = n; jobs[t].next = next;
    jobs[t].F1 = F1; jobs[t].F2 = F2; jobs[t].F3 = F3; jobs[t].do_bias = do_bias; jobs[t].do_null2 = do_null2;
    pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
  }
  for (c = 0; c < 4; c++) counters4[c] = 0;
  for (t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); nhits += jobs[t].nhits; for (c = 0; c < 4; c++) counters4[c] += jobs[t].counters[c]; }
  free(th); free(jobs); free(next);
  return nhits;
}

/* =====================================================================================
 * Round 2 additions: models from arrays, hmmscan with full results, threaded timing runs
 * for hmmscan (BASELINE configs[3]) and nhmmer (configs[4]).
 * ===================================================================================== */

/* A model built from probability arrays instead of a file: what `HMM(alphabet, M, name)` + array assignment gives a pyhmmer
 * user (plan7.pyx:2130-2200), then configured like refm_read.  t [(M+1)*7] MM MI MD IM II DM DD, mat / ins [(M+1)*K],
 * compo [K] (NULL: p7_hmm_SetComposition), consensus [M] characters (NULL: p7_hmm_SetConsensus(hmm, NULL)), evparam [6]. */
REFM *refm_from_arrays(int abc_type, int M, const float *t, const float *mat, const float *ins, const float *compo,
                       const float *evparam, int max_length, const char *name, const char *consensus, int L)
{
  REFM *m = calloc(1, sizeof(REFM));
  int k, x;
  ref_init();
  m->abc = esl_alphabet_Create(abc_type);
  m->hmm = p7_hmm_Create(M, m->abc);
  for (k = 0; k <= M; k++) {
    for (x = 0; x < 7; x++)        m->hmm->t[k][x]   = t[k * 7 + x];
    for (x = 0; x < m->abc->K; x++) { m->hmm->mat[k][x] = mat[k * m->abc->K + x]; m->hmm->ins[k][x] = ins[k * m->abc->K + x]; }
  }
  p7_hmm_SetName(m->hmm, (char *)name);
  if (compo) { for (x = 0; x < m->abc->K; x++) m->hmm->compo[x] = compo[x]; m->hmm->flags |= p7H_COMPO; }
  else p7_hmm_SetComposition(m->hmm);
  if (consensus) {
    m->hmm->consensus = malloc(M + 2);
    m->hmm->consensus[0] = ' '; memcpy(m->hmm->consensus + 1, consensus, M); m->hmm->consensus[M + 1] = '\0';
    m->hmm->flags |= p7H_CONS;
  } else p7_hmm_SetConsensus(m->hmm, NULL);
  for (x = 0; x < p7_NEVPARAM; x++) m->hmm->evparam[x] = evparam[x];
  m->hmm->flags |= p7H_STATS;
  m->hmm->max_length = max_length;
  m->hmm->nseq = 1; m->hmm->eff_nseq = 1.0f;
  m->bg = p7_bg_Create(m->abc);
  m->gm = p7_profile_Create(M, m->abc);
  m->om = p7_oprofile_Create(M, m->abc);
  p7_ProfileConfig(m->hmm, m->bg, m->gm, L, p7_LOCAL);
  p7_oprofile_Convert(m->gm, m->om);
  m->ox  = p7_omx_Create(M, 0, 400);
  m->oxb = p7_omx_Create(M, 0, 400);
  return m;
}

/* the same for n models on <nthreads> threads (the conversion costs ~0.5 ms per model): arrays concatenated model after
 * model, Ms[n]; compo [n][K] or NULL; names = n NUL-terminated strings back to back; max_length [n] or NULL; out[n]
 * receives the handles */
typedef struct { int abc_type, n, L; const int *Ms; const long *off; const float *t, *mat, *ins, *compo, *evparam; const char *cons;
                 const char *const *names; const int *maxl; REFM **out; int *next; } FA_JOB;
static void *fa_worker(void *arg)
{
  FA_JOB *j = (FA_JOB *)arg;
  int K = (j->abc_type == eslAMINO) ? 20 : 4, i;
  for (i = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED); i < j->n; i = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED)) {
    long o = j->off[i];                       /* sum of (M+1) of the models before i */
    j->out[i] = refm_from_arrays(j->abc_type, j->Ms[i], j->t + o * 7, j->mat + o * K, j->ins + o * K, j->compo ? j->compo + (long)i * K : NULL,
                                 j->evparam + (long)i * 6, j->maxl ? j->maxl[i] : 0, j->names[i], j->cons ? j->cons + (o - i) : NULL, j->L);
  }
  return NULL;
}
int refm_from_arrays_many(int abc_type, int n, const int *Ms, const float *t, const float *mat, const float *ins, const float *compo,
                          const float *evparam, const char *cons, const char *names, const int *max_length, int L, int nthreads, REFM **out)
{
  long *off = malloc(sizeof(long) * (n + 1));
  const char **nm = malloc(sizeof(char *) * (n + 1));
  pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
  FA_JOB job; int next = 0, i;
  ref_init();
  off[0] = 0; for (i = 0; i < n; i++) off[i + 1] = off[i] + Ms[i] + 1;
  for (i = 0; i < n; i++) { nm[i] = names; names += strlen(names) + 1; }
  job.abc_type = abc_type; job.n = n; job.L = L; job.Ms = Ms; job.off = off; job.t = t; job.mat = mat; job.ins = ins; job.compo = compo;
  job.evparam = evparam; job.cons = cons; job.names = nm; job.maxl = max_length; job.out = out; job.next = &next;
  for (i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, fa_worker, &job);
  for (i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  free(th); free(off); free(nm);
  return n;
}

/* hmmscan of ONE sequence against a list of models as pyhmmer runs it (Pipeline._scan_loop, plan7.pyx:6625-6677), reporting
 * thresholds opened wide so that every comparison p7_Pipeline scores to completion comes back; hit.seq = index of the MODEL. */
REF_RESULT *ref_scan(REFM **models, int nmodels, const uint8_t *dsq, long len,
                     double F1, double F2, double F3, int do_bias, int do_null2, unsigned seed)
{
  REF_RESULT  *r   = calloc(1, sizeof(REF_RESULT));
  P7_PIPELINE *pli = p7_pipeline_Create(NULL, 100, (int)len, FALSE, p7_SCAN_MODELS);
  P7_TOPHITS  *th  = p7_tophits_Create();
  ESL_SQ *sq = esl_sq_CreateDigitalFrom(models[0]->abc, "query", dsq, len, NULL, NULL, NULL);
  int *idx = malloc(sizeof(int) * (nmodels + 1));
  int t;
  pli->F1 = F1; pli->F2 = F2; pli->F3 = F3;
  pli->do_biasfilter = do_bias; pli->do_null2 = do_null2;
  pli->E = 1e300; pli->domE = 1e300; pli->incE = 1e300; pli->incdomE = 1e300;
  if (seed != 42) { esl_randomness_Init(pli->r, seed); pli->do_reseeding = pli->ddef->do_reseeding = (seed != 0); }
  p7_pli_NewSeq(pli, sq);
  for (t = 0; t < nmodels; t++) {
    REFM *m = models[t];
    uint64_t before = th->N;
    p7_oprofile_ReconfigMultihit(m->om, 400);
    p7_pli_NewModel(pli, m->om, m->bg);
    p7_bg_SetLength(m->bg, sq->n);
    p7_oprofile_ReconfigLength(m->om, sq->n);
    p7_Pipeline(pli, m->om, m->bg, sq, NULL, th);
    p7_pipeline_Reuse(pli);
    if (th->N > before) idx[before] = t;
  }
  collect(th, idx, r);
  r->counters[0] = pli->n_past_msv; r->counters[1] = pli->n_past_bias; r->counters[2] = pli->n_past_vit; r->counters[3] = pli->n_past_fwd;
  free(idx); esl_sq_Destroy(sq); p7_tophits_Destroy(th); p7_pipeline_Destroy(pli);
  return r;
}

/* ---- threaded timing run of hmmscan: the threads pull blocks of SC_BLOCK models from a shared counter; every thread owns a
 * pipeline in scan mode and a copy of the query.  (pyhmmer's own hmmscan gives ONE thread to one query, _hmmscan.py:29-37;
 * this is the reference's C pipeline given all the host's threads.)  A model is touched by one thread only. */
#define SC_BLOCK 16
typedef struct { REFM **models; int nmodels; const uint8_t *dsq; long len; int *next; double F1, F2, F3; int do_bias, do_null2;
                 long nhits; long counters[4]; } SC_JOB;
static void *sc_worker(void *arg)
{
  SC_JOB *job = (SC_JOB *)arg;
  P7_PIPELINE *pli = p7_pipeline_Create(NULL, 100, (int)job->len, FALSE, p7_SCAN_MODELS);
  P7_TOPHITS  *th  = p7_tophits_Create();
  ESL_SQ *sq = esl_sq_CreateDigitalFrom(job->models[0]->abc, "query", job->dsq, job->len, NULL, NULL, NULL);
  int t0, t;
  pli->F1 = job->F1; pli->F2 = job->F2; pli->F3 = job->F3; pli->do_biasfilter = job->do_bias; pli->do_null2 = job->do_null2;
  p7_pli_NewSeq(pli, sq);
  for (t0 = __atomic_fetch_add(job->next, SC_BLOCK, __ATOMIC_RELAXED); t0 < job->nmodels; t0 = __atomic_fetch_add(job->next, SC_BLOCK, __ATOMIC_RELAXED)) {
    int t1 = t0 + SC_BLOCK < job->nmodels ? t0 + SC_BLOCK : job->nmodels;
    for (t = t0; t < t1; t++) {
      REFM *m = job->models[t];
      p7_oprofile_ReconfigMultihit(m->om, 400);
      p7_pli_NewModel(pli, m->om, m->bg);
      p7_bg_SetLength(m->bg, sq->n);
      p7_oprofile_ReconfigLength(m->om, sq->n);
      p7_Pipeline(pli, m->om, m->bg, sq, NULL, th);
      p7_pipeline_Reuse(pli);
    }
  }
  job->nhits = th->N;
  job->counters[0] = pli->n_past_msv; job->counters[1] = pli->n_past_bias; job->counters[2] = pli->n_past_vit; job->counters[3] = pli->n_past_fwd;
  esl_sq_Destroy(sq); p7_tophits_Destroy(th); p7_pipeline_Destroy(pli);
  return NULL;
}
long ref_scan_mt(REFM **models, int nmodels, const uint8_t *dsq, long len, int nthreads,
                 double F1, double F2, double F3, int do_bias, int do_null2, long *counters4)
{
  pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
  SC_JOB *jobs = calloc(nthreads, sizeof(SC_JOB));
  int next = 0, t, c; long nhits = 0;
  for (t = 0; t < nthreads; t++) {
    jobs[t].models = models; jobs[t].nmodels = nmodels; jobs[t].dsq = dsq; jobs[t].len = len; jobs[t].next = &next;
    jobs[t].F1 = F1; jobs[t].F2 = F2; jobs[t].F3 = F3; jobs[t].do_bias = do_bias; jobs[t].do_null2 = do_null2;
    pthread_create(&th[t], NULL, sc_worker, &jobs[t]);
  }
  for (c = 0; c < 4; c++) counters4[c] = 0;
  for (t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); nhits += jobs[t].nhits; for (c = 0; c < 4; c++) counters4[c] += jobs[t].counters[c]; }
  free(th); free(jobs);
  return nhits;
}

/* ---- threaded timing run of nhmmer: the windows of ref_nhmmer's loop (block_length residues + max_length of context, both
 * strands) are pulled from a shared counter by threads that each own a pipeline, a profile clone, score data and a hit list;
 * the hit lists are merged and post-processed as in ref_nhmmer.  (pyhmmer's own nhmmer gives ONE thread to one query.)
 * Returns the number of hits left after duplicate removal; stats [6] as in ref_nhmmer. */
typedef struct { REFM *m; int nseq; const uint8_t **dsq; const long *len; long W; int strands; double F1, F2, F3; int do_bias, do_null2;
                 long nwin; const int *win_seq; const long *win_i; long *next; P7_TOPHITS *th; P7_PIPELINE *pli; } NH_JOB;
static void *nh_worker(void *arg)
{
  NH_JOB *job = (NH_JOB *)arg;
  REFM *m = job->m;
  P7_OPROFILE *om = p7_oprofile_Copy(m->om);      /* a deep copy, as nhmmer.c gives its threads: the long-target domain definition rewrites the emission scores */
  P7_BG *bg = p7_bg_Clone(m->bg);
  P7_PIPELINE *pli = p7_pipeline_Create(NULL, om->M, 100, TRUE, p7_SEARCH_SEQS);
  P7_SCOREDATA *data = p7_hmm_ScoreDataCreate(om, NULL);
  ESL_SQ *tmpsq = esl_sq_CreateDigital(m->abc);
  long C = om->max_length, W = job->W, w;
  pli->F1 = job->F1; pli->F2 = job->F2; pli->F3 = job->F3; pli->do_biasfilter = job->do_bias; pli->do_null2 = job->do_null2;
  pli->strands = (job->strands == 1) ? p7_STRAND_TOPONLY : (job->strands == 2) ? p7_STRAND_BOTTOMONLY : p7_STRAND_BOTH;
  pli->block_length = (int)W;
  pli->nseqs = 0;
  p7_pli_NewModel(pli, om, bg);
  job->th = p7_tophits_Create();
  for (w = __atomic_fetch_add(job->next, 1, __ATOMIC_RELAXED); w < job->nwin; w = __atomic_fetch_add(job->next, 1, __ATOMIC_RELAXED)) {
    int t = job->win_seq[w]; long i = job->win_i[w];
    char name[64];
    snprintf(name, sizeof name, "seq%d", t);
    tmpsq->idx = t; tmpsq->L = -1;
    esl_sq_SetAccession(tmpsq, ""); esl_sq_SetName(tmpsq, name); esl_sq_SetDesc(tmpsq, ""); esl_sq_SetSource(tmpsq, name);
    esl_sq_GrowTo(tmpsq, ESL_MIN(W + C, job->len[t]));
    tmpsq->C = (i == 0) ? 0 : ESL_MIN(C, job->len[t] - i);
    tmpsq->W = ESL_MIN(W, job->len[t] - i - tmpsq->C);
    tmpsq->n = tmpsq->C + tmpsq->W;
    tmpsq->start = i + 1;
    tmpsq->end = i + tmpsq->n;
    memcpy(tmpsq->dsq + 1, job->dsq[t] + i + 1, tmpsq->n);
    tmpsq->dsq[0] = tmpsq->dsq[tmpsq->n + 1] = eslDSQ_SENTINEL;
    p7_pli_NewSeq(pli, tmpsq);
    if (pli->strands != p7_STRAND_BOTTOMONLY) {
      pli->nres -= tmpsq->C;
      p7_Pipeline_LongTarget(pli, om, data, bg, job->th, t, tmpsq, p7_NOCOMPLEMENT, NULL, NULL, NULL);
      p7_pipeline_Reuse(pli);
    } else pli->nres -= tmpsq->n;
    if (pli->strands != p7_STRAND_TOPONLY) {
      esl_sq_ReverseComplement(tmpsq);
      p7_Pipeline_LongTarget(pli, om, data, bg, job->th, t, tmpsq, p7_COMPLEMENT, NULL, NULL, NULL);
      p7_pipeline_Reuse(pli);
      pli->nres += tmpsq->W;
    }
    esl_sq_Reuse(tmpsq);
  }
  job->pli = pli;
  esl_sq_Destroy(tmpsq); p7_hmm_ScoreDataDestroy(data); p7_bg_Destroy(bg); p7_oprofile_Destroy(om);
  return NULL;
}
long ref_nhmmer_mt(REFM *m, int nseq, const uint8_t **dsq, const long *len, long block_length, int strands, int nthreads,
                   double F1, double F2, double F3, int do_bias, int do_null2, long evalue_window, long *stats)
{
  long C = m->om->max_length, W = block_length, nwin = 0, i, w = 0, next = 0, nout;
  int t, c;
  if (C <= 0 || W <= C) return -1;
  for (t = 0; t < nseq; t++) for (i = 0; i < len[t]; i += W - C) nwin++;
  int *win_seq = malloc(sizeof(int) * (nwin + 1)); long *win_i = malloc(sizeof(long) * (nwin + 1));
  for (t = 0; t < nseq; t++) for (i = 0; i < len[t]; i += W - C) { win_seq[w] = t; win_i[w] = i; w++; }
  pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
  NH_JOB *jobs = calloc(nthreads, sizeof(NH_JOB));
  for (t = 0; t < nthreads; t++) {
    jobs[t].m = m; jobs[t].nseq = nseq; jobs[t].dsq = dsq; jobs[t].len = len; jobs[t].W = W; jobs[t].strands = strands;
    jobs[t].F1 = F1; jobs[t].F2 = F2; jobs[t].F3 = F3; jobs[t].do_bias = do_bias; jobs[t].do_null2 = do_null2;
    jobs[t].nwin = nwin; jobs[t].win_seq = win_seq; jobs[t].win_i = win_i; jobs[t].next = &next;
    pthread_create(&th[t], NULL, nh_worker, &jobs[t]);
  }
  for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  for (c = 0; c < 6; c++) stats[c] = 0;
  for (t = 1; t < nthreads; t++) {
    p7_tophits_Merge(jobs[0].th, jobs[t].th);
    p7_pipeline_Merge(jobs[0].pli, jobs[t].pli);
  }
  {
    P7_PIPELINE *pli = jobs[0].pli; P7_TOPHITS *hits = jobs[0].th;
    pli->nseqs = nseq;
    p7_tophits_ComputeNhmmerEvalues(hits, (double)pli->nres, evalue_window > 0 ? (int)evalue_window : m->om->max_length);
    p7_tophits_SortBySeqidxAndAlipos(hits);
    p7_tophits_RemoveDuplicates(hits, TRUE);
    p7_tophits_SortBySortkey(hits);
    p7_tophits_Threshold(hits, pli);
    stats[0] = pli->nres; stats[1] = nseq; stats[2] = pli->pos_past_msv; stats[3] = pli->pos_past_bias; stats[4] = pli->pos_past_vit; stats[5] = pli->pos_past_fwd;
    nout = 0;
    for (i = 0; i < (long)hits->N; i++) if (!(hits->hit[i]->flags & p7_IS_DUPLICATE)) nout++;
    if (getenv("REF_DEBUG")) for (i = 0; i < (long)hits->N; i++) fprintf(stderr, "mt hit %ld: seq %ld env %ld..%ld score %.1f flags %d\n", i, (long)hits->hit[i]->seqidx, (long)hits->hit[i]->dcl[0].ienv, (long)hits->hit[i]->dcl[0].jenv, hits->hit[i]->score, hits->hit[i]->flags);
  }
  for (t = 0; t < nthreads; t++) { p7_tophits_Destroy(jobs[t].th); p7_pipeline_Destroy(jobs[t].pli); }
  free(th); free(jobs); free(win_seq); free(win_i);
  return nout;
}
