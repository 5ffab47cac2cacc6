/* b2h.h -- C ABI of libb2h.so, the B200-native profile-HMM comparison engine.
 *
 * This is the drop-in boundary for pyhmmer's search path.  Every entry point is
 * `extern "C"`, takes plain pointers and sizes, returns an Easel-compatible int
 * status (easel.h:100-128 in the reference) and never throws across the ABI.
 * INTEGRATION.md shows the Cython stub a pyhmmer maintainer would add to bind
 * each of them.  Reference interfaces replaced (paths relative to the reference
 * checkout):
 *
 *   b2h_profile_config        p7_ProfileConfig            vendor/hmmer/src/modelconfig.c:48
 *   b2h_oprofile_convert      p7_oprofile_Convert         vendor/hmmer/src/impl_sse/p7_oprofile.c:1014
 *   b2h_profile_upload*       P7_OPROFILE tables          vendor/hmmer/src/impl_sse/impl_sse.h:75-142
 *   b2h_seqdb_create          ESL_SQ** of a DigitalSequenceBlock   src/pyhmmer/easel.pxd:313-321
 *   b2h_ssv_filter            p7_SSVFilter                vendor/hmmer/src/impl_sse/ssvfilter.c:876
 *   b2h_msv_filter            p7_MSVFilter                vendor/hmmer/src/impl_sse/msvfilter.c:74
 *   b2h_viterbi_filter        p7_ViterbiFilter            vendor/hmmer/src/impl_sse/vitfilter.c:83
 *   b2h_forward_parser        p7_ForwardParser            vendor/hmmer/src/impl_sse/fwdback.c:132
 *   b2h_backward_parser       p7_BackwardParser           vendor/hmmer/src/impl_sse/fwdback.c:236
 *   b2h_null_scores           p7_bg_NullOne, p7_bg_FilterScore   vendor/hmmer/src/p7_bg.c:357,471
 *   b2h_search / b2h_scan     Pipeline._search_loop / _scan_loop (= p7_Pipeline per target)
 *                             src/pyhmmer/plan7.pyx:6394-6453, 6625-6677; vendor/hmmer/src/p7_pipeline.c:697-936
 *
 * All DP runs in hand-written sm_100a kernels; there is NO CPU fallback: on a box
 * without a CUDA device b2h_ctx_create() fails with B2H_ECUDA.
 */
#ifndef B2H_H_INCLUDED
#define B2H_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes: the Easel values the reference's callers already switch on */
#define B2H_OK          0
#define B2H_EMEM        5
#define B2H_EINVAL     11
#define B2H_ERANGE     16
#define B2H_ENORESULT  19
#define B2H_ECUDA     100   /* CUDA runtime / device failure (no Easel equivalent) */

#define B2H_NEVPARAM    6   /* MMU MLAMBDA VMU VLAMBDA FTAU FLAMBDA (hmmer.h p7_evparams_e) */
#define B2H_NCUTOFFS    6   /* GA1 GA2 TC1 TC2 NC1 NC2 */
#define B2H_MAXABET    20
#define B2H_MAXCODE    29

typedef struct b2h_ctx     b2h_ctx;      /* one per (process, GPU): device, stream, workspaces */
typedef struct b2h_seqdb   b2h_seqdb;    /* device-resident target sequence arena              */
typedef struct b2h_profile b2h_profile;  /* device-resident optimized profile                  */

/* ------------------------------------------------------------------------------------------
 * Host-side, un-striped ("node-major") optimized profile.  Node k (1..M) lives at index k-1.
 * Transition rows follow p7o_tsc_e order BM MM IM DM MD MI II DD with the reference's
 * rotation already undone:  BM[k]: B->M_k   MM/IM/DM[k]: {M,I,D}_{k-1}->M_k
 *                           MD[k]: M_k->D_{k+1}   MI[k]: M_k->I_k   II[k]: I_k->I_k   DD[k]: D_k->D_{k+1}
 * (this is exactly de-striping rbv/rwv/twv/rfv/tfv with k = q + z*Q + 1, p7_oprofile.c:800,856,949).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t  M, K, Kp;
  int32_t  L;                 /* length the L-dependent scalars below are configured for */
  int32_t  mode_multihit;     /* 1: multihit local (nj=1), 0: unihit local (nj=0)         */
  int32_t  max_length;
  /* MSV / SSV (uint8) */
  const uint8_t *msv_cost;    /* [Kp][M] biased match costs (rbv de-striped); 255 = -inf  */
  uint8_t  tbm_b, tec_b, tjb_b, base_b, bias_b;
  float    scale_b;
  /* ViterbiFilter (int16) */
  const int16_t *vit_rsc;     /* [Kp][M]  */
  const int16_t *vit_tsc;     /* [8][M]   */
  int16_t  xw[4][2];          /* [E N J C][MOVE LOOP] */
  int16_t  base_w, ddbound_w;
  float    scale_w;
  /* Forward / Backward (fp32 odds ratios) */
  const float *fwd_rsc;       /* [Kp][M]  */
  const float *fwd_tsc;       /* [8][M]   */
  float    xf[4][2];
  /* statistics, cutoffs, composition */
  float    evparam[B2H_NEVPARAM];
  float    cutoff[B2H_NCUTOFFS];
  float    compo[B2H_MAXABET];
  float    bgf[B2H_MAXABET];  /* null-model residue frequencies (P7_BG.f) */
  const uint8_t *degen;       /* [Kp][K] alphabet degeneracy matrix (ESL_ALPHABET.degen), for the bias filter */
} b2h_oprofile_desc;

/* ---- host math that must agree bit-for-bit with the reference's libm call sites ---- */

/* HMM file value decoding: out[i] = (in[i] is '*' i.e. +inf) ? 0 : expf(-in[i])   (p7_hmmfile.c:1486-1547) */
int b2h_hmm_decode_probs(const double *neglog, float *out, size_t n);

/* The node table of a HMMER3 ASCII model (read_asc30hmm, vendor/hmmer/src/p7_hmmfile.c:1411-1500): <text> = everything between
 * the "HMM ..." column header lines and the closing "//".  Fields are "-log p" or "*" and become expf(-atof(field)), as in the
 * reference.  mat / ins [(M+1)*K] (row 0 of mat untouched), t [(M+1)*7], compo [K] (*has_compo = the COMPO line was there),
 * map [M+1], anno [(nanno-1)*M] = first character of the annotation fields after MAP (CONS, RF, MM, CS as the format has
 * them), field-major.  B2H_ERANGE: a match line does not start with its node number; B2H_EINVAL: malformed. */
int b2h_hmm_parse_body(const char *text, size_t len, int M, int K, int nanno,
                       float *compo, int32_t *has_compo, float *mat, float *ins, float *t, int64_t *map, char *anno);

/* p7_ProfileConfig (local modes only): HMM probabilities -> log-odds generic profile.
 *   t   [(M+1)*7]  MM MI MD IM II DM DD     mat [(M+1)*K]    bgf[K]
 *   degen [Kp*K]   alphabet degeneracy matrix (esl_alphabet.h: degen[x][y])
 * outputs: tsc [M*8] (nodes 0..M-1) MM IM DM BM MD DD MI II (p7p_tsc_e), msc [Kp*(M+1)] match log-odds,
 *          xsc [4*2] E N J C x LOOP MOVE (p7p_xtransitions_e: LOOP=0, MOVE=1). */
int b2h_profile_config(int M, int K, int Kp, const uint8_t *degen,
                       const float *t, const float *mat, const float *bgf,
                       int L, int multihit,
                       float *tsc, float *msc, float *xsc);

/* p7_oprofile_Convert: generic profile -> the three score systems, node-major layout.
 * Caller allocates: msv_cost[Kp*M] vit_rsc[Kp*M] vit_tsc[8*M] fwd_rsc[Kp*M] fwd_tsc[8*M];
 * scalars are written into *desc (the table pointers in *desc are left untouched). */
int b2h_oprofile_convert(int M, int K, int Kp, int L, int multihit,
                         const float *tsc, const float *msc, const float *xsc,
                         uint8_t *msv_cost, int16_t *vit_rsc, int16_t *vit_tsc,
                         float *fwd_rsc, float *fwd_tsc, b2h_oprofile_desc *desc);

/* De-stripe a reference P7_OPROFILE's SSE tables (what a Cython binding holding a P7_OPROFILE* passes):
 * rbv [Kp][Q16][16], rwv [Kp][Q8][8], twv [8*Q8][8] (7 interleaved per q, then DD), rfv [Kp][Q4][4], tfv likewise. */
int b2h_destripe_oprofile(int M, int Kp,
                          const uint8_t *rbv, const int16_t *rwv, const int16_t *twv,
                          const float *rfv, const float *tfv,
                          uint8_t *msv_cost, int16_t *vit_rsc, int16_t *vit_tsc,
                          float *fwd_rsc, float *fwd_tsc);

/* Length-dependent scalars for one target length, as p7_oprofile_ReconfigLength / p7_bg_SetLength
 * compute them (p7_oprofile.c:1095-1136, p7_bg.c:189,357).  nj = 1 (multihit) or 0 (unihit). */
typedef struct {
  uint8_t tjb_b;          /* unbiased_byteify(logf(3/(L+3)))            */
  int16_t xw_move;        /* wordify(logf(pmove))                       */
  float   pmove, ploop;   /* (2+nj)/(L+2+nj), 1-pmove                   */
  float   null1;          /* L*log(p1)+log(1-p1), p1=L/(L+1)            */
  float   p1;
  float   flt_len_a;      /* (float)L*logf(p1)      (bias filter tail)  */
  float   flt_len_b;      /* logf(1.-p1)                                */
} b2h_len_params;
int b2h_length_params(int L, float nj, b2h_len_params *out);

/* ------------------------------- device objects ------------------------------------------ */

int         b2h_ctx_create(int device, b2h_ctx **out);
void        b2h_ctx_destroy(b2h_ctx *ctx);
/* Run all subsequent work of this context on an existing CUDA stream (cudaStream_t cast to void*);
 * NULL restores the context's own stream. */
int         b2h_ctx_set_stream(b2h_ctx *ctx, void *cuda_stream);
int         b2h_ctx_synchronize(b2h_ctx *ctx);
const char *b2h_ctx_last_error(const b2h_ctx *ctx);
/* Per-stage device timing for bench.py's roofline line: when enabled, b2h_search brackets each stage's
 * launches with CUDA events on the context's stream.  b2h_ctx_stage_ms() returns the accumulated
 * milliseconds {SSV, MSV, bias, Viterbi, Forward, survivor Fwd+Bck, grouping, reserved} and optionally resets. */
int         b2h_ctx_set_profiling(b2h_ctx *ctx, int on);
int         b2h_ctx_stage_ms(b2h_ctx *ctx, double *ms8, int reset);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t    b2h_ctx_launch_count(const b2h_ctx *ctx);

/* Upload n digital sequences.  dsq[i] points at an Easel digital sequence: residues at
 * dsq[i][1..len[i]], sentinel bytes at [0] and [len+1] (esl_sq.h:100-102), i.e. ESL_SQ.dsq / ESL_SQ.n. */
int  b2h_seqdb_create(b2h_ctx *ctx, const uint8_t *const *dsq, const int64_t *len, size_t n, b2h_seqdb **out);
/* Same, from one concatenated residue buffer (no sentinels) + offsets[n+1]. */
int  b2h_seqdb_create_packed(b2h_ctx *ctx, const uint8_t *residues, const int64_t *offsets, size_t n, b2h_seqdb **out);
void b2h_seqdb_destroy(b2h_seqdb *db);
size_t  b2h_seqdb_nseq(const b2h_seqdb *db);
int64_t b2h_seqdb_nres(const b2h_seqdb *db);

int  b2h_profile_upload(b2h_ctx *ctx, const b2h_oprofile_desc *desc, b2h_profile **out);
/* The same for n profiles: the host-side table building runs on several threads, the copies are issued in order.
 * out[0..n) receives the handles; on failure nothing is left allocated. (An OptimizedProfileBlock in one call.) */
int  b2h_profile_upload_many(b2h_ctx *ctx, const b2h_oprofile_desc *const *descs, size_t n, b2h_profile **out);
void b2h_profile_destroy(b2h_profile *p);

/* ------------------------- per-stage entry points (dense outputs) ------------------------ *
 * One profile against every sequence of the database, in database order.  sc[n] and status[n]
 * are HOST buffers; status[i] is the Easel code the reference function would have returned
 * for that comparison (eslOK / eslERANGE / eslENORESULT), sc[i] the value it would have stored.
 * Each comparison is configured for its own target length first, exactly as the search loop
 * does (p7_bg_SetLength + p7_oprofile_ReconfigLength, plan7.pyx:6431-6436).                 */
int b2h_ssv_filter     (b2h_ctx*, const b2h_profile*, const b2h_seqdb*, float *sc, int32_t *status);
int b2h_msv_filter     (b2h_ctx*, const b2h_profile*, const b2h_seqdb*, float *sc, int32_t *status);
int b2h_viterbi_filter (b2h_ctx*, const b2h_profile*, const b2h_seqdb*, float *sc, int32_t *status);
int b2h_forward_parser (b2h_ctx*, const b2h_profile*, const b2h_seqdb*, float *sc, int32_t *status);
int b2h_backward_parser(b2h_ctx*, const b2h_profile*, const b2h_seqdb*, float *sc, int32_t *status);
/* null1[n] = p7_bg_NullOne; filtersc[n] = p7_bg_FilterScore after p7_bg_SetFilter(M, compo) (may be NULL) */
int b2h_null_scores    (b2h_ctx*, const b2h_profile*, const b2h_seqdb*, float *null1, float *filtersc);

/* The generic (unstriped, log-space) reference DP on the P7_PROFILE, one profile against the whole database:
 * p7_GMSV (vendor/hmmer/src/generic_msv.c:56; what Profile.msv_filter calls, plan7.pyx:8212-8253), p7_GViterbi
 * (generic_viterbi.c:64), p7_GForward and p7_GBackward (generic_fwdback.c:48,164), with p7_FLogsum's lookup table
 * (logsum.c:105) and p7_ReconfigLength per target (modelconfig.c:221).  tsc [M*8], msc [Kp*(M+1)], xsc [4*2] are the
 * outputs of b2h_profile_config; nj = 1 (multihit) or 0; nu = expected number of hits for GMSV (2.0).  Scores in nats,
 * indexed by sequence; any output may be NULL.  Bit-identical to the reference evaluated on the same host. */
int b2h_generic_scores(b2h_ctx *ctx, int M, int K, int Kp, const float *tsc, const float *msc, const float *xsc, float nj,
                       const b2h_seqdb *db, float nu, float *gmsv, float *gviterbi, float *gforward, float *gbackward);

/* p7_GDecoding (generic_decoding.c:77) of ONE comparison: posterior probabilities from the full generic Forward and Backward
 * matrices.  residues = L residue codes (no sentinels).  pp_dp [(L+1)][(M+1)][3] (M, I, D as P7_GMX.dp), pp_xmx [(L+1)][5]
 * (E N J B C); the Forward and Backward scores are returned too.  The matrices are filled in the reference's order (bit-identical
 * scores); the probabilities differ from the reference's only through expf (device vs glibc: a few ulp).  dom_btot / dom_etot /
 * dom_mocc [L+1] (each may be NULL): p7_GDomainDecoding (generic_decoding.c:207) from the same matrices' special rows. */
int b2h_generic_decoding(b2h_ctx *ctx, int M, int K, int Kp, const float *tsc, const float *msc, const float *xsc, float nj,
                         const uint8_t *residues, int L, float *pp_dp, float *pp_xmx, float *fwdsc, float *bcksc,
                         float *dom_btot, float *dom_etot, float *dom_mocc);

/* --------------------------- the fused search path (p7_Pipeline per target) ----------------- *
 * b2h_search() is what Pipeline._search_loop / _scan_loop (plan7.pyx:6394-6453, 6625-6677) do for
 * P profiles x every sequence of the database: the whole acceleration-filter cascade runs on the
 * GPU with on-device survivor compaction (SSV/MSV -> bias -> Viterbi -> Forward -> Backward);
 * only comparisons that pass F3 come back to the host, where domain definition
 * (p7_domaindef_ByPosteriorHeuristics, p7_domaindef.c:384) finishes them into hits.
 * Reporting / inclusion thresholds (E, T, Z, domZ, bit cutoffs) are NOT applied here: every
 * comparison that p7_Pipeline would have scored to completion is returned, and the caller
 * applies p7_pli_TargetReportable with the running Z exactly as the sequential loop would. */
typedef struct {
  double   F1, F2, F3;        /* P-value thresholds of the three filters (0.02, 1e-3, 1e-5)          */
  int32_t  do_biasfilter;     /* pli->do_biasfilter                                                  */
  int32_t  do_null2;          /* pli->do_null2                                                       */
  uint32_t seed;              /* RNG seed for stochastic traceback clustering (42)                   */
  int32_t  host_threads;      /* worker threads for the host-side domain definition (0 = all cores) */
  int32_t  seq_counters;      /* 1: also count the pass counters per SEQUENCE (b2h_results_seq_counters): what every
                                 scan_seq result of a multi-query hmmscan reports (plan7.pyx:6534-6677)      */
  int32_t  reserved;
} b2h_search_params;

typedef struct {              /* mirrors P7_HIT (hmmer.h:711-743) without strings */
  int32_t  profile;           /* index into the profiles[] argument                */
  int32_t  seq;               /* index into the sequence database                  */
  float    score, pre_score, sum_score;     /* bits */
  double   lnP, pre_lnP, sum_lnP;
  float    nexpected;
  int32_t  nregions, nclustered, noverlaps, nenvelopes, ndom;
  int32_t  best_domain;
  int64_t  dom_offset;        /* first domain of this hit in the domains array     */
} b2h_hit;

typedef struct {              /* mirrors P7_DOMAIN + P7_ALIDISPLAY coordinates (hmmer.h:614-628, 580-607) */
  int32_t  ienv, jenv, iali, jali;
  float    envsc, domcorrection, dombias, oasc, bitscore;
  double   lnP;
  int32_t  hmmfrom, hmmto, sqfrom, sqto;
  int32_t  N;                 /* alignment display length                          */
  int64_t  text_offset;       /* model | mline | aseq | ppline [| rfline] [| csline], each N+1 bytes, NUL-terminated */
  int32_t  has_rf, has_cs;
} b2h_domain;

typedef struct b2h_results b2h_results;

int b2h_search(b2h_ctx *ctx, const b2h_profile *const *profiles, size_t P, const b2h_seqdb *db,
               const b2h_search_params *params, b2h_results **out);
size_t            b2h_results_nhits   (const b2h_results *r);
const b2h_hit    *b2h_results_hits    (const b2h_results *r);
size_t            b2h_results_ndomains(const b2h_results *r);
const b2h_domain *b2h_results_domains (const b2h_results *r);
const char       *b2h_results_text    (const b2h_results *r, size_t *nbytes);
/* [P][4] = n_past_msv, n_past_bias, n_past_vit, n_past_fwd per profile (P7_PIPELINE counters, hmmer.h:1228-1241) */
const int64_t    *b2h_results_counters(const b2h_results *r);
/* [N][4] the same four counters per sequence of the database (summed over the profiles); NULL unless params.seq_counters */
const int64_t    *b2h_results_seq_counters(const b2h_results *r);
void              b2h_results_destroy (b2h_results *r);

/* The same search with its results handed out wave by wave.  b2h_search processes the profiles in a few waves (longest
 * models first); every comparison of a wave's profiles is final as soon as the wave is through domain definition, while
 * the GPU is already busy with the waves behind it.  pyhmmer.hmmsearch is a generator for the same reason -- results of the
 * first queries are consumed while later ones are still being searched (src/pyhmmer/hmmer/_hmmsearch.py:294-420,
 * _base.py:_BaseDispatcher.run) -- so the host-side assembly of `TopHits` (and, on several GPUs, the exchange of hit
 * records) of wave w overlaps the cascade of waves w+1...; only the last, smallest wave's post-processing is exposed.
 *   b2h_search_begin   starts the search on a driver thread of the library; *nwaves = number of waves that will come.
 *                      The context must not be used for anything else until b2h_search_end.  params.seq_counters must be 0.
 *   b2h_search_next    blocks until the next wave is complete: *out = its results (hits carry the caller's profile
 *                      indices, ordered by (profile, target); counters are [P][4] with the rows of the wave's profiles
 *                      filled; b2h_results_profiles lists them).  The caller destroys every result.  *out = NULL and the
 *                      status of the search after the last wave.
 *   b2h_search_end     joins the driver thread, frees the job; returns the status of the search. */
typedef struct b2h_search_job b2h_search_job;
int b2h_search_begin(b2h_ctx *ctx, const b2h_profile *const *profiles, size_t P, const b2h_seqdb *db,
                     const b2h_search_params *params, b2h_search_job **out, size_t *nwaves);
int b2h_search_next (b2h_search_job *job, b2h_results **out);
int b2h_search_end  (b2h_search_job *job);
const int32_t    *b2h_results_profiles(const b2h_results *r, size_t *n);

/* Diagnostic (used by the CPU-only tests of the host-side domain definition): build a profile object
 * without any device state, and run the post-Backward part of p7_Pipeline for ONE comparison from given
 * Forward/Backward parser specials ((L+1) rows of {E,N,J,B,C,SCALE}).  dsq[0..L-1] are the residues. */
int b2h_profile_create_host(const b2h_oprofile_desc *desc, b2h_profile **out);
int b2h_debug_domaindef(const b2h_profile *p, const uint8_t *dsq, int L, const float *fwd_xmx, const float *bck_xmx,
                        float fwdsc, const b2h_search_params *params, b2h_results **out);

/* Annotation lines needed to render alignments (P7_OPROFILE.consensus / rf / cs, 1..M); any may be NULL. */
int b2h_profile_set_annotation(b2h_profile *p, const char *consensus, const char *rf, const char *cs,
                               const char *alphabet_symbols /* ESL_ALPHABET.sym, Kp chars */);

/* The model mask (P7_OPROFILE.mm, the HMM file's MM line; 'm' marks a masked node), 1..M; NULL = none.  Only the long-target
 * path reads it: masked nodes keep a zero match score when the background is re-estimated (p7_oprofile.c:455). */
int b2h_profile_set_model_mask(b2h_profile *p, const char *mm);

/* Bytes copied host->device when the object was made resident (bench.py reports them as e2e.h2d_bytes_per_step). */
size_t b2h_seqdb_h2d_bytes(const b2h_seqdb *db);
size_t b2h_profile_h2d_bytes(const b2h_profile *p);

/* ---- pressed profile databases ------------------------------------------------------------------------------------
 * Bulk reader for <db>.h3f + <db>.h3p (hmmpress, format 3/f), standing in for p7_oprofile_ReadMSV / p7_oprofile_ReadRest
 * (vendor/hmmer/src/impl_sse/io.c:231, 498) when a whole database goes to the GPU.  b2h_pressed_read returns up to
 * max_models descriptors whose table pointers point into ONE malloc'ed block of node-major tables (ready for
 * b2h_profile_upload_many; desc.bgf and desc.degen are left for the caller, who knows the alphabet), plus the models'
 * strings in one NUL-separated text block (offsets in the model records, -1 = absent).  *nread = 0 at the end of the
 * database.  The three arrays are released with b2h_free.  Host only: no device is touched. */
typedef struct b2h_pressed b2h_pressed;
typedef struct {
  b2h_oprofile_desc desc;
  int32_t alphabet_type;                               /* eslRNA = 1, eslDNA = 2, eslAMINO = 3 (esl_alphabet.h) */
  int32_t reserved;
  int64_t name, acc, descr, rf, mm, cs, consensus;     /* offsets into the text block */
} b2h_pressed_model;
int  b2h_pressed_open(const char *base_path, b2h_pressed **out);
void b2h_pressed_close(b2h_pressed *h);
int  b2h_pressed_rewind(b2h_pressed *h);
const char *b2h_pressed_last_error(const b2h_pressed *h);
int  b2h_pressed_read(b2h_pressed *h, size_t max_models, b2h_pressed_model **models, size_t *nread,
                      void **block, size_t *block_bytes, char **text, size_t *text_bytes);

/* Batched p7_ProfileConfig + p7_oprofile_Convert for a block of HMM queries (what Pipeline.search_hmm does per query,
 * src/pyhmmer/plan7.pyx:5979-6013): t [(M+1)*7] and mat [(M+1)*K] as for b2h_profile_config.  The descriptors (malloc'ed
 * array of n) point into ONE malloc'ed block of node-major tables; both are released with b2h_free.  Host only. */
typedef struct {
  int32_t M, max_length;
  const float *t, *mat;
  float evparam[B2H_NEVPARAM], cutoff[B2H_NCUTOFFS], compo[B2H_MAXABET];
} b2h_hmm_desc;
int  b2h_hmm_convert_many(int K, int Kp, const uint8_t *degen, const float *bgf, int L, int multihit,
                          const b2h_hmm_desc *hmms, size_t n, int nthreads,
                          b2h_oprofile_desc **descs, void **block, size_t *block_bytes);

/* p7_Builder_MaxLength (vendor/hmmer/src/p7_builder.c:651): the window length (MAXL) of an HMM for the tail mass emit_thresh
 * (nhmmer's --w_beta, default 1e-7), from its transition probabilities t [(M+1)*7] = MM MI MD IM II DM DD.  Host only. */
int  b2h_hmm_max_length(int M, const float *t, double emit_thresh, int32_t *max_length);

/* ---- long-target (nhmmer) path, first stage ------------------------------------------------------------------------
 * p7_SSVFilter_longtarget (vendor/hmmer/src/impl_sse/msvfilter.c:256) over every sequence of <db> (= the chunks a long
 * target was cut into; LongTargetsPipeline, plan7.pyx:7542-7663), with the model's length parameters set for its
 * max_length (p7_oprofile_ReconfigMSVLength), then p7_hmm_ScoreDataComputeRest + p7_pli_ExtendAndMergeWindows(.., 0)
 * (p7_scoredata.c:313, p7_pipeline.c:323).  <raw>: the SSV diagonals {chunk, start n (1-based), model end k, length,
 * score in nats} in scan order per chunk; <merged>: the windows handed to the next stage {chunk, start, -, length}.
 * Both arrays are malloc'ed; release them with b2h_free.  B2H_EINVAL if the model carries no max_length (MAXL). */
typedef struct { int32_t seq; int32_t k; int64_t n; int32_t length; float score; } b2h_window;
int  b2h_longtarget_windows(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *db, double F1,
                            b2h_window **raw, size_t *nraw, b2h_window **merged, size_t *nmerged);
/* What that scan uses for a profile and F1 (host only): the threshold on the byte scale (msvfilter.c:289-327) and whether the
 * two-instruction cell applies (begin floor < threshold <= 256 - bias: the saturation of adds_epu8 cannot bind on a kept cell;
 * otherwise the byte arithmetic is written out in full).  Either way the diagonals are the reference's. */
int  b2h_longtarget_scan_info(const b2h_profile *p, double F1, int *sc_thresh, int *fast_cells);
void b2h_free(void *p);
/* The prefix / suffix length tables [M+1] of p7_hmm_ScoreDataComputeRest for a profile (host only; works on a profile made by
 * b2h_profile_create_host). */
int  b2h_window_lengths(const b2h_profile *p, float *prefix, float *suffix);
/* p7_pli_ExtendAndMergeWindows (p7_pipeline.c:323) on a caller-provided list, in place (host only): target_len[i] = length of the
 * chunk window i lies in; pct_overlap = 0 after SSV, 0.5 after Viterbi.  *nout = number of windows left. */
int  b2h_extend_merge_windows(const b2h_profile *p, b2h_window *windows, size_t n, const int64_t *target_len, float pct_overlap, size_t *nout);

/* The Viterbi stage of the long-target pipeline (p7_pli_postSSV_LongTarget, p7_pipeline.c:1369-1412) for the SSV windows
 * that passed the MSV and bias gates: <windows> holds one sequence per window (the window's residues); filtersc[i] = the
 * null1 score at min(window, max_length) plus the B2-scaled bias correction, as the caller computed it; active[i] = 0 skips
 * a window (NULL = all).  Every window is scanned by p7_ViterbiFilter_longtarget (impl_sse/vitfilter.c:292) with the profile
 * configured for min(window, max_length) (p7_oprofile_ReconfigRestLength) and the score threshold of P-value F2.
 * <marks>: the landmarks {seq = window, n = row i, k, length 1} in the reference's order; <out>: the windows handed to the
 * Forward stage after p7_pli_ExtendAndMergeWindows(.., 0.5) and the 80 kb cut, {seq = window, n = start inside the window,
 * length}.  Both malloc'ed (b2h_free).  B2H_EINVAL for models without max_length or above 1536 nodes. */
int  b2h_longtarget_viterbi_windows(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *windows, const float *filtersc,
                                    const uint8_t *active, double F2,
                                    b2h_window **marks, size_t *nmarks, b2h_window **out, size_t *nout);
/* The host half of the call above on its own (host only; works on a profile made by b2h_profile_create_host): <marks> is put
 * into the reference's order in place, then extended / merged / cut into <out>.  window_len[w] = length of window w. */
int  b2h_longtarget_vit_finish(const b2h_profile *p, b2h_window *marks, size_t nmarks, const int32_t *window_len, size_t nwindows,
                               b2h_window **out, size_t *nout);
/* The int16 score threshold and the N/C/J move score p7_ViterbiFilter_longtarget works with for a window whose profile is
 * configured for cfg_len (vitfilter.c:330-346; host only). */
int  b2h_longtarget_vit_threshold(const b2h_profile *p, int cfg_len, float filtersc, double F2, int32_t *thresh, int32_t *xw_move);

/* Behind the Forward gate (p7_pli_postViterbi_LongTarget, p7_pipeline.c:1113-1280), host only: for every window that passed,
 * given its residues and its Forward / Backward parser specials ((L+1) rows of {E,N,J,B,C,SCALE}, profile configured for the
 * window's length), run domain definition with its long-target branches (p7_domaindef.c:814-982: envelope-length
 * configuration, background re-estimated from the envelope, envelope trimmed to the alignment, bias = score without the
 * re-estimation) and build one hit per domain: scores corrected to the model's max_length window, coordinates mapped to the
 * target (window_start = position of the window in the chunk, seq_start = the chunk's first coordinate, the last one for the
 * complement strand).  hit.profile = index of the window; hit.seq = the window's <seq>.  Works on host profiles. */
typedef struct {
  const uint8_t *dsq; int32_t L;
  const float *fwd_xmx, *bck_xmx;
  int64_t window_start, seq_start;
  int32_t complement, seq;
  int32_t bck_own_scales;     /* 0: bck_xmx uses Forward's scale factors, as the reference's Backward does */
  int32_t reserved;
} b2h_lt_window;
int  b2h_longtarget_domains(const b2h_profile *p, const b2h_lt_window *windows, size_t n, const b2h_search_params *params,
                            b2h_results **out);
/* The same with the parser passes on the GPU: <windows> holds the windows that passed the Forward gate (one sequence each);
 * Forward and Backward parsers run on the device for all of them (profile configured for each window's length), their special
 * rows come back, domain definition runs on the host threads.  window_start / seq_start / complement / seq: [n] per window. */
int  b2h_longtarget_hits(b2h_ctx *ctx, const b2h_profile *p, const b2h_seqdb *windows, const int64_t *window_start,
                         const int64_t *seq_start, const int32_t *complement, const int32_t *seq,
                         const b2h_search_params *params, b2h_results **out);

/* The windows of a long-target search as one packed residue buffer (what LongTargetsPipeline._search_loop_longtargets hands
 * to p7_Pipeline_LongTarget window by window, plan7.pyx:7541-7663): window w = residues [offset, offset + len) of target
 * win_target[w] (0-based offset, no sentinels), reverse-complemented (esl_sq_ReverseComplement, esl_sq.c:1545: the window
 * reversed, every code mapped through comp_table[Kp]) when win_comp[w] != 0; written at out + out_off[w].  Host only,
 * <nthreads> threads (0 = all cores). */
int b2h_pack_windows(const uint8_t *const *targets, size_t nwin, const int32_t *win_target, const int64_t *win_offset,
                     const int64_t *win_len, const int32_t *win_comp, const uint8_t *comp_table, int Kp,
                     uint8_t *out, const int64_t *out_off, int nthreads);

/* Register tile the SSV kernel uses for a model of M nodes: G lanes per comparison (32/G comparisons per warp), NR packed
 * registers (2*NR nodes) per lane, and the number of 128-byte shared-memory wavefronts one DP row of one WARP moves
 * (emission loads + shuffles) -- the quantity bench.py's on-chip roofline is computed from.  B2H_EINVAL if M > 3071. */
int b2h_ssv_tile_info(int M, int *G, int *NR, double *wavefronts_per_row);

#ifdef __cplusplus
}
#endif
#endif /* B2H_H_INCLUDED */
