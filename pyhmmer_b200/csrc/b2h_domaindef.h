// b2h_domaindef.h -- host-side completion of the F3 survivors into hits (not part of the ABI).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "b2h.h"

struct b2h_profile;

struct b2h_results {
  std::vector<b2h_hit>    hits;
  std::vector<b2h_domain> doms;
  std::vector<char>       text;
  std::vector<int64_t>    counters;
  std::vector<int64_t>    seq_counters;   // [N][4], only when asked for
  std::vector<int32_t>    profiles;       // the profiles this object is final for (results of one wave: b2h_search_next)
};

struct b2h_survivor { int32_t profile, seq; float fwdsc, filtersc; int64_t fxoff = -1; };   // fxoff: row offset of the Forward specials the cascade stored (-1: none)

struct b2h_ddef_task {
  b2h_survivor       surv;
  const b2h_profile *prof;
  const uint8_t     *dsq;            // residues 1..L at dsq[0..L-1]
  int                L;
  const float       *fx, *bx;        // Forward / Backward parser specials, (L+1) rows of {E,N,J,B,C,SCALE}
  bool               bck_own_scales;
};

// One envelope handed to a rescoring backend, and what comes back (b2h_envelope.cu fills these on the GPU).
struct b2h_env_job {
  int task = 0;                    // index into the task list
  int i = 0, j = 0;                // envelope in the target, 1-based
  const float *rsc = nullptr;      // long targets: this envelope's re-estimated emission odds [Kp][M] (nullptr: the profile's)
  int cfg_len = 0;                 // length the profile is configured for (0: the task's sequence length; long targets: the envelope)
  bool fwd_only = false;           // only the Forward score is wanted
  // results
  int   status = -1;               // 0 = done; anything else: rescore this envelope on the host
  float envsc = 0.f, oasc = 0.f;
  float xn = 0.f, xc = 0.f, xj = 0.f;          // summed posteriors of N, C, J over the envelope (null2)
  std::vector<float> em, ei;       // summed posteriors of M / I per node, index k-1 (null2)
  std::vector<int32_t> trace;      // optimal-accuracy trace in traceback order: records {state, k, i, bits(postprob)}
};
struct b2h_env_backend {
  virtual ~b2h_env_backend() {}
  virtual int run(const std::vector<b2h_ddef_task> &tasks, std::vector<b2h_env_job> &jobs) = 0;
};

struct b2h_ddef_pool {
  int nthreads;
  explicit b2h_ddef_pool(int n);
  // runs every task (in parallel), appends the resulting hits to <res> in task order.  With a backend the numeric
  // rescoring of the envelopes runs there (GPU); without (host-only profiles, tests) on the pool's threads.
  int run(std::vector<b2h_ddef_task> &tasks, const b2h_search_params *prm, b2h_results *res, b2h_env_backend *backend = nullptr);
};

// Long-target (nhmmer) windows behind the Forward gate -> hits, one per domain (p7_pli_postViterbi_LongTarget); hit.profile
// carries the index of the window the hit came from.
int b2h_longtarget_domains_host(const b2h_profile *p, const b2h_lt_window *wins, size_t n, const b2h_search_params *prm, int nthreads, b2h_results *res);
// The same, the envelope rescoring on <backend> (nullptr: host) batched over all windows; db_index[w] = index of window w in the
// backend's sequence database.
int b2h_longtarget_domains_backend(const b2h_profile *p, const b2h_lt_window *wins, const int32_t *db_index, size_t n, const b2h_search_params *prm,
                                   int nthreads, b2h_env_backend *backend, b2h_results *res);
