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
};

struct b2h_survivor { int32_t profile, seq; float fwdsc, filtersc; };

struct b2h_ddef_task {
  b2h_survivor       surv;
  const b2h_profile *prof;
  const uint8_t     *dsq;            // residues 1..L at dsq[0..L-1]
  int                L;
  const float       *fx, *bx;        // Forward / Backward parser specials, (L+1) rows of {E,N,J,B,C,SCALE}
  bool               bck_own_scales;
};

struct b2h_ddef_pool {
  int nthreads;
  explicit b2h_ddef_pool(int n);
  // runs every task (in parallel), appends the resulting hits to <res> in task order
  int run(std::vector<b2h_ddef_task> &tasks, const b2h_search_params *prm, b2h_results *res);
};
