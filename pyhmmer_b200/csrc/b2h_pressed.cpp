// b2h_pressed.cpp -- bulk reader for hmmpress'ed profile databases (<db>.h3f + <db>.h3p, format 3/f).
//
// Stands in for p7_oprofile_ReadMSV / p7_oprofile_ReadRest (vendor/hmmer/src/impl_sse/io.c:231, 498) when a whole
// database goes to the GPU: models are read in batches, their SSE-striped tables are de-striped straight into ONE
// block per batch (node-major, the layout b2h_profile_upload wants) and the descriptors point into that block, so the
// host mirror wraps a batch without touching a table.  The byte layout is the x86-64 reference build's
// (little-endian, 8-byte off_t); field order follows p7_oprofile_Write (io.c:84-176).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "b2h.h"

namespace {

constexpr uint32_t FMAGIC = 0xb3e6e6f3u, PMAGIC = 0xb3e6f0f3u;    // v3f_fmagic / v3f_pmagic, io.c:46-47
constexpr int EXTRA_SB = 17;                                       // p7O_EXTRA_SB, impl_sse.h:28

int alphabet_sizes(int type, int *K, int *Kp)
{
  if (type == 3) { *K = 20; *Kp = 29; return 1; }                  // eslAMINO
  if (type == 1 || type == 2) { *K = 4; *Kp = 18; return 1; }      // eslRNA, eslDNA
  return 0;
}

template <typename T> bool rd(FILE *f, T *v, size_t n = 1) { return fread(v, sizeof(T), n, f) == n; }

bool rd_string(FILE *f, int n, std::string &s)
{
  s.assign((size_t)n + 1, '\0');
  if (fread(&s[0], 1, (size_t)n + 1, f) != (size_t)n + 1) return false;
  s.resize(n);
  return true;
}

} // namespace

struct b2h_pressed { FILE *ffp = nullptr, *pfp = nullptr; std::string err; };

extern "C" {

int b2h_pressed_open(const char *base_path, b2h_pressed **out)
{
  if (!base_path || !out) return B2H_EINVAL;
  *out = nullptr;
  b2h_pressed *h = new b2h_pressed();
  const std::string b(base_path);
  h->ffp = fopen((b + ".h3f").c_str(), "rb");
  h->pfp = fopen((b + ".h3p").c_str(), "rb");
  if (!h->ffp || !h->pfp) { if (h->ffp) fclose(h->ffp); if (h->pfp) fclose(h->pfp); delete h; return B2H_EINVAL; }
  *out = h;
  return B2H_OK;
}

void b2h_pressed_close(b2h_pressed *h)
{
  if (!h) return;
  fclose(h->ffp); fclose(h->pfp);
  delete h;
}

int b2h_pressed_rewind(b2h_pressed *h)
{
  if (!h) return B2H_EINVAL;
  rewind(h->ffp); rewind(h->pfp);
  return B2H_OK;
}

const char *b2h_pressed_last_error(const b2h_pressed *h) { return h ? h->err.c_str() : ""; }

int b2h_pressed_read(b2h_pressed *h, size_t max_models, b2h_pressed_model **models_out, size_t *nread,
                     void **block_out, size_t *block_bytes, char **text_out, size_t *text_bytes)
{
  if (!h || !models_out || !nread || !block_out || !block_bytes || !text_out || !text_bytes) return B2H_EINVAL;
  *models_out = nullptr; *block_out = nullptr; *text_out = nullptr; *nread = 0; *block_bytes = 0; *text_bytes = 0;
  struct Raw { int M, Kp; std::vector<uint8_t> rbv; std::vector<int16_t> twv, rwv; std::vector<float> tfv, rfv; };
  std::vector<b2h_pressed_model> models;
  std::vector<Raw> raws;
  std::string text;
  auto fail = [&](const char *msg) { h->err = msg; return B2H_EINVAL; };
  auto put = [&](const std::string &s) -> int64_t { const int64_t o = (int64_t)text.size(); text.append(s); text.push_back('\0'); return o; };
  size_t total = 0;                                                  // bytes of de-striped tables
  while (models.size() < max_models) {
    uint32_t magic;
    if (!rd(h->ffp, &magic)) break;                                  // clean end of the database
    if (magic != FMAGIC) return fail("bad magic in .h3f: not a 3/f pressed database (hmmpress it again with HMMER >= 3.1)");
    b2h_pressed_model m;
    memset(&m, 0, sizeof m);
    b2h_oprofile_desc &d = m.desc;
    int32_t M, atype, n;
    std::string name, name2, acc, desc;
    // ---- .h3f: the MSV part ----
    if (!rd(h->ffp, &M) || !rd(h->ffp, &atype) || !rd(h->ffp, &n) || M < 1 || n < 0 || !rd_string(h->ffp, n, name)) return fail("truncated .h3f");
    int K, Kp;
    if (!alphabet_sizes(atype, &K, &Kp)) return fail("unsupported alphabet type in pressed database");
    const int Q16 = std::max(2, (M - 1) / 16 + 1), Q8 = std::max(2, (M - 1) / 8 + 1), Q4 = std::max(2, (M - 1) / 4 + 1);
    Raw r; r.M = M; r.Kp = Kp;
    int32_t max_length; float scale_b;
    if (!rd(h->ffp, &max_length) || !rd(h->ffp, &d.tbm_b) || !rd(h->ffp, &d.tec_b) || !rd(h->ffp, &d.tjb_b) || !rd(h->ffp, &scale_b) ||
        !rd(h->ffp, &d.base_b) || !rd(h->ffp, &d.bias_b)) return fail("truncated .h3f");
    if (fseek(h->ffp, (long)Kp * (Q16 + EXTRA_SB) * 16, SEEK_CUR) != 0) return fail("truncated .h3f");      // sbv: SSV copy of rbv
    r.rbv.resize((size_t)Kp * Q16 * 16);
    uint64_t offs[3];
    if (!rd(h->ffp, r.rbv.data(), r.rbv.size()) || !rd(h->ffp, d.evparam, 6) || !rd(h->ffp, offs, 3) || !rd(h->ffp, d.compo, 20) ||
        !rd(h->ffp, &magic) || magic != FMAGIC) return fail("truncated or corrupted .h3f record");
    // ---- .h3p: the rest ----
    int32_t M2, atype2;
    if (!rd(h->pfp, &magic) || magic != PMAGIC) return fail("bad magic in .h3p");
    if (!rd(h->pfp, &M2) || !rd(h->pfp, &atype2) || !rd(h->pfp, &n) || n < 0 || !rd_string(h->pfp, n, name2)) return fail("truncated .h3p");
    if (M2 != M || atype2 != atype || name2 != name) return fail(".h3f and .h3p are out of step");
    if (!rd(h->pfp, &n) || n < 0 || (n > 0 && !rd_string(h->pfp, n, acc))) return fail("truncated .h3p");
    const bool has_acc = n > 0;
    if (!rd(h->pfp, &n) || n < 0 || (n > 0 && !rd_string(h->pfp, n, desc))) return fail("truncated .h3p");
    const bool has_desc = n > 0;
    std::string ann[4];
    for (int a = 0; a < 4; a++) { ann[a].assign((size_t)M + 2, '\0'); if (fread(&ann[a][0], 1, (size_t)M + 2, h->pfp) != (size_t)M + 2) return fail("truncated .h3p"); }
    r.twv.resize((size_t)8 * Q8 * 8); r.rwv.resize((size_t)Kp * Q8 * 8); r.tfv.resize((size_t)8 * Q4 * 4); r.rfv.resize((size_t)Kp * Q4 * 4);
    int16_t xw[8]; float xf[8], ncj, nj; int32_t mode, L;
    if (!rd(h->pfp, r.twv.data(), r.twv.size()) || !rd(h->pfp, r.rwv.data(), r.rwv.size()) || !rd(h->pfp, xw, 8) || !rd(h->pfp, &d.scale_w) ||
        !rd(h->pfp, &d.base_w) || !rd(h->pfp, &d.ddbound_w) || !rd(h->pfp, &ncj) ||
        !rd(h->pfp, r.tfv.data(), r.tfv.size()) || !rd(h->pfp, r.rfv.data(), r.rfv.size()) || !rd(h->pfp, xf, 8) ||
        !rd(h->pfp, d.cutoff, 6) || !rd(h->pfp, &nj) || !rd(h->pfp, &mode) || !rd(h->pfp, &L) ||
        !rd(h->pfp, &magic) || magic != PMAGIC) return fail("truncated or corrupted .h3p record");
    d.M = M; d.K = K; d.Kp = Kp; d.L = L; d.max_length = max_length; d.mode_multihit = nj > 0.0f; d.scale_b = scale_b;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 2; j++) { d.xw[i][j] = xw[i * 2 + j]; d.xf[i][j] = xf[i * 2 + j]; }
    m.alphabet_type = atype;
    m.name = put(name); m.acc = has_acc ? put(acc) : -1; m.descr = has_desc ? put(desc) : -1;
    int64_t *slots[4] = {&m.rf, &m.mm, &m.cs, &m.consensus};
    for (int a = 0; a < 4; a++) *slots[a] = (ann[a][1] != '\0') ? put(ann[a].substr(1, M)) : -1;   // absent annotation: NUL at position 1
    total += (size_t)Kp * M * (1 + 2 + 4) + (size_t)8 * M * (2 + 4) + 96;        // + alignment slack of the five tables
    models.push_back(m);
    raws.push_back(std::move(r));
  }
  if (models.empty()) return B2H_OK;
  uint8_t *block = (uint8_t *)malloc(total);
  b2h_pressed_model *mo = (b2h_pressed_model *)malloc(models.size() * sizeof(b2h_pressed_model));
  char *tx = (char *)malloc(std::max<size_t>(1, text.size()));
  if (!block || !mo || !tx) { free(block); free(mo); free(tx); h->err = "out of memory"; return B2H_EMEM; }
  size_t off = 0;
  auto take = [&](size_t bytes) { void *p = block + off; off = (off + bytes + 15) & ~(size_t)15; return p; };
  for (size_t i = 0; i < models.size(); i++) {
    const Raw &r = raws[i];
    b2h_oprofile_desc &d = models[i].desc;
    const size_t KM = (size_t)r.Kp * r.M, TM = (size_t)8 * r.M;
    float *fr = (float *)take(KM * 4), *ft = (float *)take(TM * 4);
    int16_t *vr = (int16_t *)take(KM * 2), *vt = (int16_t *)take(TM * 2);
    uint8_t *mc = (uint8_t *)take(KM);
    b2h_destripe_oprofile(r.M, r.Kp, r.rbv.data(), r.rwv.data(), r.twv.data(), r.rfv.data(), r.tfv.data(), mc, vr, vt, fr, ft);
    d.msv_cost = mc; d.vit_rsc = vr; d.vit_tsc = vt; d.fwd_rsc = fr; d.fwd_tsc = ft;
    mo[i] = models[i];
  }
  memcpy(tx, text.data(), text.size());
  *models_out = mo; *nread = models.size(); *block_out = block; *block_bytes = off; *text_out = tx; *text_bytes = text.size();
  return B2H_OK;
}

// Batched form of p7_ProfileConfig + p7_oprofile_Convert (b2h_profile_config, b2h_oprofile_convert) for a block of HMM
// queries: the models are converted on <nthreads> host threads into ONE block of node-major tables, descriptors pointing
// into it -- what Pipeline.search_hmm does per query (plan7.pyx:5979-6013), once per query block.
int b2h_hmm_convert_many(int K, int Kp, const uint8_t *degen, const float *bgf, int L, int multihit,
                         const b2h_hmm_desc *hmms, size_t n, int nthreads,
                         b2h_oprofile_desc **descs_out, void **block_out, size_t *block_bytes)
{
  if (!degen || !bgf || (!hmms && n) || !descs_out || !block_out || !block_bytes || K < 1 || K > B2H_MAXABET || Kp < K + 3 || Kp > B2H_MAXCODE) return B2H_EINVAL;
  *descs_out = nullptr; *block_out = nullptr; *block_bytes = 0;
  if (n == 0) return B2H_OK;
  std::vector<size_t> base(n + 1, 0);
  for (size_t i = 0; i < n; i++) {
    if (hmms[i].M < 1 || !hmms[i].t || !hmms[i].mat) return B2H_EINVAL;
    const size_t M = hmms[i].M;
    base[i + 1] = base[i] + (((size_t)Kp * M * (1 + 2 + 4) + (size_t)8 * M * (2 + 4) + 96 + 15) & ~(size_t)15);
  }
  uint8_t *block = (uint8_t *)malloc(base[n]);
  b2h_oprofile_desc *descs = (b2h_oprofile_desc *)calloc(n, sizeof(b2h_oprofile_desc));
  if (!block || !descs) { free(block); free(descs); return B2H_EMEM; }
  std::vector<int> status(n, B2H_OK);
  const int T = (int)std::min<size_t>((size_t)std::max(1, nthreads), n);
  auto work = [&](int t) {
    std::vector<float> tsc, msc; float xsc[8];
    for (size_t i = t; i < n; i += T) {
      const b2h_hmm_desc &h = hmms[i];
      const int M = h.M;
      tsc.resize((size_t)M * 8); msc.resize((size_t)Kp * (M + 1));
      int st = b2h_profile_config(M, K, Kp, degen, h.t, h.mat, bgf, L, multihit, tsc.data(), msc.data(), xsc);
      size_t off = base[i];
      auto take = [&](size_t bytes) { void *p = block + off; off = (off + bytes + 15) & ~(size_t)15; return p; };
      const size_t KM = (size_t)Kp * M, TM = (size_t)8 * M;
      float *fr = (float *)take(KM * 4), *ft = (float *)take(TM * 4);
      int16_t *vr = (int16_t *)take(KM * 2), *vt = (int16_t *)take(TM * 2);
      uint8_t *mc = (uint8_t *)take(KM);
      b2h_oprofile_desc &d = descs[i];
      if (st == B2H_OK) st = b2h_oprofile_convert(M, K, Kp, L, multihit, tsc.data(), msc.data(), xsc, mc, vr, vt, fr, ft, &d);
      d.msv_cost = mc; d.vit_rsc = vr; d.vit_tsc = vt; d.fwd_rsc = fr; d.fwd_tsc = ft;
      d.max_length = h.max_length;
      memcpy(d.evparam, h.evparam, sizeof d.evparam); memcpy(d.cutoff, h.cutoff, sizeof d.cutoff); memcpy(d.compo, h.compo, sizeof d.compo);
      for (int x = 0; x < B2H_MAXABET; x++) d.bgf[x] = x < K ? bgf[x] : 0.0f;
      d.degen = degen;
      status[i] = st;
    }
  };
  if (T <= 1) work(0);
  else { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(work, t); for (auto &x : th) x.join(); }
  for (size_t i = 0; i < n; i++) if (status[i] != B2H_OK) { const int st = status[i]; free(block); free(descs); return st; }
  *descs_out = descs; *block_out = block; *block_bytes = base[n];
  return B2H_OK;
}

} // extern "C"
