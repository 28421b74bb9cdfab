/*
 * sa_engine.cu -- batch alignment engine behind include/seqalign_b200.h.
 *
 * Owns one CUDA stream, grow-only device / pinned-host buffers and the
 * launch logic: scan the batch's alphabet, flatten scoring_t into a dense
 * table, pick a kernel variant, fill (+ direction bytes + walk in align
 * mode, in memory-bounded waves), copy results back.  There is no CPU
 * implementation of the DP in here: without a usable device every entry
 * point fails with SEQALIGN_ERR_CUDA.
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "seqalign_b200.h"
#include "sa_platform.h"
#include "sa_flatten.h"
#include "sa_kernels.cuh"
#include "sa_fast.cuh"
#include "sa_hits.cuh"
#include "sa_mats.cuh"
#include "sa_long.cuh"
#include "sa_synth.cuh"

using namespace sa;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  void release() { if(p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
  void *p = nullptr;
  size_t cap = 0;
  void release() { if(p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

} // namespace

struct seqalign_batch {
  int device = 0;
  int num_sms = 0;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t scan_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  enum { MAX_CHUNKS = 32 };
  cudaEvent_t ev_copy[MAX_CHUNKS] = {}, ev_k0[MAX_CHUNKS] = {}, ev_k1[MAX_CHUNKS] = {}, ev_scan[MAX_CHUNKS] = {};
  std::string err;
  char unk_a = 0, unk_b = 0;

  bool have_scoring = false;
  scoring_t *scoring = nullptr;
  FlatTable ft;
  unsigned scoring_version = 0;
  /* what the device tables currently hold */
  bool tables_valid = false;
  unsigned tables_version = 0;
  uint64_t tables_pres[8] = {};
  std::vector<int32_t> dev_tab32;
  std::vector<int8_t> dev_tab8;
  /* plan of the last device-resident run, reused speculatively by the next one */
  struct {
    bool valid = false;
    unsigned version = 0;
    int algo = 0;
    bool want_ends = false;
    FastPlan plan;
    int64_t max_lb = 0;
  } spec;
  int spec_hits = 0, spec_misses = 0;
  /* device-resident runs launched ahead of their verification (seqalign_batch_run_device_async) */
  struct Pending {
    int slot, algo;
    const void *a, *oa, *b, *ob;
    size_t n;
    void *ds, *dx, *dy;
    void *stream;
  };
  enum { MAX_PENDING = 4 };
  Pending pend[MAX_PENDING];
  int pend_head = 0, pend_count = 0;
  int force_mode = 0; /* 0 auto, 1 general kernel, 2 fast + per-column keys, 3 fast without end cell, 4 = 3 but int32 only, 5 = auto but int32 only */

  /* inputs on device */
  DevBuf d_seq_a, d_seq_b, d_off_a, d_off_b;
  /* scratch */
  DevBuf d_meta, d_counter, d_sub, d_forbid, d_lut, d_tab8, d_bnd, d_lbnd, d_swkey, d_bucket;
  /* score-mode results */
  DevBuf d_score, d_xend, d_yend, d_state;
  /* align-mode wave buffers */
  DevBuf d_dir, d_dir_off, d_out_a, d_out_b, d_out_off, d_walk;
  /* multi-hit mode */
  DevBuf d_m16, d_keys0, d_keys1, d_mask, d_ncand, d_which, d_nhits, d_rec;
  /* materialise mode */
  DevBuf d_mats, d_mat_off;
  std::vector<int64_t> mat_off;            /* batch materialise: first int of pair i's match plane, n+1 entries */
  /* batch materialise in waves: pairs [mat_wave[w], mat_wave[w+1]) fit the device block together; one wave
   * is resident at a time and seqalign_batch_matrices() re-runs the kernel for the wave it is asked about */
  enum { BUCKET_STREAMS = 4 };
  cudaStream_t bucket_streams[BUCKET_STREAMS] = {};
  cudaEvent_t bucket_done[BUCKET_STREAMS] = {}, bucket_ready = nullptr;
  std::vector<size_t> mat_wave;
  int mat_resident = -1;
  struct { const uint8_t *a, *b; const int64_t *off_a, *off_b; int NB; bool pack, nw; ScoreParams sp; } mat_job;
  PinBuf h_in_a, h_in_b, h_off_a, h_off_b, h_meta, h_res, h_walk, h_str_a, h_str_b;

  /* optional host destination of score-mode results (seqalign_batch_set_result_sink) */
  int32_t *sink_score = nullptr, *sink_x = nullptr, *sink_y = nullptr;
  bool last_to_sink = false;

  /* last batch */
  size_t n = 0;
  int algo = 0, mode = 0;
  std::vector<int32_t> score, xend, yend;
  std::vector<int64_t> res_off;          /* n+1, into res_a/res_b */
  std::vector<char> res_a, res_b;        /* right-aligned strings per pair (batches that took several waves) */
  const char *res_a_p = nullptr, *res_b_p = nullptr;   /* where they are: the pinned landing buffers of a one-wave batch, else res_a / res_b */
  std::vector<int32_t> aln_start, aln_len, pos_a, pos_b, len_a, len_b, status;
  /* multi-hit results */
  int32_t hit_min_score = 1, hit_max = 8;
  std::vector<int32_t> nhits, hit_rec;       /* n, n*max*8 */
  std::vector<int64_t> hit_off;              /* n+1: string offsets (per pair: max * (la+lb)) */
  std::vector<char> hit_a, hit_b;
  const char *hit_a_p = nullptr, *hit_b_p = nullptr;   /* as res_a_p / res_b_p: pinned buffers of a one-wave batch, else the vectors */
  int32_t hit_max_used = 0;

  double last_ms = 0, last_walk_ms = 0;
  int last_launches = 0;
  const char *last_kernel = "none";
};

namespace {

#define CU_TRY(call)                                                          \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if(e__ != cudaSuccess) {                                                  \
      char m__[512];                                                          \
      snprintf(m__, sizeof(m__), "CUDA error %d (%s) at %s:%d: %s", (int)e__, \
               cudaGetErrorString(e__), __FILE__, __LINE__, #call);           \
      eng->err = m__;                                                         \
      return SEQALIGN_ERR_CUDA;                                               \
    }                                                                         \
  } while(0)

int fail(seqalign_batch *eng, int code, const char *msg)
{
  eng->err = msg;
  return code;
}

int ensure_dev(seqalign_batch *eng, DevBuf &b, size_t bytes)
{
  bytes = (bytes + 255) / 256 * 256 + 256; /* slack: vector loads may over-read 16 B */
  if(b.cap >= bytes) return 0;
  b.release();
  size_t want = bytes + bytes / 4;
  cudaError_t e = cudaMalloc(&b.p, want);
  if(e != cudaSuccess) { want = bytes; e = cudaMalloc(&b.p, want); }
  if(e != cudaSuccess) {
    b.p = nullptr;
    cudaGetLastError();
    return fail(eng, SEQALIGN_ERR_NOMEM, "out of device memory");
  }
  b.cap = want;
  return 0;
}

int ensure_pin(seqalign_batch *eng, PinBuf &b, size_t bytes)
{
  bytes = (bytes + 255) / 256 * 256 + 256;
  if(b.cap >= bytes) return 0;
  b.release();
  size_t want = bytes + bytes / 4;
  if(cudaMallocHost(&b.p, want) != cudaSuccess) {
    b.p = nullptr;
    cudaGetLastError();
    return fail(eng, SEQALIGN_ERR_NOMEM, "out of pinned host memory");
  }
  b.cap = want;
  return 0;
}

#define TRY(call) do { int r__ = (call); if(r__ != 0) return r__; } while(0)

ScoreParams make_params(const scoring_t *s, int is_sw, int ncodes)
{
  ScoreParams sp;
  sp.open = (int)((unsigned)s->gap_open + (unsigned)s->gap_extend);
  sp.ext = s->gap_extend;
  sp.gap_open = s->gap_open;
  /* reference alignment.c:41 */
  sp.minv = is_sw ? 0 : (int)((unsigned)INT_MIN + (unsigned)abs(s->min_penalty));
  sp.is_sw = is_sw ? 1 : 0;
  sp.no_start = s->no_start_gap_penalty;
  sp.no_end = s->no_end_gap_penalty;
  sp.no_gaps_a = s->no_gaps_in_a;
  sp.no_gaps_b = s->no_gaps_in_b;
  sp.no_mismatches = s->no_mismatches;
  sp.ncodes = ncodes;
  return sp;
}

/* entries of a plan's profile table: rows = codes of seq_b + the padding row, columns = codes of seq_a + the padding code */
inline size_t plan_elems(int ncodes) { return (size_t)(ncodes + 1) * (ncodes + 1); }

constexpr size_t COUNTER_BYTES = 8 * (seqalign_batch::MAX_CHUNKS + 1);

struct BatchMeta {
  uint64_t pres_a[4], pres_b[4];
  int64_t max_la, max_lb, cells, max_cells, min_la, min_lb;
};

/* scan the device-resident pairs [0,n) of (off_a, off_b): alphabet, longest
 * sequences, cell count.  d_a/d_b are the buffers the offsets index into;
 * approx_bytes only sizes the grid.  scan_launch is asynchronous,
 * scan_collect waits for it and decodes the meta block. */
int scan_launch(seqalign_batch *eng, const uint8_t *d_a, const uint8_t *d_b,
                const int64_t *d_off_a, const int64_t *d_off_b, size_t n,
                int64_t approx_bytes, cudaStream_t st, int slot = 0)
{
  /* one meta block per chunk slot, so several scans can be in flight */
  TRY(ensure_dev(eng, eng->d_meta, seqalign_batch::MAX_CHUNKS * META_WORDS * 8));
  TRY(ensure_pin(eng, eng->h_meta, seqalign_batch::MAX_CHUNKS * META_WORDS * 8));
  unsigned long long *dm = (unsigned long long *)eng->d_meta.p + (size_t)slot * META_WORDS;
  CU_TRY(cudaMemsetAsync(dm, 0, META_WORDS * 8, st));
  CU_TRY(cudaMemsetAsync(dm + META_MIN_LA, 0xff, 16, st));
  int64_t work = approx_bytes / 16 + (int64_t)n;
  int grid = (int)((work + 255) / 256);
  if(grid > eng->num_sms * 8) grid = eng->num_sms * 8;
  if(grid < 1) grid = 1;
  SA_LAUNCH(scan_kernel, grid, 256, 0, st, d_a, d_b, d_off_a, d_off_b, (int64_t)n, dm);
  CU_TRY(cudaGetLastError());
  eng->last_launches++;
  CU_TRY(cudaMemcpyAsync((uint64_t *)eng->h_meta.p + (size_t)slot * META_WORDS, dm, META_WORDS * 8,
                         cudaMemcpyDeviceToHost, st));
  return 0;
}

/* decode the meta block of a finished scan */
void scan_decode(const seqalign_batch *eng, size_t n, int slot, BatchMeta *bm)
{
  const uint64_t *m = (const uint64_t *)eng->h_meta.p + (size_t)slot * META_WORDS;
  for(int i = 0; i < 4; i++) { bm->pres_a[i] = m[META_PRES_A + i]; bm->pres_b[i] = m[META_PRES_B + i]; }
  bm->max_la = (int64_t)m[META_MAX_LA];
  bm->max_lb = (int64_t)m[META_MAX_LB];
  bm->cells = (int64_t)m[META_CELLS];
  bm->max_cells = (int64_t)m[META_MAX_CELLS];
  bm->min_la = n ? (int64_t)m[META_MIN_LA] : 0;
  bm->min_lb = n ? (int64_t)m[META_MIN_LB] : 0;
}

int scan_collect(seqalign_batch *eng, size_t n, cudaStream_t st, BatchMeta *bm)
{
  CU_TRY(cudaStreamSynchronize(st));
  scan_decode(eng, n, 0, bm);
  return 0;
}

int scan_batch(seqalign_batch *eng, const uint8_t *d_a, const uint8_t *d_b,
               const int64_t *d_off_a, const int64_t *d_off_b, size_t n,
               int64_t approx_bytes, cudaStream_t st, BatchMeta *bm)
{
  TRY(scan_launch(eng, d_a, d_b, d_off_a, d_off_b, n, approx_bytes, st));
  return scan_collect(eng, n, st, bm);
}

/* flatten scoring for this batch's alphabet and upload the tables; skipped
 * when the device already holds the tables of this (scoring, alphabet) */
int upload_tables(seqalign_batch *eng, const BatchMeta &bm, cudaStream_t st)
{
  uint64_t pres[8];
  for(int i = 0; i < 4; i++) { pres[i] = bm.pres_a[i]; pres[4 + i] = bm.pres_b[i]; }
  if(eng->tables_valid && eng->tables_version == eng->scoring_version &&
     memcmp(pres, eng->tables_pres, sizeof(pres)) == 0)
    return 0;
  eng->spec.valid = false;   /* codes are about to change under any cached plan */
  flatten_scoring(eng->scoring, bm.pres_a, bm.pres_b, &eng->ft);
  const FlatTable &ft = eng->ft;
  const size_t nn = (size_t)ft.ncodes * ft.ncodes;
  TRY(ensure_dev(eng, eng->d_sub, nn * 4));
  TRY(ensure_dev(eng, eng->d_forbid, nn));
  TRY(ensure_dev(eng, eng->d_lut, 256));
  /* pageable sources: cudaMemcpyAsync returns once they are staged */
  CU_TRY(cudaMemcpyAsync(eng->d_sub.p, ft.sub.data(), nn * 4, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(eng->d_forbid.p, ft.forbid.data(), nn, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(eng->d_lut.p, ft.lut, 256, cudaMemcpyHostToDevice, st));
  eng->tables_valid = true;
  eng->tables_version = eng->scoring_version;
  memcpy(eng->tables_pres, pres, sizeof(pres));
  eng->dev_tab32.clear();
  eng->dev_tab8.clear();
  return 0;
}

/* the reference exits on the first unknown character pair it meets while
 * filling (row-major: seq_b outer, seq_a inner).  Only reached when the
 * flattened table has an unknown entry among the batch's characters. */
int check_unknown_pairs(seqalign_batch *eng, const char *seq_a, const int64_t *off_a,
                        const char *seq_b, const int64_t *off_b, size_t n)
{
  const FlatTable &ft = eng->ft;
  const int nc = ft.ncodes;
  for(size_t i = 0; i < n; i++) {
    const int64_t la = off_a[i + 1] - off_a[i], lb = off_b[i + 1] - off_b[i];
    if(la == 0 || lb == 0) continue;
    const unsigned char *a = (const unsigned char *)seq_a + off_a[i];
    const unsigned char *b = (const unsigned char *)seq_b + off_b[i];
    std::vector<uint8_t> seen_a(nc, 0);
    for(int64_t x = 0; x < la; x++) seen_a[ft.lut[a[x]]] = 1;
    for(int64_t y = 0; y < lb; y++) {
      const int cb = ft.lut[b[y]];
      bool bad = false;
      for(int ca = 0; ca < nc && !bad; ca++) bad = seen_a[ca] && ft.unknown[(size_t)cb * nc + ca];
      if(!bad) continue;
      for(int64_t x = 0; x < la; x++) {
        const int ca = ft.lut[a[x]];
        if(ft.unknown[(size_t)cb * nc + ca]) {
          eng->unk_a = (char)ft.rep[ca];
          eng->unk_b = (char)ft.rep[cb];
          char m[160];
          snprintf(m, sizeof(m), "Error: Unknown character pair (%c,%c) and match/mismatch have not been set",
                   eng->unk_a, eng->unk_b);
          eng->err = m;
          return SEQALIGN_ERR_UNKNOWN_PAIR;
        }
      }
    }
  }
  return 0;
}

int general_grid(const seqalign_batch *eng, size_t npairs)
{
  int64_t g = (int64_t)eng->num_sms * 4;
  int64_t need = ((int64_t)npairs + GEN_WARPS - 1) / GEN_WARPS;
  if(g > need) g = need;
  return g < 1 ? 1 : (int)g;
}

struct DevBatch {
  const uint8_t *a, *b;
  const int64_t *off_a, *off_b;
  size_t n;
};

template <int MODE>
int launch_general(seqalign_batch *eng, const DevBatch &db, const ScoreParams &sp,
                   int64_t pair0, int64_t npairs, const BatchMeta &bm, GenArgs extra,
                   cudaStream_t st)
{
  GenArgs A = extra;
  A.seq_a = db.a; A.seq_b = db.b; A.off_a = db.off_a; A.off_b = db.off_b;
  A.pair0 = pair0; A.npairs = npairs; A.sp = sp;
  A.sub = (const int32_t *)eng->d_sub.p;
  A.forbid = (const uint8_t *)eng->d_forbid.p;
  A.lut = (const uint8_t *)eng->d_lut.p;
  A.table_in_smem = sp.ncodes <= SMEM_TABLE_MAX_CODES;
  /* wide pairs (>= 4 strips): a whole CTA per pair, warps pipelined over the
   * column strips; otherwise a warp per pair */
  const char *cenv = getenv("SEQALIGN_COOP");
  const bool coop = cenv ? atoi(cenv) != 0 : bm.max_la > 3 * GSTRIP;
  int grid;
  if(coop) {
    int64_t g = (int64_t)eng->num_sms * 2;
    if(g > npairs) g = npairs;
    grid = g < 1 ? 1 : (int)g;
  } else {
    grid = general_grid(eng, (size_t)npairs);
  }
  A.bnd = nullptr; A.bnd_rows = 0;
  if(bm.max_la > GSTRIP) {
    A.bnd_rows = bm.max_lb + 1;
    const size_t slots = (size_t)grid * (coop ? COOP_WARPS : GEN_WARPS);
    TRY(ensure_dev(eng, eng->d_bnd, slots * (size_t)A.bnd_rows * sizeof(int4)));
    A.bnd = (int4 *)eng->d_bnd.p;
  }
  TRY(ensure_dev(eng, eng->d_counter, COUNTER_BYTES));
  CU_TRY(cudaMemsetAsync(eng->d_counter.p, 0, 8, st));
  A.counter = (unsigned long long *)eng->d_counter.p;
  if(coop) {
    const size_t smem = general_smem_bytes(sp.ncodes, A.table_in_smem, COOP_WARPS);
    SA_LAUNCH(general_coop_kernel<MODE>, grid, COOP_WARPS * 32, smem, st, A);
  } else {
    const size_t smem = general_smem_bytes(sp.ncodes, A.table_in_smem);
    SA_LAUNCH(general_kernel<MODE>, grid, GEN_WARPS * 32, smem, st, A);
  }
  CU_TRY(cudaGetLastError());
  eng->last_launches++;
  return 0;
}

/* score mode over a device-resident batch; results into d_score/d_xend/d_yend */
/* launch the specialised score kernel of `plan` (tables must be on the device) */
int launch_fast_score(seqalign_batch *eng, const FastPlan &plan, const ScoreParams &sp, const DevBatch &db,
                      int64_t max_lb, int32_t *d_score, int32_t *d_xend, int32_t *d_yend, cudaStream_t st,
                      cudaEvent_t ev0, cudaEvent_t ev1, int slot = 0, const int *d_order = nullptr, int64_t order_count = 0,
                      const int *d_range = nullptr)
{
  const size_t nn = plan_elems(sp.ncodes);
  int8_t *d_t8 = (int8_t *)eng->d_tab8.p;
  int32_t *d_t32 = (int32_t *)(d_t8 + ((nn + 15) & ~(size_t)15));
  /* one work-queue counter per slot: launches of different slots may run on different streams */
  TRY(ensure_dev(eng, eng->d_counter, COUNTER_BYTES));
  unsigned long long *d_cnt = (unsigned long long *)eng->d_counter.p + slot;
  CU_TRY(cudaMemsetAsync(d_cnt, 0, 8, st));
  FastArgs F;
  memset(&F, 0, sizeof(F));
  F.seq_a = db.a; F.seq_b = db.b; F.off_a = db.off_a; F.off_b = db.off_b;
  F.npairs = (int64_t)db.n; F.sp = sp;
  if(d_order) { F.order = d_order; F.npairs = order_count; F.range = d_range; }   /* a length bucket: pair indices through `order` */
  F.tab8 = d_t8;
  F.tab32 = d_t32;
  F.lut = (const uint8_t *)eng->d_lut.p;
  F.counter = d_cnt;
  F.score = d_score; F.xend = d_xend; F.yend = d_yend;
  F.max_lb = (int)max_lb;
  if(ev0) CU_TRY(cudaEventRecord(ev0, st));
  int r = fast_launch(plan, F, eng->num_sms, eng->smem_optin, st);
  if(r != 0) return fail(eng, SEQALIGN_ERR_CUDA, "fast kernel launch failed");
  CU_TRY(cudaGetLastError());
  if(ev1) CU_TRY(cudaEventRecord(ev1, st));
  eng->last_launches++;
  eng->last_kernel = plan.name;
  return 0;
}

/* upload the profile tables of a specialised plan unless the device holds them */
int upload_plan_tables(seqalign_batch *eng, const std::vector<int8_t> &tab8, const std::vector<int32_t> &tab32,
                       int8_t **d_t8_out, int32_t **d_t32_out, cudaStream_t st)
{
  const size_t nn = plan_elems(eng->ft.ncodes);
  TRY(ensure_dev(eng, eng->d_tab8, nn * 5 + 64));
  int8_t *d_t8 = (int8_t *)eng->d_tab8.p;
  int32_t *d_t32 = (int32_t *)(d_t8 + ((nn + 15) & ~(size_t)15));
  if(tab32 != eng->dev_tab32 || tab8 != eng->dev_tab8) {
    CU_TRY(cudaMemcpyAsync(d_t8, tab8.data(), nn, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_t32, tab32.data(), nn * 4, cudaMemcpyHostToDevice, st));
    eng->dev_tab32 = tab32;
    eng->dev_tab8 = tab8;
  }
  *d_t8_out = d_t8; *d_t32_out = d_t32;
  return 0;
}

/* wide-pair kernel over the device-resident pairs [c0, c0+m) of db; results
 * into d_score/d_xend/d_yend (absolute pair index); dir/dir_off (relative to
 * c0) only for plans with traceback flags */
int launch_long(seqalign_batch *eng, const LongPlan &plan, const ScoreParams &sp, const DevBatch &db,
                size_t c0, size_t m, int64_t max_lb, int32_t *d_score, int32_t *d_xend, int32_t *d_yend,
                uint8_t *d_dir, const int64_t *d_dir_off, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1)
{
  int8_t *d_t8 = nullptr;
  int32_t *d_t32 = nullptr;
  TRY(upload_plan_tables(eng, plan.tab8, plan.tab32, &d_t8, &d_t32, st));
  const int grid = long_grid(plan, eng->num_sms, (int64_t)m);
  LongArgs L;
  memset(&L, 0, sizeof(L));
  L.bnd_rows = (max_lb + 1 + 3) & ~(int64_t)3;
  TRY(ensure_dev(eng, eng->d_lbnd, (size_t)grid * 2 * LONG_WARPS * (size_t)L.bnd_rows * sizeof(int2)));
  L.bnd = (int2 *)eng->d_lbnd.p;
  L.seq_a = db.a; L.seq_b = db.b; L.off_a = db.off_a + c0; L.off_b = db.off_b + c0;
  L.npairs = (int64_t)m; L.sp = sp;
  L.tab8 = d_t8; L.tab32 = d_t32;
  L.lut = (const uint8_t *)eng->d_lut.p;
  L.dir = d_dir; L.dir_off = d_dir_off;
  L.score = d_score + c0;
  L.xend = d_xend ? d_xend + c0 : nullptr;
  L.yend = d_yend ? d_yend + c0 : nullptr;
  if(plan.is_sw) {
    TRY(ensure_dev(eng, eng->d_swkey, m * 8));
    L.swkey = (unsigned long long *)eng->d_swkey.p;
  }
  CU_TRY(cudaEventRecord(ev0, st));
  if(plan.is_sw) CU_TRY(cudaMemsetAsync(L.swkey, 0, m * 8, st));
  if(long_launch(plan, L, grid, st) != 0) return fail(eng, SEQALIGN_ERR_CUDA, "wide-pair kernel launch failed");
  if(plan.is_sw) {
    SA_LAUNCH(long_sw_finish_kernel, (unsigned)((m + 255) / 256), 256, 0, st, (const unsigned long long *)L.swkey, (int64_t)m,
              L.score, L.xend, L.yend);
    eng->last_launches++;
  }
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaEventRecord(ev1, st));
  eng->last_launches++;
  eng->last_kernel = plan.name;
  return 0;
}

int run_score(seqalign_batch *eng, int algo, const DevBatch &db, const BatchMeta &bm,
              int32_t *d_score, int32_t *d_xend, int32_t *d_yend, cudaStream_t st,
              cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr, FastPlan *plan_out = nullptr)
{
  if(!ev0) { ev0 = eng->ev0; ev1 = eng->ev1; }
  if(plan_out) plan_out->G = 0;
  const ScoreParams sp = make_params(eng->scoring, algo == SEQALIGN_SW, eng->ft.ncodes);
  FastPlan plan;
  const bool want_ends = (d_xend != nullptr || d_yend != nullptr) && eng->force_mode != 3 && eng->force_mode != 4;
  /* force modes 2, 4 and 5 keep to the int32 kernels */
  const bool allow16 = eng->force_mode != 4 && eng->force_mode != 2 && eng->force_mode != 5;
  const bool uniform = bm.min_la == bm.max_la && bm.min_lb == bm.max_lb;
  if(eng->force_mode != 1 && fast_plan(eng->scoring, eng->ft, sp, bm.max_la, bm.max_lb, want_ends, allow16, &plan, false, uniform)) {
    if(eng->force_mode == 2 && plan.track == TRACK_TREE) { plan.track = TRACK_COLUMN; plan.name = "fast_sw_score_endcol"; }
    const size_t nn = plan_elems(eng->ft.ncodes);
    TRY(ensure_dev(eng, eng->d_tab8, nn * 5 + 64));
    TRY(ensure_dev(eng, eng->d_counter, COUNTER_BYTES));
    int8_t *d_t8 = (int8_t *)eng->d_tab8.p;
    int32_t *d_t32 = (int32_t *)(d_t8 + ((nn + 15) & ~(size_t)15));
    if(plan.tab32 != eng->dev_tab32 || plan.tab8 != eng->dev_tab8) {
      CU_TRY(cudaMemcpyAsync(d_t8, plan.tab8.data(), nn, cudaMemcpyHostToDevice, st));
      CU_TRY(cudaMemcpyAsync(d_t32, plan.tab32.data(), nn * 4, cudaMemcpyHostToDevice, st));
      eng->dev_tab32 = plan.tab32;
      eng->dev_tab8 = plan.tab8;
    }
    /* reads of different lengths: one launch per shape class over length-sorted pairs (sa_fast.cuh "length
     * buckets") instead of one launch shaped by the longest pair */
    /* SEQALIGN_BUCKETS: "shapes" = a launch per shape class over length-sorted pairs, "rows" = one launch over
     * pairs sorted by len_b (couples and warps of equal height), unset / "0" = input order (the default: see
     * DESIGN.md 3, round-2 notes, for the measurements behind it) */
    const char *bmode = getenv("SEQALIGN_BUCKETS");
    const bool by_shape = bmode && strcmp(bmode, "shapes") == 0, by_rows = bmode && strcmp(bmode, "rows") == 0;
    const char *bmin = getenv("SEQALIGN_BUCKET_MIN");   /* smallest batch that is bucketed (fuzzing lowers it) */
    const size_t bucket_min = bmin && atoll(bmin) > 0 ? (size_t)atoll(bmin) : 4096;
    if(plan.s16 && !uniform && db.n >= bucket_min && db.n < ((size_t)1 << 31) && (by_shape || by_rows)) {
      BucketArgs B;
      memset(&B, 0, sizeof(B));
      for(const FastShape &sh : kFast16Shapes) {
        if(B.nclasses == BUCKET_MAX_CLASSES) break;
        const int w = sh.G * sh.K;
        if(by_rows && w < bm.max_la) continue;   /* one class: the batch's own shape */
        if(B.nclasses && w <= B.width[B.nclasses - 1]) continue;
        B.width[B.nclasses++] = w;
        if(w >= bm.max_la) break;
      }
      while((bm.max_lb >> B.shift) >= BUCKET_LB_BINS) B.shift++;
      const size_t nbins = (size_t)B.nclasses * BUCKET_LB_BINS;
      TRY(ensure_dev(eng, eng->d_bucket, (2 * nbins + BUCKET_MAX_CLASSES + 1) * 4 + db.n * 4));
      B.bins = (int *)eng->d_bucket.p; B.cursor = B.bins + nbins; B.class_start = B.cursor + nbins;
      B.order = B.class_start + BUCKET_MAX_CLASSES + 1;
      B.off_a = db.off_a; B.off_b = db.off_b; B.npairs = (int64_t)db.n;
      CU_TRY(cudaEventRecord(ev0, st));
      CU_TRY(cudaMemsetAsync(B.bins, 0, nbins * 4, st));
      const unsigned bgrid = (unsigned)((db.n + 255) / 256);
      SA_LAUNCH(bucket_hist_kernel, bgrid, 256, 0, st, B);
      SA_LAUNCH(bucket_scan_kernel, 1, 256, 0, st, B);
      SA_LAUNCH(bucket_scatter_kernel, bgrid, 256, 0, st, B);
      CU_TRY(cudaGetLastError());
      eng->last_launches += 3;
      /* one launch per shape class, enqueued at once on side streams: the kernels read their ranges from
       * class_start on the device (an empty class costs one launch that exits), their tails overlap */
      if(!eng->bucket_streams[0])
        for(int k = 0; k < seqalign_batch::BUCKET_STREAMS; k++) {
          CU_TRY(cudaStreamCreateWithFlags(&eng->bucket_streams[k], cudaStreamNonBlocking));
          CU_TRY(cudaEventCreateWithFlags(&eng->bucket_done[k], cudaEventDisableTiming));
        }
      if(!eng->bucket_ready) CU_TRY(cudaEventCreateWithFlags(&eng->bucket_ready, cudaEventDisableTiming));
      CU_TRY(cudaEventRecord(eng->bucket_ready, st));
      for(int k = 0; k < seqalign_batch::BUCKET_STREAMS; k++) CU_TRY(cudaStreamWaitEvent(eng->bucket_streams[k], eng->bucket_ready, 0));
      for(int c = B.nclasses - 1; c >= 0; c--) {   /* widest class first */
        FastPlan cplan;
        const int64_t cla = B.width[c] < bm.max_la ? B.width[c] : bm.max_la;
        if(!fast_plan(eng->scoring, eng->ft, sp, cla, bm.max_lb, want_ends, true, &cplan, false, false) || !cplan.s16 ||
           cplan.s16_ends != plan.s16_ends)
          cplan = plan;   /* the batch's own plan holds every pair */
        TRY(launch_fast_score(eng, cplan, sp, db, bm.max_lb, d_score, d_xend, d_yend,
                              eng->bucket_streams[c % seqalign_batch::BUCKET_STREAMS], nullptr, nullptr, 1 + c, B.order,
                              (int64_t)db.n, B.class_start + c));
      }
      for(int k = 0; k < seqalign_batch::BUCKET_STREAMS; k++) {
        CU_TRY(cudaEventRecord(eng->bucket_done[k], eng->bucket_streams[k]));
        CU_TRY(cudaStreamWaitEvent(st, eng->bucket_done[k], 0));
      }
      CU_TRY(cudaEventRecord(ev1, st));
      eng->last_kernel = plan.name;
      if(plan_out) *plan_out = plan;
      return 0;
    }
    TRY(launch_fast_score(eng, plan, sp, db, bm.max_lb, d_score, d_xend, d_yend, st, ev0, ev1));
    if(plan_out) *plan_out = plan;
    return 0;
  }
  LongPlan lplan;
  if(eng->force_mode != 1 && long_plan(eng->scoring, eng->ft, sp, bm.max_la, bm.max_lb, false, &lplan))
    return launch_long(eng, lplan, sp, db, 0, db.n, bm.max_lb, d_score, d_xend, d_yend, nullptr, nullptr, st, ev0, ev1);
  GenArgs X;
  memset(&X, 0, sizeof(X));
  X.score = d_score; X.xend = d_xend; X.yend = d_yend; X.state = nullptr;
  CU_TRY(cudaEventRecord(ev0, st));
  TRY(launch_general<MODE_SCORE>(eng, db, sp, 0, (int64_t)db.n, bm, X, st));
  CU_TRY(cudaEventRecord(ev1, st));
  eng->last_kernel = "general_score";
  return 0;
}

int64_t dir_bytes(int64_t la, int64_t lb)
{
  return (dir_stride((int)la) * lb + 15) & ~(int64_t)15;
}

/* align mode: waves of pairs bounded by direction-byte memory */
int run_align(seqalign_batch *eng, int algo, const DevBatch &db, const BatchMeta &bm,
              const int64_t *h_off_a, const int64_t *h_off_b, cudaStream_t st)
{
  const size_t n = db.n;
  const ScoreParams sp = make_params(eng->scoring, algo == SEQALIGN_SW, eng->ft.ncodes);
  TRY(ensure_dev(eng, eng->d_score, n * 4));
  TRY(ensure_dev(eng, eng->d_xend, n * 4));
  TRY(ensure_dev(eng, eng->d_yend, n * 4));
  TRY(ensure_dev(eng, eng->d_state, n * 4));
  int32_t *d_score = (int32_t *)eng->d_score.p, *d_xend = (int32_t *)eng->d_xend.p;
  int32_t *d_yend = (int32_t *)eng->d_yend.p, *d_state = (int32_t *)eng->d_state.p;

  eng->last_ms = 0;
  eng->last_walk_ms = 0;
  float ms = 0;
  /* specialised fill (score + end cell + traceback flags in one pass) when the
   * scoring shape allows, else the general kernel (SW: best cell first, its
   * direction pass stores no scores) */
  FastPlan dplan;
  const bool fast_dir = eng->force_mode != 1 &&
                        fast_plan(eng->scoring, eng->ft, sp, bm.max_la, bm.max_lb, true, false, &dplan, true);
  if(fast_dir && eng->force_mode == 2 && dplan.track == TRACK_TREE) dplan.track = TRACK_COLUMN;
  /* wide pairs / free end gaps (NW): the strip-pipelined kernel, same flag bytes */
  LongPlan lplan;
  /* wide pairs trace back through checkpoints and recomputed tiles (0.14 B/cell, fill at score speed);
   * SEQALIGN_LONG_FLAGS=1 keeps the flag bytes (1 B/cell; NW only: the flag-byte fill does not track SW's best cell) */
  const char *lf_env = getenv("SEQALIGN_LONG_FLAGS");
  const bool want_ckpt = !(lf_env && lf_env[0] == '1');
  const bool long_dir = !fast_dir && eng->force_mode != 1 &&
                        long_plan(eng->scoring, eng->ft, sp, bm.max_la, bm.max_lb, true, &lplan) && (!lplan.is_sw || want_ckpt);
  const bool long_ckpt = long_dir && want_ckpt;
  if(long_ckpt) { lplan.ckpt = true; lplan.dir = false; lplan.name = lplan.is_sw ? "long_sw_ckpt" : "long_nw_ckpt"; }
  if(algo == SEQALIGN_SW && !fast_dir && !long_dir) {
    TRY(run_score(eng, algo, db, bm, d_score, d_xend, d_yend, st));
    CU_TRY(cudaStreamSynchronize(st));
    CU_TRY(cudaEventElapsedTime(&ms, eng->ev0, eng->ev1));
    eng->last_ms += ms;
  }
  int8_t *d_t8 = nullptr;
  int32_t *d_t32 = nullptr;
  if(fast_dir) {
    const size_t nn = plan_elems(eng->ft.ncodes);
    TRY(ensure_dev(eng, eng->d_tab8, nn * 5 + 64));
    d_t8 = (int8_t *)eng->d_tab8.p;
    d_t32 = (int32_t *)(d_t8 + ((nn + 15) & ~(size_t)15));
    if(dplan.tab32 != eng->dev_tab32 || dplan.tab8 != eng->dev_tab8) {
      CU_TRY(cudaMemcpyAsync(d_t8, dplan.tab8.data(), nn, cudaMemcpyHostToDevice, st));
      CU_TRY(cudaMemcpyAsync(d_t32, dplan.tab32.data(), nn * 4, cudaMemcpyHostToDevice, st));
      eng->dev_tab32 = dplan.tab32;
      eng->dev_tab8 = dplan.tab8;
    }
  }

  size_t free_b = 0, total_b = 0;
  CU_TRY(cudaMemGetInfo(&free_b, &total_b));
  int64_t budget = (int64_t)(free_b / 2);
  const char *env = getenv("SEQALIGN_DIR_BUDGET");
  if(env) budget = atoll(env);
  if(budget > ((int64_t)64 << 30)) budget = (int64_t)64 << 30;

  eng->res_off.assign(n + 1, 0);
  for(size_t i = 0; i < n; i++)
    eng->res_off[i + 1] = eng->res_off[i] + (h_off_a[i + 1] - h_off_a[i]) + (h_off_b[i + 1] - h_off_b[i]);
  eng->aln_start.resize(n); eng->aln_len.resize(n); eng->pos_a.resize(n); eng->pos_b.resize(n);
  eng->len_a.resize(n); eng->len_b.resize(n); eng->status.resize(n);

  std::vector<int64_t> dir_off, out_off;
  size_t c0 = 0;
  while(c0 < n) {
    /* wave [c0, c1) */
    size_t c1 = c0;
    int64_t dbytes = 0;
    dir_off.clear(); out_off.clear();
    while(c1 < n) {
      const int64_t la = h_off_a[c1 + 1] - h_off_a[c1], lb = h_off_b[c1 + 1] - h_off_b[c1];
      const int64_t need = long_ckpt ? long_trace_bytes((int)la, (int)lb) : dir_bytes(la, lb);
      if(c1 > c0 && dbytes + need > budget) break;
      dir_off.push_back(dbytes);
      out_off.push_back(eng->res_off[c1] - eng->res_off[c0]);
      dbytes += need;
      c1++;
    }
    if(long_dir && c1 < n) {
      /* the strip pipeline streams pairs through its resident CTAs: a wave of
       * a whole number of pairs per CTA leaves no CTA idle at the end */
      const size_t g = (size_t)long_grid(lplan, eng->num_sms, (int64_t)n);
      if(c1 - c0 > g) {
        c1 = c0 + (c1 - c0) / g * g;
        dir_off.resize(c1 - c0); out_off.resize(c1 - c0);
        const int64_t la = h_off_a[c1] - h_off_a[c1 - 1], lb = h_off_b[c1] - h_off_b[c1 - 1];
        dbytes = dir_off.back() + (long_ckpt ? long_trace_bytes((int)la, (int)lb) : dir_bytes(la, lb));
      }
    }
    const size_t m = c1 - c0;
    const int64_t obytes = eng->res_off[c1] - eng->res_off[c0];
    TRY(ensure_dev(eng, eng->d_dir, (size_t)dbytes + 16));
    TRY(ensure_dev(eng, eng->d_dir_off, m * 8));
    TRY(ensure_dev(eng, eng->d_out_off, m * 8));
    TRY(ensure_dev(eng, eng->d_out_a, (size_t)obytes + 16));
    TRY(ensure_dev(eng, eng->d_out_b, (size_t)obytes + 16));
    /* the walks write each pair's strings right-aligned: the part in front stays untouched, and the copy
     * to the host takes the whole block (initcheck: no uninitialised byte may cross PCIe) */
    CU_TRY(cudaMemsetAsync(eng->d_out_a.p, 0, (size_t)obytes, st));
    CU_TRY(cudaMemsetAsync(eng->d_out_b.p, 0, (size_t)obytes, st));
    TRY(ensure_dev(eng, eng->d_walk, m * 4 * 7));
    TRY(ensure_pin(eng, eng->h_walk, m * 4 * 7));
    TRY(ensure_pin(eng, eng->h_str_a, (size_t)obytes + 16));
    TRY(ensure_pin(eng, eng->h_str_b, (size_t)obytes + 16));
    CU_TRY(cudaMemcpyAsync(eng->d_dir_off.p, dir_off.data(), m * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(eng->d_out_off.p, out_off.data(), m * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));

    if(fast_dir) {
      TRY(ensure_dev(eng, eng->d_counter, COUNTER_BYTES));
      CU_TRY(cudaMemsetAsync(eng->d_counter.p, 0, 8, st));
      FastArgs F;
      memset(&F, 0, sizeof(F));
      F.seq_a = db.a; F.seq_b = db.b; F.off_a = db.off_a + c0; F.off_b = db.off_b + c0;
      F.npairs = (int64_t)m; F.sp = sp;
      F.tab8 = d_t8; F.tab32 = d_t32;
      F.lut = (const uint8_t *)eng->d_lut.p;
      F.counter = (unsigned long long *)eng->d_counter.p;
      F.score = d_score + c0; F.xend = d_xend + c0; F.yend = d_yend + c0;
      F.max_lb = (int)bm.max_lb;
      F.dir = (uint8_t *)eng->d_dir.p;
      F.dir_off = (const int64_t *)eng->d_dir_off.p;
      CU_TRY(cudaEventRecord(eng->ev0, st));
      if(fast_launch(dplan, F, eng->num_sms, eng->smem_optin, st) != 0)
        return fail(eng, SEQALIGN_ERR_CUDA, "fast dir kernel launch failed");
      CU_TRY(cudaGetLastError());
      CU_TRY(cudaEventRecord(eng->ev1, st));
      eng->last_launches++;
    } else if(long_dir) {
      TRY(launch_long(eng, lplan, sp, db, c0, m, bm.max_lb, d_score, d_xend, d_yend, (uint8_t *)eng->d_dir.p,
                      (const int64_t *)eng->d_dir_off.p, st, eng->ev0, eng->ev1));
    } else {
      GenArgs X;
      memset(&X, 0, sizeof(X));
      X.score = d_score; X.xend = d_xend; X.yend = d_yend; X.state = d_state;
      X.dir = (uint8_t *)eng->d_dir.p;
      X.dir_off = (const int64_t *)eng->d_dir_off.p;
      CU_TRY(cudaEventRecord(eng->ev0, st));
      TRY(launch_general<MODE_DIR>(eng, db, sp, (int64_t)c0, (int64_t)m, bm, X, st));
      CU_TRY(cudaEventRecord(eng->ev1, st));
    }

    WalkArgs W;
    memset(&W, 0, sizeof(W));
    W.seq_a = db.a; W.seq_b = db.b; W.off_a = db.off_a; W.off_b = db.off_b;
    W.pair0 = (int64_t)c0; W.npairs = (int64_t)m; W.sp = sp;
    W.sub = (const int32_t *)eng->d_sub.p; W.lut = (const uint8_t *)eng->d_lut.p;
    W.score = d_score; W.xend = d_xend; W.yend = d_yend; W.state = d_state;
    W.dir = (const uint8_t *)eng->d_dir.p; W.dir_off = (const int64_t *)eng->d_dir_off.p;
    W.out_a = (uint8_t *)eng->d_out_a.p; W.out_b = (uint8_t *)eng->d_out_b.p;
    W.out_off = (const int64_t *)eng->d_out_off.p;
    int32_t *wk = (int32_t *)eng->d_walk.p;
    W.aln_start = wk; W.aln_len = wk + m; W.pos_a = wk + 2 * m; W.pos_b = wk + 3 * m;
    W.len_a = wk + 4 * m; W.len_b = wk + 5 * m; W.status = wk + 6 * m;
    W.fmt = (fast_dir || long_dir) ? 1 : 0;
    /* long walks: a warp per pair with a shared-memory window over the flag
     * bytes; many short ones: a thread per pair */
    const char *wenv = getenv("SEQALIGN_WALK");
    const bool tiled = wenv ? strcmp(wenv, "tiled") == 0 : bm.cells / (int64_t)n >= (1 << 20);
    if(long_ckpt) {
      int8_t *w_t8 = nullptr;
      int32_t *w_t32 = nullptr;
      TRY(upload_plan_tables(eng, lplan.tab8, lplan.tab32, &w_t8, &w_t32, st));
      if(walk_ckpt_launch(lplan, W, w_t8, w_t32,
                          walk_ckpt_grid(eng->num_sms, (int64_t)m, walk_ckpt_smem_bytes(sp.ncodes, lplan.prof32)), st) != 0)
        return fail(eng, SEQALIGN_ERR_CUDA, "recompute walk launch failed");
    } else if(tiled) {
      int wgrid = (int)((m + WT_WARPS - 1) / WT_WARPS);
      if(wgrid > eng->num_sms * 16) wgrid = eng->num_sms * 16;
      SA_LAUNCH(walk_tiled_kernel, wgrid, WT_WARPS * 32, 0, st, W);
    } else {
      int wgrid = (int)((m + 127) / 128);
      if(wgrid > eng->num_sms * 8) wgrid = eng->num_sms * 8;
      SA_LAUNCH(walk_kernel, wgrid, 128, 0, st, W);
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(eng->ev_k1[0], st));
    eng->last_launches++;

    CU_TRY(cudaMemcpyAsync(eng->h_walk.p, eng->d_walk.p, m * 4 * 7, cudaMemcpyDeviceToHost, st));
    if(obytes > 0) {
      CU_TRY(cudaMemcpyAsync(eng->h_str_a.p, eng->d_out_a.p, (size_t)obytes, cudaMemcpyDeviceToHost, st));
      CU_TRY(cudaMemcpyAsync(eng->h_str_b.p, eng->d_out_b.p, (size_t)obytes, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(cudaStreamSynchronize(st));
    CU_TRY(cudaEventElapsedTime(&ms, eng->ev0, eng->ev1));
    eng->last_ms += ms;
    CU_TRY(cudaEventElapsedTime(&ms, eng->ev1, eng->ev_k1[0]));
    eng->last_walk_ms += ms;

    const int32_t *hw = (const int32_t *)eng->h_walk.p;
    memcpy(&eng->aln_start[c0], hw, m * 4);
    memcpy(&eng->aln_len[c0], hw + m, m * 4);
    memcpy(&eng->pos_a[c0], hw + 2 * m, m * 4);
    memcpy(&eng->pos_b[c0], hw + 3 * m, m * 4);
    memcpy(&eng->len_a[c0], hw + 4 * m, m * 4);
    memcpy(&eng->len_b[c0], hw + 5 * m, m * 4);
    memcpy(&eng->status[c0], hw + 6 * m, m * 4);
    if(c0 == 0 && c1 == n) {
      /* the whole batch in one wave (the usual case): the strings are read where the copy engine left them --
       * a second pass over 0.6 GB per 2 M read pairs was a quarter of the tools' align phase */
      eng->res_a_p = (const char *)eng->h_str_a.p;
      eng->res_b_p = (const char *)eng->h_str_b.p;
    } else {
      if(c0 == 0) {
        eng->res_a.resize((size_t)eng->res_off[n] + 1);
        eng->res_b.resize((size_t)eng->res_off[n] + 1);
      }
      if(obytes > 0) {
        memcpy(&eng->res_a[(size_t)eng->res_off[c0]], eng->h_str_a.p, (size_t)obytes);
        memcpy(&eng->res_b[(size_t)eng->res_off[c0]], eng->h_str_b.p, (size_t)obytes);
      }
      eng->res_a_p = eng->res_a.data();
      eng->res_b_p = eng->res_b.data();
    }
    c0 = c1;
  }
  eng->last_kernel = fast_dir ? (algo == SEQALIGN_SW ? "fast_sw_dir+walk" : "fast_nw_dir+walk")
                     : long_ckpt ? (algo == SEQALIGN_SW ? "long_sw_ckpt+walk_recompute" : "long_nw_ckpt+walk_recompute")
                     : long_dir ? "long_nw_dir+walk"
                                : (algo == SEQALIGN_SW ? "sw_score+general_dir+walk" : "general_dir+walk");

  /* scores to host */
  eng->score.resize(n); eng->xend.resize(n); eng->yend.resize(n);
  TRY(ensure_pin(eng, eng->h_res, n * 12));
  int32_t *hr = (int32_t *)eng->h_res.p;
  CU_TRY(cudaMemcpyAsync(hr, d_score, n * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(hr + n, d_xend, n * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(hr + 2 * n, d_yend, n * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  memcpy(eng->score.data(), hr, n * 4);
  memcpy(eng->xend.data(), hr + n, n * 4);
  memcpy(eng->yend.data(), hr + 2 * n, n * 4);

  for(size_t i = 0; i < n; i++)
    if(eng->status[i] == WALK_FAIL) {
      char msg[128];
      snprintf(msg, sizeof(msg), "Program error: traceback fail (get_reverse_move), pair %zu", i);
      eng->err = msg;
      return SEQALIGN_ERR_TRACEBACK;
    }
  return 0;
}

/* multi-hit mode (SW): fill with flags + int16 match scores, candidate sort,
 * masked walks, all on the device; waves bounded by memory */
int run_hits(seqalign_batch *eng, const DevBatch &db, const BatchMeta &bm,
             const int64_t *h_off_a, const int64_t *h_off_b, cudaStream_t st)
{
  const size_t n = db.n;
  const ScoreParams sp = make_params(eng->scoring, 1, eng->ft.ncodes);
  FastPlan plan;
  if(eng->force_mode == 1 ||
     !fast_plan(eng->scoring, eng->ft, sp, bm.max_la, bm.max_lb, true, false, &plan, true) || bm.max_lb > 65535)
    return fail(eng, SEQALIGN_ERR_ARG,
                "device multi-hit needs the specialised kernel (affine gaps with gap_open <= 0, no gap/mismatch "
                "restrictions, len_a <= 512, small scores); use smith_waterman_fetch() for this input");
  plan.hits = true;
  plan.track = TRACK_NONE;
  const int maxh = eng->hit_max;
  eng->hit_max_used = maxh;
  const size_t nn = plan_elems(eng->ft.ncodes);
  TRY(ensure_dev(eng, eng->d_tab8, nn * 5 + 64));
  int8_t *d_t8 = (int8_t *)eng->d_tab8.p;
  int32_t *d_t32 = (int32_t *)(d_t8 + ((nn + 15) & ~(size_t)15));
  if(plan.tab32 != eng->dev_tab32 || plan.tab8 != eng->dev_tab8) {
    CU_TRY(cudaMemcpyAsync(d_t8, plan.tab8.data(), nn, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_t32, plan.tab32.data(), nn * 4, cudaMemcpyHostToDevice, st));
    eng->dev_tab32 = plan.tab32;
    eng->dev_tab8 = plan.tab8;
  }
  TRY(ensure_dev(eng, eng->d_score, n * 4));
  TRY(ensure_dev(eng, eng->d_counter, COUNTER_BYTES));

  eng->nhits.assign(n, 0);
  eng->hit_rec.assign(n * (size_t)maxh * 8, 0);
  eng->hit_off.assign(n + 1, 0);
  for(size_t i = 0; i < n; i++)
    eng->hit_off[i + 1] = eng->hit_off[i] + (int64_t)maxh * ((h_off_a[i + 1] - h_off_a[i]) + (h_off_b[i + 1] - h_off_b[i]));

  size_t free_b = 0, total_b = 0;
  CU_TRY(cudaMemGetInfo(&free_b, &total_b));
  int64_t budget = (int64_t)(free_b / 2);
  const char *env = getenv("SEQALIGN_DIR_BUDGET");
  if(env) budget = atoll(env);
  eng->last_ms = 0;

  std::vector<int64_t> dir_off, out_off;
  size_t c0 = 0;
  while(c0 < n) {
    size_t c1 = c0;
    int64_t cells = 0;
    dir_off.clear(); out_off.clear();
    while(c1 < n) {
      const int64_t la = h_off_a[c1 + 1] - h_off_a[c1], lb = h_off_b[c1 + 1] - h_off_b[c1];
      const int64_t need = (dir_stride((int)la) * lb + 31) & ~(int64_t)31;   /* whole mask words per pair */
      /* per cell: flags 1 + scores 2 + two key buffers 16 + mask 1/8 */
      if(c1 > c0 && (cells + need) * 20 > budget) break;
      dir_off.push_back(cells);
      out_off.push_back(eng->hit_off[c1] - eng->hit_off[c0]);
      cells += need;
      c1++;
    }
    const size_t m = c1 - c0;
    const int64_t obytes = eng->hit_off[c1] - eng->hit_off[c0];
    TRY(ensure_dev(eng, eng->d_dir, (size_t)cells + 64));
    TRY(ensure_dev(eng, eng->d_m16, (size_t)cells * 2 + 64));
    TRY(ensure_dev(eng, eng->d_keys0, (size_t)cells * 8 + 64));
    TRY(ensure_dev(eng, eng->d_keys1, (size_t)cells * 8 + 64));
    TRY(ensure_dev(eng, eng->d_mask, (size_t)cells / 8 + 64));
    TRY(ensure_dev(eng, eng->d_dir_off, m * 8));
    TRY(ensure_dev(eng, eng->d_out_off, m * 8));
    TRY(ensure_dev(eng, eng->d_ncand, m * 4));
    TRY(ensure_dev(eng, eng->d_which, m * 4));
    TRY(ensure_dev(eng, eng->d_nhits, m * 4));
    TRY(ensure_dev(eng, eng->d_rec, m * (size_t)maxh * 32));
    CU_TRY(cudaMemsetAsync(eng->d_rec.p, 0, m * (size_t)maxh * 32, st));   /* records past a pair's hit count are copied too */
    TRY(ensure_dev(eng, eng->d_out_a, (size_t)obytes + 16));
    TRY(ensure_dev(eng, eng->d_out_b, (size_t)obytes + 16));
    /* the walks write each pair's strings right-aligned: the part in front stays untouched, and the copy
     * to the host takes the whole block (initcheck: no uninitialised byte may cross PCIe) */
    CU_TRY(cudaMemsetAsync(eng->d_out_a.p, 0, (size_t)obytes, st));
    CU_TRY(cudaMemsetAsync(eng->d_out_b.p, 0, (size_t)obytes, st));
    TRY(ensure_pin(eng, eng->h_walk, m * 4 + m * (size_t)maxh * 32));
    TRY(ensure_pin(eng, eng->h_str_a, (size_t)obytes + 16));
    TRY(ensure_pin(eng, eng->h_str_b, (size_t)obytes + 16));
    CU_TRY(cudaMemcpyAsync(eng->d_dir_off.p, dir_off.data(), m * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(eng->d_out_off.p, out_off.data(), m * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemsetAsync(eng->d_mask.p, 0, (size_t)cells / 8 + 64, st));
    CU_TRY(cudaMemsetAsync(eng->d_counter.p, 0, 8, st));

    FastArgs F;
    memset(&F, 0, sizeof(F));
    F.seq_a = db.a; F.seq_b = db.b; F.off_a = db.off_a + c0; F.off_b = db.off_b + c0;
    F.npairs = (int64_t)m; F.sp = sp;
    F.tab8 = d_t8; F.tab32 = d_t32;
    F.lut = (const uint8_t *)eng->d_lut.p;
    F.counter = (unsigned long long *)eng->d_counter.p;
    F.score = (int32_t *)eng->d_score.p + c0;
    F.max_lb = (int)bm.max_lb;
    F.dir = (uint8_t *)eng->d_dir.p;
    F.dir_off = (const int64_t *)eng->d_dir_off.p;
    F.m16 = (int16_t *)eng->d_m16.p;
    CU_TRY(cudaEventRecord(eng->ev0, st));
    if(fast_launch(plan, F, eng->num_sms, eng->smem_optin, st) != 0)
      return fail(eng, SEQALIGN_ERR_CUDA, "fast hits kernel launch failed");
    CU_TRY(cudaGetLastError());
    eng->last_launches++;

    HitsArgs H;
    memset(&H, 0, sizeof(H));
    H.seq_a = db.a; H.seq_b = db.b; H.off_a = db.off_a + c0; H.off_b = db.off_b + c0;
    H.npairs = (int64_t)m; H.sp = sp;
    H.sub = (const int32_t *)eng->d_sub.p; H.lut = (const uint8_t *)eng->d_lut.p;
    H.dir = (const uint8_t *)eng->d_dir.p; H.m16 = (const int16_t *)eng->d_m16.p;
    H.dir_off = (const int64_t *)eng->d_dir_off.p;
    H.keys0 = (unsigned long long *)eng->d_keys0.p; H.keys1 = (unsigned long long *)eng->d_keys1.p;
    H.ncand = (int32_t *)eng->d_ncand.p; H.which = (int32_t *)eng->d_which.p;
    H.mask = (unsigned *)eng->d_mask.p;
    H.min_score = eng->hit_min_score; H.max_hits = maxh;
    H.nhits = (int32_t *)eng->d_nhits.p; H.rec = (int32_t *)eng->d_rec.p;
    H.out_a = (uint8_t *)eng->d_out_a.p; H.out_b = (uint8_t *)eng->d_out_b.p;
    H.out_off = (const int64_t *)eng->d_out_off.p;
    H.counter = (unsigned long long *)eng->d_counter.p;
    CU_TRY(cudaMemsetAsync(eng->d_counter.p, 0, 8, st));
    int sgrid = (int)((m + HITS_WARPS - 1) / HITS_WARPS);
    if(sgrid > eng->num_sms * 6) sgrid = eng->num_sms * 6;
    SA_LAUNCH(hits_sort_kernel, sgrid, HITS_WARPS * 32, 0, st, H);
    CU_TRY(cudaGetLastError());
    /* SEQALIGN_HITS_WALK=warp: a warp per pair (round 1); default: a thread per pair, flat state machine */
    const char *hw_env = getenv("SEQALIGN_HITS_WALK");
    if(hw_env && strcmp(hw_env, "warp") == 0) {
      int wgrid = (int)((m + HITS_WALK_WARPS - 1) / HITS_WALK_WARPS);
      if(wgrid > eng->num_sms * 16) wgrid = eng->num_sms * 16;   /* a warp per pair, 64 warps per SM */
      SA_LAUNCH(hits_walk_kernel, wgrid, HITS_WALK_WARPS * 32, 0, st, H);
    } else {
      int wgrid = (int)((m + HITS_FLAT_THREADS - 1) / HITS_FLAT_THREADS);
      if(wgrid > eng->num_sms * 16) wgrid = eng->num_sms * 16;
      SA_LAUNCH(hits_walk_flat_kernel, wgrid, HITS_FLAT_THREADS, 0, st, H);
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(eng->ev1, st));
    eng->last_launches += 2;

    int32_t *hw = (int32_t *)eng->h_walk.p;
    CU_TRY(cudaMemcpyAsync(hw, eng->d_nhits.p, m * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(hw + m, eng->d_rec.p, m * (size_t)maxh * 32, cudaMemcpyDeviceToHost, st));
    if(obytes > 0) {
      CU_TRY(cudaMemcpyAsync(eng->h_str_a.p, eng->d_out_a.p, (size_t)obytes, cudaMemcpyDeviceToHost, st));
      CU_TRY(cudaMemcpyAsync(eng->h_str_b.p, eng->d_out_b.p, (size_t)obytes, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    CU_TRY(cudaEventElapsedTime(&ms, eng->ev0, eng->ev1));
    eng->last_ms += ms;
    memcpy(&eng->nhits[c0], hw, m * 4);
    memcpy(&eng->hit_rec[c0 * (size_t)maxh * 8], hw + m, m * (size_t)maxh * 32);
    if(c0 == 0 && c1 == n) {
      eng->hit_a_p = (const char *)eng->h_str_a.p;
      eng->hit_b_p = (const char *)eng->h_str_b.p;
    } else {
      if(c0 == 0) {
        eng->hit_a.resize((size_t)eng->hit_off[n] + 1);
        eng->hit_b.resize((size_t)eng->hit_off[n] + 1);
      }
      if(obytes > 0) {
        memcpy(&eng->hit_a[(size_t)eng->hit_off[c0]], eng->h_str_a.p, (size_t)obytes);
        memcpy(&eng->hit_b[(size_t)eng->hit_off[c0]], eng->h_str_b.p, (size_t)obytes);
      }
      eng->hit_a_p = eng->hit_a.data();
      eng->hit_b_p = eng->hit_b.data();
    }
    c0 = c1;
  }
  /* scores: best hit of each pair */
  for(size_t i = 0; i < n; i++) {
    eng->score[i] = eng->nhits[i] > 0 ? eng->hit_rec[i * (size_t)maxh * 8] : 0;
    eng->xend[i] = eng->yend[i] = 0;
  }
  eng->last_kernel = "fast_sw_hits+sort+walk";
  return 0;
}

/* batch materialise: the three matrices of every pair stay in device memory,
 * seqalign_batch_matrices() copies one pair's planes out */
/* the materialise kernel over the pairs of wave w; their matrices fill the device block from its start */
int run_mats_wave(seqalign_batch *eng, int w, cudaStream_t st)
{
  const size_t first = eng->mat_wave[w], count = eng->mat_wave[w + 1] - first;
  MatsArgs M;
  memset(&M, 0, sizeof(M));
  M.seq_a = eng->mat_job.a; M.seq_b = eng->mat_job.b;
  M.off_a = eng->mat_job.off_a + first; M.off_b = eng->mat_job.off_b + first;
  M.npairs = (int64_t)count; M.sp = eng->mat_job.sp;
  M.sub = (const int32_t *)eng->d_sub.p; M.lut = (const uint8_t *)eng->d_lut.p;
  /* mat_off stays absolute; the block pointer is moved back by the wave's first offset instead */
  M.mats = (int32_t *)eng->d_mats.p - eng->mat_off[first];
  M.mat_off = (const int64_t *)eng->d_mat_off.p + first;
  M.score = (int32_t *)eng->d_score.p + first;
  M.counter = (unsigned long long *)eng->d_counter.p;
  CU_TRY(cudaMemsetAsync(eng->d_counter.p, 0, 8, st));
  if(mats_launch(eng->mat_job.NB, eng->mat_job.pack, eng->mat_job.nw, M, eng->ft.ncodes, eng->num_sms, eng->smem_optin, st) != 0)
    return fail(eng, SEQALIGN_ERR_CUDA, "materialise kernel launch failed");
  CU_TRY(cudaGetLastError());
  eng->last_launches++;
  eng->mat_resident = w;
  return 0;
}

int run_mats(seqalign_batch *eng, const DevBatch &db, const BatchMeta &bm,
             const int64_t *h_off_a, const int64_t *h_off_b, cudaStream_t st)
{
  const size_t n = db.n;
  const scoring_t *s = eng->scoring;
  const bool nw = eng->algo == SEQALIGN_NW;
  const ScoreParams sp = make_params(s, !nw, eng->ft.ncodes);
  const int NB = mats_blocks(bm.max_la);
  long nw_pen = 0;               /* NW: largest |score| one step can add */
  if(nw) {
    /* the NW rows leave the sentinel of alignment.c:41 out of their maxima: exact only while none of
     * the reference's own `min + penalty` sums wraps and every real score stays far above MATS_NEG */
    const long slack = labs((long)s->min_penalty);
    long pen = labs((long)sp.open);
    if(labs((long)sp.ext) > pen) pen = labs((long)sp.ext);
    if(labs((long)s->gap_open) > pen) pen = labs((long)s->gap_open);
    if(labs((long)eng->ft.min_sub) > pen) pen = labs((long)eng->ft.min_sub);
    if(labs((long)eng->ft.max_sub) > pen) pen = labs((long)eng->ft.max_sub);
    nw_pen = pen;
    if(slack + sp.open < 0 || slack + sp.ext < 0 || slack + eng->ft.min_sub < 0 ||
       (long)(bm.max_la + bm.max_lb + 2) * pen > (1L << 27))
      return fail(eng, SEQALIGN_ERR_ARG,
                  "batch materialise (NW) needs affine gaps with gap_open <= 0, penalties within min_penalty, no gap/mismatch "
                  "restrictions and len_a <= 511; use aligner_align() for this input");
  }
  if(eng->force_mode == 1 || (sp.no_end && !nw) || sp.no_gaps_a || sp.no_gaps_b || sp.no_mismatches || s->gap_open > 0 ||
     s->gap_extend > 0 || eng->ft.any_unknown || NB == 0 || eng->ft.min_sub < -32768 || eng->ft.max_sub > 32767 ||
     (long)bm.max_la * (eng->ft.max_sub > 0 ? eng->ft.max_sub : 0) > (1L << 28) || sp.open < -(1 << 20))
    return fail(eng, SEQALIGN_ERR_ARG,
                "batch materialise needs affine gaps (gap_open <= 0, no free end gaps, no gap/mismatch "
                "restrictions) and len_a <= 511; use aligner_align() for this input");
  eng->mat_off.assign(n + 1, 0);
  for(size_t i = 0; i < n; i++)
    eng->mat_off[i + 1] = eng->mat_off[i] + 3 * ((h_off_a[i + 1] - h_off_a[i]) + 1) * ((h_off_b[i + 1] - h_off_b[i]) + 1);
  /* waves: as many whole pairs as fit three quarters of what the device can give (SEQALIGN_MATS_BUDGET: bytes) */
  size_t free_b = 0, total_b = 0;
  CU_TRY(cudaMemGetInfo(&free_b, &total_b));
  size_t budget = (free_b + eng->d_mats.cap) / 4 * 3;
  const char *benv = getenv("SEQALIGN_MATS_BUDGET");
  if(benv && atoll(benv) > 0) budget = (size_t)atoll(benv);
  eng->mat_wave.assign(1, 0);
  size_t largest = 0;
  for(size_t i = 0; i < n;) {
    size_t j = i;
    while(j < n && (size_t)(eng->mat_off[j + 1] - eng->mat_off[i]) * 4 <= budget) j++;
    if(j == i) return fail(eng, SEQALIGN_ERR_NOMEM, "the matrices of one pair do not fit device memory (12 bytes per cell)");
    if((size_t)(eng->mat_off[j] - eng->mat_off[i]) > largest) largest = (size_t)(eng->mat_off[j] - eng->mat_off[i]);
    eng->mat_wave.push_back(j);
    i = j;
  }
  TRY(ensure_dev(eng, eng->d_mats, largest * 4 + 64));
  TRY(ensure_dev(eng, eng->d_mat_off, (n + 1) * 8));
  TRY(ensure_dev(eng, eng->d_score, n * 4));
  TRY(ensure_dev(eng, eng->d_counter, COUNTER_BYTES));
  CU_TRY(cudaMemcpyAsync(eng->d_mat_off.p, eng->mat_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
  /* packed 16-bit prefix scans when every scan value (score + x*|ext|) fits */
  const long shortest = (long)(bm.max_la < bm.max_lb ? bm.max_la : bm.max_lb);
  bool pack = !nw && shortest * (eng->ft.max_sub > 0 ? eng->ft.max_sub : 0) - 512L * sp.ext < 32000 && !getenv("SEQALIGN_MATS_NOPACK");
  /* NW with packed scans wherever every score and scan value fits int16 (timed on a B200: 69 vs 67 % of the HBM
   * roofline at 150x150, 73 vs 58 % at 400x400 protein; profiles/mats_nw_r02l.jsonl) */
  if(nw && !getenv("SEQALIGN_MATS_NOPACK") && (long)(bm.max_la + bm.max_lb + 2) * nw_pen - 512L * sp.ext < 32000) pack = true;
  eng->mat_job.a = db.a; eng->mat_job.b = db.b; eng->mat_job.off_a = db.off_a; eng->mat_job.off_b = db.off_b;
  eng->mat_job.NB = NB; eng->mat_job.pack = pack; eng->mat_job.nw = nw; eng->mat_job.sp = sp;
  eng->mat_resident = -1;
  const int nwaves = (int)eng->mat_wave.size() - 1;
  CU_TRY(cudaEventRecord(eng->ev0, st));
  /* every wave once for the scores; the first one again at the end when there are several, so that a reader
   * going through the pairs in order starts on a resident wave */
  for(int w = 0; w < nwaves; w++) TRY(run_mats_wave(eng, w, st));
  if(nwaves > 1) TRY(run_mats_wave(eng, 0, st));
  CU_TRY(cudaEventRecord(eng->ev1, st));
  TRY(ensure_pin(eng, eng->h_res, n * 4));
  CU_TRY(cudaMemcpyAsync(eng->h_res.p, eng->d_score.p, n * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  float ms = 0;
  CU_TRY(cudaEventElapsedTime(&ms, eng->ev0, eng->ev1));
  eng->last_ms = ms;
  memcpy(eng->score.data(), eng->h_res.p, n * 4);
  eng->last_kernel = nw ? (pack ? "mats_nw_packed" : "mats_nw") : pack ? "mats_sw_packed" : "mats_sw";
  return 0;
}

/* h_off_a == NULL: every pair is ula x ulb (seqalign_batch_submit_uniform), nothing per pair is read on the host */
int submit_common(seqalign_batch *eng, int algo, int mode, const char *h_a, const int64_t *h_off_a,
                  const char *h_b, const int64_t *h_off_b, size_t n, int64_t ula = -1, int64_t ulb = -1,
                  const uint8_t *dev_a = nullptr, const uint8_t *dev_b = nullptr)
{
  /* dev_a / dev_b: the sequences already lie in HBM (seqalign_batch_submit_reads); h_a / h_b are null then */
  const bool resident = dev_a != nullptr;
  std::vector<char> back_a, back_b;   /* host copies, only made when an unknown character pair must be located */
  auto host_copy = [&](int64_t total_a, int64_t total_b) -> int {
    if(!resident || h_a) return 0;
    back_a.resize((size_t)total_a + 1); back_b.resize((size_t)total_b + 1);
    if(total_a) CU_TRY(cudaMemcpy(back_a.data(), dev_a, (size_t)total_a, cudaMemcpyDeviceToHost));
    if(total_b) CU_TRY(cudaMemcpy(back_b.data(), dev_b, (size_t)total_b, cudaMemcpyDeviceToHost));
    h_a = back_a.data(); h_b = back_b.data();
    return 0;
  };
  while(eng->pend_count > 0) TRY(seqalign_batch_run_device_wait(eng));   /* runs launched ahead finish first */
  eng->err.clear();
  {
    /* an error some earlier runtime call of this thread left behind would otherwise be pinned on the first launch below */
    const cudaError_t stale = cudaGetLastError();
    if(stale != cudaSuccess)
      return fail(eng, SEQALIGN_ERR_CUDA, (std::string("CUDA error pending before this submit: ") + cudaGetErrorString(stale)).c_str());
  }
  eng->n = 0;
  eng->last_launches = 0;
  eng->last_ms = 0;
  if(!eng->have_scoring) return fail(eng, SEQALIGN_ERR_ARG, "seqalign_batch_set_scoring() has not been called");
  if((algo != SEQALIGN_NW && algo != SEQALIGN_SW) || (mode != SEQALIGN_MODE_SCORE && mode != SEQALIGN_MODE_ALIGN && mode != SEQALIGN_MODE_SCORE_ONLY &&
      mode != SEQALIGN_MODE_HITS && mode != SEQALIGN_MODE_MATS) ||
     (mode == SEQALIGN_MODE_HITS && algo != SEQALIGN_SW))
    return fail(eng, SEQALIGN_ERR_ARG, "bad algo/mode");
  const bool score_mode = mode == SEQALIGN_MODE_SCORE || mode == SEQALIGN_MODE_SCORE_ONLY;
  std::vector<int64_t> made_a, made_b;
  const bool given_uniform = h_off_a == nullptr;
  if(given_uniform && !score_mode) {
    /* the other modes walk the offsets on the host anyway */
    made_a.resize(n + 1); made_b.resize(n + 1);
    for(size_t i = 0; i <= n; i++) { made_a[i] = (int64_t)i * ula; made_b[i] = (int64_t)i * ulb; }
    h_off_a = made_a.data(); h_off_b = made_b.data();
  }
  eng->algo = algo; eng->mode = mode;
  const bool to_sink = score_mode && eng->sink_score != nullptr;
  eng->last_to_sink = to_sink;
  if(!to_sink) { eng->score.assign(n, 0); eng->xend.assign(n, 0); eng->yend.assign(n, 0); }
  if(n == 0) return 0;
  cudaStream_t st = eng->stream;
  bool uniform_batch = true;   /* every pair la0 x lb0: the device can make the offsets itself */
  int64_t la0 = ula, lb0 = ulb;
  if(h_off_a) {
    la0 = h_off_a[1] - h_off_a[0]; lb0 = h_off_b[1] - h_off_b[0];
    for(size_t i = 0; i < n; i++) {
      const int64_t la = h_off_a[i + 1] - h_off_a[i], lb = h_off_b[i + 1] - h_off_b[i];
      if(la < 0 || lb < 0 || la > (1 << 30) || lb > (1 << 30))
        return fail(eng, SEQALIGN_ERR_ARG, "bad offsets / sequence too long");
      uniform_batch = uniform_batch && la == la0 && lb == lb0;
    }
  } else if(la0 < 0 || lb0 < 0 || la0 > (1 << 30) || lb0 > (1 << 30)) {
    return fail(eng, SEQALIGN_ERR_ARG, "bad sequence length");
  }
  /* offset of pair i in the host buffers */
  auto hoa = [&](size_t i) -> int64_t { return h_off_a ? h_off_a[i] : (int64_t)i * la0; };
  auto hob = [&](size_t i) -> int64_t { return h_off_b ? h_off_b[i] : (int64_t)i * lb0; };
  const int64_t total_a = hoa(n), total_b = hob(n);
  if(!resident) {
    TRY(ensure_dev(eng, eng->d_seq_a, (size_t)total_a + 32));
    TRY(ensure_dev(eng, eng->d_seq_b, (size_t)total_b + 32));
  }
  TRY(ensure_dev(eng, eng->d_off_a, (n + 1) * 8));
  TRY(ensure_dev(eng, eng->d_off_b, (n + 1) * 8));
  const uint8_t *d_a = resident ? dev_a : (const uint8_t *)eng->d_seq_a.p, *d_b = resident ? dev_b : (const uint8_t *)eng->d_seq_b.p;
  const int64_t *d_oa = (const int64_t *)eng->d_off_a.p, *d_ob = (const int64_t *)eng->d_off_b.p;

  if(mode == SEQALIGN_MODE_SCORE || mode == SEQALIGN_MODE_SCORE_ONLY) {
    /* Pipelined over chunks of pairs: the copy stream pushes chunk c+1 over
     * PCIe while the compute stream scans and aligns chunk c.  Every chunk is
     * a self-contained pass (own alphabet scan; tables are cached). */
    const bool ends = mode == SEQALIGN_MODE_SCORE;
    cudaStream_t cs = eng->copy_stream;
    TRY(ensure_dev(eng, eng->d_score, n * 4));
    TRY(ensure_dev(eng, eng->d_xend, n * 4));
    TRY(ensure_dev(eng, eng->d_yend, n * 4));
    if(!to_sink) TRY(ensure_pin(eng, eng->h_res, n * 12));
    /* results land in the pinned staging block, or straight in the caller's sink */
    int32_t *hr_s = to_sink ? eng->sink_score : (int32_t *)eng->h_res.p;
    int32_t *hr_x = to_sink ? eng->sink_x : (int32_t *)eng->h_res.p + n;
    int32_t *hr_y = to_sink ? eng->sink_y : (int32_t *)eng->h_res.p + 2 * n;
    /* ~16 MB per chunk: small chunks leave the persistent DP kernel with a
     * fraction of a wave at its tail (measured: 4 x 25k pairs of 150 bp cost
     * 0.79 ms of kernel time, 2 x 50k 0.56 ms, one launch 0.54 ms) */
    int nchunks = (int)((total_a - hoa(0) + total_b - hob(0)) / (16 << 20)) + 1;
    if(nchunks > seqalign_batch::MAX_CHUNKS) nchunks = seqalign_batch::MAX_CHUNKS;
    if((size_t)nchunks > n / 2048 + 1) nchunks = (int)(n / 2048 + 1);
    const char *env = getenv("SEQALIGN_CHUNKS");
    if(env && atoi(env) >= 1 && atoi(env) <= seqalign_batch::MAX_CHUNKS) nchunks = atoi(env);
    if((size_t)nchunks > n) nchunks = (int)n;
    if(uniform_batch && (n >= 4096 || given_uniform)) {
      int ogrid = (int)((n + 256) / 256);
      if(ogrid > eng->num_sms * 8) ogrid = eng->num_sms * 8;
      SA_LAUNCH(uniform_offsets_kernel, ogrid, 256, 0, cs, (int64_t *)eng->d_off_a.p, (int64_t *)eng->d_off_b.p,
                (int64_t)n, la0, lb0);
      CU_TRY(cudaGetLastError());
      eng->last_launches++;
    } else {
      CU_TRY(cudaMemcpyAsync(eng->d_off_a.p, h_off_a, (n + 1) * 8, cudaMemcpyHostToDevice, cs));
      CU_TRY(cudaMemcpyAsync(eng->d_off_b.p, h_off_b, (n + 1) * 8, cudaMemcpyHostToDevice, cs));
    }
    size_t bounds[seqalign_batch::MAX_CHUNKS + 1];
    for(int c = 0; c <= nchunks; c++) bounds[c] = n * (size_t)c / (size_t)nchunks;
    for(int c = 0; c < nchunks; c++) {
      const int64_t a0 = hoa(bounds[c]), a1 = hoa(bounds[c + 1]);
      const int64_t b0 = hob(bounds[c]), b1 = hob(bounds[c + 1]);
      if(a1 > a0 && !resident) CU_TRY(cudaMemcpyAsync((uint8_t *)eng->d_seq_a.p + a0, h_a + a0, (size_t)(a1 - a0), cudaMemcpyHostToDevice, cs));
      if(b1 > b0 && !resident) CU_TRY(cudaMemcpyAsync((uint8_t *)eng->d_seq_b.p + b0, h_b + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, cs));
      CU_TRY(cudaEventRecord(eng->ev_copy[c], cs));
      /* the alphabet/shape scan of a chunk runs on its own stream as soon as
       * the chunk has landed, next to the DP kernel of the chunk before it */
      const size_t c0 = bounds[c], m = bounds[c + 1] - c0;
      if(m == 0) continue;
      CU_TRY(cudaStreamWaitEvent(eng->scan_stream, eng->ev_copy[c], 0));
      TRY(scan_launch(eng, d_a, d_b, d_oa + c0, d_ob + c0, m,
                      hoa(c0 + m) - hoa(c0) + hob(c0 + m) - hob(c0), eng->scan_stream, c));
      CU_TRY(cudaEventRecord(eng->ev_scan[c], eng->scan_stream));
    }
    for(int c = 0; c < nchunks; c++) {
      const size_t c0 = bounds[c], m = bounds[c + 1] - c0;
      if(m == 0) continue;
      DevBatch db;
      db.a = d_a; db.b = d_b; db.off_a = d_oa + c0; db.off_b = d_ob + c0; db.n = m;
      BatchMeta bm;
      CU_TRY(cudaEventSynchronize(eng->ev_scan[c]));
      scan_decode(eng, m, c, &bm);
      CU_TRY(cudaStreamWaitEvent(st, eng->ev_copy[c], 0));
      TRY(upload_tables(eng, bm, st));
      if(eng->ft.any_unknown) {
        TRY(host_copy(total_a, total_b));
        if(h_off_a) TRY(check_unknown_pairs(eng, h_a, h_off_a + c0, h_b, h_off_b + c0, m));
        else {
          std::vector<int64_t> ta(m + 1), tb(m + 1);
          for(size_t i = 0; i <= m; i++) { ta[i] = (int64_t)(c0 + i) * la0; tb[i] = (int64_t)(c0 + i) * lb0; }
          TRY(check_unknown_pairs(eng, h_a, ta.data(), h_b, tb.data(), m));
        }
      }
      int32_t *ds = (int32_t *)eng->d_score.p + c0, *dx = (int32_t *)eng->d_xend.p + c0, *dy = (int32_t *)eng->d_yend.p + c0;
      TRY(run_score(eng, algo, db, bm, ds, ends ? dx : nullptr, ends ? dy : nullptr, st, eng->ev_k0[c], eng->ev_k1[c]));
      CU_TRY(cudaMemcpyAsync(hr_s + c0, ds, m * 4, cudaMemcpyDeviceToHost, st));
      if(ends && hr_x) CU_TRY(cudaMemcpyAsync(hr_x + c0, dx, m * 4, cudaMemcpyDeviceToHost, st));
      if(ends && hr_y) CU_TRY(cudaMemcpyAsync(hr_y + c0, dy, m * 4, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(cudaStreamSynchronize(st));
    eng->last_ms = 0;
    for(int c = 0; c < nchunks; c++) {
      if(bounds[c + 1] == bounds[c]) continue;
      float ms = 0;
      CU_TRY(cudaEventElapsedTime(&ms, eng->ev_k0[c], eng->ev_k1[c]));
      eng->last_ms += ms;
    }
    if(!to_sink) {
      memcpy(eng->score.data(), hr_s, n * 4);
      if(ends) {
        memcpy(eng->xend.data(), hr_x, n * 4);
        memcpy(eng->yend.data(), hr_y, n * 4);
      } else if(algo == SEQALIGN_NW) {
        for(size_t i = 0; i < n; i++) {
          eng->xend[i] = (int32_t)(hoa(i + 1) - hoa(i));
          eng->yend[i] = (int32_t)(hob(i + 1) - hob(i));
        }
      }
    } else if(!ends) {
      /* score only into a sink: the end-cell arrays, if given, read like the internal ones */
      for(size_t i = 0; i < n; i++) {
        if(eng->sink_x) eng->sink_x[i] = algo == SEQALIGN_NW ? (int32_t)(hoa(i + 1) - hoa(i)) : 0;
        if(eng->sink_y) eng->sink_y[i] = algo == SEQALIGN_NW ? (int32_t)(hob(i + 1) - hob(i)) : 0;
      }
    }
  } else {
    if(total_a && !resident) CU_TRY(cudaMemcpyAsync(eng->d_seq_a.p, h_a, (size_t)total_a, cudaMemcpyHostToDevice, st));
    if(total_b && !resident) CU_TRY(cudaMemcpyAsync(eng->d_seq_b.p, h_b, (size_t)total_b, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(eng->d_off_a.p, h_off_a, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(eng->d_off_b.p, h_off_b, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    DevBatch db;
    db.a = d_a; db.b = d_b; db.off_a = d_oa; db.off_b = d_ob; db.n = n;
    BatchMeta bm;
    TRY(scan_batch(eng, db.a, db.b, db.off_a, db.off_b, n, total_a - hoa(0) + total_b - hob(0), st, &bm));
    TRY(upload_tables(eng, bm, st));
    if(eng->ft.any_unknown) { TRY(host_copy(total_a, total_b)); TRY(check_unknown_pairs(eng, h_a, h_off_a, h_b, h_off_b, n)); }
    if(mode == SEQALIGN_MODE_MATS) TRY(run_mats(eng, db, bm, h_off_a, h_off_b, st));
    else if(mode == SEQALIGN_MODE_HITS) TRY(run_hits(eng, db, bm, h_off_a, h_off_b, st));
    else TRY(run_align(eng, algo, db, bm, h_off_a, h_off_b, st));
  }
  eng->n = n;
  return 0;
}

} // namespace

/* =========================================================================
 * C-ABI
 */
extern "C" {

int seqalign_device_count(void)
{
  int n = 0;
  if(cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for(int i = 0; i < n; i++) {
    cudaDeviceProp p;
    if(cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ok++;
  }
  return ok;
}

const char *seqalign_version(void) { return "seqalign_b200 0.1 (sm_100a)"; }

int seqalign_shared_alloc(int device, size_t bytes, void **d_ptr, seqalign_ipc_handle_t *handle)
{
  static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(seqalign_ipc_handle_t), "handle size");
  if(!d_ptr || !handle) return SEQALIGN_ERR_ARG;
  *d_ptr = nullptr;
  cudaIpcMemHandle_t h;
  if(cudaSetDevice(device) != cudaSuccess || cudaMalloc(d_ptr, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return SEQALIGN_ERR_NOMEM; }
  if(cudaIpcGetMemHandle(&h, *d_ptr) != cudaSuccess) { cudaGetLastError(); cudaFree(*d_ptr); *d_ptr = nullptr; return SEQALIGN_ERR_CUDA; }
  memcpy(handle->bytes, &h, sizeof(h));
  return SEQALIGN_OK;
}

int seqalign_shared_free(int device, void *d_ptr)
{
  if(cudaSetDevice(device) != cudaSuccess || cudaFree(d_ptr) != cudaSuccess) { cudaGetLastError(); return SEQALIGN_ERR_CUDA; }
  return SEQALIGN_OK;
}

int seqalign_shared_open(int device, const seqalign_ipc_handle_t *handle, void **d_ptr)
{
  if(!d_ptr || !handle) return SEQALIGN_ERR_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle->bytes, sizeof(h));
  /* opened with `device` current: the mapping is made for this device's kernels */
  if(cudaSetDevice(device) != cudaSuccess ||
     cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return SEQALIGN_ERR_CUDA; }
  return SEQALIGN_OK;
}

int seqalign_shared_close(int device, void *d_ptr)
{
  if(cudaSetDevice(device) != cudaSuccess || cudaIpcCloseMemHandle(d_ptr) != cudaSuccess) { cudaGetLastError(); return SEQALIGN_ERR_CUDA; }
  return SEQALIGN_OK;
}

int seqalign_enable_peer_access(int device, int peer)
{
  if(device == peer) return SEQALIGN_OK;
  int can = 0;
  if(cudaSetDevice(device) != cudaSuccess || cudaDeviceCanAccessPeer(&can, device, peer) != cudaSuccess || !can) {
    cudaGetLastError();
    return SEQALIGN_ERR_CUDA;
  }
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if(e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return SEQALIGN_ERR_CUDA; }
  cudaGetLastError();
  return SEQALIGN_OK;
}

const char *seqalign_last_create_error(void) { return g_create_error.c_str(); }

void *seqalign_host_alloc(size_t bytes)
{
  void *p = nullptr;
  if(cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}

void seqalign_host_free(void *p)
{
  if(p && cudaFreeHost(p) != cudaSuccess) cudaGetLastError();
}

int seqalign_synth_batch(int device, int kind, uint64_t seed, int64_t first_pair, int64_t npairs,
                         int len_a, int len_b, void *d_seq_a, void *d_seq_b, void *stream)
{
  if(npairs < 0 || len_a < 0 || len_b < 0 || (kind != 0 && kind != 1) || !d_seq_a || !d_seq_b) return SEQALIGN_ERR_ARG;
  if(npairs == 0) return SEQALIGN_OK;
  if(cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return SEQALIGN_ERR_CUDA; }
  SynthArgs A;
  A.seed = seed; A.first_pair = first_pair; A.npairs = npairs; A.len_a = len_a; A.len_b = len_b;
  A.k = kind == 0 ? 4 : 20;
  /* substitution / indel probabilities as 16-bit thresholds: DNA 0.05 / 0.01, protein 0.15 / 0.02 */
  A.t_sub = kind == 0 ? 3277u : 9830u;
  A.t_indel = kind == 0 ? 655u : 1311u;
  A.seq_a = (uint8_t *)d_seq_a; A.seq_b = (uint8_t *)d_seq_b;
  int64_t grid = (npairs + 127) / 128;
  if(grid > 148 * 32) grid = 148 * 32;
  cudaStream_t st = (cudaStream_t)stream;
  SA_LAUNCH(synth_kernel, (int)grid, 128, 0, st, A);
  if(cudaGetLastError() != cudaSuccess) return SEQALIGN_ERR_CUDA;
  if(!stream && cudaStreamSynchronize(st) != cudaSuccess) { cudaGetLastError(); return SEQALIGN_ERR_CUDA; }
  return SEQALIGN_OK;
}

seqalign_batch_t *seqalign_batch_create(int device)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if(e != cudaSuccess || n == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                     " (this library has no CPU path)";
    cudaGetLastError();
    return nullptr;
  }
  if(device < 0 || device >= n) { g_create_error = "device index out of range"; return nullptr; }
  cudaDeviceProp p;
  if(cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&p, device) != cudaSuccess) {
    g_create_error = "cannot select device";
    return nullptr;
  }
  if(p.major != 10) {
    g_create_error = std::string("device '") + p.name + "' is not sm_100: kernels are built for sm_100a only";
    return nullptr;
  }
  seqalign_batch *eng = new seqalign_batch();
  eng->device = device;
  eng->num_sms = p.multiProcessorCount;
  eng->smem_optin = p.sharedMemPerBlockOptin;
  bool ok = cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&eng->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&eng->scan_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreate(&eng->ev0) == cudaSuccess && cudaEventCreate(&eng->ev1) == cudaSuccess;
  for(int i = 0; ok && i < seqalign_batch::MAX_CHUNKS; i++)
    ok = cudaEventCreate(&eng->ev_copy[i]) == cudaSuccess && cudaEventCreate(&eng->ev_k0[i]) == cudaSuccess &&
         cudaEventCreate(&eng->ev_k1[i]) == cudaSuccess &&
         cudaEventCreate(&eng->ev_scan[i]) == cudaSuccess;
  if(!ok) {
    g_create_error = "cannot create stream/events";
    delete eng;
    return nullptr;
  }
  eng->scoring = (scoring_t *)malloc(sizeof(scoring_t));
  return eng;
}

void seqalign_batch_destroy(seqalign_batch_t *eng)
{
  if(!eng) return;
  while(eng->pend_count > 0) { if(seqalign_batch_run_device_wait(eng) != 0) break; }
  cudaSetDevice(eng->device);
  DevBuf *d[] = {&eng->d_seq_a, &eng->d_seq_b, &eng->d_off_a, &eng->d_off_b, &eng->d_meta, &eng->d_counter,
                 &eng->d_sub, &eng->d_forbid, &eng->d_lut, &eng->d_tab8, &eng->d_bnd, &eng->d_score,
                 &eng->d_xend, &eng->d_yend, &eng->d_state, &eng->d_dir, &eng->d_dir_off, &eng->d_out_a,
                 &eng->d_out_b, &eng->d_out_off, &eng->d_walk, &eng->d_mats, &eng->d_m16, &eng->d_keys0,
                 &eng->d_keys1, &eng->d_mask, &eng->d_ncand, &eng->d_which, &eng->d_nhits, &eng->d_rec, &eng->d_lbnd, &eng->d_mat_off, &eng->d_swkey, &eng->d_bucket};
  for(DevBuf *b : d) b->release();
  PinBuf *h[] = {&eng->h_in_a, &eng->h_in_b, &eng->h_off_a, &eng->h_off_b, &eng->h_meta, &eng->h_res,
                 &eng->h_walk, &eng->h_str_a, &eng->h_str_b};
  for(PinBuf *b : h) b->release();
  for(int i = 0; i < seqalign_batch::MAX_CHUNKS; i++) {
    if(eng->ev_copy[i]) cudaEventDestroy(eng->ev_copy[i]);
    if(eng->ev_k0[i]) cudaEventDestroy(eng->ev_k0[i]);
    if(eng->ev_k1[i]) cudaEventDestroy(eng->ev_k1[i]);
    if(eng->ev_scan[i]) cudaEventDestroy(eng->ev_scan[i]);
  }
  for(int k = 0; k < seqalign_batch::BUCKET_STREAMS; k++) {
    if(eng->bucket_streams[k]) cudaStreamDestroy(eng->bucket_streams[k]);
    if(eng->bucket_done[k]) cudaEventDestroy(eng->bucket_done[k]);
  }
  if(eng->bucket_ready) cudaEventDestroy(eng->bucket_ready);
  if(eng->copy_stream) cudaStreamDestroy(eng->copy_stream);
  if(eng->scan_stream) cudaStreamDestroy(eng->scan_stream);
  if(eng->ev0) cudaEventDestroy(eng->ev0);
  if(eng->ev1) cudaEventDestroy(eng->ev1);
  if(eng->stream) cudaStreamDestroy(eng->stream);
  free(eng->scoring);
  delete eng;
}

const char *seqalign_batch_error(const seqalign_batch_t *eng) { return eng ? eng->err.c_str() : "null engine"; }

int seqalign_batch_set_scoring(seqalign_batch_t *eng, const scoring_t *scoring)
{
  if(eng) while(eng->pend_count > 0) { if(seqalign_batch_run_device_wait(eng) != 0) break; }
  if(!eng || !scoring) return SEQALIGN_ERR_ARG;
  if(!eng->have_scoring || memcmp(eng->scoring, scoring, sizeof(scoring_t)) != 0) {
    memcpy(eng->scoring, scoring, sizeof(scoring_t));
    eng->scoring_version++;
  }
  eng->have_scoring = true;
  return 0;
}

void seqalign_batch_force_general(seqalign_batch_t *eng, int on) { if(eng) eng->force_mode = on; }

int seqalign_batch_submit_packed(seqalign_batch_t *eng, int algo, int mode,
                                 const char *seq_a, const int64_t *off_a,
                                 const char *seq_b, const int64_t *off_b, size_t n)
{
  if(!eng) return SEQALIGN_ERR_ARG;
  if(n > 0 && (!off_a || !off_b)) return fail(eng, SEQALIGN_ERR_ARG, "null offsets");
  CU_TRY(cudaSetDevice(eng->device));
  if(n > 0 && (off_a[0] != 0 || off_b[0] != 0)) return fail(eng, SEQALIGN_ERR_ARG, "offsets must start at 0");
  return submit_common(eng, algo, mode, seq_a, off_a, seq_b, off_b, n);
}

int seqalign_batch_submit(seqalign_batch_t *eng, int algo, int mode,
                          const char *const *seq_a, const size_t *len_a,
                          const char *const *seq_b, const size_t *len_b, size_t n)
{
  if(!eng) return SEQALIGN_ERR_ARG;
  CU_TRY(cudaSetDevice(eng->device));
  size_t ta = 0, tb = 0;
  for(size_t i = 0; i < n; i++) { ta += len_a[i]; tb += len_b[i]; }
  TRY(ensure_pin(eng, eng->h_in_a, ta + 16));
  TRY(ensure_pin(eng, eng->h_in_b, tb + 16));
  TRY(ensure_pin(eng, eng->h_off_a, (n + 1) * 8));
  TRY(ensure_pin(eng, eng->h_off_b, (n + 1) * 8));
  char *pa = (char *)eng->h_in_a.p, *pb = (char *)eng->h_in_b.p;
  int64_t *oa = (int64_t *)eng->h_off_a.p, *ob = (int64_t *)eng->h_off_b.p;
  oa[0] = ob[0] = 0;
  for(size_t i = 0; i < n; i++) {
    memcpy(pa + oa[i], seq_a[i], len_a[i]);
    memcpy(pb + ob[i], seq_b[i], len_b[i]);
    oa[i + 1] = oa[i] + (int64_t)len_a[i];
    ob[i + 1] = ob[i] + (int64_t)len_b[i];
  }
  return submit_common(eng, algo, mode, pa, oa, pb, ob, n);
}

size_t seqalign_batch_size(const seqalign_batch_t *eng) { return eng ? eng->n : 0; }

int seqalign_batch_submit_reads(seqalign_batch_t *eng, int algo, int mode, const seqalign_reads_t *ra, int side_a,
                                const seqalign_reads_t *rb, int side_b, size_t first, size_t n)
{
  if(!eng) return SEQALIGN_ERR_ARG;
  if(!ra || !rb || side_a < 0 || side_a > 1 || side_b < 0 || side_b > 1) return fail(eng, SEQALIGN_ERR_ARG, "bad reads / side");
  if(seqalign_reads_device(ra) != eng->device || seqalign_reads_device(rb) != eng->device)
    return fail(eng, SEQALIGN_ERR_ARG, "reads and engine live on different devices");
  if(first + n > seqalign_reads_count(ra, side_a) || first + n > seqalign_reads_count(rb, side_b))
    return fail(eng, SEQALIGN_ERR_ARG, "more pairs asked for than records decoded");
  CU_TRY(cudaSetDevice(eng->device));
  if(n == 0) return submit_common(eng, algo, mode, "", nullptr, "", nullptr, 0, 0, 0);
  /* offsets stay absolute (relative to the start of the side's buffer): the kernels add them to the base pointer */
  return submit_common(eng, algo, mode, nullptr, seqalign_reads_offsets(ra, side_a) + first, nullptr,
                       seqalign_reads_offsets(rb, side_b) + first, n, -1, -1,
                       (const uint8_t *)seqalign_reads_device_seq(ra, side_a), (const uint8_t *)seqalign_reads_device_seq(rb, side_b));
}

int seqalign_batch_set_result_sink(seqalign_batch_t *eng, int32_t *score, int32_t *x_end, int32_t *y_end)
{
  if(!eng) return SEQALIGN_ERR_ARG;
  eng->sink_score = score;
  eng->sink_x = score ? x_end : nullptr;
  eng->sink_y = score ? y_end : nullptr;
  return 0;
}

int seqalign_batch_submit_uniform(seqalign_batch_t *eng, int algo, int mode, const char *seq_a, size_t len_a,
                                  const char *seq_b, size_t len_b, size_t n)
{
  if(!eng) return SEQALIGN_ERR_ARG;
  CU_TRY(cudaSetDevice(eng->device));
  return submit_common(eng, algo, mode, seq_a, nullptr, seq_b, nullptr, n, (int64_t)len_a, (int64_t)len_b);
}

int seqalign_batch_scores(seqalign_batch_t *eng, int32_t *score)
{
  if(!eng || !score) return SEQALIGN_ERR_ARG;
  if(eng->last_to_sink) return fail(eng, SEQALIGN_ERR_ARG, "the results of the last submit went to the result sink");
  memcpy(score, eng->score.data(), eng->n * 4);
  return 0;
}

int seqalign_batch_ends(seqalign_batch_t *eng, int32_t *score, int32_t *x_end, int32_t *y_end)
{
  if(!eng) return SEQALIGN_ERR_ARG;
  if(eng->last_to_sink) return fail(eng, SEQALIGN_ERR_ARG, "the results of the last submit went to the result sink");
  if(score) memcpy(score, eng->score.data(), eng->n * 4);
  if(x_end) memcpy(x_end, eng->xend.data(), eng->n * 4);
  if(y_end) memcpy(y_end, eng->yend.data(), eng->n * 4);
  return 0;
}

int seqalign_batch_alignment(seqalign_batch_t *eng, size_t i, alignment_t *out)
{
  if(!eng || !out) return SEQALIGN_ERR_ARG;
  if(eng->mode != SEQALIGN_MODE_ALIGN || i >= eng->n) return fail(eng, SEQALIGN_ERR_ARG, "no alignment for this index");
  if(eng->status[i] == WALK_FAIL) return fail(eng, SEQALIGN_ERR_TRACEBACK, "traceback fail");
  if(eng->status[i] == WALK_NOHIT) return 0;
  const size_t len = (size_t)eng->aln_len[i];
  /* grow like alignment_ensure_capacity (reference alignment.c:219-233) */
  if(out->capacity < len + 1) {
    size_t cap = ROUNDUP2POW(len + 1);
    out->result_a = (char *)realloc(out->result_a, cap);
    out->result_b = (char *)realloc(out->result_b, cap);
    out->capacity = cap;
    if(!out->result_a || !out->result_b) return fail(eng, SEQALIGN_ERR_NOMEM, "Out of memory");
  }
  const size_t base = (size_t)eng->res_off[i] + (size_t)eng->aln_start[i];
  memcpy(out->result_a, eng->res_a_p + base, len);
  memcpy(out->result_b, eng->res_b_p + base, len);
  out->result_a[len] = out->result_b[len] = '\0';
  out->length = len;
  out->score = eng->score[i];
  if(eng->algo == SEQALIGN_SW) {
    out->pos_a = (size_t)eng->pos_a[i]; out->pos_b = (size_t)eng->pos_b[i];
    out->len_a = (size_t)eng->len_a[i]; out->len_b = (size_t)eng->len_b[i];
  }
  return 1;
}

static bool can_speculate(const seqalign_batch *eng, int algo, bool want_ends)
{
  return eng->spec.valid && eng->spec.version == eng->scoring_version && eng->spec.algo == algo &&
         eng->spec.want_ends == want_ends && eng->tables_valid && eng->tables_version == eng->scoring_version &&
         eng->spec.plan.tab32 == eng->dev_tab32 && eng->spec.plan.tab8 == eng->dev_tab8 &&
         eng->force_mode == 0 && !getenv("SEQALIGN_NO_SPECULATION");
}

/* does the batch whose scan is in `bm` fit the plan it was launched with? */
static bool speculation_held(seqalign_batch *eng, int algo, bool want_ends, const BatchMeta &bm)
{
  const ScoreParams sp = make_params(eng->scoring, algo == SEQALIGN_SW, eng->ft.ncodes);
  uint64_t pres[8];
  bool subset = true;
  for(int i = 0; i < 4; i++) { pres[i] = bm.pres_a[i]; pres[4 + i] = bm.pres_b[i]; }
  for(int i = 0; i < 8; i++) subset = subset && (pres[i] & ~eng->tables_pres[i]) == 0;
  FastPlan now;
  const bool uniform = bm.min_la == bm.max_la && bm.min_lb == bm.max_lb;
  const FastPlan &old = eng->spec.plan;
  /* (a larger shape than needed, or int32 where 16 bits would do, is still exact) */
  return subset && bm.max_lb <= eng->spec.max_lb &&
         fast_plan(eng->scoring, eng->ft, sp, bm.max_la, bm.max_lb, want_ends, true, &now, false, uniform) &&
         now.G * now.K <= old.G * old.K && (now.s16 || !old.s16) && (old.pad_row || !now.pad_row) &&
         (old.track != TRACK_TREE || bm.max_lb <= 2047);
}

int seqalign_batch_run_device_wait(seqalign_batch_t *eng)
{
  if(!eng) return SEQALIGN_ERR_ARG;
  if(eng->pend_count == 0) return 0;
  CU_TRY(cudaSetDevice(eng->device));
  const seqalign_batch::Pending p = eng->pend[eng->pend_head];
  eng->pend_head = (eng->pend_head + 1) % seqalign_batch::MAX_PENDING;
  eng->pend_count--;
  CU_TRY(cudaEventSynchronize(eng->ev_scan[p.slot]));
  BatchMeta bm;
  scan_decode(eng, p.n, p.slot, &bm);
  CU_TRY(cudaEventSynchronize(eng->ev_k1[p.slot]));
  const bool want_ends = (p.dx != nullptr || p.dy != nullptr);
  if(eng->spec.valid && speculation_held(eng, p.algo, want_ends, bm)) {
    float ms = 0;
    CU_TRY(cudaEventElapsedTime(&ms, eng->ev_k0[p.slot], eng->ev_k1[p.slot]));
    eng->last_ms = ms;
    eng->spec_hits++;
    return 0;
  }
  /* wrong guess: this run and everything launched after it with the same plan is redone, in order */
  eng->spec_misses++;
  eng->spec.valid = false;
  seqalign_batch::Pending redo[seqalign_batch::MAX_PENDING + 1];
  int nredo = 0;
  redo[nredo++] = p;
  while(eng->pend_count > 0) {
    redo[nredo] = eng->pend[eng->pend_head];
    eng->pend_head = (eng->pend_head + 1) % seqalign_batch::MAX_PENDING;
    eng->pend_count--;
    CU_TRY(cudaEventSynchronize(eng->ev_k1[redo[nredo].slot]));
    CU_TRY(cudaEventSynchronize(eng->ev_scan[redo[nredo].slot]));
    nredo++;
  }
  for(int i = 0; i < nredo; i++)
    TRY(seqalign_batch_run_device(eng, redo[i].algo, redo[i].a, redo[i].oa, redo[i].b, redo[i].ob, redo[i].n,
                                  redo[i].ds, redo[i].dx, redo[i].dy, redo[i].stream));
  return 0;
}

static int drain_pending(seqalign_batch *eng)
{
  while(eng->pend_count > 0) TRY(seqalign_batch_run_device_wait(eng));
  return 0;
}

int seqalign_batch_run_device_async(seqalign_batch_t *eng, int algo,
                                    const void *d_seq_a, const void *d_off_a,
                                    const void *d_seq_b, const void *d_off_b,
                                    size_t n, void *d_score, void *d_x_end, void *d_y_end, void *stream)
{
  if(!eng) return SEQALIGN_ERR_ARG;
  if(!eng->have_scoring) return fail(eng, SEQALIGN_ERR_ARG, "seqalign_batch_set_scoring() has not been called");
  if((((uintptr_t)d_seq_a) | ((uintptr_t)d_seq_b)) & 15) return fail(eng, SEQALIGN_ERR_ARG, "d_seq_a / d_seq_b must be 16-byte aligned");
  const bool want_ends = (d_x_end != nullptr || d_y_end != nullptr) && eng->force_mode != 3 && eng->force_mode != 4;
  if(n == 0 || !can_speculate(eng, algo, want_ends)) {
    /* nothing to guess from yet (first run, new scoring, ...): the blocking call, which also learns the plan */
    TRY(drain_pending(eng));
    return seqalign_batch_run_device(eng, algo, d_seq_a, d_off_a, d_seq_b, d_off_b, n, d_score, d_x_end, d_y_end, stream);
  }
  if(eng->pend_count == seqalign_batch::MAX_PENDING) TRY(seqalign_batch_run_device_wait(eng));
  if(!can_speculate(eng, algo, want_ends)) {   /* the wait may have found a wrong guess */
    TRY(drain_pending(eng));
    return seqalign_batch_run_device(eng, algo, d_seq_a, d_off_a, d_seq_b, d_off_b, n, d_score, d_x_end, d_y_end, stream);
  }
  eng->err.clear();
  CU_TRY(cudaSetDevice(eng->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : eng->stream;
  const int slot = (eng->pend_head + eng->pend_count) % seqalign_batch::MAX_PENDING;
  DevBatch db;
  db.a = (const uint8_t *)d_seq_a; db.b = (const uint8_t *)d_seq_b;
  db.off_a = (const int64_t *)d_off_a; db.off_b = (const int64_t *)d_off_b;
  db.n = n;
  const ScoreParams sp = make_params(eng->scoring, algo == SEQALIGN_SW, eng->ft.ncodes);
  cudaStream_t cs = eng->copy_stream;
  CU_TRY(cudaEventRecord(eng->ev_copy[slot], st));
  TRY(launch_fast_score(eng, eng->spec.plan, sp, db, eng->spec.max_lb, (int32_t *)d_score,
                        (int32_t *)d_x_end, (int32_t *)d_y_end, st, eng->ev_k0[slot], eng->ev_k1[slot], 1 + slot));
  CU_TRY(cudaStreamWaitEvent(cs, eng->ev_copy[slot], 0));
  TRY(scan_launch(eng, db.a, db.b, db.off_a, db.off_b, n, (int64_t)n * 256, cs, slot));
  CU_TRY(cudaEventRecord(eng->ev_scan[slot], cs));
  seqalign_batch::Pending &q = eng->pend[slot];
  q.slot = slot; q.algo = algo; q.a = d_seq_a; q.oa = d_off_a; q.b = d_seq_b; q.ob = d_off_b; q.n = n;
  q.ds = d_score; q.dx = want_ends ? d_x_end : nullptr; q.dy = want_ends ? d_y_end : nullptr; q.stream = stream;
  eng->pend_count++;
  return 0;
}

int seqalign_batch_run_device(seqalign_batch_t *eng, int algo,
                              const void *d_seq_a, const void *d_off_a,
                              const void *d_seq_b, const void *d_off_b,
                              size_t n, void *d_score, void *d_x_end, void *d_y_end, void *stream)
{
  if(!eng) return SEQALIGN_ERR_ARG;
  if(eng->pend_count > 0) TRY(drain_pending(eng));
  eng->err.clear();
  eng->last_launches = 0;
  if(!eng->have_scoring) return fail(eng, SEQALIGN_ERR_ARG, "seqalign_batch_set_scoring() has not been called");
  if(n == 0) return 0;
  /* the kernels align their 16-byte bulk copies on the OFFSET (seq + (off & ~15)) */
  if((((uintptr_t)d_seq_a) | ((uintptr_t)d_seq_b)) & 15) return fail(eng, SEQALIGN_ERR_ARG, "d_seq_a / d_seq_b must be 16-byte aligned");
  CU_TRY(cudaSetDevice(eng->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : eng->stream;
  DevBatch db;
  db.a = (const uint8_t *)d_seq_a; db.b = (const uint8_t *)d_seq_b;
  db.off_a = (const int64_t *)d_off_a; db.off_b = (const int64_t *)d_off_b;
  db.n = n;
  const bool want_ends = (d_x_end != nullptr || d_y_end != nullptr) && eng->force_mode != 3 && eng->force_mode != 4;
  const int64_t approx_bytes = (int64_t)n * 256;
  BatchMeta bm;
  float ms = 0;

  /* Speculative path: batches of a stream usually share alphabet and shape
   * with their predecessor, so launch the DP kernel right away with the
   * previous plan and tables while the scan of THIS batch runs next to it;
   * the scan result is checked afterwards and the batch is redone the slow
   * way if the guess was wrong (the kernels turn pairs that do not fit the
   * plan into empty ones, so a wrong guess is harmless). */
  if(can_speculate(eng, algo, want_ends)) {
    const ScoreParams sp = make_params(eng->scoring, algo == SEQALIGN_SW, eng->ft.ncodes);
    /* the scan runs on the side stream, next to the DP kernel */
    cudaStream_t cs = eng->copy_stream;
    /* the DP kernel goes first: the scan's handful of launches would otherwise sit in front of it */
    CU_TRY(cudaEventRecord(eng->ev_copy[0], st));
    TRY(launch_fast_score(eng, eng->spec.plan, sp, db, eng->spec.max_lb, (int32_t *)d_score,
                          (int32_t *)d_x_end, (int32_t *)d_y_end, st, eng->ev0, eng->ev1));
    CU_TRY(cudaStreamWaitEvent(cs, eng->ev_copy[0], 0));
    TRY(scan_launch(eng, db.a, db.b, db.off_a, db.off_b, n, approx_bytes, cs));
    TRY(scan_collect(eng, n, cs, &bm));
    CU_TRY(cudaStreamSynchronize(st));
    if(speculation_held(eng, algo, want_ends, bm)) {
      CU_TRY(cudaEventElapsedTime(&ms, eng->ev0, eng->ev1));
      eng->last_ms = ms;
      eng->spec_hits++;
      return 0;
    }
    eng->spec_misses++;
    eng->spec.valid = false;
    eng->last_launches = 0;
  } else {
    TRY(scan_batch(eng, db.a, db.b, db.off_a, db.off_b, n, approx_bytes, st, &bm));
  }

  TRY(upload_tables(eng, bm, st));
  if(eng->ft.any_unknown) {
    int64_t tot[2];
    CU_TRY(cudaMemcpy(&tot[0], (const int64_t *)d_off_a + n, 8, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(&tot[1], (const int64_t *)d_off_b + n, 8, cudaMemcpyDeviceToHost));
    std::vector<char> ha((size_t)tot[0] + 1), hb((size_t)tot[1] + 1);
    std::vector<int64_t> oa(n + 1), ob(n + 1);
    CU_TRY(cudaMemcpy(ha.data(), d_seq_a, (size_t)tot[0], cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(hb.data(), d_seq_b, (size_t)tot[1], cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(oa.data(), d_off_a, (n + 1) * 8, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(ob.data(), d_off_b, (n + 1) * 8, cudaMemcpyDeviceToHost));
    TRY(check_unknown_pairs(eng, ha.data(), oa.data(), hb.data(), ob.data(), n));
  }
  FastPlan used;
  TRY(run_score(eng, algo, db, bm, (int32_t *)d_score, (int32_t *)d_x_end, (int32_t *)d_y_end, st,
                nullptr, nullptr, &used));
  CU_TRY(cudaStreamSynchronize(st));
  CU_TRY(cudaEventElapsedTime(&ms, eng->ev0, eng->ev1));
  eng->last_ms = ms;
  if(used.G != 0 && !eng->ft.any_unknown) {
    eng->spec.valid = true;
    eng->spec.version = eng->scoring_version;
    eng->spec.algo = algo;
    eng->spec.want_ends = want_ends;
    eng->spec.plan = used;
    eng->spec.max_lb = bm.max_lb;
  } else {
    eng->spec.valid = false;
  }
  return 0;
}

int seqalign_fill_matrices(seqalign_batch_t *eng, const char *seq_a, size_t len_a,
                           const char *seq_b, size_t len_b, int is_sw,
                           int32_t *match, int32_t *gap_a, int32_t *gap_b)
{
  if(!eng) return SEQALIGN_ERR_ARG;
  while(eng->pend_count > 0) TRY(seqalign_batch_run_device_wait(eng));
  eng->err.clear();
  eng->last_launches = 0;
  if(!eng->have_scoring) return fail(eng, SEQALIGN_ERR_ARG, "seqalign_batch_set_scoring() has not been called");
  if(len_a > (1u << 30) || len_b > (1u << 30)) return fail(eng, SEQALIGN_ERR_ARG, "sequence too long");
  CU_TRY(cudaSetDevice(eng->device));
  cudaStream_t st = eng->stream;
  const int64_t off_a[2] = {0, (int64_t)len_a}, off_b[2] = {0, (int64_t)len_b};
  TRY(ensure_dev(eng, eng->d_seq_a, len_a + 32));
  TRY(ensure_dev(eng, eng->d_seq_b, len_b + 32));
  TRY(ensure_dev(eng, eng->d_off_a, 16));
  TRY(ensure_dev(eng, eng->d_off_b, 16));
  if(len_a) CU_TRY(cudaMemcpyAsync(eng->d_seq_a.p, seq_a, len_a, cudaMemcpyHostToDevice, st));
  if(len_b) CU_TRY(cudaMemcpyAsync(eng->d_seq_b.p, seq_b, len_b, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(eng->d_off_a.p, off_a, 16, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(eng->d_off_b.p, off_b, 16, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaStreamSynchronize(st));
  DevBatch db;
  db.a = (const uint8_t *)eng->d_seq_a.p; db.b = (const uint8_t *)eng->d_seq_b.p;
  db.off_a = (const int64_t *)eng->d_off_a.p; db.off_b = (const int64_t *)eng->d_off_b.p;
  db.n = 1;
  BatchMeta bm;
  TRY(scan_batch(eng, db.a, db.b, db.off_a, db.off_b, 1, off_a[1] + off_b[1], st, &bm));
  TRY(upload_tables(eng, bm, st));
  if(eng->ft.any_unknown) TRY(check_unknown_pairs(eng, seq_a, off_a, seq_b, off_b, 1));

  const ScoreParams sp = make_params(eng->scoring, is_sw, eng->ft.ncodes);
  const int64_t pitch = ((int64_t)len_a + 4 + 3) & ~(int64_t)3;
  const size_t mat_ints = (size_t)pitch * (len_b + 1) + 16;
  TRY(ensure_dev(eng, eng->d_mats, mat_ints * 4 * 3));
  GenArgs X;
  memset(&X, 0, sizeof(X));
  X.mat_m = (int32_t *)eng->d_mats.p;
  X.mat_ga = X.mat_m + mat_ints;
  X.mat_gb = X.mat_ga + mat_ints;
  X.pitch = pitch;
  CU_TRY(cudaEventRecord(eng->ev0, st));
  TRY(launch_general<MODE_MATS>(eng, db, sp, 0, 1, bm, X, st));
  CU_TRY(cudaEventRecord(eng->ev1, st));
  const size_t w = (len_a + 1) * 4;
  int32_t *dst[3] = {match, gap_a, gap_b};
  int32_t *src[3] = {X.mat_m, X.mat_ga, X.mat_gb};
  for(int k = 0; k < 3; k++)
    CU_TRY(cudaMemcpy2DAsync(dst[k], w, src[k] + 3, (size_t)pitch * 4, w, len_b + 1, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  float ms = 0;
  CU_TRY(cudaEventElapsedTime(&ms, eng->ev0, eng->ev1));
  eng->last_ms = ms;
  eng->last_kernel = "general_mats";
  return 0;
}

int seqalign_batch_matrices(seqalign_batch_t *eng, size_t i, int32_t *match, int32_t *gap_a, int32_t *gap_b)
{
  if(!eng || !match || !gap_a || !gap_b) return SEQALIGN_ERR_ARG;
  if(eng->mode != SEQALIGN_MODE_MATS || i >= eng->n) return fail(eng, SEQALIGN_ERR_ARG, "no matrices for this index");
  CU_TRY(cudaSetDevice(eng->device));
  const size_t cells = (size_t)(eng->mat_off[i + 1] - eng->mat_off[i]) / 3;
  /* the wave holding pair i: resident, or made resident by running the kernel over it again (the inputs of
   * the submit are still on the device; 12 bytes per cell at HBM speed) */
  int w = eng->mat_resident;
  if(w < 0 || i < eng->mat_wave[w] || i >= eng->mat_wave[w + 1]) {
    w = (int)(std::upper_bound(eng->mat_wave.begin(), eng->mat_wave.end(), i) - eng->mat_wave.begin()) - 1;
    TRY(run_mats_wave(eng, w, eng->stream));
    CU_TRY(cudaStreamSynchronize(eng->stream));
  }
  const int32_t *src = (const int32_t *)eng->d_mats.p + (eng->mat_off[i] - eng->mat_off[eng->mat_wave[w]]);
  CU_TRY(cudaMemcpy(match, src, cells * 4, cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(gap_a, src + cells, cells * 4, cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(gap_b, src + 2 * cells, cells * 4, cudaMemcpyDeviceToHost));
  return 0;
}

int seqalign_batch_set_hit_limits(seqalign_batch_t *eng, size_t max_hits, int32_t min_score)
{
  if(!eng || max_hits < 1 || max_hits > 4096) return SEQALIGN_ERR_ARG;
  eng->hit_max = (int32_t)max_hits;
  eng->hit_min_score = min_score;
  return 0;
}

size_t seqalign_batch_hit_count(const seqalign_batch_t *eng, size_t i)
{
  if(!eng || eng->mode != SEQALIGN_MODE_HITS || i >= eng->n) return 0;
  return (size_t)eng->nhits[i];
}

int seqalign_batch_hit(seqalign_batch_t *eng, size_t i, size_t h, alignment_t *out)
{
  if(!eng || !out) return SEQALIGN_ERR_ARG;
  if(eng->mode != SEQALIGN_MODE_HITS || i >= eng->n) return fail(eng, SEQALIGN_ERR_ARG, "no hits for this index");
  if(h >= (size_t)eng->nhits[i]) return 0;
  const int32_t *r = &eng->hit_rec[(i * (size_t)eng->hit_max_used + h) * 8];
  const size_t len = (size_t)r[5];
  if(out->capacity < len + 1) {
    size_t cap = ROUNDUP2POW(len + 1);
    out->result_a = (char *)realloc(out->result_a, cap);
    out->result_b = (char *)realloc(out->result_b, cap);
    out->capacity = cap;
    if(!out->result_a || !out->result_b) return fail(eng, SEQALIGN_ERR_NOMEM, "Out of memory");
  }
  const size_t per = (size_t)(eng->hit_off[i + 1] - eng->hit_off[i]) / (size_t)eng->hit_max_used;
  const size_t base = (size_t)eng->hit_off[i] + h * per + (size_t)r[6];
  memcpy(out->result_a, eng->hit_a_p + base, len);
  memcpy(out->result_b, eng->hit_b_p + base, len);
  out->result_a[len] = out->result_b[len] = '\0';
  out->length = len;
  out->score = r[0];
  out->pos_a = (size_t)r[1]; out->pos_b = (size_t)r[2];
  out->len_a = (size_t)r[3]; out->len_b = (size_t)r[4];
  return 1;
}

void seqalign_batch_speculation_stats(const seqalign_batch_t *eng, int *hits, int *misses)
{
  if(!eng) return;
  if(hits) *hits = eng->spec_hits;
  if(misses) *misses = eng->spec_misses;
}

void seqalign_batch_unknown_pair(const seqalign_batch_t *eng, char *a, char *b)
{
  if(!eng) return;
  if(a) *a = eng->unk_a;
  if(b) *b = eng->unk_b;
}

double seqalign_batch_last_kernel_ms(const seqalign_batch_t *eng) { return eng ? eng->last_ms : 0; }
double seqalign_batch_last_walk_ms(const seqalign_batch_t *eng) { return eng ? eng->last_walk_ms : 0; }
int seqalign_batch_last_launches(const seqalign_batch_t *eng) { return eng ? eng->last_launches : 0; }
const char *seqalign_batch_last_kernel(const seqalign_batch_t *eng) { return eng ? eng->last_kernel : "none"; }

} /* extern "C" */
