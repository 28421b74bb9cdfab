/*
 * sa_batch.h -- read-ahead buffer of sequence pairs shared by the batching
 * tools.  Replaces the one-pair-at-a-time callback loop of the reference
 * (src/alignment_cmdline.c:611-622): pairs accumulate here and are handed to
 * the engine as one batch.
 */
#ifndef SA_BATCH_H
#define SA_BATCH_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <zlib.h>
#include "sa_cli.h"
#include "seqalign_b200.h"

/* SEQALIGN_CLI_TIMING=1: seconds spent reading, aligning and printing, on stderr at exit */
static double sa_t_read = 0, sa_t_align = 0, sa_t_print = 0, sa_t_init = 0, sa_t_start = 0;
static inline double sa_now(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static inline void sa_timing_report(void)
{
  const char *e = getenv("SEQALIGN_CLI_TIMING");
  if(e && e[0] == '1')
    fprintf(stderr, "timing: init %.3f s, read %.3f s, align %.3f s, print %.3f s, total %.3f s\n", sa_t_init, sa_t_read,
            sa_t_align, sa_t_print, sa_now() - sa_t_start);
}

/* ---- stdout of the batch loop ----------------------------------------------------------------
 * A pair's lines are put together in one buffer by memcpy and hand-written integer formatting and
 * handed to stdio a megabyte at a time: through printf / fputs / putc the print phase was the
 * longest phase of a large run (1 us per pair, 2 s for 2 M pairs, the alignment itself 0.5 s).
 * Anything else that writes to stdout (the library's matrix, colour and spacer printers, the
 * interactive prompt) is preceded by sa_out_sync(); exit() reaches it through atexit(). */
static char *sa_ob = NULL;
static size_t sa_ob_n = 0, sa_ob_cap = 0;
static inline void sa_out_sync(void)
{
  if(sa_ob_n) fwrite(sa_ob, 1, sa_ob_n, stdout);
  sa_ob_n = 0;
}
static inline void sa_out_flush(void) { sa_out_sync(); fflush(stdout); }
static inline char *sa_out_room(size_t n)
{
  if(sa_ob_n + n > sa_ob_cap) {
    sa_out_sync();
    if(n > sa_ob_cap) {
      /* SEQALIGN_CLI_OUTBUF=<bytes>: first size of the buffer (the tests replay every recorded invocation with 16) */
      const char *env = sa_ob_cap ? NULL : getenv("SEQALIGN_CLI_OUTBUF");
      size_t cap = sa_ob_cap ? sa_ob_cap : (env && atol(env) > 0 ? (size_t)atol(env) : ((size_t)1 << 20));
      while(cap < n) cap *= 2;
      char *nb = (char *)realloc(sa_ob, cap);
      if(!nb) { fprintf(stderr, "Error: Out of memory\n"); exit(EXIT_FAILURE); }
      sa_ob = nb; sa_ob_cap = cap;
    }
  }
  return sa_ob + sa_ob_n;
}
static inline void sa_out_mem(const char *s, size_t n) { memcpy(sa_out_room(n), s, n); sa_ob_n += n; }
static inline void sa_out_str(const char *s) { sa_out_mem(s, strlen(s)); }
#define sa_out_lit(s) sa_out_mem("" s, sizeof(s) - 1)
static inline void sa_out_chr(char c) { *sa_out_room(1) = c; sa_ob_n++; }
static inline void sa_out_fill(char c, size_t n) { if(n) { memset(sa_out_room(n), c, n); sa_ob_n += n; } }
static inline void sa_out_ulong(unsigned long v)
{
  char tmp[24];
  int k = 24;
  do { tmp[--k] = (char)('0' + v % 10); v /= 10; } while(v);
  sa_out_mem(tmp + k, (size_t)(24 - k));
}
static inline void sa_out_long(long v)
{
  if(v < 0) { sa_out_chr('-'); sa_out_ulong(0ul - (unsigned long)v); }
  else sa_out_ulong((unsigned long)v);
}
static inline void sa_out_init(void) { sa_out_room(1); atexit(sa_out_sync); }

/* The engine (CUDA context, streams) is created on a helper thread while the main thread parses
 * options, opens the first input and inflates its first chunk: context creation is the largest fixed
 * cost of a run (0.4 s and more).  sa_engine_wait() joins it before the first use. */
static pthread_t sa_eng_thread;
static int sa_eng_pending = 0;
static seqalign_batch_t **sa_eng_slot = NULL;
static const scoring_t *sa_eng_scoring = NULL;
/* --gpus N (N > 1): the batches of the main engine are cut over N devices by seqalign_multi_* (one
 * engine + one host thread per device, every device fed over its own PCIe link); the engine on device
 * 0 stays for decoding and for the paths that go pair by pair */
static int sa_gpus = 1;
static seqalign_multi_t *sa_multi = NULL;
static void *sa_engine_main(void *arg)
{
  (void)arg;
  const double t0 = sa_now();
  *sa_eng_slot = seqalign_batch_create(0);
  if(*sa_eng_slot && seqalign_batch_set_scoring(*sa_eng_slot, sa_eng_scoring) != SEQALIGN_OK) {
    fprintf(stderr, "Error: %s\n", seqalign_batch_error(*sa_eng_slot));
    exit(EXIT_FAILURE);
  }
  if(*sa_eng_slot && sa_gpus != 1) {
    /* SEQALIGN_CLI_DEVICES=0,0,1: an explicit device list (a device may appear twice: two engines on it) */
    int devs[64], nd = 0;
    const char *list = getenv("SEQALIGN_CLI_DEVICES");
    for(const char *q = list; q && *q && nd < 64;) {
      devs[nd++] = atoi(q);
      q = strchr(q, ',');
      if(q) q++;
    }
    sa_multi = nd ? seqalign_multi_create(devs, nd) : seqalign_multi_create(NULL, sa_gpus);   /* <= 0: every usable device */
    if(!sa_multi || seqalign_multi_set_scoring(sa_multi, sa_eng_scoring) != SEQALIGN_OK) {
      fprintf(stderr, "Error: --gpus: %s\n", sa_multi ? seqalign_multi_error(sa_multi) : seqalign_last_create_error());
      exit(EXIT_FAILURE);
    }
  }
  sa_t_init = sa_now() - t0;
  return NULL;
}
static inline void sa_engine_start(seqalign_batch_t **slot, const scoring_t *scoring)
{
  sa_eng_slot = slot; sa_eng_scoring = scoring;
  if(pthread_create(&sa_eng_thread, NULL, sa_engine_main, NULL) == 0) sa_eng_pending = 1;
  else sa_engine_main(NULL);
}
static inline void sa_engine_wait(void)
{
  if(sa_eng_pending) { pthread_join(sa_eng_thread, NULL); sa_eng_pending = 0; }
  if(sa_eng_slot && !*sa_eng_slot) { fprintf(stderr, "Error: %s\n", seqalign_last_create_error()); exit(EXIT_FAILURE); }
}

typedef struct {
  char **a, **b, **name_a, **name_b; /* owned copies; names NULL when the record had none */
  size_t *la, *lb;
  size_t n, cap, bytes;
  /* a batch decoded on the device (sa_for_each_batch_dev): the sequences lie in HBM as records
   * [dev_first, dev_first + n) of two decoded sides; a[] / b[] stay NULL until sa_pairs_host() */
  const seqalign_reads_t *dev_a, *dev_b;
  int dev_side_a, dev_side_b;
  size_t dev_first;
  struct sa_dev_chunk *chunk;
} sa_pairs;

/* host copies of a decoded chunk's packed sides, fetched at most once per chunk and only when a
 * tool has to print or re-align the sequences themselves */
struct sa_dev_chunk {
  seqalign_reads_t *ra, *rb;
  char *host_a, *host_b;
  size_t cap_a, cap_b;
  size_t text_a, text_b;   /* bytes of text the sides were decoded from: no side is longer */
  int have_a, have_b;
};

static inline char *sa_dup(const char *s, size_t n)
{
  char *d = malloc(n + 1);
  if(!d) { fprintf(stderr, "%s:%i: Out of memory\n", __FILE__, __LINE__); exit(EXIT_FAILURE); }
  memcpy(d, s, n);
  d[n] = '\0';
  return d;
}

static inline void sa_pairs_reserve(sa_pairs *p, size_t n);

static inline void sa_pairs_add(sa_pairs *p, const sa_record *r1, const sa_record *r2)
{
  sa_pairs_reserve(p, p->n + 1);
  const size_t i = p->n++;
  p->a[i] = sa_dup(r1->seq.b, r1->seq.len); p->la[i] = r1->seq.len;
  p->b[i] = sa_dup(r2->seq.b, r2->seq.len); p->lb[i] = r2->seq.len;
  p->name_a[i] = r1->name.len ? sa_dup(r1->name.b, r1->name.len) : NULL;
  p->name_b[i] = r2->name.len ? sa_dup(r2->name.b, r2->name.len) : NULL;
  p->bytes += r1->seq.len + r2->seq.len;
}

static inline void sa_pairs_clear(sa_pairs *p)
{
  for(size_t i = 0; i < p->n; i++) { free(p->a[i]); free(p->b[i]); free(p->name_a[i]); free(p->name_b[i]); }
  p->n = 0; p->bytes = 0;
  p->dev_a = p->dev_b = NULL; p->chunk = NULL;
}

static inline void sa_pairs_reserve(sa_pairs *p, size_t n)
{
  if(n <= p->cap) return;
  size_t cap = p->cap ? p->cap : 1024;
  while(cap < n) cap *= 2;
  const size_t old = p->cap;
  p->cap = cap;
  p->a = realloc(p->a, p->cap * sizeof(char *)); p->b = realloc(p->b, p->cap * sizeof(char *));
  p->name_a = realloc(p->name_a, p->cap * sizeof(char *)); p->name_b = realloc(p->name_b, p->cap * sizeof(char *));
  p->la = realloc(p->la, p->cap * sizeof(size_t)); p->lb = realloc(p->lb, p->cap * sizeof(size_t));
  if(!p->a || !p->b || !p->name_a || !p->name_b || !p->la || !p->lb) {
    fprintf(stderr, "%s:%i: Out of memory\n", __FILE__, __LINE__); exit(EXIT_FAILURE);
  }
  for(size_t i = old; i < cap; i++) p->a[i] = p->b[i] = p->name_a[i] = p->name_b[i] = NULL;
}

/* one submit for both kinds of batch */
static inline int sa_submit(seqalign_batch_t *eng, int algo, int mode, const sa_pairs *p)
{
  if(p->dev_a)
    return seqalign_batch_submit_reads(eng, algo, mode, p->dev_a, p->dev_side_a, p->dev_b, p->dev_side_b, p->dev_first, p->n);
  return seqalign_batch_submit(eng, algo, mode, (const char *const *)p->a, p->la, (const char *const *)p->b, p->lb, p->n);
}

/* NUL-terminated host copies of a device-decoded batch's sequences (no-op for host batches) */
static inline void sa_pairs_host(sa_pairs *p)
{
  if(!p->dev_a || p->n == 0 || p->a[0]) return;
  struct sa_dev_chunk *c = p->chunk;
  for(int side = 0; side < 2; side++) {
    int *have = side ? &c->have_b : &c->have_a;
    char **host = side ? &c->host_b : &c->host_a;
    size_t *cap = side ? &c->cap_b : &c->cap_a;
    const seqalign_reads_t *r = side ? p->dev_b : p->dev_a;
    const int rs = side ? p->dev_side_b : p->dev_side_a;
    if(!*have) {
      const size_t need = (side ? c->text_b : c->text_a) + 64;
      if(*cap < need) { free(*host); *host = malloc(need); *cap = need; }
      if(!*host || seqalign_reads_fetch((seqalign_reads_t *)r, rs, *host) != SEQALIGN_OK) {
        fprintf(stderr, "Error: %s\n", seqalign_reads_error(r)); exit(EXIT_FAILURE);
      }
      *have = 1;
    }
    const int64_t *off = seqalign_reads_offsets(r, rs) + p->dev_first;
    for(size_t i = 0; i < p->n; i++) {
      char *d = sa_dup(*host + off[i], (size_t)(off[i + 1] - off[i]));
      if(side) p->b[i] = d; else p->a[i] = d;
    }
  }
}

static inline void sa_pairs_free(sa_pairs *p)
{
  sa_pairs_clear(p);
  free(p->a); free(p->b); free(p->name_a); free(p->name_b); free(p->la); free(p->lb);
  memset(p, 0, sizeof(*p));
}

/* ---- the tool's main engine: one device, or --gpus N through seqalign_multi_* -------------- */
static char *sa_pack_a = NULL, *sa_pack_b = NULL;
static int64_t *sa_pack_oa = NULL, *sa_pack_ob = NULL;
static size_t sa_pack_cap = 0, sa_pack_ncap = 0;

static inline int sa_main_submit(int algo, int mode, sa_pairs *p)
{
  if(!sa_multi) return sa_submit(*sa_eng_slot, algo, mode, p);
  /* packed host arrays for the multi-device call: a device-decoded chunk comes back to the host once
   * (every device then pulls its own range of it), host batches are packed here */
  if(p->dev_a) {
    struct sa_dev_chunk *c = p->chunk;
    for(int side = 0; side < 2; side++) {
      int *have = side ? &c->have_b : &c->have_a;
      char **host = side ? &c->host_b : &c->host_a;
      size_t *cap = side ? &c->cap_b : &c->cap_a;
      const seqalign_reads_t *r = side ? p->dev_b : p->dev_a;
      if(*have) continue;
      const size_t need = (side ? c->text_b : c->text_a) + 64;
      if(*cap < need) { free(*host); *host = malloc(need); *cap = need; }
      if(!*host || seqalign_reads_fetch((seqalign_reads_t *)r, side ? p->dev_side_b : p->dev_side_a, *host) != SEQALIGN_OK) {
        fprintf(stderr, "Error: %s\n", seqalign_reads_error(r)); exit(EXIT_FAILURE);
      }
      *have = 1;
    }
    return seqalign_multi_submit_packed(sa_multi, algo, mode, c->host_a, seqalign_reads_offsets(p->dev_a, p->dev_side_a) + p->dev_first,
                                        c->host_b, seqalign_reads_offsets(p->dev_b, p->dev_side_b) + p->dev_first, p->n);
  }
  size_t ta = 0, tb = 0;
  for(size_t i = 0; i < p->n; i++) { ta += p->la[i]; tb += p->lb[i]; }
  if(ta + tb + 64 > sa_pack_cap) {
    sa_pack_cap = (ta + tb + 64) * 2;
    free(sa_pack_a); free(sa_pack_b);
    sa_pack_a = malloc(sa_pack_cap); sa_pack_b = malloc(sa_pack_cap);
  }
  if(p->n + 1 > sa_pack_ncap) {
    sa_pack_ncap = (p->n + 1) * 2;
    free(sa_pack_oa); free(sa_pack_ob);
    sa_pack_oa = malloc(sa_pack_ncap * sizeof(int64_t)); sa_pack_ob = malloc(sa_pack_ncap * sizeof(int64_t));
  }
  if(!sa_pack_a || !sa_pack_b || !sa_pack_oa || !sa_pack_ob) { fprintf(stderr, "%s:%i: Out of memory\n", __FILE__, __LINE__); exit(EXIT_FAILURE); }
  ta = tb = 0;
  for(size_t i = 0; i < p->n; i++) {
    sa_pack_oa[i] = (int64_t)ta; sa_pack_ob[i] = (int64_t)tb;
    memcpy(sa_pack_a + ta, p->a[i], p->la[i]); ta += p->la[i];
    memcpy(sa_pack_b + tb, p->b[i], p->lb[i]); tb += p->lb[i];
  }
  sa_pack_oa[p->n] = (int64_t)ta; sa_pack_ob[p->n] = (int64_t)tb;
  return seqalign_multi_submit_packed(sa_multi, algo, mode, sa_pack_a, sa_pack_oa, sa_pack_b, sa_pack_ob, p->n);
}
static inline int sa_main_alignment(size_t i, alignment_t *out)
{ return sa_multi ? seqalign_multi_alignment(sa_multi, i, out) : seqalign_batch_alignment(*sa_eng_slot, i, out); }
static inline size_t sa_main_hit_count(size_t i)
{ return sa_multi ? seqalign_multi_hit_count(sa_multi, i) : seqalign_batch_hit_count(*sa_eng_slot, i); }
static inline int sa_main_hit(size_t i, size_t h, alignment_t *out)
{ return sa_multi ? seqalign_multi_hit(sa_multi, i, h, out) : seqalign_batch_hit(*sa_eng_slot, i, h, out); }
static inline int sa_main_set_hit_limits(size_t max_hits, int32_t min_score)
{ return sa_multi ? seqalign_multi_set_hit_limits(sa_multi, max_hits, min_score) : seqalign_batch_set_hit_limits(*sa_eng_slot, max_hits, min_score); }
static inline const char *sa_main_error(void)
{ return sa_multi ? seqalign_multi_error(sa_multi) : seqalign_batch_error(*sa_eng_slot); }
static inline void sa_main_destroy(void)
{
  if(sa_multi) seqalign_multi_destroy(sa_multi);
  sa_multi = NULL;
  if(sa_eng_slot && *sa_eng_slot) seqalign_batch_destroy(*sa_eng_slot);
  free(sa_pack_a); free(sa_pack_b); free(sa_pack_oa); free(sa_pack_ob);
}

/* a batch is full at this many pairs or bytes of sequence */
/* 262,144 pairs per submit: the per-submit costs (synchronisations, table checks, small copies) were visible at
 * 65,536 -- 2 M pairs of 150 bp: align phase 0.70-0.98 s against 0.55-0.66 s (profiles/cli_batches_r02p.jsonl) */
#define SA_BATCH_MAX_PAIRS ((size_t)1 << 18)
#define SA_BATCH_MAX_BYTES ((size_t)256 << 20)

/* Read every pair of one input (path2 == NULL: consecutive records of path1)
 * and call flush() whenever the read-ahead buffer is full, and at the end.
 * interactive: flush after every pair (request/response protocol of the perl
 * wrappers, reference perl/NeedlemanWunsch.pm:182-210).  Messages as
 * reference src/alignment_cmdline.c:584-632. */
static inline void sa_for_each_batch_from(const char *path1, const char *path2, int interactive, int buffered,
                                          size_t max_pairs, sa_pairs *pairs, void (*flush)(sa_pairs *, sa_reader *),
                                          unsigned long skip)
{
  sa_reader *r1 = sa_reader_open(path1, buffered), *r2 = r1;
  if(!r1) { fprintf(stderr, "Alignment Error: couldn't open file %s\n", path1); fflush(stderr); return; }
  if(path2 && !(r2 = sa_reader_open(path2, buffered))) {
    fprintf(stderr, "Alignment Error: couldn't open file %s\n", path1); fflush(stderr);
    return;
  }
  sa_record rec1, rec2;
  memset(&rec1, 0, sizeof(rec1)); memset(&rec2, 0, sizeof(rec2));
  unsigned long count = 0;
  double t0 = sa_now();
  for(; sa_reader_next(r1, &rec1) > 0; count++) {
    if(count < skip) {   /* pairs the device path has already handled */
      if(sa_reader_next(r2, &rec2) <= 0) break;
      continue;
    }
    if(sa_reader_next(r2, &rec2) <= 0) {
      flush(pairs, r1);
      fprintf(stderr, "Alignment Error: Odd number of sequences - I read in pairs!\n"); fflush(stderr);
      break;
    }
    sa_pairs_add(pairs, &rec1, &rec2);
    if(interactive || pairs->n >= max_pairs || pairs->bytes >= SA_BATCH_MAX_BYTES) {
      sa_t_read += sa_now() - t0;
      flush(pairs, r1);
      t0 = sa_now();
    }
  }
  sa_t_read += sa_now() - t0;
  flush(pairs, r1);
  if(count == 0) { fprintf(stderr, "Alignment Warning: empty input\n"); fflush(stderr); }
  sa_reader_close(r1);
  if(path2) sa_reader_close(r2);
  sa_str_free(&rec1.name); sa_str_free(&rec1.seq); sa_str_free(&rec2.name); sa_str_free(&rec2.seq);
}

static inline void sa_for_each_batch(const char *path1, const char *path2, int interactive, int buffered,
                                     size_t max_pairs, sa_pairs *pairs, void (*flush)(sa_pairs *, sa_reader *))
{
  sa_for_each_batch_from(path1, path2, interactive, buffered, max_pairs, pairs, flush, 0);
}

/* ---- the same loop with the records decoded on the device ---------------------------------
 * (SURVEY.md 8 f-4).  The file is read (and inflated, zlib) in chunks of text into pinned memory;
 * each chunk goes to the GPU once, seqalign_reads_decode() turns it into packed sides + offsets in
 * HBM, and flush() aligns sub-batches of it in place.  The tail the decoder holds back (a record
 * that may continue in the next chunk) is carried to the front of the buffer.
 * Returns 0 when the input was handled to its end; 1 when the decoder declined a chunk (text
 * outside its grammar) after *done pairs had been flushed: the caller goes on with the host reader,
 * skipping that many pairs; -1 when the path could not be used at all (nothing read, nothing printed). */
typedef struct { gzFile gz; char *buf; size_t len, cap; int eof; seqalign_reads_t *reads; } sa_dev_file;

static inline void sa_dev_fill(sa_dev_file *f)
{
  while(!f->eof && f->len < f->cap) {
    const size_t room = f->cap - f->len;
    const int n = gzread(f->gz, f->buf + f->len, room > ((size_t)1 << 30) ? 1u << 30 : (unsigned)room);
    if(n <= 0) f->eof = 1; else f->len += (size_t)n;
  }
  if(!f->eof) {
    /* is the input exactly at its end?  then this chunk is the last one */
    const int c = gzgetc(f->gz);
    if(c == -1) f->eof = 1; else gzungetc(c, f->gz);
  }
}

static inline int sa_dev_grow(sa_dev_file *f)
{
  const size_t cap = f->cap * 2;
  if(cap > ((size_t)1 << 30)) return 0;
  char *nb = seqalign_host_alloc(cap + 64);
  if(!nb) return 0;
  memcpy(nb, f->buf, f->len);
  seqalign_host_free(f->buf);
  f->buf = nb; f->cap = cap;
  return 1;
}

static inline int sa_for_each_batch_dev(const char *path1, const char *path2, int device, size_t max_pairs, int want_names,
                                        sa_pairs *pairs, void (*flush)(sa_pairs *, sa_reader *), unsigned long *done)
{
  *done = 0;
  const char *env = getenv("SEQALIGN_CLI_CHUNK_MB");
  size_t cap = (size_t)(env && atoi(env) > 0 ? atoi(env) : 64) << 20;
  if(env && atoi(env) < 0) cap = (size_t)(-atoi(env));   /* negative: bytes (tests cut chunks inside records) */
  sa_dev_file f[2];
  memset(f, 0, sizeof(f));
  const int nf = path2 ? 2 : 1;
  int ok = 1;
  for(int k = 0; k < nf && ok; k++) {
    f[k].gz = gzopen(k ? path2 : path1, "r");
    if(f[k].gz) gzbuffer(f[k].gz, 1 << 20);
    ok = f[k].gz != NULL;
  }
  if(ok) sa_engine_wait();   /* the pinned buffers and the reads objects need the context */
  for(int k = 0; k < nf && ok; k++) {
    f[k].cap = cap;
    f[k].buf = seqalign_host_alloc(cap + 64);
    f[k].reads = f[k].buf ? seqalign_reads_create(device) : NULL;
    ok = f[k].buf && f[k].reads;
  }
  int rc = ok ? 0 : -1;
  struct sa_dev_chunk chunk;
  memset(&chunk, 0, sizeof(chunk));
  while(ok) {
    double t0 = sa_now();
    size_t cnt[2] = {0, 0};
    int declined = 0;
    for(int k = 0; k < nf; k++) {
      sa_dev_fill(&f[k]);
      const int drc = seqalign_reads_decode(f[k].reads, f[k].buf, f[k].len, f[k].eof, nf == 1);
      if(drc == SEQALIGN_ERR_IRREGULAR) { declined = 1; break; }
      if(drc != SEQALIGN_OK) { fprintf(stderr, "Error: %s\n", seqalign_reads_error(f[k].reads)); exit(EXIT_FAILURE); }
      cnt[k] = seqalign_reads_records(f[k].reads);
    }
    sa_t_read += sa_now() - t0;
    if(declined) { rc = 1; break; }
    const size_t navail = nf == 1 ? cnt[0] / 2 : (cnt[0] < cnt[1] ? cnt[0] : cnt[1]);
    if(navail == 0) {
      /* not one whole pair in the buffer: more text needed (a record longer than the chunk), or the end */
      int grew = 0, stuck = 0;
      for(int k = 0; k < nf; k++)
        if(!f[k].eof && cnt[k] < (size_t)(nf == 1 ? 2 : 1)) { if(sa_dev_grow(&f[k])) grew = 1; else stuck = 1; }
      if(stuck) { rc = *done == 0 ? -1 : 1; break; }
      if(grew) continue;
    }
    chunk.ra = f[0].reads; chunk.rb = f[nf - 1].reads; chunk.have_a = chunk.have_b = 0;
    chunk.text_a = f[0].len; chunk.text_b = f[nf - 1].len;
    for(size_t first = 0; first < navail; first += max_pairs) {
      const size_t n = navail - first < max_pairs ? navail - first : max_pairs;
      t0 = sa_now();
      sa_pairs_clear(pairs);
      sa_pairs_reserve(pairs, n);
      const int64_t *oa = seqalign_reads_offsets(f[0].reads, 0) + first;
      const int64_t *ob = seqalign_reads_offsets(f[nf - 1].reads, nf == 1 ? 1 : 0) + first;
      for(size_t i = 0; i < n; i++) {
        pairs->la[i] = (size_t)(oa[i + 1] - oa[i]);
        pairs->lb[i] = (size_t)(ob[i + 1] - ob[i]);
        pairs->a[i] = pairs->b[i] = pairs->name_a[i] = pairs->name_b[i] = NULL;
        if(want_names) {
          size_t pos = 0, len = 0;
          seqalign_reads_name(f[0].reads, nf == 1 ? 2 * (first + i) : first + i, &pos, &len);
          if(len) pairs->name_a[i] = sa_dup(f[0].buf + pos, len);
          seqalign_reads_name(f[nf - 1].reads, nf == 1 ? 2 * (first + i) + 1 : first + i, &pos, &len);
          if(len) pairs->name_b[i] = sa_dup(f[nf - 1].buf + pos, len);
        }
      }
      pairs->n = n;
      pairs->dev_a = f[0].reads; pairs->dev_side_a = 0;
      pairs->dev_b = f[nf - 1].reads; pairs->dev_side_b = nf == 1 ? 1 : 0;
      pairs->dev_first = first;
      pairs->chunk = &chunk;
      sa_t_read += sa_now() - t0;
      flush(pairs, NULL);
      sa_pairs_clear(pairs);
    }
    *done += navail;
    /* what is left of each buffer goes to its front */
    for(int k = 0; k < nf; k++) {
      /* record_start(i) is where record i starts, or (i = number of complete records) the held-back tail */
      size_t cut = seqalign_reads_record_start(f[k].reads, nf == 1 ? 2 * navail : navail);
      if(cut > f[k].len) cut = f[k].len;
      memmove(f[k].buf, f[k].buf + cut, f[k].len - cut);
      f[k].len -= cut;
    }
    /* end of the input (messages as reference src/alignment_cmdline.c:613-620) */
    if(nf == 1) {
      if(f[0].eof) {
        if(cnt[0] & 1) { fprintf(stderr, "Alignment Error: Odd number of sequences - I read in pairs!\n"); fflush(stderr); }
        break;
      }
    } else {
      if(f[0].eof && cnt[0] == navail) break;                 /* first file exhausted: done, whatever the second still holds */
      if(f[1].eof && cnt[1] == navail && cnt[0] > navail) {   /* a record of the first file without a partner */
        fprintf(stderr, "Alignment Error: Odd number of sequences - I read in pairs!\n"); fflush(stderr);
        break;
      }
    }
  }
  if(rc == 0 && *done == 0) { fprintf(stderr, "Alignment Warning: empty input\n"); fflush(stderr); }
  for(int k = 0; k < nf; k++) {
    if(f[k].reads) seqalign_reads_destroy(f[k].reads);
    if(f[k].buf) seqalign_host_free(f[k].buf);
    if(f[k].gz) gzclose(f[k].gz);
  }
  free(chunk.host_a); free(chunk.host_b);
  return rc;
}

/* One input of a tool: regular files go through the device decoder (SEQALIGN_CLI_DECODE=host keeps them
 * on the host reader); stdin and anything the decoder declines go through the host reader. */
static inline void sa_read_input(const char *path1, const char *path2, int interactive, int device, size_t max_pairs,
                                 int want_names, sa_pairs *pairs, void (*flush)(sa_pairs *, sa_reader *))
{
  const char *bp = getenv("SEQALIGN_CLI_BATCH_PAIRS");   /* pairs per engine submit (default: the tool's own) */
  if(bp && atol(bp) > 0) max_pairs = (size_t)atol(bp);
  const char *mode = getenv("SEQALIGN_CLI_DECODE");
  const int dev_ok = !interactive && strcmp(path1, "-") != 0 && (!path2 || strcmp(path2, "-") != 0) &&
                     !(mode && strcmp(mode, "host") == 0);
  unsigned long done = 0;
  if(dev_ok) {
    const int rc = sa_for_each_batch_dev(path1, path2, device, max_pairs, want_names, pairs, flush, &done);
    if(rc == 0) return;
    if(rc < 0) done = 0;
    if(mode && strcmp(mode, "device") == 0) { fprintf(stderr, "Error: the device decoder declined this input\n"); exit(EXIT_FAILURE); }
  }
  sa_for_each_batch_from(path1, path2, interactive, !interactive, max_pairs, pairs, flush, done);
}

#endif
