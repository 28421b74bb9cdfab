/*
 * sa_multi.c -- one batch over several GPUs of one node, from plain C.
 *
 * Implements the seqalign_multi_* entry points of include/seqalign_b200.h:
 * one engine per device, the batch cut into contiguous pair ranges balanced
 * by cell count, one host thread per device for the duration of a submit.
 * Pairs are independent (the reference loops over them one by one,
 * src/alignment_cmdline.c:611-622), so nothing is exchanged between devices:
 * every device pulls its range over its own PCIe link and leaves its scores
 * directly in the caller-visible result arrays (seqalign_batch_set_result_sink),
 * alignments and hit lists stay with the engine that made them and are
 * fetched by global pair index.  No NCCL: one process, N devices.  (One
 * process PER device, the torch.distributed shape, is seqalign.distributed.)
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "seqalign_b200.h"

typedef struct {
  struct seqalign_multi *m;
  int index, device;
  seqalign_batch_t *eng;
  pthread_t th;
  int started;
  int rc;
  size_t first, count;        /* pair range of the last submit */
  int64_t *off_a, *off_b;     /* rebased offsets of the range (packed submits) */
  size_t off_cap;
} sa_worker;

struct seqalign_multi {
  int n;
  sa_worker *w;
  /* job being submitted */
  int algo, mode;
  const char *seq_a, *seq_b;
  const int64_t *in_off_a, *in_off_b;   /* NULL: uniform */
  size_t ula, ulb;
  /* results of the last submit */
  size_t total;
  int32_t *score, *xend, *yend;
  size_t res_cap;
  char err[640];
};

static int score_mode(int mode) { return mode == SEQALIGN_MODE_SCORE || mode == SEQALIGN_MODE_SCORE_ONLY; }

static void *worker_main(void *arg)
{
  sa_worker *w = (sa_worker *)arg;
  struct seqalign_multi *m = w->m;
  w->rc = 0;
  if(w->count == 0 && !score_mode(m->mode)) {
    /* an empty submit still resets the engine's "last batch" */
    w->rc = seqalign_batch_submit_packed(w->eng, m->algo, m->mode, NULL, NULL, NULL, NULL, 0);
    return NULL;
  }
  if(m->mode == SEQALIGN_MODE_SCORE)
    seqalign_batch_set_result_sink(w->eng, m->score + w->first, m->xend + w->first, m->yend + w->first);
  else if(m->mode == SEQALIGN_MODE_SCORE_ONLY)
    seqalign_batch_set_result_sink(w->eng, m->score + w->first, NULL, NULL);
  if(!m->in_off_a) {
    w->rc = seqalign_batch_submit_uniform(w->eng, m->algo, m->mode, m->seq_a + w->first * m->ula, m->ula,
                                          m->seq_b + w->first * m->ulb, m->ulb, w->count);
  } else {
    if(w->off_cap < w->count + 1) {
      free(w->off_a); free(w->off_b);
      w->off_cap = w->count + 1 + w->count / 4;
      w->off_a = malloc(w->off_cap * sizeof(int64_t));
      w->off_b = malloc(w->off_cap * sizeof(int64_t));
      if(!w->off_a || !w->off_b) { w->off_cap = 0; w->rc = SEQALIGN_ERR_NOMEM; return NULL; }
    }
    const int64_t a0 = m->in_off_a[w->first], b0 = m->in_off_b[w->first];
    for(size_t i = 0; i <= w->count; i++) {
      w->off_a[i] = m->in_off_a[w->first + i] - a0;
      w->off_b[i] = m->in_off_b[w->first + i] - b0;
    }
    w->rc = seqalign_batch_submit_packed(w->eng, m->algo, m->mode, m->seq_a + a0, w->off_a, m->seq_b + b0, w->off_b, w->count);
  }
  if(score_mode(m->mode)) seqalign_batch_set_result_sink(w->eng, NULL, NULL, NULL);
  else if(w->rc == 0 && w->count) w->rc = seqalign_batch_ends(w->eng, m->score + w->first, m->xend + w->first, m->yend + w->first);
  return NULL;
}

seqalign_multi_t *seqalign_multi_create(const int *devices, int n_devices)
{
  if(n_devices <= 0) n_devices = seqalign_device_count();
  if(n_devices <= 0) return NULL;   /* message: seqalign_last_create_error() of the first engine attempt below */
  struct seqalign_multi *m = calloc(1, sizeof(*m));
  if(!m) return NULL;
  m->w = calloc((size_t)n_devices, sizeof(sa_worker));
  if(!m->w) { free(m); return NULL; }
  m->n = n_devices;
  for(int i = 0; i < n_devices; i++) {
    sa_worker *w = &m->w[i];
    w->m = m; w->index = i; w->device = devices ? devices[i] : i;
    w->eng = seqalign_batch_create(w->device);
    if(!w->eng) { seqalign_multi_destroy(m); return NULL; }
  }
  return m;
}

void seqalign_multi_destroy(seqalign_multi_t *m)
{
  if(!m) return;
  for(int i = 0; i < m->n; i++) {
    if(m->w[i].eng) seqalign_batch_destroy(m->w[i].eng);
    free(m->w[i].off_a); free(m->w[i].off_b);
  }
  free(m->w);
  seqalign_host_free(m->score); seqalign_host_free(m->xend); seqalign_host_free(m->yend);
  free(m);
}

int seqalign_multi_devices(const seqalign_multi_t *m) { return m ? m->n : 0; }
const char *seqalign_multi_error(const seqalign_multi_t *m) { return m ? m->err : "null handle"; }

int seqalign_multi_set_scoring(seqalign_multi_t *m, const scoring_t *scoring)
{
  if(!m || !scoring) return SEQALIGN_ERR_ARG;
  for(int i = 0; i < m->n; i++) {
    const int rc = seqalign_batch_set_scoring(m->w[i].eng, scoring);
    if(rc != 0) return rc;
  }
  return 0;
}

int seqalign_multi_set_hit_limits(seqalign_multi_t *m, size_t max_hits, int32_t min_score)
{
  if(!m) return SEQALIGN_ERR_ARG;
  for(int i = 0; i < m->n; i++) {
    const int rc = seqalign_batch_set_hit_limits(m->w[i].eng, max_hits, min_score);
    if(rc != 0) return rc;
  }
  return 0;
}

/* contiguous ranges with near-equal sum(len_a*len_b + 1) */
static void cut_ranges(struct seqalign_multi *m, size_t n)
{
  const int nw = m->n;
  size_t *bounds = calloc((size_t)nw + 1, sizeof(size_t));
  if(!m->in_off_a) {
    for(int r = 0; r <= nw; r++) bounds[r] = n * (size_t)r / (size_t)nw;
  } else {
    long double total = 0;
    for(size_t i = 0; i < n; i++)
      total += (long double)(m->in_off_a[i + 1] - m->in_off_a[i]) * (long double)(m->in_off_b[i + 1] - m->in_off_b[i]) + 1;
    long double acc = 0;
    int r = 1;
    for(size_t i = 0; i < n && r < nw; i++) {
      acc += (long double)(m->in_off_a[i + 1] - m->in_off_a[i]) * (long double)(m->in_off_b[i + 1] - m->in_off_b[i]) + 1;
      while(r < nw && acc > total * r / nw) bounds[r++] = i;   /* pair i is the first past target r */
    }
    while(r < nw) bounds[r++] = n;
    bounds[nw] = n;
    for(int k = 1; k <= nw; k++) if(bounds[k] < bounds[k - 1]) bounds[k] = bounds[k - 1];
  }
  for(int r = 0; r < nw; r++) { m->w[r].first = bounds[r]; m->w[r].count = bounds[r + 1] - bounds[r]; }
  free(bounds);
}

static int run_job(struct seqalign_multi *m, size_t n)
{
  m->err[0] = '\0';
  m->total = 0;
  if(m->res_cap < n + 1) {
    /* page-locked: the engines' device->host copies write the scores here directly */
    seqalign_host_free(m->score); seqalign_host_free(m->xend); seqalign_host_free(m->yend);
    m->res_cap = n + 1 + n / 4;
    m->score = seqalign_host_alloc(m->res_cap * sizeof(int32_t));
    m->xend = seqalign_host_alloc(m->res_cap * sizeof(int32_t));
    m->yend = seqalign_host_alloc(m->res_cap * sizeof(int32_t));
    if(!m->score || !m->xend || !m->yend) { m->res_cap = 0; snprintf(m->err, sizeof(m->err), "Out of memory"); return SEQALIGN_ERR_NOMEM; }
  }
  cut_ranges(m, n);
  for(int i = 0; i < m->n; i++) {
    sa_worker *w = &m->w[i];
    w->started = 0;
    if(w->count == 0 && score_mode(m->mode)) { w->rc = 0; continue; }
    if(m->n == 1 || pthread_create(&w->th, NULL, worker_main, w) != 0) worker_main(w);   /* no thread: run it here */
    else w->started = 1;
  }
  int rc = 0;
  for(int i = 0; i < m->n; i++) {
    sa_worker *w = &m->w[i];
    if(w->started) pthread_join(w->th, NULL);
    if(w->rc != 0 && rc == 0) {
      /* the first failing range in pair order is the one the reference would have met first */
      rc = w->rc;
      snprintf(m->err, sizeof(m->err), "%s", seqalign_batch_error(w->eng));
    }
  }
  if(rc == 0) m->total = n;
  return rc;
}

int seqalign_multi_submit_packed(seqalign_multi_t *m, int algo, int mode, const char *seq_a, const int64_t *off_a,
                                 const char *seq_b, const int64_t *off_b, size_t n)
{
  if(!m || (n > 0 && (!off_a || !off_b))) return SEQALIGN_ERR_ARG;
  m->algo = algo; m->mode = mode; m->seq_a = seq_a; m->seq_b = seq_b; m->in_off_a = off_a; m->in_off_b = off_b;
  static const int64_t zero[1] = {0};
  if(n == 0) { m->in_off_a = zero; m->in_off_b = zero; }
  return run_job(m, n);
}

int seqalign_multi_submit_uniform(seqalign_multi_t *m, int algo, int mode, const char *seq_a, size_t len_a,
                                  const char *seq_b, size_t len_b, size_t n)
{
  if(!m) return SEQALIGN_ERR_ARG;
  m->algo = algo; m->mode = mode; m->seq_a = seq_a; m->seq_b = seq_b; m->in_off_a = m->in_off_b = NULL;
  m->ula = len_a; m->ulb = len_b;
  return run_job(m, n);
}

size_t seqalign_multi_size(const seqalign_multi_t *m) { return m ? m->total : 0; }

int seqalign_multi_ends(seqalign_multi_t *m, int32_t *score, int32_t *x_end, int32_t *y_end)
{
  if(!m) return SEQALIGN_ERR_ARG;
  if(score) memcpy(score, m->score, m->total * sizeof(int32_t));
  if(m->mode == SEQALIGN_MODE_SCORE_ONLY) {
    /* no end cells were computed: they read like the single-engine call's (0,0 for SW, the lengths for NW) */
    for(size_t i = 0; i < m->total; i++) {
      const int nw = m->algo == SEQALIGN_NW;
      if(x_end) x_end[i] = !nw ? 0 : (int32_t)(m->in_off_a ? m->in_off_a[i + 1] - m->in_off_a[i] : (int64_t)m->ula);
      if(y_end) y_end[i] = !nw ? 0 : (int32_t)(m->in_off_b ? m->in_off_b[i + 1] - m->in_off_b[i] : (int64_t)m->ulb);
    }
    return 0;
  }
  if(x_end) memcpy(x_end, m->xend, m->total * sizeof(int32_t));
  if(y_end) memcpy(y_end, m->yend, m->total * sizeof(int32_t));
  return 0;
}

int seqalign_multi_scores(seqalign_multi_t *m, int32_t *score) { return seqalign_multi_ends(m, score, NULL, NULL); }

/* engine and local index that hold pair i of the last submit */
static sa_worker *locate(seqalign_multi_t *m, size_t i, size_t *local)
{
  if(!m || i >= m->total) return NULL;
  for(int r = 0; r < m->n; r++)
    if(i >= m->w[r].first && i < m->w[r].first + m->w[r].count) { *local = i - m->w[r].first; return &m->w[r]; }
  return NULL;
}

int seqalign_multi_where(seqalign_multi_t *m, size_t i, int *device, size_t *local_index)
{
  size_t l = 0;
  sa_worker *w = locate(m, i, &l);
  if(!w) return SEQALIGN_ERR_ARG;
  if(device) *device = w->device;
  if(local_index) *local_index = l;
  return 0;
}

int seqalign_multi_alignment(seqalign_multi_t *m, size_t i, alignment_t *out)
{
  size_t l = 0;
  sa_worker *w = locate(m, i, &l);
  if(!w) return SEQALIGN_ERR_ARG;
  const int rc = seqalign_batch_alignment(w->eng, l, out);
  if(rc < 0) snprintf(m->err, sizeof(m->err), "%s", seqalign_batch_error(w->eng));
  return rc;
}

size_t seqalign_multi_hit_count(seqalign_multi_t *m, size_t i)
{
  size_t l = 0;
  sa_worker *w = locate(m, i, &l);
  return w ? seqalign_batch_hit_count(w->eng, l) : 0;
}

int seqalign_multi_hit(seqalign_multi_t *m, size_t i, size_t h, alignment_t *out)
{
  size_t l = 0;
  sa_worker *w = locate(m, i, &l);
  if(!w) return SEQALIGN_ERR_ARG;
  const int rc = seqalign_batch_hit(w->eng, l, h, out);
  if(rc < 0) snprintf(m->err, sizeof(m->err), "%s", seqalign_batch_error(w->eng));
  return rc;
}

int seqalign_multi_matrices(seqalign_multi_t *m, size_t i, int32_t *match, int32_t *gap_a, int32_t *gap_b)
{
  size_t l = 0;
  sa_worker *w = locate(m, i, &l);
  if(!w) return SEQALIGN_ERR_ARG;
  const int rc = seqalign_batch_matrices(w->eng, l, match, gap_a, gap_b);
  if(rc < 0) snprintf(m->err, sizeof(m->err), "%s", seqalign_batch_error(w->eng));
  return rc;
}

void seqalign_multi_unknown_pair(const seqalign_multi_t *m, char *a, char *b)
{
  if(!m) return;
  for(int r = 0; r < m->n; r++)
    if(m->w[r].rc == SEQALIGN_ERR_UNKNOWN_PAIR) { seqalign_batch_unknown_pair(m->w[r].eng, a, b); return; }
}

double seqalign_multi_last_kernel_ms(const seqalign_multi_t *m)
{
  double mx = 0;
  if(m) for(int r = 0; r < m->n; r++) { const double v = seqalign_batch_last_kernel_ms(m->w[r].eng); if(v > mx) mx = v; }
  return mx;
}
