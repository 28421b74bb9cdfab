/*
 * sa_sw.c -- Smith-Waterman front-end of the seq-align C API.
 *
 * Implements include/smith_waterman.h.  smith_waterman_align2() only binds the
 * pair (the matrices of the embedded aligner_t, which smith_waterman_get_aligner()
 * exposes, reference smith_waterman.c:126-129, are filled when somebody reads
 * them: sa_alignment.c "deferred matrices").  Every smith_waterman_fetch() is
 * served by the GPU: the first one by the engine's align mode (best cell under
 * the reference's hit order + walk kernel: all `--maxhits 1` needs), the
 * following ones by its multi-hit mode (SEQALIGN_MODE_HITS: candidate sort and
 * masked walks on the device, reference smith_waterman.c:152-161, 165-277), whose
 * list is copied into the aligner and grown eightfold whenever the caller
 * fetches past its end.
 * Only scoring shapes the device multi-hit stage refuses (no_gaps_in_a/b,
 * no_mismatches, free end gaps: SEQALIGN_ERR_ARG) iterate on the host over the
 * materialised match matrix, with the reference's semantics: candidates with
 * match>0 in (score desc, x asc, y asc) order, a visited mask, a hit dropped
 * as soon as its walk meets a marked cell.  The mask is cleared completely
 * for every pair (the reference clears a quarter of it, smith_waterman.c:149
 * -- see DESIGN.md "known upstream defects").
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "smith_waterman.h"
#include "seqalign_b200.h"
#include "sa_host.h"

typedef struct { score_t score; unsigned x, y; } sw_cand_t;

struct sw_aligner_t
{
  aligner_t aligner;
  /* hit iteration state */
  uint32_t *mask;            /* one bit per cell */
  size_t mask_words;
  sw_cand_t *cands;          /* sorted candidates, built on demand */
  size_t ncands, cand_cap, next_cand;
  int have_cands;
  size_t fetched;            /* hits returned so far for this pair */
  /* hit list made by the device (copied out of the engine, which other aligners share) */
  alignment_t **list;
  size_t list_n, list_cap, list_alloc;
  int use_host;              /* the device stage refused this scoring shape */
};

sw_aligner_t *smith_waterman_new()
{
  sw_aligner_t *sw = calloc(1, sizeof(sw_aligner_t));
  if(!sw) { fprintf(stderr, "%s:%i: Out of memory\n", __FILE__, __LINE__); exit(EXIT_FAILURE); }
  return sw;
}

void smith_waterman_free(sw_aligner_t *sw)
{
  aligner_destroy(&sw->aligner);
  for(size_t i = 0; i < sw->list_alloc; i++) alignment_free(sw->list[i]);
  free(sw->list);
  free(sw->mask);
  free(sw->cands);
  free(sw);
}

aligner_t *smith_waterman_get_aligner(sw_aligner_t *sw)
{
  return &sw->aligner;
}

void smith_waterman_align(const char *a, const char *b, const scoring_t *scoring, sw_aligner_t *sw)
{
  smith_waterman_align2(a, b, strlen(a), strlen(b), scoring, sw);
}

void smith_waterman_align2(const char *a, const char *b, size_t len_a, size_t len_b,
                           const scoring_t *scoring, sw_aligner_t *sw)
{
  sa_host_bind(&sw->aligner, a, b, len_a, len_b, scoring, 1);
  sw->have_cands = 0;
  sw->ncands = sw->next_cand = 0;
  sw->fetched = 0;
  sw->list_n = sw->list_cap = 0;
  sw->use_host = 0;
}

/* total order of hits: score desc, x asc, y asc */
static int cand_cmp(const void *pa, const void *pb)
{
  const sw_cand_t *p = pa, *q = pb;
  if(p->score != q->score) return p->score > q->score ? -1 : 1;
  if(p->x != q->x) return p->x < q->x ? -1 : 1;
  if(p->y != q->y) return p->y < q->y ? -1 : 1;
  return 0;
}

static void build_candidates(sw_aligner_t *sw)
{
  const aligner_t *al = &sw->aligner;
  sa_host_materialise(al);
  const size_t w = al->score_width, cells = w * al->score_height;
  size_t words = (cells + 31) / 32;
  if(words > sw->mask_words) {
    sw->mask = realloc(sw->mask, words * sizeof(uint32_t));
    sw->mask_words = words;
  }
  size_t n = 0;
  for(size_t k = 0; k < cells; k++) n += al->match_scores[k] > 0;
  if(n > sw->cand_cap) {
    sw->cands = realloc(sw->cands, n * sizeof(sw_cand_t));
    sw->cand_cap = n;
  }
  if((words && !sw->mask) || (n && !sw->cands)) {
    fprintf(stderr, "%s:%i: Out of memory\n", __FILE__, __LINE__);
    exit(EXIT_FAILURE);
  }
  memset(sw->mask, 0, words * sizeof(uint32_t));
  n = 0;
  for(size_t k = 0; k < cells; k++)
    if(al->match_scores[k] > 0) {
      sw->cands[n].score = al->match_scores[k];
      sw->cands[n].x = (unsigned)(k % w);
      sw->cands[n].y = (unsigned)(k / w);
      n++;
    }
  qsort(sw->cands, n, sizeof(sw_cand_t), cand_cmp);
  sw->ncands = n;
  sw->next_cand = 0;
  sw->have_cands = 1;
}

/* walk one candidate back over the matrices (reference smith_waterman.c:
 * 165-258).  Returns 0 if the walk runs into an already used cell.  With
 * result == NULL the path is only marked. */
static int follow_candidate(sw_aligner_t *sw, const sw_cand_t *c, alignment_t *result)
{
  const aligner_t *al = &sw->aligner;
  sa_host_materialise(al);
  size_t x = c->x, y = c->y, k = (size_t)c->y * al->score_width + c->x;
  enum Matrix m = MATCH;
  score_t s = c->score;
  size_t length = 0;
  for(;; length++) {
    if(bitset32_get(sw->mask, k)) return 0;
    bitset32_set(sw->mask, k);
    if(s == 0) break;
    alignment_reverse_move(&m, &s, &x, &y, &k, al);
  }
  if(!result) return 1;

  alignment_ensure_capacity(result, length);
  result->length = length;
  x = c->x; y = c->y; k = (size_t)c->y * al->score_width + c->x;
  m = MATCH; s = c->score;
  for(size_t i = length; s > 0; ) {
    i--;
    result->result_a[i] = m == GAP_A ? '-' : al->seq_a[x - 1];
    result->result_b[i] = m == GAP_B ? '-' : al->seq_b[y - 1];
    alignment_reverse_move(&m, &s, &x, &y, &k, al);
  }
  result->result_a[length] = result->result_b[length] = '\0';
  result->score = c->score;
  result->pos_a = x; result->pos_b = y;
  result->len_a = c->x - x; result->len_b = c->y - y;
  return 1;
}

static void copy_alignment(alignment_t *dst, const alignment_t *src)
{
  alignment_ensure_capacity(dst, src->length);
  memcpy(dst->result_a, src->result_a, src->length + 1);
  memcpy(dst->result_b, src->result_b, src->length + 1);
  dst->length = src->length; dst->score = src->score;
  dst->pos_a = src->pos_a; dst->pos_b = src->pos_b; dst->len_a = src->len_a; dst->len_b = src->len_b;
}

/* (re)build the device hit list with room for `cap` hits; 0 if the engine's multi-hit stage
 * does not take this scoring shape */
static int build_device_list(sw_aligner_t *sw, size_t cap)
{
  const aligner_t *al = &sw->aligner;
  if(cap > 4096) return 0;   /* beyond the engine's per-pair list (seqalign_batch_set_hit_limits): the host iteration goes on from here */
  seqalign_batch_t *eng = sa_host_engine();
  const size_t la = al->score_width - 1, lb = al->score_height - 1;
  seqalign_batch_set_scoring(eng, al->scoring);
  if(seqalign_batch_set_hit_limits(eng, cap, 1) != SEQALIGN_OK) return 0;
  const int rc = seqalign_batch_submit(eng, SEQALIGN_SW, SEQALIGN_MODE_HITS, &al->seq_a, &la, &al->seq_b, &lb, 1);
  if(rc == SEQALIGN_ERR_ARG || rc == SEQALIGN_ERR_NOMEM) return 0;
  sa_host_check(eng, rc);
  const size_t n = seqalign_batch_hit_count(eng, 0);
  if(n > sw->list_alloc) {
    sw->list = realloc(sw->list, n * sizeof(alignment_t *));
    if(!sw->list) { fprintf(stderr, "%s:%i: Out of memory\n", __FILE__, __LINE__); exit(EXIT_FAILURE); }
    for(size_t i = sw->list_alloc; i < n; i++) sw->list[i] = alignment_create(la + lb + 1);
    sw->list_alloc = n;
  }
  alignment_t *tmp = alignment_create(la + lb + 1);
  for(size_t i = 0; i < n; i++) {
    sa_host_check(eng, seqalign_batch_hit(eng, 0, i, tmp));
    copy_alignment(sw->list[i], tmp);
  }
  alignment_free(tmp);
  sw->list_n = n;
  sw->list_cap = cap;
  return 1;
}

static int fetch_on_host(sw_aligner_t *sw, alignment_t *result)
{
  const aligner_t *al = &sw->aligner;
  if(!sw->have_cands) {
    /* replay the hits already handed out (their paths mark cells), then continue */
    build_candidates(sw);
    size_t replay = sw->fetched;
    while(replay > 0 && sw->next_cand < sw->ncands) {
      const sw_cand_t *c = &sw->cands[sw->next_cand++];
      const size_t k = (size_t)c->y * al->score_width + c->x;
      if(!bitset32_get(sw->mask, k) && follow_candidate(sw, c, NULL)) replay--;
    }
  }
  while(sw->next_cand < sw->ncands) {
    const sw_cand_t *c = &sw->cands[sw->next_cand++];
    const size_t k = (size_t)c->y * al->score_width + c->x;
    if(!bitset32_get(sw->mask, k) && follow_candidate(sw, c, result)) {
      sw->fetched++;
      return 1;
    }
  }
  return 0;
}

int smith_waterman_fetch(sw_aligner_t *sw, alignment_t *result)
{
  const aligner_t *al = &sw->aligner;
  if(sw->fetched == 0) {
    /* first hit: best cell + walk (batch of one, align mode; every scoring shape) */
    seqalign_batch_t *eng = sa_host_engine();
    const size_t la = al->score_width - 1, lb = al->score_height - 1;
    seqalign_batch_set_scoring(eng, al->scoring);
    sa_host_check(eng, seqalign_batch_submit(eng, SEQALIGN_SW, SEQALIGN_MODE_ALIGN,
                                             &al->seq_a, &la, &al->seq_b, &lb, 1));
    const int rc = seqalign_batch_alignment(eng, 0, result);
    sa_host_check(eng, rc);
    sw->fetched = 1;
    return rc;
  }
  if(!sw->use_host) {
    /* later hits: the device's list, made with room to spare and grown when the caller reads past it */
    if(sw->list_cap == 0 || (sw->fetched >= sw->list_n && sw->list_n == sw->list_cap)) {
      size_t cap = sw->list_cap ? sw->list_cap * 8 : 8;
      while(cap <= sw->fetched) cap *= 8;
      if(!build_device_list(sw, cap)) sw->use_host = 1;
    }
    if(!sw->use_host) {
      if(sw->fetched >= sw->list_n) return 0;
      copy_alignment(result, sw->list[sw->fetched]);
      sw->fetched++;
      return 1;
    }
  }
  return fetch_on_host(sw, result);
}
