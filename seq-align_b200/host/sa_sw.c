/*
 * sa_sw.c -- Smith-Waterman front-end of the seq-align C API.
 *
 * Implements include/smith_waterman.h.  smith_waterman_align2() fills on the
 * GPU (matrices materialised into the embedded aligner_t, which
 * smith_waterman_get_aligner() exposes, reference smith_waterman.c:126-129).
 * The first smith_waterman_fetch() -- the only one `--maxhits 1` and the
 * batch path need -- is served by the GPU: best cell under the reference's
 * hit order + walk kernel.  Further hits are iterated on the host over the
 * materialised match matrix with the reference's semantics
 * (smith_waterman.c:152-161, 165-277): candidates with match>0 in (score
 * desc, x asc, y asc) order, a visited mask, a hit dropped as soon as its
 * walk meets a marked cell.  The mask is cleared completely for every pair
 * (the reference clears a quarter of it, smith_waterman.c:149 -- see
 * DESIGN.md "known upstream defects").
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
  aligner_align(&sw->aligner, a, b, len_a, len_b, scoring, 1);
  sw->have_cands = 0;
  sw->ncands = sw->next_cand = 0;
  sw->fetched = 0;
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

int smith_waterman_fetch(sw_aligner_t *sw, alignment_t *result)
{
  const aligner_t *al = &sw->aligner;
  if(sw->fetched == 0) {
    /* first hit: GPU best cell + GPU walk (batch of one, align mode) */
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
  if(!sw->have_cands) {
    /* replay hit 0 on the host to mark its path, then continue from hit 1 */
    build_candidates(sw);
    if(sw->ncands == 0) return 0;
    follow_candidate(sw, &sw->cands[0], NULL);
    sw->next_cand = 1;
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
