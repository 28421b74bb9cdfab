/*
 * sa_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see sa_oracle.h).
 *
 * Restates, function by function, the reference algorithm of
 * noporpoise/seq-align for the DP hot path.  Not product code: nothing under
 * seq-align_b200/ may call into this file.
 *
 * Parity pinned against the reference's golden vectors and the compiled
 * reference (oracle/_ref) -- see tests/test_oracle.py.
 */
#include "sa_oracle.h"

#include <ctype.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>

enum { ST_M = 0, ST_GA = 1, ST_GB = 2 };

static inline int imin(int p, int q) { return p < q ? p : q; }
static inline int imax(int p, int q) { return p > q ? p : q; }
static inline int imax3(int p, int q, int r) { return imax(imax(p, q), r); }

size_t orc_scoring_sizeof(void) { return sizeof(orc_scoring_t); }

/* ref: src/alignment_scoring.c:21-55 (scoring_init).  min/max_penalty only
 * include the gap terms when at least one sequence may carry gaps (:51). */
void orc_scoring_init(orc_scoring_t *s, int match, int mismatch,
                      int gap_open, int gap_extend,
                      int no_start_gap_penalty, int no_end_gap_penalty,
                      int no_gaps_in_a, int no_gaps_in_b,
                      int no_mismatches, int case_sensitive)
{
  memset(s, 0, sizeof(*s));
  s->gap_open = gap_open;
  s->gap_extend = gap_extend;
  s->no_start_gap_penalty = !!no_start_gap_penalty;
  s->no_end_gap_penalty = !!no_end_gap_penalty;
  s->no_gaps_in_a = !!no_gaps_in_a;
  s->no_gaps_in_b = !!no_gaps_in_b;
  s->no_mismatches = !!no_mismatches;
  s->use_match_mismatch = 1;
  s->match = match;
  s->mismatch = mismatch;
  s->case_sensitive = !!case_sensitive;
  s->min_penalty = imin(match, mismatch);
  s->max_penalty = imax(match, mismatch);
  if(!no_gaps_in_a || !no_gaps_in_b) {
    int first = gap_open + gap_extend;
    s->min_penalty = imin(s->min_penalty, imin(first, gap_extend));
    s->max_penalty = imax(s->max_penalty, imax(first, gap_extend));
  }
}

static void widen_bounds(orc_scoring_t *s, int score)
{
  s->min_penalty = imin(s->min_penalty, score);
  s->max_penalty = imax(s->max_penalty, score);
}

/* ref: src/alignment_scoring.c:57-64 -- folds case unless case_sensitive */
void orc_scoring_add_wildcard(orc_scoring_t *s, int c, int score)
{
  c &= 0xff;
  if(!s->case_sensitive) c = tolower(c);
  s->is_wild[c] = 1;
  s->wild_score[c] = score;
  widen_bounds(s, score);
}

/* ref: src/alignment_scoring.c:66-72 -- does NOT fold case */
void orc_scoring_add_mutation(orc_scoring_t *s, int a, int b, int score)
{
  a &= 0xff; b &= 0xff;
  s->has_swap[a][b] = 1;
  s->swap_score[a][b] = score;
  widen_bounds(s, score);
}

/* ref: src/alignment_scoring.c:74-97 -- this one folds case; scores is a
 * len x len table read as scores[j*len + i] for letters (i, j). */
void orc_scoring_add_mutations(orc_scoring_t *s, const char *letters,
                               const int *scores, int use_match_mismatch)
{
  size_t n = strlen(letters);
  for(size_t i = 0; i < n; i++) {
    int ca = (unsigned char)letters[i];
    if(!s->case_sensitive) ca = tolower(ca);
    for(size_t j = 0; j < n; j++) {
      int cb = (unsigned char)letters[j];
      if(!s->case_sensitive) cb = tolower(cb);
      orc_scoring_add_mutation(s, ca, cb, scores[j * n + i]);
    }
  }
  s->use_match_mismatch = !!use_match_mismatch;
}

void orc_scoring_poke(orc_scoring_t *s, int match, int mismatch,
                      int gap_open, int gap_extend)
{
  s->match = match;
  s->mismatch = mismatch;
  s->gap_open = gap_open;
  s->gap_extend = gap_extend;
}

/* ref: src/alignment_scoring.c:115-129 (_scoring_check_wildcards): the
 * smaller wildcard score of the two characters wins. */
static int wildcard_score(const orc_scoring_t *s, int a, int b, int *score)
{
  int found = 0, best = 0;
  if(s->is_wild[a]) { best = s->wild_score[a]; found = 1; }
  if(s->is_wild[b]) {
    best = found ? imin(best, s->wild_score[b]) : s->wild_score[b];
    found = 1;
  }
  /* the reference uses INT_MAX as "none"; a wildcard scored INT_MAX would
   * read as "none" there too, keep that corner identical */
  if(found && best == INT_MAX) found = 0;
  *score = found ? best : 0;
  return found;
}

/* ref: src/alignment_scoring.c:133-182 (scoring_lookup) */
int orc_scoring_lookup(const orc_scoring_t *s, int a, int b,
                       int *score, int *is_match)
{
  a &= 0xff; b &= 0xff;
  if(!s->case_sensitive) { a = tolower(a); b = tolower(b); }
  *is_match = (a == b);

  if(s->no_mismatches && !*is_match) {            /* :148-153 */
    *is_match = wildcard_score(s, a, b, score);
    return 0;
  }
  if(s->has_swap[a][b]) {                         /* :156-160 */
    *score = s->swap_score[a][b];
    return 0;
  }
  if(wildcard_score(s, a, b, score)) {            /* :165-169 */
    *is_match = 1;
    return 0;
  }
  if(s->use_match_mismatch) {                     /* :172-176 */
    *score = *is_match ? s->match : s->mismatch;
    return 0;
  }
  return -1;                                      /* :179-181 exit(1) */
}

/* ref: src/alignment.c:28-168 (alignment_fill_matrices).
 * x indexes seq_a (columns), y indexes seq_b (rows). */
int orc_fill(const orc_scoring_t *s, const char *a, size_t la,
             const char *b, size_t lb, int is_sw,
             int *mm, int *ga, int *gb)
{
  const size_t w = la + 1;
  const int open = s->gap_open + s->gap_extend;    /* :38 */
  const int ext = s->gap_extend;                   /* :39 */
  const int floor_ = is_sw ? 0 : INT_MIN + abs(s->min_penalty); /* :41 */

  mm[0] = ga[0] = gb[0] = 0;                       /* :47-49 */
  for(size_t x = 1; x <= la; x++) {                /* row 0 */
    if(is_sw) { mm[x] = ga[x] = gb[x] = 0; continue; }        /* :53-54 */
    mm[x] = ga[x] = floor_;                                    /* :63-66 */
    gb[x] = s->no_start_gap_penalty ? 0
            : s->gap_open + (int)x * s->gap_extend;            /* :67-68 */
  }
  for(size_t y = 1; y <= lb; y++) {                /* column 0 */
    size_t k = y * w;
    if(is_sw) { mm[k] = ga[k] = gb[k] = floor_; continue; }    /* :55-56 */
    mm[k] = gb[k] = floor_;                                    /* :74,79 */
    ga[k] = s->no_start_gap_penalty ? 0
            : s->gap_open + (int)y * s->gap_extend;            /* :77-78 */
  }

  for(size_t y = 1; y <= lb; y++) {
    for(size_t x = 1; x <= la; x++) {
      const size_t c = y * w + x;
      const size_t up = c - w, left = c - 1, diag = up - 1;
      int sub, is_match;
      if(orc_scoring_lookup(s, a[x - 1], b[y - 1], &sub, &is_match) != 0)
        return -2;

      /* match state: from the diagonal, :101-116 */
      if(s->no_mismatches && !is_match)
        mm[c] = floor_;
      else
        mm[c] = imax(imax3(mm[diag], ga[diag], gb[diag]) + sub, floor_);

      /* gap in a: consumes a character of b, predecessor above, :122-137 */
      if(x == la && s->no_end_gap_penalty)
        ga[c] = imax3(mm[up], ga[up], gb[up]);
      else if(!s->no_gaps_in_a || x == la)
        ga[c] = imax(imax3(mm[up] + open, ga[up] + ext, gb[up] + open), floor_);
      else
        ga[c] = floor_;

      /* gap in b: consumes a character of a, predecessor left, :140-155 */
      if(y == lb && s->no_end_gap_penalty)
        gb[c] = imax3(mm[left], ga[left], gb[left]);
      else if(!s->no_gaps_in_b || y == lb)
        gb[c] = imax(imax3(mm[left] + open, ga[left] + open, gb[left] + ext), floor_);
      else
        gb[c] = floor_;
    }
  }
  return 0;
}

/* ref: src/alignment.c:244-350 (alignment_reverse_move) */
int orc_reverse_move(const orc_scoring_t *s, const char *a, size_t la,
                     const char *b, size_t lb,
                     const int *mm, const int *ga, const int *gb,
                     int *state, int *score, size_t *px, size_t *py)
{
  const size_t w = la + 1;
  size_t x = *px, y = *py;
  int sub, is_match;
  if(orc_scoring_lookup(s, a[x - 1], b[y - 1], &sub, &is_match) != 0) return -2;

  int a_open = s->gap_open + s->gap_extend, a_ext = s->gap_extend;
  int b_open = a_open, b_ext = a_ext;
  if(s->no_end_gap_penalty) {                      /* :265-268 */
    if(x == la) a_open = a_ext = 0;
    if(y == lb) b_open = b_ext = 0;
  }
  if(s->no_start_gap_penalty) {                    /* :269-272 (dead: x,y>0) */
    if(x == 0) a_open = a_ext = 0;
    if(y == 0) b_open = b_ext = 0;
  }

  int from_m, from_ga, from_gb;
  switch(*state) {                                 /* :276-307 */
    case ST_M:  from_m = from_ga = from_gb = sub; x--; y--; break;
    case ST_GA: from_m = a_open; from_ga = a_ext; from_gb = a_open; y--; break;
    case ST_GB: from_m = b_open; from_ga = b_open; from_gb = b_ext; x--; break;
    default: return -3;
  }
  const size_t c = y * w + x;
  const int ok_a = !s->no_gaps_in_a || x == 0 || x == la;
  const int ok_b = !s->no_gaps_in_b || y == 0 || y == lb;

  if(ok_a && ga[c] + from_ga == *score)      { *state = ST_GA; *score = ga[c]; }
  else if(ok_b && gb[c] + from_gb == *score) { *state = ST_GB; *score = gb[c]; }
  else if(mm[c] + from_m == *score)          { *state = ST_M;  *score = mm[c]; }
  else return -1;                                  /* :328-349 */
  *px = x; *py = y;
  return 0;
}

static int *alloc3(size_t la, size_t lb, int **ga, int **gb)
{
  size_t n = (la + 1) * (lb + 1);
  int *base = (int *)malloc(3 * n * sizeof(int));
  if(!base) return NULL;
  *ga = base + n;
  *gb = base + 2 * n;
  return base;
}

/* ref: src/needleman_wunsch.c:53-66 -- ">=" chain, GA beats GB beats M */
static void nw_end_state(const int *mm, const int *ga, const int *gb,
                         size_t last, int *state, int *score)
{
  *state = ST_M; *score = mm[last];
  if(gb[last] >= *score) { *state = ST_GB; *score = gb[last]; }
  if(ga[last] >= *score) { *state = ST_GA; *score = ga[last]; }
}

/* ref: src/needleman_wunsch.c:34-145 (needleman_wunsch_align2) */
int orc_nw_align(const orc_scoring_t *s, const char *a, size_t la,
                 const char *b, size_t lb, orc_alignment_t *out)
{
  int *ga, *gb, *mm = alloc3(la, lb, &ga, &gb);
  if(!mm) return -4;
  int rc = orc_fill(s, a, la, b, lb, 0, mm, ga, gb);
  if(rc) { free(mm); return rc; }

  int state, score;
  nw_end_state(mm, ga, gb, (la + 1) * (lb + 1) - 1, &state, &score);
  out->score = score;

  /* build right-to-left into the tail of the buffers, then shift (:79-145) */
  size_t cap = la + lb, n = 0, x = la, y = lb;
  char *ra = out->result_a, *rb = out->result_b;
  while(x > 0 && y > 0) {
    char ca = '-', cb = '-';
    if(state != ST_GA) ca = a[x - 1];
    if(state != ST_GB) cb = b[y - 1];
    n++;
    ra[cap - n] = ca; rb[cap - n] = cb;
    rc = orc_reverse_move(s, a, la, b, lb, mm, ga, gb, &state, &score, &x, &y);
    if(rc) { free(mm); return rc; }
  }
  for(; y > 0; y--) { n++; ra[cap - n] = '-'; rb[cap - n] = b[y - 1]; }   /* :117-123 */
  for(; x > 0; x--) { n++; ra[cap - n] = a[x - 1]; rb[cap - n] = '-'; }   /* :126-132 */
  memmove(ra, ra + cap - n, n);
  memmove(rb, rb + cap - n, n);
  ra[n] = rb[n] = '\0';
  out->length = n;
  out->pos_a = out->pos_b = 0;
  out->len_a = la; out->len_b = lb;
  free(mm);
  return 0;
}

int orc_nw_score(const orc_scoring_t *s, const char *a, size_t la,
                 const char *b, size_t lb, int *score)
{
  int *ga, *gb, *mm = alloc3(la, lb, &ga, &gb);
  if(!mm) return -4;
  int rc = orc_fill(s, a, la, b, lb, 0, mm, ga, gb);
  if(!rc) {
    int st;
    nw_end_state(mm, ga, gb, (la + 1) * (lb + 1) - 1, &st, score);
  }
  free(mm);
  return rc;
}

/* Hit order, ref: src/smith_waterman.c:71-86 + glibc's stable qsort_r:
 * score descending, then x ascending, then (stability over ascending
 * indices) y ascending. */
typedef struct { int score; unsigned x, y; } hit_key_t;

static int hit_cmp(const void *pa, const void *pb)
{
  const hit_key_t *p = (const hit_key_t *)pa, *q = (const hit_key_t *)pb;
  if(p->score != q->score) return p->score > q->score ? -1 : 1;
  if(p->x != q->x) return p->x < q->x ? -1 : 1;
  if(p->y != q->y) return p->y < q->y ? -1 : 1;
  return 0;
}

int orc_sw_best(const orc_scoring_t *s, const char *a, size_t la,
                const char *b, size_t lb,
                int *score, size_t *x_end, size_t *y_end)
{
  int *ga, *gb, *mm = alloc3(la, lb, &ga, &gb);
  if(!mm) return -4;
  int rc = orc_fill(s, a, la, b, lb, 1, mm, ga, gb);
  if(rc) { free(mm); return rc; }
  hit_key_t best = {0, 0, 0};
  for(size_t y = 0; y <= lb; y++)
    for(size_t x = 0; x <= la; x++) {
      int v = mm[y * (la + 1) + x];
      if(v <= 0) continue;                         /* :152-156 */
      hit_key_t k = {v, (unsigned)x, (unsigned)y};
      if(best.score == 0 || hit_cmp(&k, &best) < 0) best = k;
    }
  *score = best.score; *x_end = best.x; *y_end = best.y;
  free(mm);
  return 0;
}

/* ref: src/smith_waterman.c:137-277 (align2 + fetch + _follow_hit), on a
 * fresh visited mask (the reference's reused-aligner stale mask, :149, is a
 * bug and is not part of the contract). */
long orc_sw_hits(const orc_scoring_t *s, const char *a, size_t la,
                 const char *b, size_t lb, size_t max_hits,
                 orc_alignment_t *hits, char *pool_a, char *pool_b,
                 size_t stride)
{
  const size_t w = la + 1, cells = w * (lb + 1);
  int *ga, *gb, *mm = alloc3(la, lb, &ga, &gb);
  if(!mm) return -4;
  int rc = orc_fill(s, a, la, b, lb, 1, mm, ga, gb);
  if(rc) { free(mm); return rc; }

  size_t nk = 0;
  hit_key_t *keys = (hit_key_t *)malloc((cells ? cells : 1) * sizeof(*keys));
  unsigned char *seen = (unsigned char *)calloc(cells ? cells : 1, 1);
  for(size_t c = 0; c < cells; c++)
    if(mm[c] > 0) {
      keys[nk].score = mm[c]; keys[nk].x = (unsigned)(c % w);
      keys[nk].y = (unsigned)(c / w); nk++;
    }
  qsort(keys, nk, sizeof(*keys), hit_cmp);   /* total order: no ties left */

  size_t nh = 0;
  for(size_t k = 0; k < nk && nh < max_hits; k++) {
    size_t x = keys[k].x, y = keys[k].y;
    if(seen[y * w + x]) continue;                  /* :270 */

    /* pass 1 (:187-199): mark the path, measure it, abort on a seen cell */
    int state = ST_M, score = mm[y * w + x];
    size_t len = 0, ok = 1;
    for(;; len++) {
      size_t c = y * w + x;
      if(seen[c]) { ok = 0; break; }
      seen[c] = 1;
      if(score == 0) break;
      rc = orc_reverse_move(s, a, la, b, lb, mm, ga, gb, &state, &score, &x, &y);
      if(rc) goto done;
    }
    if(!ok) continue;

    /* pass 2 (:217-244): emit characters right to left */
    orc_alignment_t *h = &hits[nh];
    h->result_a = pool_a + nh * stride;
    h->result_b = pool_b + nh * stride;
    x = keys[k].x; y = keys[k].y; state = ST_M; score = keys[k].score;
    for(size_t i = len; score > 0; ) {
      i--;
      h->result_a[i] = (state == ST_GA) ? '-' : a[x - 1];
      h->result_b[i] = (state == ST_GB) ? '-' : b[y - 1];
      rc = orc_reverse_move(s, a, la, b, lb, mm, ga, gb, &state, &score, &x, &y);
      if(rc) goto done;
    }
    h->result_a[len] = h->result_b[len] = '\0';
    h->length = len;
    h->score = keys[k].score;                      /* :246-257 */
    h->pos_a = x; h->pos_b = y;
    h->len_a = keys[k].x - x; h->len_b = keys[k].y - y;
    nh++;
  }
  rc = 0;
done:
  free(keys); free(seen); free(mm);
  return rc ? rc : (long)nh;
}

int orc_batch_sw_best(const orc_scoring_t *s, size_t n,
                      const char *seq_a, const long long *off_a,
                      const char *seq_b, const long long *off_b,
                      int *score, int *x_end, int *y_end)
{
  for(size_t i = 0; i < n; i++) {
    size_t xe, ye;
    int rc = orc_sw_best(s, seq_a + off_a[i], (size_t)(off_a[i + 1] - off_a[i]),
                         seq_b + off_b[i], (size_t)(off_b[i + 1] - off_b[i]),
                         &score[i], &xe, &ye);
    if(rc) return rc;
    x_end[i] = (int)xe; y_end[i] = (int)ye;
  }
  return 0;
}

int orc_batch_nw_score(const orc_scoring_t *s, size_t n,
                       const char *seq_a, const long long *off_a,
                       const char *seq_b, const long long *off_b,
                       int *score)
{
  for(size_t i = 0; i < n; i++) {
    int rc = orc_nw_score(s, seq_a + off_a[i], (size_t)(off_a[i + 1] - off_a[i]),
                          seq_b + off_b[i], (size_t)(off_b[i + 1] - off_b[i]),
                          &score[i]);
    if(rc) return rc;
  }
  return 0;
}

/* =========================================================================
 * Sequence-file reader (see sa_oracle.h).  A memory buffer stands in for the
 * reference's gzFile + StreamBuffer pair (libs/seq_file/stream_buffer.h:221-315):
 * getc / ungetc / readline / chomp below are those of the buffered variants.
 */
typedef struct { const char *t; size_t n, pos; } orc_text;
typedef struct { char *b; size_t len; } orc_buf;   /* caller-sized: never longer than the text */

/* ref: stream_buffer.h:230-238 (getc_buf): the byte as a (signed) char, -1 at the end */
static int rd_getc(orc_text *s) { return s->pos < s->n ? (int)(signed char)s->t[s->pos++] : -1; }
/* ref: stream_buffer.h:245-258: the byte goes back in front of the buffer; pushing back the -1 of
 * the end of input leaves a byte that reads as -1 again, i.e. nothing changes */
static void rd_ungetc(orc_text *s, int c) { if(c != -1 && s->pos > 0) s->pos--; }
/* ref: stream_buffer.h:295-315: append up to and including the next '\n'; bytes read */
static size_t rd_readline(orc_text *s, orc_buf *b)
{
  size_t got = 0;
  while(s->pos < s->n) {
    const char c = s->t[s->pos++];
    b->b[b->len++] = c; got++;
    if(c == '\n') break;
  }
  return got;
}
/* ref: stream_buffer.h:56-61 */
static void rd_chomp(orc_buf *b) { while(b->len && (b->b[b->len - 1] == '\n' || b->b[b->len - 1] == '\r')) b->len--; }

/* ref: seq_file.h:245-272 */
static int rd_fastq(orc_text *s, orc_buf *name, orc_buf *seq, orc_buf *qual, size_t *name_pos)
{
  int c = rd_getc(s);
  if(c == -1) return 0;
  *name_pos = s->pos;
  if(c != '@' || rd_readline(s, name) == 0) return -1;
  rd_chomp(name);
  while((c = rd_getc(s)) != '+') {
    if(c == -1) return -1;
    if(c != '\r' && c != '\n') {
      seq->b[seq->len++] = (char)c;
      if(rd_readline(s, seq) == 0) return -1;
      rd_chomp(seq);
    }
  }
  while((c = rd_getc(s)) != -1 && c != '\n') {}
  if(c == -1) return -1;
  do {
    if(rd_readline(s, qual) > 0) rd_chomp(qual);
    else return 1;
  } while(qual->len < seq->len);
  while((c = rd_getc(s)) != -1 && c != '@') {}
  rd_ungetc(s, c);
  return 1;
}

/* ref: seq_file.h:274-295 */
static int rd_fasta(orc_text *s, orc_buf *name, orc_buf *seq, size_t *name_pos)
{
  int c = rd_getc(s);
  if(c == -1) return 0;
  *name_pos = s->pos;
  if(c != '>' || rd_readline(s, name) == 0) return -1;
  rd_chomp(name);
  while((c = rd_getc(s)) != '>') {
    if(c == -1) return 1;
    if(c != '\r' && c != '\n') {
      seq->b[seq->len++] = (char)c;
      const long nread = (long)rd_readline(s, seq);
      rd_chomp(seq);
      if(nread <= 0) return 1;
    }
  }
  rd_ungetc(s, c);
  return 1;
}

/* ref: seq_file.h:298-309; the skipline of :303 is the H6 no-op */
static int rd_plain(orc_text *s, orc_buf *seq)
{
  int c;
  while((c = rd_getc(s)) != -1 && isspace(c)) {}
  if(c == -1) return 0;
  seq->b[seq->len++] = (char)c;
  rd_readline(s, seq);
  rd_chomp(seq);
  return 1;
}

long orc_read_records(const char *text, size_t n, size_t max_rec, char *seq_out, long long *seq_off,
                      long long *name_pos, long long *name_len, long long *rec_pos, int *fmt, int *last)
{
  orc_text s = {text, n, 0};
  char *nb = (char *)malloc(n + 2), *qb = (char *)malloc(n + 2);
  long count = 0;
  size_t used = 0;
  int st = 0;
  seq_off[0] = 0;
  for(;;) {
    /* ref: seq_file.h:311-323, run for every record */
    int c;
    while((c = rd_getc(&s)) != -1 && isspace(c)) {}
    if(c == -1) { st = 0; break; }
    const int f = c == '@' ? 4 : c == '>' ? 2 : 1;
    rd_ungetc(&s, c);
    const size_t start = s.pos;
    orc_buf name = {nb, 0}, qual = {qb, 0}, seq = {seq_out + used, 0};
    size_t npos = start;
    st = f == 4 ? rd_fastq(&s, &name, &seq, &qual, &npos) : f == 2 ? rd_fasta(&s, &name, &seq, &npos) : rd_plain(&s, &seq);
    if(st <= 0) break;
    if((size_t)count == max_rec) { st = -2; break; }
    fmt[count] = f;
    rec_pos[count] = (long long)start;
    name_pos[count] = (long long)npos;
    name_len[count] = (long long)name.len;
    used += seq.len;
    seq_off[++count] = (long long)used;
  }
  free(nb); free(qb);
  if(last) *last = st;
  return count;
}
