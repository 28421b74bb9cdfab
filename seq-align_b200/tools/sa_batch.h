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
#include "sa_cli.h"

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

typedef struct {
  char **a, **b, **name_a, **name_b; /* owned copies; names NULL when the record had none */
  size_t *la, *lb;
  size_t n, cap, bytes;
} sa_pairs;

static inline char *sa_dup(const char *s, size_t n)
{
  char *d = malloc(n + 1);
  if(!d) { fprintf(stderr, "%s:%i: Out of memory\n", __FILE__, __LINE__); exit(EXIT_FAILURE); }
  memcpy(d, s, n);
  d[n] = '\0';
  return d;
}

static inline void sa_pairs_add(sa_pairs *p, const sa_record *r1, const sa_record *r2)
{
  if(p->n == p->cap) {
    p->cap = p->cap ? 2 * p->cap : 1024;
    p->a = realloc(p->a, p->cap * sizeof(char *)); p->b = realloc(p->b, p->cap * sizeof(char *));
    p->name_a = realloc(p->name_a, p->cap * sizeof(char *)); p->name_b = realloc(p->name_b, p->cap * sizeof(char *));
    p->la = realloc(p->la, p->cap * sizeof(size_t)); p->lb = realloc(p->lb, p->cap * sizeof(size_t));
    if(!p->a || !p->b || !p->name_a || !p->name_b || !p->la || !p->lb) {
      fprintf(stderr, "%s:%i: Out of memory\n", __FILE__, __LINE__); exit(EXIT_FAILURE);
    }
  }
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
}

static inline void sa_pairs_free(sa_pairs *p)
{
  sa_pairs_clear(p);
  free(p->a); free(p->b); free(p->name_a); free(p->name_b); free(p->la); free(p->lb);
  memset(p, 0, sizeof(*p));
}

/* a batch is full at this many pairs or bytes of sequence */
#define SA_BATCH_MAX_PAIRS ((size_t)1 << 16)
#define SA_BATCH_MAX_BYTES ((size_t)256 << 20)

/* Read every pair of one input (path2 == NULL: consecutive records of path1)
 * and call flush() whenever the read-ahead buffer is full, and at the end.
 * interactive: flush after every pair (request/response protocol of the perl
 * wrappers, reference perl/NeedlemanWunsch.pm:182-210).  Messages as
 * reference src/alignment_cmdline.c:584-632. */
static inline void sa_for_each_batch(const char *path1, const char *path2, int interactive, int buffered,
                                     size_t max_pairs, sa_pairs *pairs, void (*flush)(sa_pairs *, sa_reader *))
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

#endif
