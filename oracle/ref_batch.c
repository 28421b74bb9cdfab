/*
 * ref_batch.c -- batch driver around the UNMODIFIED reference library
 * (TEST / BENCH INFRASTRUCTURE ONLY).
 *
 * Compiled by oracle/Makefile together with the reference's own sources,
 * taken where they lie under /root/reference, into oracle/_ref/ref_batch.
 * It is the "reference" CPU baseline of bench.py (--impl reference and the
 * cpu_baseline leg) and a second checker for the tests.  The reference is
 * single-threaded; this driver runs one independent aligner object per
 * worker thread over disjoint pair ranges (the reference keeps no global
 * state in its library layers).
 *
 * usage: ref_batch MODE THREADS IN.bin OUT.bin
 *   MODE  fill  aligner_align() only, then the score is read from the
 *               matrices (SW: max match score with the (x asc, y asc) tie
 *               rule; NW: max of the three end cells)
 *         full  the complete public call: SW = fresh sw_aligner_t +
 *               smith_waterman_align + first smith_waterman_fetch;
 *               NW = needleman_wunsch_align
 *
 * IN.bin  : int64 header[16] = {magic, n, is_sw, preset, match, mismatch,
 *           gap_open, gap_extend, no_start, no_end, no_gaps_a, no_gaps_b,
 *           no_mismatches, case_sensitive, poke, 0}
 *           int64 off_a[n+1], int64 off_b[n+1], bytes seq_a, bytes seq_b
 *           preset: 0 scoring_init(args) ; 1 BLOSUM62 ; 2 PAM30 ; 3 PAM70 ;
 *           4 BLOSUM80 ; 5 DNA_hybridization.  poke=1: scoring_system_default
 *           then overwrite match/mismatch/gap_open/gap_extend in place, the
 *           way sw_cmdline.c:37-46 does.
 * OUT.bin : int32 score[n], int32 x_end[n], int32 y_end[n]
 *           (full mode additionally: int32 aln_len[n], then for every pair
 *           result_a and result_b, each NUL-terminated, back to back)
 * stdout  : one JSON line {"seconds":..,"cells":..,"threads":..}
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "needleman_wunsch.h"
#include "smith_waterman.h"

#define RB_MAGIC 0x5345514252454631LL

typedef struct {
  int64_t n, is_sw;
  const int64_t *off_a, *off_b;
  const char *seq_a, *seq_b;
  const scoring_t *scoring;
  int full;
  int32_t *score, *x_end, *y_end, *aln_len;
  char **str_a, **str_b;
} job_t;

typedef struct { job_t *job; int64_t lo, hi; } slice_t;

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void run_fill(job_t *j, int64_t lo, int64_t hi)
{
  aligner_t al;
  aligner_init(&al);
  for(int64_t i = lo; i < hi; i++) {
    size_t la = (size_t)(j->off_a[i + 1] - j->off_a[i]);
    size_t lb = (size_t)(j->off_b[i + 1] - j->off_b[i]);
    aligner_align(&al, j->seq_a + j->off_a[i], j->seq_b + j->off_b[i],
                  la, lb, j->scoring, (char)j->is_sw);
    size_t w = la + 1, cells = w * (lb + 1);
    if(j->is_sw) {
      int best = 0; size_t bx = 0, by = 0;
      for(size_t c = 0; c < cells; c++) {
        int v = al.match_scores[c];
        if(v <= 0) continue;
        size_t x = c % w, y = c / w;
        if(v > best || (v == best && (x < bx || (x == bx && y < by)))) {
          best = v; bx = x; by = y;
        }
      }
      j->score[i] = best; j->x_end[i] = (int32_t)bx; j->y_end[i] = (int32_t)by;
    } else {
      int s = al.match_scores[cells - 1];
      if(al.gap_b_scores[cells - 1] >= s) s = al.gap_b_scores[cells - 1];
      if(al.gap_a_scores[cells - 1] >= s) s = al.gap_a_scores[cells - 1];
      j->score[i] = s; j->x_end[i] = (int32_t)la; j->y_end[i] = (int32_t)lb;
    }
  }
  aligner_destroy(&al);
}

static void run_full(job_t *j, int64_t lo, int64_t hi)
{
  alignment_t *res = alignment_create(256);
  nw_aligner_t *nw = j->is_sw ? NULL : needleman_wunsch_new();
  for(int64_t i = lo; i < hi; i++) {
    size_t la = (size_t)(j->off_a[i + 1] - j->off_a[i]);
    size_t lb = (size_t)(j->off_b[i + 1] - j->off_b[i]);
    const char *a = j->seq_a + j->off_a[i], *b = j->seq_b + j->off_b[i];
    if(j->is_sw) {
      /* fresh aligner per pair: the reused-aligner mask is stale upstream */
      sw_aligner_t *sw = smith_waterman_new();
      smith_waterman_align2(a, b, la, lb, j->scoring, sw);
      if(smith_waterman_fetch(sw, res)) {
        j->score[i] = res->score;
        j->x_end[i] = (int32_t)(res->pos_a + res->len_a);
        j->y_end[i] = (int32_t)(res->pos_b + res->len_b);
      } else {
        res->result_a[0] = res->result_b[0] = '\0'; res->length = 0;
        j->score[i] = 0; j->x_end[i] = j->y_end[i] = 0;
      }
      smith_waterman_free(sw);
    } else {
      needleman_wunsch_align2(a, b, la, lb, j->scoring, nw, res);
      j->score[i] = res->score; j->x_end[i] = (int32_t)la; j->y_end[i] = (int32_t)lb;
    }
    j->aln_len[i] = (int32_t)res->length;
    if(j->str_a) {
      j->str_a[i] = strdup(res->result_a);
      j->str_b[i] = strdup(res->result_b);
    }
  }
  if(nw) needleman_wunsch_free(nw);
  alignment_free(res);
}

static void *worker(void *p)
{
  slice_t *s = (slice_t *)p;
  if(s->job->full) run_full(s->job, s->lo, s->hi);
  else run_fill(s->job, s->lo, s->hi);
  return NULL;
}

int main(int argc, char **argv)
{
  if(argc != 5 && argc != 6) {
    fprintf(stderr, "usage: ref_batch fill|full THREADS IN.bin OUT.bin [nostrings]\n");
    return 2;
  }
  int full = strcmp(argv[1], "full") == 0;
  int threads = atoi(argv[2]);
  int keep_strings = full && argc == 5;
  if(threads < 1) threads = 1;

  FILE *f = fopen(argv[3], "rb");
  if(!f) { perror(argv[3]); return 1; }
  int64_t hdr[16];
  if(fread(hdr, sizeof(hdr), 1, f) != 1 || hdr[0] != RB_MAGIC) {
    fprintf(stderr, "ref_batch: bad header\n"); return 1;
  }
  int64_t n = hdr[1];
  int64_t *off_a = malloc((n + 1) * sizeof(int64_t));
  int64_t *off_b = malloc((n + 1) * sizeof(int64_t));
  if(fread(off_a, sizeof(int64_t), n + 1, f) != (size_t)(n + 1) ||
     fread(off_b, sizeof(int64_t), n + 1, f) != (size_t)(n + 1)) return 1;
  char *seq_a = malloc(off_a[n] + 1), *seq_b = malloc(off_b[n] + 1);
  if(fread(seq_a, 1, off_a[n], f) != (size_t)off_a[n] ||
     fread(seq_b, 1, off_b[n], f) != (size_t)off_b[n]) return 1;
  fclose(f);

  static scoring_t scoring;
  switch(hdr[3]) {
    case 1: scoring_system_BLOSUM62(&scoring); break;
    case 2: scoring_system_PAM30(&scoring); break;
    case 3: scoring_system_PAM70(&scoring); break;
    case 4: scoring_system_BLOSUM80(&scoring); break;
    case 5: scoring_system_DNA_hybridization(&scoring); break;
    default:
      if(hdr[14]) {
        scoring_system_default(&scoring);
        scoring.match = (int)hdr[4]; scoring.mismatch = (int)hdr[5];
        scoring.gap_open = (int)hdr[6]; scoring.gap_extend = (int)hdr[7];
      } else {
        scoring_init(&scoring, (int)hdr[4], (int)hdr[5], (int)hdr[6], (int)hdr[7],
                     hdr[8], hdr[9], hdr[10], hdr[11], hdr[12], hdr[13]);
      }
  }
  if(hdr[3] != 0) {           /* presets still take the positional flags */
    scoring.no_start_gap_penalty = hdr[8]; scoring.no_end_gap_penalty = hdr[9];
  }

  job_t job = {0};
  job.n = n; job.is_sw = hdr[2]; job.off_a = off_a; job.off_b = off_b;
  job.seq_a = seq_a; job.seq_b = seq_b; job.scoring = &scoring; job.full = full;
  job.score = calloc(n, 4); job.x_end = calloc(n, 4); job.y_end = calloc(n, 4);
  job.aln_len = calloc(n, 4);
  if(keep_strings) { job.str_a = calloc(n, sizeof(char *)); job.str_b = calloc(n, sizeof(char *)); }

  pthread_t *tid = malloc(threads * sizeof(*tid));
  slice_t *sl = malloc(threads * sizeof(*sl));
  double t0 = now_s();
  for(int t = 0; t < threads; t++) {
    sl[t].job = &job;
    sl[t].lo = n * t / threads;
    sl[t].hi = n * (t + 1) / threads;
    pthread_create(&tid[t], NULL, worker, &sl[t]);
  }
  for(int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
  double dt = now_s() - t0;

  double cells = 0;
  for(int64_t i = 0; i < n; i++)
    cells += (double)(off_a[i + 1] - off_a[i]) * (double)(off_b[i + 1] - off_b[i]);

  f = fopen(argv[4], "wb");
  if(!f) { perror(argv[4]); return 1; }
  fwrite(job.score, 4, n, f); fwrite(job.x_end, 4, n, f); fwrite(job.y_end, 4, n, f);
  if(full) {
    fwrite(job.aln_len, 4, n, f);
    if(keep_strings)
      for(int64_t i = 0; i < n; i++) {
        fwrite(job.str_a[i], 1, strlen(job.str_a[i]) + 1, f);
        fwrite(job.str_b[i], 1, strlen(job.str_b[i]) + 1, f);
      }
  }
  fclose(f);
  printf("{\"seconds\": %.6f, \"cells\": %.0f, \"threads\": %d, \"mode\": \"%s\"}\n",
         dt, cells, threads, argv[1]);
  return 0;
}
