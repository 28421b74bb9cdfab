/*
 * nw_main.c -- `needleman_wunsch` command-line tool on the B200 batch engine.
 *
 * Same flags and the same bytes on stdout as the reference tool (reference
 * src/tools/nw_cmdline.c:78-149 for the output layout, :158-196 for main);
 * the difference is the loop: pairs read from files are aligned as batches
 * (one launch sequence per batch, score + traceback on the device) instead of
 * one needleman_wunsch_align() call per pair.  --stdin keeps the per-pair
 * request/response rhythm the perl wrapper depends on.  --printmatrices over
 * several pairs takes the batch materialise mode (SEQALIGN_MODE_MATS, NW rows);
 * single pairs and scoring shapes that mode refuses go through the single-pair API.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <stdio.h>
#include <stdio_ext.h>
#include <stdlib.h>
#include <string.h>

#include "needleman_wunsch.h"
#include "seqalign_b200.h"
#include "sa_cli.h"
#include "sa_batch.h"

static sa_opts opt;
static scoring_t scoring;
static nw_aligner_t *nw;
static alignment_t *result;
static seqalign_batch_t *eng;
static seqalign_batch_t *mats_eng;   /* --printmatrices: the batch's matrices live on a second engine */

/* "Br1:/Br2:" layout of --zam (reference nw_cmdline.c:36-75) */
static void print_zam(void)
{
  int mismatches = 0, indels = 0;
  for(char *p = result->result_a, *q = result->result_b; *p; p++, q++) {
    if(*p == '-') *p = '_';
    if(*q == '-') *q = '_';
  }
  const size_t len = strlen(result->result_a);
  sa_out_lit("Br1:"); sa_out_mem(result->result_a, len); sa_out_lit("\n    ");
  char *mid = sa_out_room(len);
  for(size_t i = 0; i < len; i++) {
    const char x = result->result_a[i], y = result->result_b[i];
    if(x == '_' || y == '_') { mid[i] = ' '; indels++; }
    else if((scoring.case_sensitive && x != y) || tolower(x) != tolower(y)) { mid[i] = '*'; mismatches++; }
    else mid[i] = '|';
  }
  sa_ob_n += len;
  sa_out_lit("\nBr2:"); sa_out_str(result->result_b); sa_out_chr('\n');
  sa_out_long(mismatches); sa_out_chr(' '); sa_out_long(indels); sa_out_lit("\n\n");
}

/* one aligned pair from `result` (reference nw_cmdline.c:94-148) */
static void print_pair(const char *name_a, const char *name_b)
{
  if(opt.zam) { print_zam(); if(opt.interactive) sa_out_flush(); return; }
  if(opt.print_fasta && name_a) { sa_out_str(name_a); sa_out_chr('\n'); }
  if(opt.print_fasta && opt.print_pretty && name_b) { sa_out_str(name_b); sa_out_chr('\n'); }
  if(opt.print_colour) { sa_out_sync(); alignment_colour_print_against(result->result_a, result->result_b, scoring.case_sensitive); }
  else sa_out_str(result->result_a);
  sa_out_chr('\n');
  if(opt.print_pretty) { sa_out_sync(); alignment_print_spacer(result->result_a, result->result_b, &scoring); sa_out_chr('\n'); }
  else if(opt.print_fasta && name_b) { sa_out_str(name_b); sa_out_chr('\n'); }
  if(opt.print_colour) { sa_out_sync(); alignment_colour_print_against(result->result_b, result->result_a, scoring.case_sensitive); }
  else sa_out_str(result->result_b);
  sa_out_chr('\n');
  if(opt.print_scores) { sa_out_lit("score: "); sa_out_long(result->score); sa_out_chr('\n'); }
  sa_out_chr('\n');
  /* the reference flushes after every pair; a batch is flushed once (same
   * bytes), --stdin keeps the per-pair flush its callers wait for */
  if(opt.interactive) sa_out_flush();
}

/* single-pair API: fills nw's matrices too (for --printmatrices), and reports
 * unknown character pairs exactly where the reference would stop */
static void align_single(const char *a, const char *b, const char *name_a, const char *name_b)
{
  needleman_wunsch_align(a, b, &scoring, nw, result);
  if(opt.print_matrices && !opt.zam) { sa_out_sync(); alignment_print_matrices(nw); }
  print_pair(name_a, name_b);
}

/* --printmatrices over a batch: pair i's three matrices from the batch materialise mode, dressed
 * as the aligner_t aligner_align() would leave behind, for alignment_print_matrices */
static void print_batch_matrices(size_t i, const char *a, size_t la, const char *b, size_t lb)
{
  aligner_t tmp;
  memset(&tmp, 0, sizeof(tmp));
  const size_t cells = (la + 1) * (lb + 1);
  tmp.scoring = &scoring; tmp.seq_a = a; tmp.seq_b = b;
  tmp.score_width = la + 1; tmp.score_height = lb + 1; tmp.capacity = cells;
  tmp.match_scores = malloc(cells * sizeof(score_t));
  tmp.gap_a_scores = malloc(cells * sizeof(score_t));
  tmp.gap_b_scores = malloc(cells * sizeof(score_t));
  if(!tmp.match_scores || !tmp.gap_a_scores || !tmp.gap_b_scores ||
     seqalign_batch_matrices(mats_eng, i, tmp.match_scores, tmp.gap_a_scores, tmp.gap_b_scores) != SEQALIGN_OK) {
    fprintf(stderr, "Error: %s\n", seqalign_batch_error(mats_eng));
    exit(EXIT_FAILURE);
  }
  sa_out_sync();
  alignment_print_matrices(&tmp);
  free(tmp.match_scores); free(tmp.gap_a_scores); free(tmp.gap_b_scores);
}

static void align_batch(sa_pairs *p)
{
  const size_t n = p->n;
  const size_t *la = p->la, *lb = p->lb;
  char *const *name_a = p->name_a, *const *name_b = p->name_b;
  if(n == 0) return;
  sa_engine_wait();
  int rc = SEQALIGN_ERR_ARG;
  double t0 = sa_now();
  /* matrices that nobody prints (--zam) are not made; several pairs with --printmatrices take the
   * batch materialise mode on a second engine when the scoring shape allows, else pair by pair */
  const int want_mats = opt.print_matrices && !opt.zam;
  int with_mats = 0;
  if(want_mats && n > 1) {
    if(!mats_eng) {
      mats_eng = seqalign_batch_create(0);
      if(mats_eng) seqalign_batch_set_scoring(mats_eng, &scoring);
    }
    with_mats = mats_eng && sa_submit(mats_eng, SEQALIGN_NW, SEQALIGN_MODE_MATS, p) == SEQALIGN_OK;
  }
  if(!opt.print_matrices || with_mats) rc = sa_main_submit(SEQALIGN_NW, SEQALIGN_MODE_ALIGN, p);
  sa_t_align += sa_now() - t0;
  if(rc == SEQALIGN_OK) {
    t0 = sa_now();
    if(with_mats) sa_pairs_host(p);   /* the matrix printer shows the sequences */
    for(size_t i = 0; i < n; i++) {
      if(with_mats) print_batch_matrices(i, p->a[i], la[i], p->b[i], lb[i]);
      alignment_ensure_capacity(result, la[i] + lb[i]);
      rc = sa_main_alignment(i, result);
      if(rc < 0) { fprintf(stderr, "Error: %s\n", sa_main_error()); exit(EXIT_FAILURE); }
      print_pair(name_a ? name_a[i] : NULL, name_b ? name_b[i] : NULL);
    }
    sa_t_print += sa_now() - t0;
    return;
  }
  if(!opt.print_matrices && rc != SEQALIGN_ERR_UNKNOWN_PAIR) {
    fprintf(stderr, "Error: %s\n", sa_main_error());
    exit(EXIT_FAILURE);
  }
  /* pair by pair: prints everything up to the offending pair, then the
   * reference's "Unknown character pair" message and exit(EXIT_FAILURE) */
  sa_pairs_host(p);
  for(size_t i = 0; i < n; i++) align_single(p->a[i], p->b[i], name_a ? name_a[i] : NULL, name_b ? name_b[i] : NULL);
}

static void flush_pairs(sa_pairs *p, sa_reader *r)
{
  (void)r;
  align_batch(p);
  sa_out_flush();
  sa_pairs_clear(p);
}

int main(int argc, char **argv)
{
  scoring_system_default(&scoring);
  sa_cli_parse(argc, argv, &scoring, SA_TOOL_NW, &opt);

  if(!opt.interactive) {
    /* millions of short writes per run: a 1 MB buffer and no per-call locking (the helper thread that brings up
     * the engine never touches stdout); --stdin keeps line-by-line answers */
    setvbuf(stdout, NULL, _IOFBF, 1 << 20);
    __fsetlocking(stdout, FSETLOCKING_BYCALLER);
  }
  sa_out_init();
  sa_t_start = sa_now();
  sa_gpus = opt.gpus_set ? opt.gpus : 1;
  sa_engine_start(&eng, &scoring);   /* the CUDA context comes up while the first input is opened and read */
  nw = needleman_wunsch_new();
  result = alignment_create(256);

  if(opt.seq1) {
    sa_pairs one;
    memset(&one, 0, sizeof(one));
    sa_pairs_reserve(&one, 1);
    one.a[0] = sa_dup(opt.seq1, strlen(opt.seq1)); one.la[0] = strlen(opt.seq1);
    one.b[0] = sa_dup(opt.seq2, strlen(opt.seq2)); one.lb[0] = strlen(opt.seq2);
    one.n = 1;
    align_batch(&one);
    sa_out_flush();
    sa_pairs_free(&one);
  }
  sa_pairs pairs;
  memset(&pairs, 0, sizeof(pairs));
  for(size_t i = 0; i < opt.nfiles; i++) {
    const char *f1 = opt.files[i].path1, *f2 = opt.files[i].path2;
    if(f1 && *f1 == '\0' && !f2) f1 = "-";
    sa_read_input(f1, f2, opt.interactive, 0, SA_BATCH_MAX_PAIRS, opt.print_fasta, &pairs, flush_pairs);
  }
  sa_pairs_free(&pairs);
  needleman_wunsch_free(nw);
  alignment_free(result);
  sa_engine_wait();
  sa_main_destroy();
  if(mats_eng) seqalign_batch_destroy(mats_eng);
  sa_cli_free(&opt);
  sa_timing_report();
  return EXIT_SUCCESS;
}
