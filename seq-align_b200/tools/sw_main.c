/*
 * sw_main.c -- `smith_waterman` command-line tool on the B200 batch engine.
 *
 * Same flags and stdout as the reference tool (reference
 * src/tools/sw_cmdline.c:125-314 for the per-pair output, :323-362 for main).
 * Pairs read from files are aligned as batches in the engine's multi-hit mode
 * (fill, candidate sort, masked walks all on the device: the whole of
 * smith_waterman_align2 + the fetch loop, reference smith_waterman.c:137-277);
 * the tool then prints each pair's hits down to its own --minscore default.
 * The single-pair API is used where the batch mode does not apply: --stdin
 * (hits are fetched one keystroke at a time), --printmatrices, scoring shapes
 * outside the specialised kernel, and pairs whose hit list filled the
 * per-pair cap.  Every pair gets a fresh visited mask (the reference's reuse
 * of a partly cleared mask across pairs, smith_waterman.c:149, is not
 * reproduced -- DESIGN.md "known upstream defects").
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdio_ext.h>
#include <stdlib.h>
#include <string.h>

#include "smith_waterman.h"
#include "seqalign_b200.h"
#include "sa_cli.h"
#include "sa_batch.h"

static sa_opts opt;
static scoring_t scoring;
static sw_aligner_t *sw;
static alignment_t *result;
static seqalign_batch_t *eng;
static seqalign_batch_t *mats_eng;   /* --printmatrices: the batch's matrices live on a second engine */
static size_t alignment_index = 0;
static int wait_on_keystroke = 0;
static sa_reader *prompt_input = NULL;

/* hits kept per pair by the device when --maxhits is absent or large (a pair
 * that fills its list is redone through the single-pair API), and pairs per
 * batch: the engine reserves cap x (len_a + len_b) bytes of strings per pair */
#define HIT_CAP 8
#define SW_BATCH_PAIRS ((size_t)1 << 14)

static size_t zmax(size_t a, size_t b) { return a > b ? a : b; }
static size_t zmin(size_t a, size_t b) { return a < b ? a : b; }

/* one sequence line of a hit (reference sw_cmdline.c:49-82) */
static void print_part(const char *row, const char *other, size_t pos, size_t len, const char *whole,
                       size_t spaces_left, size_t spaces_right, size_t ctx_left, size_t ctx_right)
{
  sa_out_lit("  ");
  sa_out_fill(' ', spaces_left);
  if(opt.print_colour) {
    /* the library's colour printer writes to stdout itself */
    sa_out_sync();
    if(ctx_left > 0) {
      fputs(align_col_context, stdout);
      printf("%.*s", (int)ctx_left, whole + pos - ctx_left);
      fputs(align_col_stop, stdout);
    }
    alignment_colour_print_against(row, other, scoring.case_sensitive);
    if(ctx_right > 0) {
      fputs(align_col_context, stdout);
      printf("%.*s", (int)ctx_right, whole + pos + len);
      fputs(align_col_stop, stdout);
    }
  } else {
    if(ctx_left > 0) sa_out_mem(whole + pos - ctx_left, strnlen(whole + pos - ctx_left, ctx_left));
    sa_out_str(row);
    if(ctx_right > 0) sa_out_mem(whole + pos + len, strnlen(whole + pos + len, ctx_right));
  }
  sa_out_fill(' ', spaces_right);
  sa_out_lit("  [pos: "); sa_out_long((long)pos);
  sa_out_lit("; len: "); sa_out_ulong((unsigned long)len);
  sa_out_lit("]\n");
}

/* the hit in `result` (reference sw_cmdline.c:219-306) */
static void print_hit(const char *seq_a, const char *seq_b, size_t len_a, size_t len_b, size_t hit_index)
{
  sa_out_lit("hit "); sa_out_ulong(alignment_index); sa_out_chr('.'); sa_out_ulong(hit_index);
  sa_out_lit(" score: "); sa_out_long(result->score); sa_out_chr('\n');
  size_t ctx_l = 0, ctx_r = 0, ls_a = 0, ls_b = 0, rs_a = 0, rs_b = 0;
  if(opt.context) {
    ctx_l = zmin(zmax(result->pos_a, result->pos_b), opt.context);
    const size_t rem_a = len_a - (result->pos_a + result->len_a), rem_b = len_b - (result->pos_b + result->len_b);
    ctx_r = zmin(zmax(rem_a, rem_b), opt.context);
    ls_a = ctx_l > result->pos_a ? ctx_l - result->pos_a : 0;
    ls_b = ctx_l > result->pos_b ? ctx_l - result->pos_b : 0;
    rs_a = ctx_r > rem_a ? ctx_r - rem_a : 0;
    rs_b = ctx_r > rem_b ? ctx_r - rem_b : 0;
  }
  print_part(result->result_a, result->result_b, result->pos_a, result->len_a, seq_a, ls_a, rs_a, ctx_l - ls_a, ctx_r - rs_a);
  if(opt.print_pretty) {
    const size_t ml = zmax(ls_a, ls_b), mr = zmax(rs_a, rs_b);
    sa_out_lit("  ");
    sa_out_fill(' ', ml);
    sa_out_fill('.', ctx_l - ml);
    sa_out_sync();
    alignment_print_spacer(result->result_a, result->result_b, &scoring);
    sa_out_fill('.', ctx_r - mr);
    sa_out_fill(' ', mr);
    sa_out_chr('\n');
  }
  print_part(result->result_b, result->result_a, result->pos_b, result->len_b, seq_b, ls_b, rs_b, ctx_l - ls_b, ctx_r - rs_b);
  sa_out_chr('\n');
  /* flushed per hit where someone waits for it; a batch is flushed once (same bytes) */
  if(opt.interactive) sa_out_flush();
}

/* pair header up to the blank line (reference sw_cmdline.c:157-190) */
static void print_header(const char *seq_a, const char *seq_b, const char *name_a, const char *name_b,
                         size_t len_a, size_t len_b, aligner_t *matrices)
{
  sa_out_lit("== Alignment "); sa_out_ulong(alignment_index);
  sa_out_lit(" lengths ("); sa_out_ulong((unsigned long)len_a); sa_out_lit(", "); sa_out_ulong((unsigned long)len_b);
  sa_out_lit("):\n");
  if(matrices) { sa_out_sync(); alignment_print_matrices(matrices); }
  if(opt.print_fasta && name_a) { sa_out_str(name_a); sa_out_chr('\n'); }
  if(opt.print_seq) { sa_out_str(seq_a); sa_out_chr('\n'); }
  if(opt.print_fasta && name_b) { sa_out_str(name_b); sa_out_chr('\n'); }
  if(opt.print_seq) { sa_out_str(seq_b); sa_out_chr('\n'); }
  sa_out_chr('\n');
}

/* --minscore default of a pair (reference sw_cmdline.c:192-202) */
static int pair_min_score(size_t len_a, size_t len_b)
{
  if(opt.min_score_set) return opt.min_score;
  if(wait_on_keystroke) return 0;
  const double frac = 0.2 * (double)zmin(len_a, len_b);
  return (int)(scoring.match * (frac >= 2 ? frac : 2));
}

/* interactive prompt between hits (reference sw_cmdline.c:84-122) */
static int next_hit_wanted(void)
{
  if(!wait_on_keystroke) return 1;
  int r = 0, answered = 0, next = 0;
  while(!answered) {
    sa_out_lit("next [h]it or [a]lignment: ");
    sa_out_flush();
    while((r = sa_reader_getc(prompt_input)) != -1 && r != '\n' && r != '\r') {
      if(r == 'h' || r == 'H') { next = 1; answered = 1; }
      else if(r == 'a' || r == 'A') { next = 0; answered = 1; }
    }
    if(r == -1) { sa_out_chr('\n'); exit(EXIT_SUCCESS); }
  }
  return next;
}

static int rejects_pair(size_t len_a, size_t len_b, const char *name_a, const char *name_b)
{
  if((name_a || name_b) && wait_on_keystroke) {
    fprintf(stderr, "Error: Interactive input takes seq only (no FASTA/FASTQ) '%s:%s'\n", name_a, name_b);
    fflush(stderr);
    exit(EXIT_FAILURE);
  }
  if(len_a == 0 || len_b == 0) {
    fprintf(stderr, "Error: Sequences must have length > 0\n");
    fflush(stderr);
    if(opt.print_fasta && name_a && name_b) fprintf(stderr, "%s\n%s\n", name_a, name_b);
    fflush(stderr);
    return 1;
  }
  return 0;
}

/* single-pair API: smith_waterman_align + fetch loop, hit by hit */
static void align_single(const char *seq_a, const char *seq_b, const char *name_a, const char *name_b)
{
  if(rejects_pair(strlen(seq_a), strlen(seq_b), name_a, name_b)) return;
  smith_waterman_align(seq_a, seq_b, &scoring, sw);
  aligner_t *al = smith_waterman_get_aligner(sw);
  const size_t len_a = al->score_width - 1, len_b = al->score_height - 1;
  print_header(seq_a, seq_b, name_a, name_b, len_a, len_b, opt.print_matrices ? al : NULL);
  const int min_score = pair_min_score(len_a, len_b);
  sa_out_flush();
  size_t hit_index = 0;
  while(next_hit_wanted() && smith_waterman_fetch(sw, result) && result->score >= min_score &&
        (!opt.max_hits_set || hit_index < opt.max_hits))
    print_hit(seq_a, seq_b, len_a, len_b, hit_index++);
  sa_out_lit("==\n");
  sa_out_flush();
  alignment_index++;
}

static void align_batch(sa_pairs *p)
{
  const size_t n = p->n;
  const size_t *la = p->la, *lb = p->lb;
  char *const *name_a = p->name_a, *const *name_b = p->name_b;
  if(n == 0) return;
  sa_engine_wait();
  int batch_ok = !wait_on_keystroke;
  size_t cap = HIT_CAP;
  /* --printmatrices: all three matrices of every pair from the batch materialise mode
   * (device resident, copied out pair by pair while printing); the single-pair API remains
   * the way out for scoring shapes that mode does not take */
  int with_mats = 0;
  if(batch_ok && opt.print_matrices) {
    if(!mats_eng) {
      mats_eng = seqalign_batch_create(0);
      if(mats_eng) seqalign_batch_set_scoring(mats_eng, &scoring);
    }
    with_mats = mats_eng && n > 1 &&
                sa_submit(mats_eng, SEQALIGN_SW, SEQALIGN_MODE_MATS, p) == SEQALIGN_OK;
    if(!with_mats) batch_ok = 0;
  }
  /* --maxhits 1: only the first fetch matters, and that is what align mode delivers (best cell
   * under the hit order + traceback, one fill pass, every scoring shape): no candidate sort */
  const int first_only = batch_ok && opt.max_hits_set && opt.max_hits == 1;
  if(first_only) {
    const double t0 = sa_now();
    const int rc = sa_main_submit(SEQALIGN_SW, SEQALIGN_MODE_ALIGN, p);
    sa_t_align += sa_now() - t0;
    if(rc == SEQALIGN_ERR_UNKNOWN_PAIR) batch_ok = 0;
    else if(rc != SEQALIGN_OK) { fprintf(stderr, "Error: %s\n", sa_main_error()); exit(EXIT_FAILURE); }
  } else if(batch_ok) {
    /* empty sequences are reported at print time; the engine sees them as pairs without hits */
    int min_all = 0, have = 0;
    for(size_t i = 0; i < n; i++) {
      if(la[i] == 0 || lb[i] == 0) continue;
      const int m = pair_min_score(la[i], lb[i]);
      if(!have || m < min_all) { min_all = m; have = 1; }
    }
    if(opt.max_hits_set && opt.max_hits < HIT_CAP) cap = opt.max_hits ? opt.max_hits : 1;
    sa_main_set_hit_limits(cap, min_all < 1 ? 1 : min_all);
    const double t0 = sa_now();
    const int rc = sa_main_submit(SEQALIGN_SW, SEQALIGN_MODE_HITS, p);
    sa_t_align += sa_now() - t0;
    if(rc == SEQALIGN_ERR_ARG || rc == SEQALIGN_ERR_UNKNOWN_PAIR) batch_ok = 0; /* single-pair API handles both */
    else if(rc != SEQALIGN_OK) { fprintf(stderr, "Error: %s\n", sa_main_error()); exit(EXIT_FAILURE); }
  }
  const double t1 = sa_now();
  /* the sequences themselves are only needed on the host to print them (--printseq, --context, the matrix
   * printer) or to go pair by pair; a device-decoded batch fetches them then, and only then */
  if(!batch_ok || opt.print_seq || opt.context || with_mats) sa_pairs_host(p);
  char *const *a = p->a, *const *b = p->b;
  for(size_t i = 0; i < n; i++) {
    const char *na = name_a ? name_a[i] : NULL, *nb = name_b ? name_b[i] : NULL;
    if(!batch_ok) { align_single(a[i], b[i], na, nb); continue; }
    if(rejects_pair(la[i], lb[i], na, nb)) continue;
    const int min_score = pair_min_score(la[i], lb[i]);
    aligner_t tmp;
    aligner_t *mats = NULL;
    if(with_mats) {
      /* an aligner_t exactly as aligner_align() would leave it, for alignment_print_matrices */
      memset(&tmp, 0, sizeof(tmp));
      const size_t cells = (la[i] + 1) * (lb[i] + 1);
      tmp.scoring = &scoring; tmp.seq_a = a[i]; tmp.seq_b = b[i];
      tmp.score_width = la[i] + 1; tmp.score_height = lb[i] + 1; tmp.capacity = cells;
      tmp.match_scores = malloc(cells * sizeof(score_t));
      tmp.gap_a_scores = malloc(cells * sizeof(score_t));
      tmp.gap_b_scores = malloc(cells * sizeof(score_t));
      if(!tmp.match_scores || !tmp.gap_a_scores || !tmp.gap_b_scores ||
         seqalign_batch_matrices(mats_eng, i, tmp.match_scores, tmp.gap_a_scores, tmp.gap_b_scores) != SEQALIGN_OK) {
        fprintf(stderr, "Error: %s\n", seqalign_batch_error(mats_eng));
        exit(EXIT_FAILURE);
      }
      mats = &tmp;
    }
    size_t nh = 0;
    const size_t want = opt.max_hits_set ? opt.max_hits : (size_t)-1;
    if(!first_only) {
      nh = sa_main_hit_count(i);
      /* the device list is complete unless it is full and the caller wants more */
      if(nh == cap && want > cap) {
        sa_main_hit(i, nh - 1, result);
        if(result->score >= min_score) {
          if(mats) { free(tmp.match_scores); free(tmp.gap_a_scores); free(tmp.gap_b_scores); }
          sa_pairs_host(p);
          align_single(p->a[i], p->b[i], na, nb);
          continue;
        }
      }
    }
    print_header(a[i], b[i], na, nb, la[i], lb[i], mats);
    if(mats) { free(tmp.match_scores); free(tmp.gap_a_scores); free(tmp.gap_b_scores); }
    if(first_only) {
      alignment_ensure_capacity(result, la[i] + lb[i]);
      const int got = sa_main_alignment(i, result);
      if(got < 0) { fprintf(stderr, "Error: %s\n", sa_main_error()); exit(EXIT_FAILURE); }
      if(got == 1 && result->score >= min_score) print_hit(a[i], b[i], la[i], lb[i], 0);
    } else {
      size_t hit_index = 0;
      for(size_t h = 0; h < nh && hit_index < want; h++) {
        if(sa_main_hit(i, h, result) != 1) break;
        if(result->score < min_score) break;
        print_hit(a[i], b[i], la[i], lb[i], hit_index++);
      }
    }
    sa_out_lit("==\n");
    alignment_index++;
  }
  sa_out_flush();
  sa_t_print += sa_now() - t1;
}

static void flush_pairs(sa_pairs *p, sa_reader *r)
{
  prompt_input = r;
  align_batch(p);
  sa_pairs_clear(p);
}

int main(int argc, char **argv)
{
  /* smith_waterman's own defaults (reference sw_cmdline.c:37-46) */
  scoring_system_default(&scoring);
  scoring.match = 2;
  scoring.mismatch = -2;
  scoring.gap_open = -2;
  scoring.gap_extend = -1;
  sa_cli_parse(argc, argv, &scoring, SA_TOOL_SW, &opt);

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
  sw = smith_waterman_new();
  result = alignment_create(256);

  if(opt.seq1) {
    sa_pairs one;
    memset(&one, 0, sizeof(one));
    sa_pairs_reserve(&one, 1);
    one.a[0] = sa_dup(opt.seq1, strlen(opt.seq1)); one.la[0] = strlen(opt.seq1);
    one.b[0] = sa_dup(opt.seq2, strlen(opt.seq2)); one.lb[0] = strlen(opt.seq2);
    one.n = 1;
    align_batch(&one);
    sa_pairs_free(&one);
  }
  sa_pairs pairs;
  memset(&pairs, 0, sizeof(pairs));
  for(size_t i = 0; i < opt.nfiles; i++) {
    const char *f1 = opt.files[i].path1, *f2 = opt.files[i].path2;
    if(f1 && *f1 == '\0' && !f2) { wait_on_keystroke = 1; f1 = "-"; }
    /* --maxhits 1 runs in align mode (one string pair per pair): large batches; the multi-hit mode reserves
     * HIT_CAP string pairs and 16 bytes of candidate keys per cell and stays at SW_BATCH_PAIRS */
    const size_t batch_pairs = (opt.max_hits_set && opt.max_hits == 1 && !opt.print_matrices) ? SA_BATCH_MAX_PAIRS : SW_BATCH_PAIRS;
    sa_read_input(f1, f2, opt.interactive, 0, batch_pairs, opt.print_fasta, &pairs, flush_pairs);
  }
  sa_pairs_free(&pairs);
  smith_waterman_free(sw);
  alignment_free(result);
  sa_engine_wait();
  sa_main_destroy();
  if(mats_eng) seqalign_batch_destroy(mats_eng);
  sa_cli_free(&opt);
  sa_timing_report();
  return EXIT_SUCCESS;
}
