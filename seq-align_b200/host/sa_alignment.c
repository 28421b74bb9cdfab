/*
 * sa_alignment.c -- aligner_t / alignment_t lifecycle, the DP entry point and
 * the printers of the seq-align C API (host code around the GPU engine).
 *
 * Implements include/alignment.h.  aligner_align() is where the reference
 * ran its host loop (src/alignment.c:170-193 -> :28-168); here it hands the
 * pair to the batch engine in materialise mode.  There is no host DP: if no
 * B200 is usable the call prints the engine's message and exits, following
 * the reference's error convention (stderr + exit(EXIT_FAILURE)).
 */
#include <ctype.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "alignment.h"
#include "seqalign_b200.h"
#include "sa_host.h"

const char align_col_mismatch[] = "\033[92m";
const char align_col_indel[] = "\033[91m";
const char align_col_context[] = "\033[95m";
const char align_col_stop[] = "\033[0m";

/* The engine of the single-pair API.  The reference's L0-L2 code has no global
 * state: distinct aligner objects may be used from distinct threads
 * (SURVEY.md 8b "Threading").  An engine keeps per-submit state (device and
 * pinned buffers, the last results), so every calling thread gets its own,
 * created on first use and destroyed when the thread exits.  SEQALIGN_DEVICE
 * selects the device (the batch API is the multi-GPU path). */
static pthread_key_t g_engine_key;
static pthread_once_t g_engine_once = PTHREAD_ONCE_INIT;

static void engine_key_destroy(void *p) { seqalign_batch_destroy((seqalign_batch_t *)p); }
static void engine_key_make(void) { pthread_key_create(&g_engine_key, engine_key_destroy); }

seqalign_batch_t *sa_host_engine(void)
{
  pthread_once(&g_engine_once, engine_key_make);
  seqalign_batch_t *eng = pthread_getspecific(g_engine_key);
  if(!eng) {
    const char *dev = getenv("SEQALIGN_DEVICE");
    eng = seqalign_batch_create(dev ? atoi(dev) : 0);
    if(!eng) {
      fprintf(stderr, "seq-align (B200): %s\n", seqalign_last_create_error());
      exit(EXIT_FAILURE);
    }
    pthread_setspecific(g_engine_key, eng);
  }
  return eng;
}

void sa_host_check(seqalign_batch_t *eng, int rc)
{
  if(rc >= 0) return;
  if(rc == SEQALIGN_ERR_TRACEBACK) {
    /* reference alignment.c:341-347 */
    fprintf(stderr,
"Program error: traceback fail (get_reverse_move)\n"
"This may be due to an integer overflow if your sequences are long or scores\n"
"are large. If this is the case using smaller scores or shorter sequences may\n"
"work around this problem.  \n"
"  If you think this is a bug, please report it to: turner.isaac@gmail.com\n");
  } else {
    /* unknown pair: the engine's message is the reference's text */
    fprintf(stderr, "%s\n", seqalign_batch_error(eng));
  }
  exit(EXIT_FAILURE);
}

/* ---- deferred matrices ---------------------------------------------------
 * needleman_wunsch_align* and smith_waterman_align* get their results (score, gapped strings, hit
 * lists) straight from the engine's align / multi-hit modes; the three (len_a+1) x (len_b+1) int32
 * matrices of the reference's aligner_t (12 bytes per cell over PCIe: 1.2 GB for one 10k x 10k pair)
 * are only needed by callers that LOOK at them.  Everything in this library that does --
 * alignment_print_matrices, alignment_reverse_move, the host hit iteration -- fills them on demand
 * (sa_host_materialise).  aligner_align() itself, the reference's seam, always fills at once.
 * Code that reads aligner->match_scores[] directly after needleman_wunsch_align() asks for the
 * reference's eager behaviour with seqalign_host_eager_matrices(1) or SEQALIGN_EAGER_MATRICES=1. */
#define SA_DEFER_SLOTS 64
static struct { const aligner_t *al; char is_sw; } g_defer[SA_DEFER_SLOTS];
static volatile int g_deferred = 0;   /* entries in use: lets the hot callers skip the lock */
static pthread_mutex_t g_defer_lock = PTHREAD_MUTEX_INITIALIZER;
static int g_eager = -1;

void seqalign_host_eager_matrices(int on) { g_eager = on ? 1 : 0; }

static int eager_matrices(void)
{
  if(g_eager < 0) {
    const char *e = getenv("SEQALIGN_EAGER_MATRICES");
    g_eager = e && e[0] == '1';
  }
  return g_eager;
}

static int defer_forget(const aligner_t *al)
{
  int was = -1;
  if(g_deferred == 0) return was;
  pthread_mutex_lock(&g_defer_lock);
  for(int i = 0; i < SA_DEFER_SLOTS; i++)
    if(g_defer[i].al == al) { was = g_defer[i].is_sw; g_defer[i].al = NULL; g_deferred--; break; }
  pthread_mutex_unlock(&g_defer_lock);
  return was;
}

void sa_host_bind(aligner_t *aligner, const char *seq_a, const char *seq_b, size_t len_a, size_t len_b,
                  const scoring_t *scoring, char is_sw)
{
  if(eager_matrices()) { aligner_align(aligner, seq_a, seq_b, len_a, len_b, scoring, is_sw); return; }
  defer_forget(aligner);
  aligner->scoring = scoring;
  aligner->seq_a = seq_a;
  aligner->seq_b = seq_b;
  aligner->score_width = len_a + 1;
  aligner->score_height = len_b + 1;
  pthread_mutex_lock(&g_defer_lock);
  int slot = -1;
  for(int i = 0; i < SA_DEFER_SLOTS && slot < 0; i++) if(!g_defer[i].al) slot = i;
  if(slot >= 0) { g_defer[slot].al = aligner; g_defer[slot].is_sw = is_sw; g_deferred++; }
  pthread_mutex_unlock(&g_defer_lock);
  /* more aligners with pending matrices than slots: this one is filled now */
  if(slot < 0) aligner_align(aligner, seq_a, seq_b, len_a, len_b, scoring, is_sw);
}

void sa_host_materialise(const aligner_t *al)
{
  const int is_sw = defer_forget(al);
  if(is_sw < 0) return;   /* matrices are current */
  aligner_t *w = (aligner_t *)al;
  aligner_align(w, w->seq_a, w->seq_b, w->score_width - 1, w->score_height - 1, w->scoring, (char)is_sw);
}

void aligner_align(aligner_t *aligner, const char *seq_a, const char *seq_b,
                   size_t len_a, size_t len_b, const scoring_t *scoring, char is_sw)
{
  defer_forget(aligner);
  aligner->scoring = scoring;
  aligner->seq_a = seq_a;
  aligner->seq_b = seq_b;
  aligner->score_width = len_a + 1;
  aligner->score_height = len_b + 1;

  /* grow like the reference (alignment.c:181-190): power-of-two cells */
  const size_t cells = aligner->score_width * aligner->score_height;
  if(aligner->capacity < cells) {
    aligner->capacity = ROUNDUP2POW(cells);
    const size_t mem = sizeof(score_t) * aligner->capacity;
    aligner->match_scores = realloc(aligner->match_scores, mem);
    aligner->gap_a_scores = realloc(aligner->gap_a_scores, mem);
    aligner->gap_b_scores = realloc(aligner->gap_b_scores, mem);
    if(!aligner->match_scores || !aligner->gap_a_scores || !aligner->gap_b_scores) {
      fprintf(stderr, "%s:%i: Out of memory\n", __FILE__, __LINE__);
      exit(EXIT_FAILURE);
    }
  }

  seqalign_batch_t *eng = sa_host_engine();
  seqalign_batch_set_scoring(eng, scoring);
  sa_host_check(eng, seqalign_fill_matrices(eng, seq_a, len_a, seq_b, len_b, is_sw,
                                            aligner->match_scores, aligner->gap_a_scores,
                                            aligner->gap_b_scores));
}

void aligner_destroy(aligner_t *aligner)
{
  defer_forget(aligner);
  if(aligner->capacity > 0) {
    free(aligner->match_scores);
    free(aligner->gap_a_scores);
    free(aligner->gap_b_scores);
  }
}

alignment_t *alignment_create(size_t capacity)
{
  capacity = ROUNDUP2POW(capacity);
  alignment_t *r = malloc(sizeof(alignment_t));
  r->result_a = malloc(capacity);
  r->result_b = malloc(capacity);
  if(!r->result_a || !r->result_b) {
    fprintf(stderr, "%s:%i: Out of memory\n", __FILE__, __LINE__);
    exit(EXIT_FAILURE);
  }
  r->capacity = capacity;
  r->length = 0;
  r->result_a[0] = r->result_b[0] = '\0';
  r->pos_a = r->pos_b = r->len_a = r->len_b = 0;
  r->score = 0;
  return r;
}

void alignment_ensure_capacity(alignment_t *r, size_t strlength)
{
  size_t capacity = strlength + 1;
  if(r->capacity >= capacity) return;
  capacity = ROUNDUP2POW(capacity);
  r->result_a = realloc(r->result_a, capacity);
  r->result_b = realloc(r->result_b, capacity);
  r->capacity = capacity;
  if(!r->result_a || !r->result_b) {
    fprintf(stderr, "%s:%i: Out of memory\n", __FILE__, __LINE__);
    exit(EXIT_FAILURE);
  }
}

void alignment_free(alignment_t *r)
{
  free(r->result_a);
  free(r->result_b);
  free(r);
}

/* One backward step over the materialised matrices, for callers that walk
 * themselves (and for SW hit iteration).  Reference alignment.c:244-350:
 * candidates are tested in the order GAP_A, GAP_B, MATCH by equality with
 * the current score. */
void alignment_reverse_move(enum Matrix *curr_matrix, score_t *curr_score,
                            size_t *score_x, size_t *score_y,
                            size_t *arr_index, const aligner_t *al)
{
  if(g_deferred) sa_host_materialise(al);
  const scoring_t *sc = al->scoring;
  const size_t la = al->score_width - 1, lb = al->score_height - 1;
  const size_t seq_x = *score_x - 1, seq_y = *score_y - 1;
  bool is_match;
  int sub;
  scoring_lookup(sc, al->seq_a[seq_x], al->seq_b[seq_y], &sub, &is_match);

  /* penalties of leaving a gap state; free in the last column / row */
  int open_a = sc->gap_extend + sc->gap_open, ext_a = sc->gap_extend;
  int open_b = open_a, ext_b = ext_a;
  if(sc->no_end_gap_penalty) {
    if(*score_x == la) open_a = ext_a = 0;
    if(*score_y == lb) open_b = ext_b = 0;
  }
  if(sc->no_start_gap_penalty) {
    if(*score_x == 0) open_a = ext_a = 0;
    if(*score_y == 0) open_b = ext_b = 0;
  }

  int from_m, from_ga, from_gb;
  switch(*curr_matrix) {
    case MATCH:
      from_m = from_ga = from_gb = sub;
      (*score_x)--; (*score_y)--;
      *arr_index -= al->score_width + 1;
      break;
    case GAP_A:
      from_m = from_gb = open_a; from_ga = ext_a;
      (*score_y)--;
      *arr_index -= al->score_width;
      break;
    case GAP_B:
      from_m = from_ga = open_b; from_gb = ext_b;
      (*score_x)--;
      (*arr_index)--;
      break;
    default:
      fprintf(stderr, "Program error: invalid matrix in get_reverse_move()\n");
      fprintf(stderr, "Please submit a bug report to: turner.isaac@gmail.com\n");
      exit(EXIT_FAILURE);
  }

  const bool ok_a = !sc->no_gaps_in_a || *score_x == 0 || *score_x == la;
  const bool ok_b = !sc->no_gaps_in_b || *score_y == 0 || *score_y == lb;
  const size_t k = *arr_index;
  if(ok_a && al->gap_a_scores[k] + from_ga == *curr_score) {
    *curr_matrix = GAP_A; *curr_score = al->gap_a_scores[k];
  } else if(ok_b && al->gap_b_scores[k] + from_gb == *curr_score) {
    *curr_matrix = GAP_B; *curr_score = al->gap_b_scores[k];
  } else if(al->match_scores[k] + from_m == *curr_score) {
    *curr_matrix = MATCH; *curr_score = al->match_scores[k];
  } else {
    alignment_print_matrices(al);
    fprintf(stderr, "[%s:%zu,%zu]: %i [ismatch: %i] '%c' '%c'\n",
            MATRIX_NAME(*curr_matrix), *score_x, *score_y, *curr_score, is_match,
            al->seq_a[seq_x], al->seq_b[seq_y]);
    fprintf(stderr, " Penalties match: %i gap_open: %i gap_extend: %i\n", from_m, from_ga, from_gb);
    fprintf(stderr, " Expected MATCH: %i GAP_A: %i GAP_B: %i\n",
            al->match_scores[k], al->gap_a_scores[k], al->gap_b_scores[k]);
    sa_host_check(NULL, SEQALIGN_ERR_TRACEBACK);
  }
}

static void print_matrix(const char *name, const score_t *m, size_t w, size_t h)
{
  printf("%s:\n", name);
  for(size_t j = 0; j < h; j++) {
    printf("%3i:", (int)j);
    for(size_t i = 0; i < w; i++) printf("\t%3i", (int)m[j * w + i]);
    putc('\n', stdout);
  }
}

/* reference alignment.c:353-403 (byte-for-byte the same text) */
void alignment_print_matrices(const aligner_t *al)
{
  if(g_deferred) sa_host_materialise(al);
  printf("seq_a: %.*s\nseq_b: %.*s\n", (int)al->score_width - 1, al->seq_a,
         (int)al->score_height - 1, al->seq_b);
  print_matrix("match_scores", al->match_scores, al->score_width, al->score_height);
  print_matrix("gap_a_scores", al->gap_a_scores, al->score_width, al->score_height);
  print_matrix("gap_b_scores", al->gap_b_scores, al->score_width, al->score_height);
  printf("match: %i mismatch: %i gapopen: %i gapexend: %i\n", al->scoring->match,
         al->scoring->mismatch, al->scoring->gap_open, al->scoring->gap_extend);
  printf("\n");
}

/* reference alignment.c:405-452: red while b has a gap, green on a mismatch
 * column; escape codes are switched only on state changes */
void alignment_colour_print_against(const char *aln_a, const char *aln_b, char case_sensitive)
{
  int red = 0, green = 0;
  for(int i = 0; aln_a[i] != '\0'; i++) {
    const int indel = aln_b[i] == '-';
    if(indel && !red) { fputs(align_col_indel, stdout); red = 1; }
    else if(!indel && red) { red = 0; fputs(align_col_stop, stdout); }

    const int differ = case_sensitive ? aln_a[i] != aln_b[i]
                                      : tolower(aln_a[i]) != tolower(aln_b[i]);
    const int mismatch = differ && aln_a[i] != '-' && aln_b[i] != '-';
    if(mismatch && !green) { fputs(align_col_mismatch, stdout); green = 1; }
    else if(!mismatch && green) { green = 0; fputs(align_col_stop, stdout); }

    putc(aln_a[i], stdout);
  }
  if(green || red) fputs(align_col_stop, stdout);
}

/* reference alignment.c:455-474: ' ' for a gap, '|' match, '*' mismatch */
void alignment_print_spacer(const char *aln_a, const char *aln_b, const scoring_t *scoring)
{
  for(int i = 0; aln_a[i] != '\0'; i++) {
    char c = '*';
    if(aln_a[i] == '-' || aln_b[i] == '-') c = ' ';
    else if(aln_a[i] == aln_b[i] ||
            (!scoring->case_sensitive && tolower(aln_a[i]) == tolower(aln_b[i]))) c = '|';
    putc(c, stdout);
  }
}
