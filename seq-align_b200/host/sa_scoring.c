/*
 * sa_scoring.c -- scoring model of the seq-align C API (host code).
 *
 * Implements include/alignment_scoring.h.  Behaviour follows reference
 * src/alignment_scoring.c (cited per function); the GPU never sees
 * scoring_t: csrc/sa_flatten.h turns it into a dense table per batch.
 */
#include <ctype.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "alignment_scoring.h"
#include "sa_scoring_tables.h"

static void widen(scoring_t *s, int v)
{
  if(v < s->min_penalty) s->min_penalty = v;
  if(v > s->max_penalty) s->max_penalty = v;
}

/* reference alignment_scoring.c:21-55.  Only the two bitsets are cleared;
 * min/max_penalty cover the gap terms only if some gap is allowed (:51). */
void scoring_init(scoring_t *s, int match, int mismatch, int gap_open, int gap_extend,
                  bool no_start_gap_penalty, bool no_end_gap_penalty,
                  bool no_gaps_in_a, bool no_gaps_in_b,
                  bool no_mismatches, bool case_sensitive)
{
  s->gap_open = gap_open;
  s->gap_extend = gap_extend;
  s->no_start_gap_penalty = no_start_gap_penalty;
  s->no_end_gap_penalty = no_end_gap_penalty;
  s->no_gaps_in_a = no_gaps_in_a;
  s->no_gaps_in_b = no_gaps_in_b;
  s->no_mismatches = no_mismatches;
  s->use_match_mismatch = 1;
  s->match = match;
  s->mismatch = mismatch;
  s->case_sensitive = case_sensitive;
  memset(s->wildcards, 0, sizeof(s->wildcards));
  memset(s->swap_set, 0, sizeof(s->swap_set));

  s->min_penalty = match < mismatch ? match : mismatch;
  s->max_penalty = match > mismatch ? match : mismatch;
  if(!no_gaps_in_a || !no_gaps_in_b) {
    widen(s, gap_open + gap_extend);
    widen(s, gap_extend);
  }
}

/* reference alignment_scoring.c:57-64: folds case unless case_sensitive */
void scoring_add_wildcard(scoring_t *s, char c, int score)
{
  if(!s->case_sensitive) c = (char)tolower(c);
  set_wildcard_bit(s, c);
  s->wildscores[(size_t)c] = score;
  widen(s, score);
}

/* reference alignment_scoring.c:66-72: no case folding here */
void scoring_add_mutation(scoring_t *s, char a, char b, int score)
{
  s->swap_scores[(size_t)a][(size_t)b] = score;
  set_swap_bit(s, a, b);
  widen(s, score);
}

/* reference alignment_scoring.c:74-97: scores is len x len, the score of
 * (str[i], str[j]) sits at scores[j*len + i]; letters are case-folded */
void scoring_add_mutations(scoring_t *s, const char *str, const int *scores,
                           char use_match_mismatch)
{
  const size_t len = strlen(str);
  for(size_t i = 0; i < len; i++) {
    const char a = s->case_sensitive ? str[i] : (char)tolower(str[i]);
    for(size_t j = 0; j < len; j++) {
      const char b = s->case_sensitive ? str[j] : (char)tolower(str[j]);
      scoring_add_mutation(s, a, b, scores[j * len + i]);
    }
  }
  s->use_match_mismatch = use_match_mismatch;
}

/* reference alignment_scoring.c:99-112 (text is part of the CLI surface) */
void scoring_print(const scoring_t *s)
{
  printf("scoring:\n");
  printf("  match: %i; mismatch: %i; (use_match_mismatch: %i)\n",
         s->match, s->mismatch, s->use_match_mismatch);
  printf("  gap_open: %i; gap_extend: %i;\n", s->gap_open, s->gap_extend);
  printf("  no_gaps_in_a: %i; no_gaps_in_b: %i; no_mismatches: %i;\n",
         s->no_gaps_in_a, s->no_gaps_in_b, s->no_mismatches);
  printf("  no_start_gap_penalty: %i; no_end_gap_penalty: %i;\n",
         s->no_start_gap_penalty, s->no_end_gap_penalty);
}

/* reference alignment_scoring.c:115-129 */
static bool wildcard_score(const scoring_t *s, char a, char b, int *score)
{
  int best = INT_MAX;
  if(get_wildcard_bit(s, a)) best = s->wildscores[(size_t)a];
  if(get_wildcard_bit(s, b) && s->wildscores[(size_t)b] < best) best = s->wildscores[(size_t)b];
  if(best != INT_MAX) { *score = best; return true; }
  *score = 0;
  return false;
}

/* reference alignment_scoring.c:133-182.  Host-side single lookup, kept for
 * API users and the host traceback step; the kernels use the flattened
 * table built with the same rules (csrc/sa_flatten.h). */
void scoring_lookup(const scoring_t *s, char a, char b, int *score, bool *is_match)
{
  if(!s->case_sensitive) { a = (char)tolower(a); b = (char)tolower(b); }
  *is_match = (a == b);

  if(s->no_mismatches && !*is_match) {
    *is_match = wildcard_score(s, a, b, score);
    return;
  }
  if(get_swap_bit(s, a, b)) {
    *score = s->swap_scores[(size_t)a][(size_t)b];
    return;
  }
  if(wildcard_score(s, a, b, score)) {
    *is_match = 1;
    return;
  }
  if(s->use_match_mismatch) {
    *score = *is_match ? s->match : s->mismatch;
    return;
  }
  fprintf(stderr, "Error: Unknown character pair (%c,%c) and "
                  "match/mismatch have not been set\n", a, b);
  exit(EXIT_FAILURE);
}

/* load an n x n published matrix: T[i][j] scores (letters[i], letters[j]) */
static void load_table(scoring_t *s, const char *letters, const signed char *t,
                       char use_match_mismatch)
{
  const size_t n = strlen(letters);
  for(size_t i = 0; i < n; i++) {
    const char a = s->case_sensitive ? letters[i] : (char)tolower(letters[i]);
    for(size_t j = 0; j < n; j++) {
      const char b = s->case_sensitive ? letters[j] : (char)tolower(letters[j]);
      scoring_add_mutation(s, a, b, t[i * n + j]);
    }
  }
  s->use_match_mismatch = use_match_mismatch;
}

/* reference alignment_scoring.c:306-377: match/mismatch/gap defaults of each
 * system, then the matrix; case-insensitive, no free end gaps */
void scoring_system_PAM30(scoring_t *s)
{
  scoring_init(s, 1, -17, -9, -1, 0, 0, 0, 0, 0, 0);
  load_table(s, sa_amino_letters, &sa_table_PAM30[0][0], 1);
}

void scoring_system_PAM70(scoring_t *s)
{
  scoring_init(s, 1, -11, -10, -1, 0, 0, 0, 0, 0, 0);
  load_table(s, sa_amino_letters, &sa_table_PAM70[0][0], 1);
}

void scoring_system_BLOSUM80(scoring_t *s)
{
  scoring_init(s, 1, -8, -10, -1, 0, 0, 0, 0, 0, 0);
  load_table(s, sa_amino_letters, &sa_table_BLOSUM80[0][0], 1);
}

void scoring_system_BLOSUM62(scoring_t *s)
{
  scoring_init(s, 1, -4, -10, -1, 0, 0, 0, 0, 0, 0);
  load_table(s, sa_amino_letters, &sa_table_BLOSUM62[0][0], 1);
}

/* reference alignment_scoring.c:364-377: no match/mismatch fallback */
void scoring_system_DNA_hybridization(scoring_t *s)
{
  scoring_init(s, 0, 0, -10, -10, 0, 0, 0, 0, 0, 0);
  load_table(s, sa_dna_letters, &sa_table_DNA_HYBRIDIZATION[0][0], 0);
}

/* reference alignment_scoring.c:380-392 */
void scoring_system_default(scoring_t *s)
{
  scoring_init(s, 1, -2, -4, -1, 0, 0, 0, 0, 0, 0);
}
