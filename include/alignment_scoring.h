/*
 * alignment_scoring.h -- scoring model of the seq-align C API (B200 build).
 *
 * Drop-in for reference src/alignment_scoring.h:16-81.  Callers allocate
 * scoring_t themselves (on the stack, see reference examples/nw_example.c:43)
 * and poke its fields directly (reference src/alignment_cmdline.c:401-439), so
 * the field order, types and sizes below ARE the ABI: sizeof(scoring_t) is
 * 271,428 bytes with gap_open at offset 0, wildcards at 28, swap_set at 60,
 * wildscores at 8252, swap_scores at 9276, min_penalty at 271,420.
 * tests/test_abi.py checks these offsets against the compiled reference.
 *
 * The functions are implemented in seq-align_b200/host/sa_scoring.c (host
 * code; the GPU never sees scoring_t -- the batch engine flattens it into a
 * dense code table, see include/seqalign_b200.h).
 */
#ifndef ALIGNMENT_SCORING_HEADER_SEEN
#define ALIGNMENT_SCORING_HEADER_SEEN

#include <inttypes.h>
#include <stdbool.h>
#include <limits.h>

typedef int score_t;
#define SCORE_MIN INT_MIN

typedef struct
{
  /* a gap of length N costs gap_open + N*gap_extend */
  int gap_open, gap_extend;
  /* Needleman-Wunsch: leading / trailing gaps are free */
  bool no_start_gap_penalty, no_end_gap_penalty;
  /* restrictions: forbid gaps inside a / inside b / forbid mismatches */
  bool no_gaps_in_a, no_gaps_in_b, no_mismatches;
  /* fall back to match/mismatch for pairs without a swap score */
  bool use_match_mismatch;
  int match, mismatch;
  bool case_sensitive;
  /* wildcard characters pair with anything at a fixed score; swap_set marks
   * the (a,b) pairs that have an explicit substitution score */
  uint32_t wildcards[256/32], swap_set[256][256/32];
  score_t wildscores[256], swap_scores[256][256];
  /* extremes over every penalty known at scoring_init / add time */
  int min_penalty, max_penalty;
} scoring_t;

#ifndef bitset32_get
  #define bitset32_get(arr,idx)   (((arr)[(idx)>>5] >> ((idx)&31)) & 0x1)
  #define bitset32_set(arr,idx)   ((arr)[(idx)>>5] |=   (1u<<((idx)&31)))   /* 1u: bit 31 of a signed 1 is undefined behaviour (UBSan) */
  #define bitset32_clear(arr,idx) ((arr)[(idx)>>5] &=  ~(1u<<((idx)&31)))
#endif

#define get_wildcard_bit(scoring,c) bitset32_get((scoring)->wildcards,c)
#define set_wildcard_bit(scoring,c) bitset32_set((scoring)->wildcards,c)
#define get_swap_bit(scoring,a,b) bitset32_get((scoring)->swap_set[(size_t)(a)],b)
#define set_swap_bit(scoring,a,b) bitset32_set((scoring)->swap_set[(size_t)(a)],b)
#define scoring_is_wildcard(scoring,c) (get_wildcard_bit(scoring,c))

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/alignment_scoring.c:21-55: clears the tables, stores the
 * penalties and flags, derives min_penalty / max_penalty */
void scoring_init(scoring_t *model, int match, int mismatch, int gap_open, int gap_extend, bool free_start_gaps, bool free_end_gaps, bool forbid_gaps_in_a, bool forbid_gaps_in_b, bool forbid_mismatches, bool case_sensitive);

/* reference src/alignment_scoring.c:57-64: `wild` pairs with every character at `score` */
void scoring_add_wildcard(scoring_t *model, char wild, int score);

/* reference src/alignment_scoring.c:66-72: explicit score for the ordered pair (from, to); no case folding */
void scoring_add_mutation(scoring_t *model, char from, char to, int score);

/* reference src/alignment_scoring.c:74-97 (exported there, not declared):
 * a square table over the characters of `alphabet`, row-major */
void scoring_add_mutations(scoring_t *model, const char *alphabet, const int *table, char use_match_mismatch);

/* reference src/alignment_scoring.c:99-112 */
void scoring_print(const scoring_t *model);

/* reference src/alignment_scoring.c:133-182: case fold, swap table, wildcards,
 * match/mismatch; prints "Unknown character pair" and exits when nothing applies */
void scoring_lookup(const scoring_t *model, char a, char b, int *score_out, bool *is_match_out);

/* built-in systems, reference src/alignment_scoring.c:306-392 (tables in host/sa_scoring_tables.h) */
void scoring_system_PAM30(scoring_t *model);
void scoring_system_PAM70(scoring_t *model);
void scoring_system_BLOSUM80(scoring_t *model);
void scoring_system_BLOSUM62(scoring_t *model);
void scoring_system_DNA_hybridization(scoring_t *model);
void scoring_system_default(scoring_t *model);

#ifdef __cplusplus
}
#endif

#endif
