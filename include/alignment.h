/*
 * alignment.h -- aligner_t / alignment_t and the DP entry point of the
 * seq-align C API (B200 build).
 *
 * Drop-in for reference src/alignment.h:14-79.  aligner_align() is the seam:
 * in the reference it runs alignment_fill_matrices() (src/alignment.c:28-168)
 * on the host; here it ships the pair to the GPU, runs the sm_100a fill kernel
 * in materialise mode and copies the three int32 matrices back into the
 * aligner_t buffers, so every caller that reads match_scores / gap_a_scores /
 * gap_b_scores (alignment_reverse_move, alignment_print_matrices, user code)
 * sees exactly the reference's numbers.  aligner_t and alignment_t are
 * caller-visible (stack allocated, fields read directly: reference
 * examples/nw_example.c:58-60, src/tools/tests.c:82-93) so their layout is ABI.
 */
#ifndef ALIGNMENT_HEADER_SEEN
#define ALIGNMENT_HEADER_SEEN

#include <string.h>
#include <stddef.h>
#include "alignment_scoring.h"

#ifndef ROUNDUP2POW
  #define ROUNDUP2POW(x) sa_roundup_pow2_u64(x)
  static inline size_t sa_roundup_pow2_u64(unsigned long long v)
  {
    /* smallest power of two >= v (v=0 -> 0), as reference alignment.h:14-21 */
    int shift;
    v--;
    for(shift = 1; shift < 64; shift <<= 1) v |= v >> shift;
    return (size_t)(v + 1);
  }
#endif

typedef struct
{
  const scoring_t *scoring;
  const char *seq_a, *seq_b;           /* borrowed, not copied */
  size_t score_width, score_height;    /* len_a+1, len_b+1 */
  score_t *match_scores, *gap_a_scores, *gap_b_scores; /* row-major, owned */
  size_t capacity;                     /* cells, power of two */
} aligner_t;

typedef struct
{
  char *result_a, *result_b;           /* NUL-terminated gapped strings */
  size_t capacity, length;
  size_t pos_a, pos_b;                 /* 0-based start of a local hit */
  size_t len_a, len_b;                 /* bases consumed by a local hit */
  score_t score;
} alignment_t;

enum Matrix { MATCH, GAP_A, GAP_B };

#define MATRIX_NAME(x) ((x) == MATCH ? "MATCH" : ((x) == GAP_A ? "GAP_A" : "GAP_B"))

#ifdef __cplusplus
extern "C" {
#endif

extern const char align_col_mismatch[], align_col_indel[], align_col_context[],
                  align_col_stop[];

#define aligner_init(a) (memset(a, 0, sizeof(aligner_t)))

/* reference src/alignment.c:170-193: bind the pair, grow the three buffers to
 * the next power of two, fill.  Here the fill is the GPU kernel in materialise
 * mode and the matrices are copied back into the buffers. */
void aligner_align(aligner_t *self, const char *a, const char *b, size_t n_a, size_t n_b, const scoring_t *model, char is_sw);

/* reference src/alignment.c:195-202: frees the matrices, zeroes the struct */
void aligner_destroy(aligner_t *self);

/* reference src/alignment.c:205-240: result buffers, grown in powers of two;
 * "Out of memory" on stderr and exit(EXIT_FAILURE) when realloc fails */
alignment_t *alignment_create(size_t capacity);
void alignment_ensure_capacity(alignment_t *out, size_t columns);
void alignment_free(alignment_t *out);

/* reference src/alignment.c:244-350: one backward step over the materialised
 * matrices, predecessor chosen by equality in the order GAP_A, GAP_B, MATCH
 * (host side, for callers that walk themselves; the batch engine takes the
 * same decisions at fill time) */
void alignment_reverse_move(enum Matrix *state, score_t *score, size_t *x, size_t *y, size_t *cell_index, const aligner_t *self);

/* reference src/alignment.c:353-474: text output, byte for byte */
void alignment_print_matrices(const aligner_t *self);
void alignment_colour_print_against(const char *row, const char *other_row, char case_sensitive);
void alignment_print_spacer(const char *row_a, const char *row_b, const scoring_t *model);

#ifdef __cplusplus
}
#endif

#endif
