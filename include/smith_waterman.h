/*
 * smith_waterman.h -- local alignment front-end (B200 build).
 *
 * Drop-in for reference src/smith_waterman.h:15-39 (same names, argument
 * order and types).  sw_aligner_t stays opaque; besides the aligner_t the
 * reference keeps there, this build hides the hit-iteration state of the pair
 * in it.  The two sequences and the scoring model must stay alive and
 * unchanged between smith_waterman_align() and the last
 * smith_waterman_fetch(), exactly as upstream (they are borrowed, not copied).
 */
#ifndef SMITH_WATERMAN_HEADER_SEEN
#define SMITH_WATERMAN_HEADER_SEEN

#include "seq_align.h"
#include "alignment.h"

typedef struct sw_aligner_t sw_aligner_t;

#ifdef __cplusplus
extern "C" {
#endif

/* exits with "Out of memory" on failure, like upstream */
sw_aligner_t *smith_waterman_new();
void smith_waterman_free(sw_aligner_t *handle);

/* the embedded aligner_t: score_width / score_height and the three matrices
 * of the last smith_waterman_align (reference smith_waterman.c:126-129) */
aligner_t *smith_waterman_get_aligner(sw_aligner_t *handle);

/* fill for a NUL-terminated pair; resets the hit iteration */
void smith_waterman_align(const char *a, const char *b, const scoring_t *model, sw_aligner_t *handle);

/* the same with explicit lengths */
void smith_waterman_align2(const char *a, const char *b, size_t n_a, size_t n_b, const scoring_t *model, sw_aligner_t *handle);

/* next local hit in (score desc, x asc, y asc) order, skipping hits whose
 * walk runs into an earlier one; 1 if `out` was filled, 0 when there are none left */
int smith_waterman_fetch(sw_aligner_t *handle, alignment_t *out);

#ifdef __cplusplus
}
#endif

#endif
