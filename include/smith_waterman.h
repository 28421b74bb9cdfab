/*
 * smith_waterman.h -- local alignment front-end (B200 build).
 *
 * Drop-in for reference src/smith_waterman.h:15-39.  sw_aligner_t stays
 * opaque; besides the aligner_t the reference keeps there, this build hides
 * the device-side state of the pair in it.  seq_a, seq_b and scoring must
 * stay alive and unchanged between smith_waterman_align() and the last
 * smith_waterman_fetch(), exactly as upstream.
 */
#ifndef SMITH_WATERMAN_HEADER_SEEN
#define SMITH_WATERMAN_HEADER_SEEN

#include "seq_align.h"
#include "alignment.h"

typedef struct sw_aligner_t sw_aligner_t;

#ifdef __cplusplus
extern "C" {
#endif

sw_aligner_t *smith_waterman_new();
void smith_waterman_free(sw_aligner_t *sw_aligner);

aligner_t *smith_waterman_get_aligner(sw_aligner_t *sw);

void smith_waterman_align(const char *seq_a, const char *seq_b,
                          const scoring_t *scoring, sw_aligner_t *sw);

void smith_waterman_align2(const char *seq_a, const char *seq_b,
                           size_t len_a, size_t len_b,
                           const scoring_t *scoring, sw_aligner_t *sw);

/* next local hit in (score desc, x asc, y asc) order; 1 if one was written */
int smith_waterman_fetch(sw_aligner_t *sw, alignment_t *result);

#ifdef __cplusplus
}
#endif

#endif
