/*
 * needleman_wunsch.h -- global alignment front-end (B200 build).
 *
 * Drop-in for reference src/needleman_wunsch.h:16-32 (same names, argument
 * order and types; callers compile unchanged).  The fill, the end-state
 * choice (src/needleman_wunsch.c:53-66) and the traceback
 * (src/needleman_wunsch.c:79-145) all run on the GPU: the fill kernel emits
 * one traceback byte per cell, a walk kernel follows them and writes the two
 * gapped strings, which are copied into alignment_t.
 */
#ifndef NEEDLEMAN_WUNSCH_HEADER_SEEN
#define NEEDLEMAN_WUNSCH_HEADER_SEEN

#include "seq_align.h"
#include "alignment.h"

/* the reference's Needleman-Wunsch aligner is its plain aligner_t */
typedef aligner_t nw_aligner_t;

#ifdef __cplusplus
extern "C" {
#endif

/* zeroed aligner; its matrices grow on demand and are filled when something reads them
 * (deferred matrices: seqalign_host_eager_matrices() / SEQALIGN_EAGER_MATRICES in INTEGRATION.md) */
nw_aligner_t *needleman_wunsch_new();

/* releases the matrices and the aligner itself */
void needleman_wunsch_free(nw_aligner_t *aligner);

/* NUL-terminated sequences: strlen() both, then needleman_wunsch_align2 */
void needleman_wunsch_align(const char *seq_a, const char *seq_b, const scoring_t *model, nw_aligner_t *aligner, alignment_t *out);

/* explicit lengths (no NUL needed).  out->result_a / result_b / length / score
 * are filled; pos_* and len_* are left alone, as upstream */
void needleman_wunsch_align2(const char *seq_a, const char *seq_b, size_t n_a, size_t n_b, const scoring_t *model, nw_aligner_t *aligner, alignment_t *out);

#ifdef __cplusplus
}
#endif

#endif
