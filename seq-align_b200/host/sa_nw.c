/*
 * sa_nw.c -- Needleman-Wunsch front-end of the seq-align C API.
 *
 * Implements include/needleman_wunsch.h on top of the batch engine.  Where
 * the reference filled on the host and walked back with
 * alignment_reverse_move (src/needleman_wunsch.c:34-145), this submits a
 * batch of one in align mode: fill + direction bytes + walk all run on the
 * GPU and the gapped strings come back ready.  The aligner_t matrices are
 * filled when a caller looks at them (`needleman_wunsch --printmatrices` through
 * alignment_print_matrices); seqalign_host_eager_matrices(1) fills them at once.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "needleman_wunsch.h"
#include "seqalign_b200.h"
#include "sa_host.h"

nw_aligner_t *needleman_wunsch_new()
{
  return calloc(1, sizeof(nw_aligner_t));
}

void needleman_wunsch_free(nw_aligner_t *nw)
{
  aligner_destroy(nw);
  free(nw);
}

void needleman_wunsch_align(const char *a, const char *b, const scoring_t *scoring,
                            nw_aligner_t *nw, alignment_t *result)
{
  needleman_wunsch_align2(a, b, strlen(a), strlen(b), scoring, nw, result);
}

void needleman_wunsch_align2(const char *a, const char *b, size_t len_a, size_t len_b,
                             const scoring_t *scoring, nw_aligner_t *nw, alignment_t *result)
{
  /* one fill: score, traceback bytes (or checkpoints) and walk on the device; the matrices of the
   * aligner_t follow only if somebody reads them (sa_alignment.c "deferred matrices") */
  sa_host_bind(nw, a, b, len_a, len_b, scoring, 0);
  seqalign_batch_t *eng = sa_host_engine();
  seqalign_batch_set_scoring(eng, scoring);
  sa_host_check(eng, seqalign_batch_submit(eng, SEQALIGN_NW, SEQALIGN_MODE_ALIGN,
                                           &a, &len_a, &b, &len_b, 1));
  /* same growth as the reference: room for len_a+len_b columns */
  alignment_ensure_capacity(result, len_a + len_b);
  sa_host_check(eng, seqalign_batch_alignment(eng, 0, result));
}
