/* sa_host.h -- internals shared by the host-side API files */
#ifndef SA_HOST_H
#define SA_HOST_H

#include "seqalign_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* lazily created process-wide engine of the single-pair API; exits with a
 * message if no B200 is usable (there is no CPU path) */
seqalign_batch_t *sa_host_engine(void);
/* map an engine error to the reference's stderr text + exit(EXIT_FAILURE) */
void sa_host_check(seqalign_batch_t *eng, int rc);
/* bind a pair to an aligner_t the way aligner_align() does, leaving the three matrices to be filled
 * when something reads them (sa_host_materialise); fills at once in eager mode */
void sa_host_bind(aligner_t *aligner, const char *seq_a, const char *seq_b, size_t len_a, size_t len_b,
                  const scoring_t *scoring, char is_sw);
void sa_host_materialise(const aligner_t *al);

#ifdef __cplusplus
}
#endif

#endif
