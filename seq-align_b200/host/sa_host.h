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

#ifdef __cplusplus
}
#endif

#endif
