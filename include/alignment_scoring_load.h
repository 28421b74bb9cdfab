/*
 * alignment_scoring_load.h -- substitution matrix / pair-score file loaders
 * of the seq-align C API (B200 build).
 *
 * Drop-in for reference src/alignment_scoring_load.h:14-18; file formats as
 * read by reference src/alignment_scoring_load.c:39-306.  The caller opens
 * and closes the gzFile (reference src/alignment_cmdline.c:343-349); file_path
 * is only used in error messages.  Errors print "Error: substitution matrix :
 * ..." / "Error: substitution pairs : ..." on stderr and exit(EXIT_FAILURE),
 * like the reference.  Every entry lands in scoring->swap_scores through
 * scoring_add_mutation(), from where the engine flattens it for the device.
 */
#ifndef ALIGNMENT_SCORING_LOAD_HEADER_SEEN
#define ALIGNMENT_SCORING_LOAD_HEADER_SEEN

#include <zlib.h>
#include "alignment_scoring.h"

#ifdef __cplusplus
extern "C" {
#endif

void align_scoring_load_matrix(gzFile file, const char *file_path,
                               scoring_t *scoring, char case_sensitive);

void align_scoring_load_pairwise(gzFile file, const char *file_path,
                                 scoring_t *scoring, char case_sensitive);

#ifdef __cplusplus
}
#endif

#endif
