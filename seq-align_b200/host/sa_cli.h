/*
 * sa_cli.h -- shared pieces of the batching command-line tools
 * (needleman_wunsch, smith_waterman, lcs): option parsing, sequence file
 * reading, scoring file loading.
 *
 * These are the B200-side replacements of the reference's host front-end
 * (SURVEY.md 8f-2 / 8f-3):
 *   option parser      reference src/alignment_cmdline.c:179-485 (cmdline_new)
 *   file loop          reference src/alignment_cmdline.c:578-640 (align_from_file)
 *   sequence reader    reference libs/seq_file/seq_file.h:245-325 (FASTA / FASTQ / plain, gzip)
 *   scoring loaders    reference src/alignment_scoring_load.c:39-306
 * Same flags, same stdout bytes; what changes is the shape of the loop: pairs
 * are read ahead and aligned as ONE batch per launch instead of one
 * aligner_align() call per pair.
 */
#ifndef SA_CLI_H
#define SA_CLI_H

#include <stddef.h>
#include "alignment_scoring.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { SA_TOOL_NW = 0, SA_TOOL_SW = 1, SA_TOOL_LCS = 2 }; /* LCS: neither the NW-only nor the SW-only flags */

typedef struct {
  const char *path1, *path2; /* path2 == NULL: pairs come from consecutive records of path1 */
} sa_file_pair;

typedef struct {
  int tool;
  int case_sensitive;     /* only steers the scoring file loaders, as in the reference */
  int print_seq, print_matrices, print_scores, print_fasta, print_pretty, print_colour;
  int zam, interactive;
  int min_score, min_score_set;
  unsigned max_hits; int max_hits_set;
  unsigned context;
  int gpus, gpus_set;     /* --gpus <n>: devices the batches are cut over (0 = all); not a flag of the reference */
  const char *seq1, *seq2; /* pair given on the command line */
  sa_file_pair *files; size_t nfiles, files_cap;
} sa_opts;

/* parse argv into opts + scoring (scoring holds the tool's defaults on entry);
 * prints "Error: ..." + usage and exits on bad input, like cmdline_new() */
void sa_cli_parse(int argc, char **argv, scoring_t *scoring, int tool, sa_opts *o);
void sa_cli_free(sa_opts *o);

/* growable string */
typedef struct { char *b; size_t len, cap; } sa_str;
void sa_str_free(sa_str *s);

typedef struct { sa_str name, seq; } sa_record;

typedef struct sa_reader sa_reader;
/* path "-" is stdin.  buffered = 0 reads stdin byte by byte (interactive
 * protocol: nothing beyond the current record is consumed) */
sa_reader *sa_reader_open(const char *path, int buffered);
void sa_reader_close(sa_reader *r);
/* 1 = record read, 0 = end of input, -1 = syntax error / truncated record */
int sa_reader_next(sa_reader *r, sa_record *rec);
/* one raw byte from the same input (-1 at end): the interactive SW prompt reads its answers here */
int sa_reader_getc(sa_reader *r);

/* --substitution_matrix / --substitution_pairs files (gzip ok) */
void sa_load_matrix(const char *path, scoring_t *scoring, int case_sensitive);
void sa_load_pairs(const char *path, scoring_t *scoring, int case_sensitive);
/* the same on a gzFile the caller opened and will close (path: for messages) */
void sa_load_matrix_gz(void *gz, const char *path, scoring_t *scoring, int case_sensitive);
void sa_load_pairs_gz(void *gz, const char *path, scoring_t *scoring, int case_sensitive);

#ifdef __cplusplus
}
#endif

#endif
