/*
 * alignment_cmdline.h -- command-line layer of the seq-align C API (B200 build):
 * option parsing into cmdline_t and the file loop that feeds pairs of records
 * to an alignment callback.
 *
 * Drop-in for reference src/alignment_cmdline.h:22-72 (implemented in
 * src/alignment_cmdline.c:32-640, part of the reference's libalign.a,
 * Makefile:17-26).  cmdline_t is caller-visible (the reference's tools read
 * cmd->print_pretty, cmd->min_score ... directly: src/tools/nw_cmdline.c:81-143,
 * src/tools/sw_cmdline.c:192-217), so its field order and types are ABI.
 *
 * read_t is the record type of the reference's sequence reader
 * (libs/seq_file/seq_file.h:61-73, a separate header-only library the reference
 * pulls in as a git submodule).  When the caller has included seq_file.h first
 * its own definition is used; otherwise an ABI-identical declaration of the
 * fields a callback may touch is supplied here, so this header stands alone.
 */
#ifndef ALIGNMENT_CMDLINE_HEADER_SEEN
#define ALIGNMENT_CMDLINE_HEADER_SEEN

#include <stdarg.h>
#include <stdbool.h>
#include <stddef.h>
#include "alignment.h"

#ifndef _SEQ_FILE_HEADER
/* layout of reference libs/seq_file/seq_file.h:61-73 */
typedef struct {
  char *b;           /* NUL-terminated text */
  size_t end, size;  /* length, capacity */
} seq_buf_t;

typedef struct read_struct read_t;
struct read_struct {
  seq_buf_t name, seq, qual; /* name: the header line without its leading '>' / '@' */
  void *bam;
  read_t *next;
  bool from_sam;
};
#endif

#ifdef __cplusplus
extern "C" {
#endif

enum SeqAlignCmdType {SEQ_ALIGN_SW_CMD, SEQ_ALIGN_NW_CMD, SEQ_ALIGN_LCS_CMD};

typedef struct
{
  /* input files: pair i is (file_paths1[i], file_paths2[i]); file_paths2[i]
   * NULL = both records of a pair come from file_paths1[i] */
  size_t file_list_length, file_list_capacity;
  char **file_paths1, **file_paths2;

  bool case_sensitive;
  int match, mismatch, gap_open, gap_extend;

  /* Smith-Waterman */
  score_t min_score;
  unsigned int print_context, max_hits_per_alignment;
  bool min_score_set, max_hits_per_alignment_set;
  bool print_seq;

  /* Needleman-Wunsch */
  bool freestartgap_set, freeendgap_set;
  bool print_matrices, print_scores;
  bool zam_stle_output;

  /* --stdin: answer pair by pair, read stdin unbuffered and without zlib */
  bool interactive;

  bool print_fasta, print_pretty, print_colour;

  bool no_gaps_in1, no_gaps_in2;
  bool no_mismatches;

  /* pair given on the command line (borrowed from argv) */
  const char *seq1, *seq2;
} cmdline_t;

/* 1 and *result set if the WHOLE string is a number in range, else 0 */
char parse_entire_int(char *str, int *result);
char parse_entire_uint(char *str, unsigned int *result);

/* Parses argv into a new cmdline_t and into *scoring (which holds the tool's
 * default scores on entry).  Bad input prints "Error: ..." and the usage text
 * on stderr and exits with EXIT_FAILURE. */
cmdline_t *cmdline_new(int argc, char **argv, scoring_t *scoring,
                       enum SeqAlignCmdType cmd_type);
void cmdline_free(cmdline_t *cmd);

void cmdline_add_files(cmdline_t *cmd, char *p1, char *p2);
size_t cmdline_get_num_of_file_pairs(cmdline_t *cmd);
char *cmdline_get_file1(cmdline_t *cmd, size_t i);
char *cmdline_get_file2(cmdline_t *cmd, size_t i);

/* Reads records (FASTA, FASTQ or one sequence per line; gzip when use_zlib)
 * two at a time from path1 (path2 == NULL) or one from each file, and calls
 * align(r1, r2) per pair, in input order.  path "-" is stdin. */
void align_from_file(const char *path1, const char *path2,
                     void (align)(read_t *r1, read_t *r2),
                     bool use_zlib);

#ifdef __cplusplus
}
#endif

#endif
