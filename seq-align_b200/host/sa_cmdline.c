/*
 * sa_cmdline.c -- the reference's command-line layer as exported library
 * functions (include/alignment_cmdline.h, include/alignment_scoring_load.h),
 * so that libalign is a superset of the reference's src/libalign.a
 * (reference Makefile:17-26 bundles every src/ file into it) and the reference's
 * own tool mains link against this library alone.
 *
 *   parse_entire_int / _uint      reference src/alignment_cmdline.c:32-66
 *   cmdline_new / cmdline_free    reference src/alignment_cmdline.c:179-539
 *   cmdline_add_files, getters    reference src/alignment_cmdline.c:542-573
 *   align_from_file               reference src/alignment_cmdline.c:578-640
 *   align_scoring_load_matrix     reference src/alignment_scoring_load.c:39-233
 *   align_scoring_load_pairwise   reference src/alignment_scoring_load.c:236-306
 *
 * All of them are thin adapters over sa_cli.c (option table, record reader,
 * loaders).  align_from_file() keeps the reference's one-pair-per-callback
 * contract; callers that want the batched GPU loop use the tools in
 * seq-align_b200/tools or seqalign_b200.h directly.
 */
#define _GNU_SOURCE
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "alignment_cmdline.h"
#include "alignment_scoring_load.h"
#include "sa_cli.h"

char parse_entire_int(char *str, int *result)
{
  char *end = str;
  const long v = strtol(str, &end, 10);
  if(v > INT_MAX || v < INT_MIN || end != str + strlen(str)) return 0;
  *result = (int)v;
  return 1;
}

char parse_entire_uint(char *str, unsigned int *result)
{
  char *end = str;
  const unsigned long v = strtoul(str, &end, 10);
  if(v > UINT_MAX || end != str + strlen(str)) return 0;
  *result = (unsigned int)v;
  return 1;
}

static void *must(void *p)
{
  if(!p) { fprintf(stderr, "%s:%i: Out of memory\n", __FILE__, __LINE__); exit(EXIT_FAILURE); }
  return p;
}

cmdline_t *cmdline_new(int argc, char **argv, scoring_t *scoring, enum SeqAlignCmdType cmd_type)
{
  sa_opts o;
  const int tool = cmd_type == SEQ_ALIGN_SW_CMD ? SA_TOOL_SW : cmd_type == SEQ_ALIGN_NW_CMD ? SA_TOOL_NW : SA_TOOL_LCS;
  sa_cli_parse(argc, argv, scoring, tool, &o); /* exits with the usage text on bad input */

  cmdline_t *cmd = must(calloc(1, sizeof(*cmd)));
  cmd->file_list_capacity = 256;
  while(cmd->file_list_capacity < o.nfiles) cmd->file_list_capacity *= 2;
  cmd->file_paths1 = must(malloc(sizeof(char *) * cmd->file_list_capacity));
  cmd->file_paths2 = must(malloc(sizeof(char *) * cmd->file_list_capacity));
  for(size_t i = 0; i < o.nfiles; i++) cmdline_add_files(cmd, (char *)o.files[i].path1, (char *)o.files[i].path2);

  cmd->case_sensitive = o.case_sensitive;
  cmd->min_score = o.min_score;
  cmd->min_score_set = o.min_score_set;
  cmd->print_context = o.context;
  cmd->max_hits_per_alignment = o.max_hits;
  cmd->max_hits_per_alignment_set = o.max_hits_set;
  cmd->print_seq = o.print_seq;
  cmd->print_matrices = o.print_matrices;
  cmd->print_scores = o.print_scores;
  cmd->zam_stle_output = o.zam;
  cmd->interactive = o.interactive;
  cmd->print_fasta = o.print_fasta;
  cmd->print_pretty = o.print_pretty;
  cmd->print_colour = o.print_colour;
  cmd->seq1 = o.seq1;
  cmd->seq2 = o.seq2;
  /* match .. gap_extend, freestartgap_set, freeendgap_set, no_gaps_in1/2, no_mismatches stay 0: the
   * reference's parser writes those choices into *scoring only (src/alignment_cmdline.c:257-283) */
  sa_cli_free(&o);
  return cmd;
}

void cmdline_free(cmdline_t *cmd)
{
  free(cmd->file_paths1);
  free(cmd->file_paths2);
  free(cmd);
}

void cmdline_add_files(cmdline_t *cmd, char *p1, char *p2)
{
  if(cmd->file_list_length == cmd->file_list_capacity) {
    cmd->file_list_capacity = cmd->file_list_capacity ? 2 * cmd->file_list_capacity : 256;
    cmd->file_paths1 = must(realloc(cmd->file_paths1, sizeof(char *) * cmd->file_list_capacity));
    cmd->file_paths2 = must(realloc(cmd->file_paths2, sizeof(char *) * cmd->file_list_capacity));
  }
  cmd->file_paths1[cmd->file_list_length] = p1;
  cmd->file_paths2[cmd->file_list_length] = p2;
  cmd->file_list_length++;
}

size_t cmdline_get_num_of_file_pairs(cmdline_t *cmd) { return cmd->file_list_length; }
char *cmdline_get_file1(cmdline_t *cmd, size_t i) { return cmd->file_paths1[i]; }
char *cmdline_get_file2(cmdline_t *cmd, size_t i) { return cmd->file_paths2[i]; }

/* view of an sa_record through the reference's record type */
static void as_read(read_t *r, const sa_record *rec)
{
  memset(r, 0, sizeof(*r));
  r->name.b = rec->name.b; r->name.end = rec->name.len; r->name.size = rec->name.cap;
  r->seq.b = rec->seq.b; r->seq.end = rec->seq.len; r->seq.size = rec->seq.cap;
  static char empty[1] = "";
  r->qual.b = empty; /* qualities are parsed over but not kept: no alignment path reads them */
}

void align_from_file(const char *path1, const char *path2, void (align)(read_t *r1, read_t *r2), bool use_zlib)
{
  /* use_zlib false and path "-": stdin byte by byte, nothing read beyond the current record
   * (the perl wrappers' request / response protocol) */
  sa_reader *r1 = sa_reader_open(path1, use_zlib), *r2 = r1;
  if(!r1) { fprintf(stderr, "Alignment Error: couldn't open file %s\n", path1); fflush(stderr); return; }
  if(path2 && !(r2 = sa_reader_open(path2, use_zlib))) {
    fprintf(stderr, "Alignment Error: couldn't open file %s\n", path1); fflush(stderr); /* path1: as the reference prints it */
    sa_reader_close(r1);
    return;
  }
  sa_record rec1, rec2;
  memset(&rec1, 0, sizeof(rec1)); memset(&rec2, 0, sizeof(rec2));
  unsigned long pairs = 0;
  for(; sa_reader_next(r1, &rec1) > 0; pairs++) {
    if(sa_reader_next(r2, &rec2) <= 0) {
      fprintf(stderr, "Alignment Error: Odd number of sequences - I read in pairs!\n"); fflush(stderr);
      break;
    }
    read_t a, b;
    as_read(&a, &rec1); as_read(&b, &rec2);
    align(&a, &b);
  }
  if(pairs == 0) { fprintf(stderr, "Alignment Warning: empty input\n"); fflush(stderr); }
  sa_reader_close(r1);
  if(path2) sa_reader_close(r2);
  sa_str_free(&rec1.name); sa_str_free(&rec1.seq); sa_str_free(&rec2.name); sa_str_free(&rec2.seq);
}

void align_scoring_load_matrix(gzFile file, const char *file_path, scoring_t *scoring, char case_sensitive)
{
  sa_load_matrix_gz(file, file_path, scoring, case_sensitive);
}

void align_scoring_load_pairwise(gzFile file, const char *file_path, scoring_t *scoring, char case_sensitive)
{
  sa_load_pairs_gz(file, file_path, scoring, case_sensitive);
}
