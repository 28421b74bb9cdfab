/*
 * lcs_main.c -- `lcs`: repeated substrings of one sequence, longest first.
 * The reference tool (src/tools/lcs_cmdline.c:32-70) aligns the sequence
 * against itself with Smith-Waterman, gaps and mismatches forbidden and case
 * sensitive, and prints every hit that lies above the diagonal.  The fill
 * runs on the GPU through the single-pair API (this scoring shape is outside
 * the batch multi-hit mode).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "smith_waterman.h"

int main(int argc, char **argv)
{
  if(argc != 2) {
    fprintf(stderr, "%s [options] <sequence>\n", argv[0]);
    fprintf(stderr, "  Print substrings in decreasing order of length\n");
    return EXIT_FAILURE;
  }
  const char *seq = argv[1];
  scoring_t scoring;
  scoring_init(&scoring, 1, -1, -4, -1, false, false, true, true, true, true);
  sw_aligner_t *sw = smith_waterman_new();
  alignment_t *aln = alignment_create(strlen(seq) + 1);
  smith_waterman_align(seq, seq, &scoring, sw);
  while(smith_waterman_fetch(sw, aln)) {
    if(aln->pos_a < aln->pos_b) {
      fputs(aln->result_a, stdout);
      printf(" [%zu,%zu]\n", aln->pos_a, aln->pos_b);
    }
  }
  smith_waterman_free(sw);
  alignment_free(aln);
  return EXIT_SUCCESS;
}
