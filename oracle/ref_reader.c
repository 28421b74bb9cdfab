/*
 * ref_reader.c -- TEST INFRASTRUCTURE: dumps the records the reference's own
 * sequence reader (libs/seq_file/seq_file.h, included from where it lies under
 * /root/reference, never copied) returns for a file, opened exactly as the
 * reference's align_from_file() opens it (src/alignment_cmdline.c:570-596:
 * seq_open(path), buffered, zlib).  Used to pin oracle/sa_oracle.c's restatement
 * of the reader (orc_read_records) and to record tests/golden/reader_vectors.json.
 *
 *   ref_reader <file>        one line per record:  <status> <name_len> <seq_len>\n<name bytes>\n<seq bytes>\n
 *                            then "end <last seq_read() return value>"
 */
#include <stdio.h>
#include <stdlib.h>
#include "seq_file/seq_file.h"

int main(int argc, char **argv)
{
  if(argc != 2) { fprintf(stderr, "usage: ref_reader <file>\n"); return 2; }
  seq_file_t *sf = seq_open(argv[1]);
  if(!sf) { printf("open failed\n"); return 1; }
  read_t r;
  seq_read_alloc(&r);
  int s;
  while((s = seq_read(sf, &r)) > 0) {
    printf("rec %zu %zu\n", r.name.end, r.seq.end);
    fwrite(r.name.b, 1, r.name.end, stdout); putc('\n', stdout);
    fwrite(r.seq.b, 1, r.seq.end, stdout); putc('\n', stdout);
  }
  printf("end %d fmt %d\n", s, (int)sf->format);
  seq_close(sf);
  seq_read_dealloc(&r);
  return 0;
}
