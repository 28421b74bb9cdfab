/*
 * sa_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C restatement of the seq-align hot path (three-matrix affine-gap
 * fill, traceback, SW hit iteration, scoring lookup).  It exists so the CUDA
 * path can be checked bit-for-bit.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * library (seq-align_b200/) never links or calls anything in oracle/.
 *
 * Parity is PINNED: tests/test_oracle.py checks this restatement against
 *   (1) the golden vectors the reference's own tests and README hold
 *       (src/tools/tests.c:65-268, README.md:71-74, README.md:118-145), and
 *   (2) the unmodified reference compiled into oracle/_ref/libalign_ref.so,
 *       differentially on seeded random inputs, and through the committed
 *       fixtures in tests/golden/ generated from that library.
 *
 * Citations are file:line relative to the reference checkout.
 */
#ifndef SA_ORACLE_H
#define SA_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Scoring model, restating scoring_t (src/alignment_scoring.h:19-40) with
 * byte-wide flags instead of bitsets; semantics identical. */
typedef struct orc_scoring {
  int gap_open, gap_extend;
  int no_start_gap_penalty, no_end_gap_penalty;
  int no_gaps_in_a, no_gaps_in_b, no_mismatches;
  int use_match_mismatch, match, mismatch;
  int case_sensitive;
  int min_penalty, max_penalty;
  unsigned char is_wild[256];
  int wild_score[256];
  unsigned char has_swap[256][256];
  int swap_score[256][256];
} orc_scoring_t;

size_t orc_scoring_sizeof(void);

void orc_scoring_init(orc_scoring_t *s, int match, int mismatch,
                      int gap_open, int gap_extend,
                      int no_start_gap_penalty, int no_end_gap_penalty,
                      int no_gaps_in_a, int no_gaps_in_b,
                      int no_mismatches, int case_sensitive);
void orc_scoring_add_wildcard(orc_scoring_t *s, int c, int score);
void orc_scoring_add_mutation(orc_scoring_t *s, int a, int b, int score);
void orc_scoring_add_mutations(orc_scoring_t *s, const char *letters,
                               const int *scores, int use_match_mismatch);
/* Direct field pokes, as the CLI does after scoring_init
 * (src/alignment_cmdline.c:401-439, src/tools/sw_cmdline.c:42-45):
 * min_penalty/max_penalty are NOT recomputed. */
void orc_scoring_poke(orc_scoring_t *s, int match, int mismatch,
                      int gap_open, int gap_extend);

/* returns 0 and sets score/is_match; returns -1 for the reference's fatal
 * "Unknown character pair" case (score/is_match then undefined). */
int orc_scoring_lookup(const orc_scoring_t *s, int a, int b,
                       int *score, int *is_match);

/* Fill the three (lb+1) x (la+1) row-major matrices. index = y*(la+1)+x. */
int orc_fill(const orc_scoring_t *s, const char *a, size_t la,
             const char *b, size_t lb, int is_sw,
             int *match, int *gap_a, int *gap_b);

/* One backward step. state: 0 MATCH, 1 GAP_A, 2 GAP_B.
 * Returns 0, or -1 on the reference's "traceback fail". */
int orc_reverse_move(const orc_scoring_t *s, const char *a, size_t la,
                     const char *b, size_t lb,
                     const int *match, const int *gap_a, const int *gap_b,
                     int *state, int *score, size_t *x, size_t *y);

typedef struct orc_alignment {
  char *result_a, *result_b; /* caller-owned, capacity >= la+lb+1 */
  size_t length;
  size_t pos_a, pos_b, len_a, len_b;
  int score;
} orc_alignment_t;

/* Needleman-Wunsch: fill + end-state choice + traceback.
 * Returns 0, -1 traceback fail, -2 unknown character pair. */
int orc_nw_align(const orc_scoring_t *s, const char *a, size_t la,
                 const char *b, size_t lb, orc_alignment_t *out);

/* Needleman-Wunsch final score only (max of the three end cells). */
int orc_nw_score(const orc_scoring_t *s, const char *a, size_t la,
                 const char *b, size_t lb, int *score);

/* Smith-Waterman best cell: max match score, ties -> smallest x then
 * smallest y (glibc qsort_r is a stable merge sort; indices are generated
 * in ascending order).  score 0 => no hit, x_end = y_end = 0. */
int orc_sw_best(const orc_scoring_t *s, const char *a, size_t la,
                const char *b, size_t lb,
                int *score, size_t *x_end, size_t *y_end);

/* Smith-Waterman hit iteration on a FRESH aligner (fresh visited mask).
 * Writes up to max_hits hits; strings of hit i are placed at
 * pool_a + i*stride / pool_b + i*stride (stride >= la+lb+1), NUL-terminated.
 * Returns number of hits written (>=0) or negative error as above. */
long orc_sw_hits(const orc_scoring_t *s, const char *a, size_t la,
                 const char *b, size_t lb, size_t max_hits,
                 orc_alignment_t *hits, char *pool_a, char *pool_b,
                 size_t stride);

/* Batch helpers used by bench.py's cpu_baseline leg ("port" kind) and tests.
 * seqs are packed back to back; off_a/off_b have n+1 entries. */
int orc_batch_sw_best(const orc_scoring_t *s, size_t n,
                      const char *seq_a, const long long *off_a,
                      const char *seq_b, const long long *off_b,
                      int *score, int *x_end, int *y_end);
int orc_batch_nw_score(const orc_scoring_t *s, size_t n,
                       const char *seq_a, const long long *off_a,
                       const char *seq_b, const long long *off_b,
                       int *score);

/* ---- sequence-file reader ------------------------------
 * Restates the record grammar of the reference's reader as align_from_file()
 * drives it (src/alignment_cmdline.c:570-622: seq_open(path) = zlib + 1 MiB
 * stream buffer, seq_read() per record): libs/seq_file/seq_file.h:311-323
 * (_read_unknown), :245-272 (FASTQ), :274-295 (FASTA), :298-309 (plain).
 * NB seq_read() calls sf->readfunc, which is only ever set to the "unknown
 * format" reader (:97, :434-437; :318-320 re-point origreadfunc, not readfunc),
 * so EVERY record starts with the white-space skip and the format choice of
 * :315-320; the format is per record, not per file.
 *
 * Parity PINNED against the reference's reader itself: oracle/ref_reader.c
 * includes the unmodified seq_file.h and dumps its records
 * (tests/test_reader.py, tests/golden/reader_vectors.json).
 *
 * Known hazard, excluded from parity (H6): for a non-newline white-space
 * character in front of a record the buffered reader calls the UNbuffered
 * skipline on the file behind its buffer (:426-427 pass _sf_gzskipline to the
 * buffered _read_unknown) -- a no-op once the whole file is in the buffer
 * (files <= 1 MiB), a loss of not-yet-buffered bytes otherwise.  Restated here as
 * the no-op.  Bytes >= 0x80 are outside the contract (0xFF reads as EOF, H4).
 *
 * text[0..n) -> records.  Sequence bytes of record i: seq[seq_off[i]..seq_off[i+1]);
 * its name (header line without '>' / '@', chomped): text[name_pos[i] ..+name_len[i]);
 * rec_pos[i] = text offset of the record's first character; fmt[i] = 1 plain, 2 FASTA,
 * 4 FASTQ (seq_format values, seq_file.h:34-39).  Arrays hold max_rec (+1 for seq_off)
 * entries, seq holds n bytes.  Returns the number of records; *last = what the
 * final seq_read() returned (0 end of input, -1 syntax error / truncated record). */
long orc_read_records(const char *text, size_t n, size_t max_rec,
                      char *seq, long long *seq_off,
                      long long *name_pos, long long *name_len,
                      long long *rec_pos, int *fmt, int *last);

#ifdef __cplusplus
}
#endif
#endif
