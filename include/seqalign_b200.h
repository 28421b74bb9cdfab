/*
 * seqalign_b200.h -- C-ABI of the B200-native batch alignment engine.
 *
 * The reference (noporpoise/seq-align) has no batch API: its only entry to
 * the DP is aligner_align() (reference src/alignment.h:56-59, one pair per
 * call, src/alignment.c:170-193), driven pair-by-pair from
 * src/alignment_cmdline.c:611-622.  A GPU needs thousands of independent
 * pairs per launch, so this header adds the batch entry points a maintainer
 * would bind (SURVEY.md 8b "NEW (additive) batch entry points"); the classic
 * single-pair functions in alignment.h / needleman_wunsch.h /
 * smith_waterman.h are implemented on top of them as batches of one.
 *
 * Plain C: pointers, sizes and ints only.  No CUDA or torch types appear in
 * any signature (device pointers and streams travel as void* / const void*).
 * The library fails loudly (error code + message, never a CPU fallback) when
 * no sm_100 device is usable.
 *
 * Each function cites the reference code whose work it takes over.
 */
#ifndef SEQALIGN_B200_H
#define SEQALIGN_B200_H

#include <stddef.h>
#include <stdint.h>
#include "alignment.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct seqalign_batch seqalign_batch_t;

/* algorithm: is_sw argument of aligner_align (reference src/alignment.h:59) */
#define SEQALIGN_NW 0
#define SEQALIGN_SW 1

/* what a submit computes */
#define SEQALIGN_MODE_SCORE 0 /* score (+ best SW cell): fill only           */
#define SEQALIGN_MODE_ALIGN 1 /* + traceback: gapped strings, pos/len fields */
#define SEQALIGN_MODE_HITS 3  /* SW: every local hit in the reference's order,
                                 up to the limits of seqalign_batch_set_hit_limits */
#define SEQALIGN_MODE_MATS 4  /* SW or NW: the three DP matrices of every pair, kept
                                 on the device (seqalign_batch_matrices)         */
#define SEQALIGN_MODE_SCORE_ONLY 2 /* scores only (x_end/y_end = 0): lets the
                                      engine use its packed 16-bit SW kernel   */

/* error codes (negative returns) */
#define SEQALIGN_OK 0
#define SEQALIGN_ERR_CUDA (-1)      /* CUDA runtime / no usable device         */
#define SEQALIGN_ERR_UNKNOWN_PAIR (-2) /* reference alignment_scoring.c:179-181 */
#define SEQALIGN_ERR_TRACEBACK (-3) /* reference alignment.c:328-349           */
#define SEQALIGN_ERR_ARG (-4)
#define SEQALIGN_ERR_NOMEM (-5)
#define SEQALIGN_ERR_IRREGULAR (-6) /* seqalign_reads_decode: text outside the device grammar */

/* number of usable sm_100 devices (0 if none / no driver) */
int seqalign_device_count(void);
const char *seqalign_version(void);

/* One engine per device and host thread.  NULL on failure (message through
 * seqalign_last_create_error()). */
seqalign_batch_t *seqalign_batch_create(int device);
void seqalign_batch_destroy(seqalign_batch_t *eng);
const char *seqalign_last_create_error(void);
/* message of the last failing call on this engine ("" if none) */
const char *seqalign_batch_error(const seqalign_batch_t *eng);

/* Snapshot the scoring model.  Takes over the per-cell scoring_lookup()
 * (reference src/alignment.c:98 -> src/alignment_scoring.c:133-182): the
 * 256x256 table, wildcards, case folding and match/mismatch fallback are
 * flattened into a dense code table on the host, once per batch.
 * min_penalty is carried verbatim (it defines the NW sentinel,
 * reference src/alignment.c:41). */
int seqalign_batch_set_scoring(seqalign_batch_t *eng, const scoring_t *scoring);

/* Align n independent pairs: replaces n calls of aligner_align()
 * (+ the traceback of needleman_wunsch_align2 / the first
 * smith_waterman_fetch in MODE_ALIGN).  Sequences are raw characters with
 * explicit lengths, no NUL needed (as aligner_align).  Blocking: results are
 * on the host when it returns. */
int seqalign_batch_submit(seqalign_batch_t *eng, int algo, int mode,
                          const char *const *seq_a, const size_t *len_a,
                          const char *const *seq_b, const size_t *len_b,
                          size_t n);

/* Same, sequences packed back to back: pair i is
 * seq_a[off_a[i] .. off_a[i+1]) vs seq_b[off_b[i] .. off_b[i+1]);
 * off_* have n+1 entries.  One host->device copy per array. */
int seqalign_batch_submit_packed(seqalign_batch_t *eng, int algo, int mode,
                                 const char *seq_a, const int64_t *off_a,
                                 const char *seq_b, const int64_t *off_b,
                                 size_t n);

/* Fixed-length read sets (every BASELINE config): pair i is seq_a[i*len_a ..
 * (i+1)*len_a) vs seq_b[i*len_b .. (i+1)*len_b).  No offset arrays exist on the
 * host, none cross PCIe (the device makes its own) and nothing per pair is
 * touched by the CPU in the score modes.  Same results and accessors as
 * seqalign_batch_submit_packed. */
int seqalign_batch_submit_uniform(seqalign_batch_t *eng, int algo, int mode,
                                  const char *seq_a, size_t len_a,
                                  const char *seq_b, size_t len_b, size_t n);

/* Where the score modes (SEQALIGN_MODE_SCORE / _SCORE_ONLY) of the following
 * submits leave their results: host arrays of n int32 each, written by the
 * device->host copies themselves (pinned memory: no CPU copy at all).  x_end /
 * y_end may be NULL.  score == NULL switches back to the engine's own arrays
 * (seqalign_batch_scores / _ends, which fail with SEQALIGN_ERR_ARG after a
 * submit into a sink).  The arrays must stay valid until the submit returns. */
int seqalign_batch_set_result_sink(seqalign_batch_t *eng, int32_t *score, int32_t *x_end, int32_t *y_end);

/* Results of the last submit (host arrays of n entries each).
 * score: NW = max of the three matrices at [len_a,len_b]
 *        (reference needleman_wunsch.c:53-66); SW = best match score, 0 if
 *        the pair has no hit.
 * x_end/y_end: SW = 1-based cell of the best hit under the reference's hit
 *        order (score desc, x asc, y asc: smith_waterman.c:71-86 + stable
 *        glibc qsort_r); 0,0 if no hit.  NW = len_a,len_b. */
int seqalign_batch_scores(seqalign_batch_t *eng, int32_t *score);
int seqalign_batch_ends(seqalign_batch_t *eng, int32_t *score,
                        int32_t *x_end, int32_t *y_end);
size_t seqalign_batch_size(const seqalign_batch_t *eng);

/* MODE_ALIGN: copy pair i of the last submit into a reference alignment_t
 * (grown with alignment_ensure_capacity).  NW fills result_a/result_b/
 * length/score like needleman_wunsch_align2 (needleman_wunsch.c:34-145); SW
 * fills them plus pos_a/pos_b/len_a/len_b like the first
 * smith_waterman_fetch on a fresh aligner (smith_waterman.c:165-277).
 * Returns 1 if an alignment was written, 0 if the SW pair has no hit,
 * negative on error. */
int seqalign_batch_alignment(seqalign_batch_t *eng, size_t i, alignment_t *out);

/* MODE_HITS (Smith-Waterman): the whole hit iteration of
 * smith_waterman_align2 + repeated smith_waterman_fetch on a fresh aligner
 * (reference src/smith_waterman.c:152-161, 165-277) runs on the device:
 * candidates with match score >= min_score (and > 0) sorted by score desc,
 * x asc, y asc; walked back in that order through a visited mask; a walk that
 * meets a marked cell is dropped.  At most max_hits hits per pair are kept
 * (the reference's CLI stops the same way: src/tools/sw_cmdline.c:214-217).
 * Defaults: max_hits 8, min_score 1.  Available for the scoring shapes of the
 * specialised kernel (SEQALIGN_ERR_ARG otherwise). */
int seqalign_batch_set_hit_limits(seqalign_batch_t *eng, size_t max_hits, int32_t min_score);
size_t seqalign_batch_hit_count(const seqalign_batch_t *eng, size_t i);
/* hit h of pair i into a reference alignment_t (all fields); 1 if written */
int seqalign_batch_hit(seqalign_batch_t *eng, size_t i, size_t h, alignment_t *out);

/* MODE_MATS (Smith-Waterman or Needleman-Wunsch): aligner_align() for a whole batch -- the match /
 * gap_a / gap_b matrices of every pair exactly as alignment_fill_matrices
 * leaves them (reference src/alignment.c:28-168), (len_a+1)*(len_b+1) ints
 * each, index = y*(len_a+1)+x, borders included.  They stay in device memory
 * (12 bytes per cell: SEQALIGN_ERR_NOMEM if the batch does not fit); this
 * copies pair i's three planes into host arrays.  seqalign_batch_scores()
 * gives the best match score per pair (SW) / the max of the three matrices
 * at [len_a][len_b] (NW; the border cells carry the reference's INT_MIN-based
 * sentinel of alignment.c:41 exactly).  Available for the scoring shapes of
 * the specialised kernel with len_a <= 511: affine gaps with gap_open <= 0, no
 * gap / mismatch restrictions; free start / end gaps for NW only
 * (SEQALIGN_ERR_ARG otherwise: use seqalign_fill_matrices pair by pair). */
int seqalign_batch_matrices(seqalign_batch_t *eng, size_t i, int32_t *match,
                            int32_t *gap_a, int32_t *gap_b);

/* Device-resident variant (score mode): all pointers are device memory on
 * the engine's device, stream is a cudaStream_t (NULL = engine's own).
 * d_x_end/d_y_end may be NULL (score only: lets the engine use its packed
 * 16-bit kernel).  Returns when the results are in d_score.  d_seq_a and
 * d_seq_b must be 16-byte aligned (SEQALIGN_ERR_ARG otherwise) and readable up
 * to the next 16-byte boundary past their end: the kernels stage sequences
 * with 16-byte bulk copies aligned on the offset.  A shard of a larger batch
 * is passed as the batch's base pointers plus a window of its offset arrays
 * (offsets are absolute), never as a pointer into the middle of the buffer. */
int seqalign_batch_run_device(seqalign_batch_t *eng, int algo,
                              const void *d_seq_a, const void *d_off_a,
                              const void *d_seq_b, const void *d_off_b,
                              size_t n, void *d_score, void *d_x_end,
                              void *d_y_end, void *stream);

/* Let the kernels of `device` read buffers that live on `peer` (same node,
 * NVLink / NVSwitch): cudaDeviceEnablePeerAccess.  With it the device
 * pointers handed to seqalign_batch_run_device may be peer memory, e.g. a
 * batch held by another rank and mapped through CUDA IPC -- the DP kernels'
 * bulk loads then pull the sequences across NVLink while they compute and no
 * scatter step is needed (seqalign.distributed.align_sharded_peer).  The
 * reference has no counterpart (single process, single thread). */
int seqalign_enable_peer_access(int device, int peer);

/* A batch that one process holds in the HBM of its GPU and other processes
 * (one per GPU, same node) align in place: the owner allocates the buffer
 * and passes the 64-byte handle to its peers by any means (MPI, a pipe,
 * torch.distributed); a peer opens it ON ITS OWN DEVICE, which maps the
 * owner's memory for that device's kernels (cudaIpcOpenMemHandle with lazy
 * peer access, NVLink / NVSwitch).  Pointers into the mapping are valid
 * arguments of seqalign_batch_run_device on the peer. */
typedef struct { unsigned char bytes[64]; } seqalign_ipc_handle_t;
int seqalign_shared_alloc(int device, size_t bytes, void **d_ptr, seqalign_ipc_handle_t *handle);
int seqalign_shared_free(int device, void *d_ptr);
int seqalign_shared_open(int device, const seqalign_ipc_handle_t *handle, void **d_ptr);
int seqalign_shared_close(int device, void *d_ptr);

/* The same without waiting: when the engine has a plan to guess from (the
 * previous run's: same scoring, algorithm and outputs) the kernels are only
 * ENQUEUED on `stream` and the call returns; up to 4 runs may be outstanding,
 * so a caller that streams batches keeps the GPU busy across the host's launch
 * and synchronisation latency.  seqalign_batch_run_device_wait() completes the
 * OLDEST outstanding run: its results are in its d_score when it returns (a run
 * whose batch did not fit the guessed plan -- and every run enqueued after it
 * -- is redone right there).  Without a plan to guess from (first run, new
 * scoring) the call is the blocking one and wait() has nothing to do.  Any
 * other entry point of the engine first waits for everything outstanding. */
int seqalign_batch_run_device_async(seqalign_batch_t *eng, int algo,
                                    const void *d_seq_a, const void *d_off_a,
                                    const void *d_seq_b, const void *d_off_b,
                                    size_t n, void *d_score, void *d_x_end,
                                    void *d_y_end, void *stream);
int seqalign_batch_run_device_wait(seqalign_batch_t *eng);

/* seqalign_batch_run_device launches speculatively with the previous run's
 * plan (same scoring, algorithm and outputs) and verifies against this
 * batch's scan afterwards; these count how often the guess held / was redone. */
void seqalign_batch_speculation_stats(const seqalign_batch_t *eng, int *hits, int *misses);

/* Materialise mode, one pair: the literal aligner_align() contract
 * (reference src/alignment.c:28-168).  match/gap_a/gap_b are host arrays of
 * (len_a+1)*(len_b+1) ints, row-major, index = y*(len_a+1)+x, borders
 * included. */
int seqalign_fill_matrices(seqalign_batch_t *eng,
                           const char *seq_a, size_t len_a,
                           const char *seq_b, size_t len_b, int is_sw,
                           int32_t *match, int32_t *gap_a, int32_t *gap_b);

/* characters of the first unknown pair after SEQALIGN_ERR_UNKNOWN_PAIR
 * (case-folded, as the reference prints them) */
void seqalign_batch_unknown_pair(const seqalign_batch_t *eng, char *a, char *b);

/* Page-locked host memory for callers without CUDA headers: input buffers and
 * result sinks allocated here travel over PCIe by DMA without a staging copy
 * (cudaMallocHost / cudaFreeHost).  NULL on failure. */
/* Classic single-pair API (needleman_wunsch_align*, smith_waterman_align*): the three matrices of
 * the caller's aligner_t are by default filled only when something in this library reads them
 * (alignment_print_matrices, alignment_reverse_move, aligner_align itself always fills).  on = 1
 * restores the reference's behaviour for callers that index aligner->match_scores[] themselves:
 * filled on every call (12 bytes per cell over PCIe).  Also: environment SEQALIGN_EAGER_MATRICES=1. */
void seqalign_host_eager_matrices(int on);

void *seqalign_host_alloc(size_t bytes);
void seqalign_host_free(void *p);

/* ---- sequence-file text -> packed records on the device (csrc/sa_decode.cu) ----
 * Takes over the reference's record reader for whole chunks of file text:
 * libs/seq_file/seq_file.h:245-325 (FASTA / FASTQ / one sequence per line,
 * format chosen from the first non-blank character) as driven by
 * align_from_file (src/alignment_cmdline.c:578-640).  The text crosses PCIe
 * once; the decoded sequences stay in HBM in the engine's own batch layout
 * and seqalign_batch_submit_reads() aligns them in place.  Inflating gzip
 * stays with the caller (zlib on the host, as in the reference).
 *
 * seqalign_reads_decode(r, text, bytes, final, split)
 *   text   bytes of the file, starting at a record boundary (the start of the
 *          file, or what seqalign_reads_record_start() said to carry over)
 *   final  1 = the text ends the input; 0 = more follows, the trailing record
 *          is held back (it may be incomplete)
 *   split  0 = every record goes to side 0; 1 = records alternate side 0,
 *          side 1 (pairs from consecutive records of one file)
 *   returns SEQALIGN_OK, or SEQALIGN_ERR_IRREGULAR when the text leaves the
 *   grammar the device takes (wrapped FASTQ, '>' / '@' records inside a
 *   one-per-line file ...: sa_decode.cu lists it) -- nothing was decoded and
 *   the input has to be read by the host reader (align_from_file), which
 *   implements the reference's full grammar.
 * Accessors after a successful decode:
 *   _format          SEQALIGN_FMT_* of the chunk (values of seq_format, seq_file.h:34-39)
 *   _records         complete records
 *   _count(side)     complete records on a side
 *   _offsets(side)   host array, offsets of every record of the side in its packed
 *                    buffer, one closing entry behind the last record
 *   _record_start(i) offset in `text` where record i starts; i == _records: where the
 *                    held-back tail starts (carry text[that..] to the next call)
 *   _name(i)         span of record i's name inside `text` (header line without its
 *                    '>' / '@', line end dropped)
 *   _fetch(side,out) packed sequence bytes of a side to host memory
 *   _device_seq / _device_offsets   the same buffers in HBM (const void *)
 */
typedef struct seqalign_reads seqalign_reads_t;
#define SEQALIGN_FMT_PLAIN 1
#define SEQALIGN_FMT_FASTA 2
#define SEQALIGN_FMT_FASTQ 4
seqalign_reads_t *seqalign_reads_create(int device);
void seqalign_reads_destroy(seqalign_reads_t *r);
const char *seqalign_reads_error(const seqalign_reads_t *r);
int seqalign_reads_decode(seqalign_reads_t *r, const char *text, size_t bytes, int final, int split);
int seqalign_reads_format(const seqalign_reads_t *r);
size_t seqalign_reads_records(const seqalign_reads_t *r);
size_t seqalign_reads_count(const seqalign_reads_t *r, int side);
const int64_t *seqalign_reads_offsets(const seqalign_reads_t *r, int side);
size_t seqalign_reads_record_start(const seqalign_reads_t *r, size_t i);
int seqalign_reads_name(const seqalign_reads_t *r, size_t i, size_t *pos, size_t *len);
int seqalign_reads_fetch(seqalign_reads_t *r, int side, char *out);
const void *seqalign_reads_device_seq(const seqalign_reads_t *r, int side);
const void *seqalign_reads_device_offsets(const seqalign_reads_t *r, int side);
int seqalign_reads_device(const seqalign_reads_t *r);
double seqalign_reads_last_ms(const seqalign_reads_t *r);   /* H2D of the text + all decode kernels, CUDA events */

/* Align records [first, first + n) of side_a of `ra` against records
 * [first, first + n) of side_b of `rb` (the same object for pairs from one
 * file), reading the sequences where the decoder left them in HBM.  Same
 * modes, same result calls as seqalign_batch_submit_packed() (pair i of the
 * submit is record first + i).  Engine and reads objects must be on the same
 * device. */
int seqalign_batch_submit_reads(seqalign_batch_t *eng, int algo, int mode,
                                const seqalign_reads_t *ra, int side_a,
                                const seqalign_reads_t *rb, int side_b, size_t first, size_t n);

/* ---- one batch over several GPUs of one node, one process (host/sa_multi.c) ----
 * The reference's loop over pairs (src/alignment_cmdline.c:611-622) sharded by
 * pair index: one engine per device, contiguous pair ranges balanced by cell
 * count, one host thread per device for the duration of a submit, every device
 * fed over its own PCIe link.  Pairs are independent, so nothing is exchanged
 * between devices; results are addressed by the pair's index in the submitted
 * batch, exactly like the single-engine calls.  The offset arrays must stay
 * valid until the results have been read (seqalign_multi_ends in score-only
 * mode reads the lengths from them).  devices == NULL: devices 0..n-1;
 * n_devices <= 0: every usable device.  A device may be listed twice (two
 * engines on it). */
typedef struct seqalign_multi seqalign_multi_t;
seqalign_multi_t *seqalign_multi_create(const int *devices, int n_devices);
void seqalign_multi_destroy(seqalign_multi_t *m);
int seqalign_multi_devices(const seqalign_multi_t *m);
const char *seqalign_multi_error(const seqalign_multi_t *m);
int seqalign_multi_set_scoring(seqalign_multi_t *m, const scoring_t *scoring);
int seqalign_multi_set_hit_limits(seqalign_multi_t *m, size_t max_hits, int32_t min_score);
int seqalign_multi_submit_packed(seqalign_multi_t *m, int algo, int mode,
                                 const char *seq_a, const int64_t *off_a,
                                 const char *seq_b, const int64_t *off_b, size_t n);
int seqalign_multi_submit_uniform(seqalign_multi_t *m, int algo, int mode,
                                  const char *seq_a, size_t len_a,
                                  const char *seq_b, size_t len_b, size_t n);
size_t seqalign_multi_size(const seqalign_multi_t *m);
int seqalign_multi_scores(seqalign_multi_t *m, int32_t *score);
int seqalign_multi_ends(seqalign_multi_t *m, int32_t *score, int32_t *x_end, int32_t *y_end);
int seqalign_multi_alignment(seqalign_multi_t *m, size_t i, alignment_t *out);
size_t seqalign_multi_hit_count(seqalign_multi_t *m, size_t i);
int seqalign_multi_hit(seqalign_multi_t *m, size_t i, size_t h, alignment_t *out);
int seqalign_multi_matrices(seqalign_multi_t *m, size_t i, int32_t *match, int32_t *gap_a, int32_t *gap_b);
/* which device aligned pair i of the last submit, and its index there */
int seqalign_multi_where(seqalign_multi_t *m, size_t i, int *device, size_t *local_index);
void seqalign_multi_unknown_pair(const seqalign_multi_t *m, char *a, char *b);
double seqalign_multi_last_kernel_ms(const seqalign_multi_t *m);   /* slowest device */

/* Synthetic pairs of SURVEY.md 8d, generated on the device: pairs
 * [first_pair, first_pair + npairs) of the counter-based stream `seed`,
 * fixed lengths, kind 0 = DNA (ACGT; seq_b = seq_a with 5 % substitutions, 1 %
 * insertions, 1 % deletions), kind 1 = protein (20 letters; 15 % / 2 % / 2 %).
 * A pair depends only on (seed, pair index), so any rank can make any shard
 * of a job (csrc/sa_synth.cuh; seqalign/synth.py is the same function in
 * numpy).  d_seq_a / d_seq_b: device buffers of npairs*len_a / npairs*len_b
 * bytes.  The reference reads its pairs from files
 * (src/alignment_cmdline.c:578-640) and has no generator.  stream NULL: the
 * call returns when the buffers are filled. */
int seqalign_synth_batch(int device, int kind, uint64_t seed, int64_t first_pair, int64_t npairs,
                         int len_a, int len_b, void *d_seq_a, void *d_seq_b, void *stream);

/* Instrumentation for bench.py: device time (ms, CUDA events on the
 * engine's stream) and launch count of the DP kernels of the last
 * submit/run, and the name of the kernel variant that ran. */
double seqalign_batch_last_kernel_ms(const seqalign_batch_t *eng);
double seqalign_batch_last_walk_ms(const seqalign_batch_t *eng);   /* traceback walk kernels of an align-mode submit */
int seqalign_batch_last_launches(const seqalign_batch_t *eng);
const char *seqalign_batch_last_kernel(const seqalign_batch_t *eng);

/* kernel selection knob for tests: 0 = automatic; 1 = general kernel only;
 * 2 = specialised kernel with per-column end-cell keys; 3 = specialised
 * kernel without end-cell tracking (x_end/y_end come back 0; packed 16-bit
 * kernel when the batch qualifies); 4 = like 3 but int32 arithmetic only;
 * 5 = automatic, but int32 arithmetic only. */
void seqalign_batch_force_general(seqalign_batch_t *eng, int on);

#ifdef __cplusplus
}
#endif

#endif
