/*
 * seq_align.h -- version macros of the seq-align C API as provided by the
 * B200-native implementation.
 *
 * Replaces: reference src/seq_align.h:12-13 (same macro names and values so
 * that callers testing SEQ_ALIGN_VERSION keep working).
 */
#ifndef SEQ_ALIGN_HEADER_SEEN
#define SEQ_ALIGN_HEADER_SEEN

#define SEQ_ALIGN_VERSION_STR "1.0.0"
#define SEQ_ALIGN_VERSION 0x100

/* set by this implementation only: lets callers detect the GPU build */
#define SEQ_ALIGN_B200 1

#endif
