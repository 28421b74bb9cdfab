/*
 * alignment_macros.h -- the small helper macros callers of the seq-align C API expect to find.
 *
 * Same names and meaning as the reference's src/alignment_macros.h:11-24 (its tools include that header,
 * src/tools/sw_cmdline.c:23, so a drop-in needs them); written here from the documented behaviour.
 * The score matrices are row-major with x, the position in seq_a, running fastest:
 *     cell (x, y) of a matrix `width` cells wide sits at  y * width + x.
 */
#ifndef SEQALIGN_B200_ALIGNMENT_MACROS_H
#define SEQALIGN_B200_ALIGNMENT_MACROS_H

/* (x, y) <-> linear index; the row offset is computed in unsigned long, like the matrices' size_t fields */
#define ARR_2D_INDEX(width, x, y)      ((unsigned long)(y) * (width) + (x))
#define ARR_2D_X(index, width)         ((index) % (width))
#define ARR_2D_Y(index, width)         ((index) / (width))
#define ARR_LOOKUP(arr, width, x, y)   ((arr)[ARR_2D_INDEX(width, x, y)])

/* stringify */
#define QUOTE(token)                   #token

/* order statistics of two, three and four values; ties keep the first argument, arguments may be
 * evaluated more than once */
#define MIN2(a, b)                     ((b) < (a) ? (b) : (a))
#define MAX2(a, b)                     ((b) > (a) ? (b) : (a))
#define MIN3(a, b, c)                  MIN2(MIN2(a, b), c)
#define MAX3(a, b, c)                  MAX2(MAX2(a, b), c)
#define MAX4(a, b, c, d)               MAX2(MAX2(a, b), MAX2(c, d))

/* |a - b| without going through a signed difference (safe for unsigned operands) */
#define ABSDIFF(a, b)                  ((a) < (b) ? (b) - (a) : (a) - (b))

#endif
