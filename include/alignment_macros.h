/*
 * alignment_macros.h -- index / min / max helper macros of the seq-align C API.
 *
 * Drop-in for reference src/alignment_macros.h:11-24 (the reference's tools
 * include it: src/tools/sw_cmdline.c:23).  Matrices are row-major with x (the
 * position in seq_a) running fastest: index = y * width + x.
 */
#ifndef ALIGNMENT_MACROS_HEADER_SEEN
#define ALIGNMENT_MACROS_HEADER_SEEN

#define ARR_2D_INDEX(width,i,j) (((unsigned long)(j)*(width)) + (i))
#define ARR_LOOKUP(arr,width,i,j) arr[ARR_2D_INDEX((width),(i),(j))]
#define ARR_2D_X(arr_index, arr_width) ((arr_index) % (arr_width))
#define ARR_2D_Y(arr_index, arr_width) ((arr_index) / (arr_width))

#define QUOTE(str) #str

#define MAX2(x,y) ((x) >= (y) ? (x) : (y))
#define MIN2(x,y) ((x) <= (y) ? (x) : (y))
#define MAX3(x,y,z) MAX2(MAX2(x,y),z)
#define MIN3(x,y,z) MIN2(MIN2(x,y),z)
#define MAX4(w,x,y,z) MAX2(MAX2(w,x),MAX2(y,z))

#define ABSDIFF(a,b) ((a) > (b) ? (a)-(b) : (b)-(a))

#endif
