/*
 * sa_decode.cu -- sequence-file text -> packed records, on the device
 * ("input decode on device").
 *
 * Takes over, for whole chunks of file text at a time, what the reference does
 * record by record on one host thread: libs/seq_file/seq_file.h:311-323 (white
 * space skip + format choice), :274-295 (FASTA), :245-272 (FASTQ), :298-309
 * (one sequence per line), driven by align_from_file (src/alignment_cmdline.c:
 * 578-640).  The text goes host -> HBM once; what comes out is exactly the
 * engine's batch layout -- sequence bytes packed back to back with int64
 * offsets -- so the DP kernels read it in place (seqalign_batch_submit_reads),
 * plus a small record table (where each record starts in the text, where its
 * name is) that goes back to the host.
 *
 * Grammar taken by the device ("regular" text; anything else is DECLINED with
 * SEQALIGN_ERR_IRREGULAR and the caller reads that input with the host reader,
 * host/sa_cli.c, which implements the full grammar):
 *   FASTA  every layout: header lines start with '>' (after any '\r's), every
 *          other non-empty line is sequence, line ends are \n or \r\n.
 *   plain  one record per line holding a non-blank character; leading white
 *          space and the line end are dropped.  Declined: a record starting
 *          with '>' or '@' (the reference re-chooses the format per record).
 *   FASTQ  four lines per record: '@'name, sequence (not empty, not starting
 *          with '+'), '+'..., quality at least as long as the sequence.
 *          Declined: wrapped sequence / quality lines, blank lines between
 *          records, truncated records.
 *
 * Kernels (all HBM-bound byte work; no tensor cores, nothing to reshape):
 *   nl_count / nl_fill   newline index of the text, one 8 KB tile per CTA
 *   classify             one thread per line: kind, content span, grammar checks
 *   contrib              which side (pairs from one file alternate A, B) and how many bytes
 *   emit                 one warp per line: content bytes to their place in the packed
 *                        side buffer; record table entries
 *   scan3_*              exclusive prefix sums over tiles / lines (reduce, single-CTA
 *                        scan of the partial sums, rescan + add)
 * Algorithmic bytes per text byte: 1 read + ~0.9 written (+ 4 per line of
 * index); DESIGN.md 3 K9 has the measured traffic.
 */
#include <string>
#include <vector>
#include <string.h>
#include <ctype.h>

#include "sa_platform.h"
#include "seqalign_b200.h"

namespace sa {

constexpr int DEC_TPB = 256;
constexpr int DEC_BYTES_PER_THREAD = 32;
constexpr int DEC_TILE = DEC_TPB * DEC_BYTES_PER_THREAD;   /* 8 KB of text per CTA */
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_BLOCK = DEC_TPB * SCAN_ITEMS;            /* 1024 items per CTA */

enum { LK_NONE = 0, LK_REC = 1, LK_DATA = 2 };   /* a plain record line is LK_REC | LK_DATA */

/* 0x80 in every byte of w that equals c */
__host__ __device__ __forceinline__ unsigned eq_bytes(unsigned w, unsigned c4)
{
  const unsigned x = w ^ c4;
  return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);
}

/* exclusive prefix sum of one int per thread over a CTA of DEC_TPB threads; *total = the CTA's sum */
__device__ __forceinline__ int block_exclusive_scan(int v, int *total)
{
  __shared__ int s_warp[DEC_TPB / 32 + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if(lane >= o) incl += t;
  }
  __syncthreads();   /* s_warp may still be read by a previous call */
  if(lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if(threadIdx.x == 0) {
    int run = 0;
    for(int w = 0; w < DEC_TPB / 32; w++) { const int t = s_warp[w]; s_warp[w] = run; run += t; }
    s_warp[DEC_TPB / 32] = run;
  }
  __syncthreads();
  *total = s_warp[DEC_TPB / 32];
  return s_warp[wid] + incl - v;
}

/* ---- generic exclusive scan of an int32 array (three launches) ---------- */
__global__ void __launch_bounds__(DEC_TPB) scan3_reduce(const int *__restrict__ in, int64_t n, int *__restrict__ sums)
{
  const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v = 0;
#pragma unroll
  for(int k = 0; k < SCAN_ITEMS; k++) if(base + k < n) v += in[base + k];
  int total;
  block_exclusive_scan(v, &total);
  if(threadIdx.x == 0) sums[blockIdx.x] = total;
}

/* one CTA: sums[0..nb) -> exclusive prefix in place, grand total to *total */
__global__ void __launch_bounds__(DEC_TPB) scan3_sums(int *__restrict__ sums, int64_t nb, int64_t *__restrict__ total)
{
  int64_t carry = 0;
  for(int64_t b0 = 0; b0 < nb; b0 += DEC_TPB) {
    const int64_t i = b0 + threadIdx.x;
    const int v = i < nb ? sums[i] : 0;
    int tot;
    const int ex = block_exclusive_scan(v, &tot);
    if(i < nb) sums[i] = (int)(carry + ex);
    carry += tot;
  }
  if(threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(DEC_TPB) scan3_apply(const int *__restrict__ in, int64_t n, const int *__restrict__ sums,
                                                       int *__restrict__ out)
{
  const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
  int item[SCAN_ITEMS], v = 0;
#pragma unroll
  for(int k = 0; k < SCAN_ITEMS; k++) { item[k] = base + k < n ? in[base + k] : 0; v += item[k]; }
  int total;
  int run = block_exclusive_scan(v, &total) + sums[blockIdx.x];
#pragma unroll
  for(int k = 0; k < SCAN_ITEMS; k++) {
    if(base + k < n) out[base + k] = run;
    run += item[k];
  }
}

/* ---- newline index ------------------------------------------------------ */
/* text is padded with zero bytes up to a multiple of DEC_TILE */
__device__ __forceinline__ int count_nl_32(const uint4 a, const uint4 b)
{
  const unsigned NL = 0x0a0a0a0au;
  return __popc(eq_bytes(a.x, NL)) + __popc(eq_bytes(a.y, NL)) + __popc(eq_bytes(a.z, NL)) + __popc(eq_bytes(a.w, NL)) +
         __popc(eq_bytes(b.x, NL)) + __popc(eq_bytes(b.y, NL)) + __popc(eq_bytes(b.z, NL)) + __popc(eq_bytes(b.w, NL));
}

__global__ void __launch_bounds__(DEC_TPB) nl_count_kernel(const uint8_t *__restrict__ text, int *__restrict__ tile_counts)
{
  const uint4 *p = (const uint4 *)(text + (int64_t)blockIdx.x * DEC_TILE + threadIdx.x * DEC_BYTES_PER_THREAD);
  const int c = count_nl_32(p[0], p[1]);
  int total;
  block_exclusive_scan(c, &total);
  if(threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(DEC_TPB) nl_fill_kernel(const uint8_t *__restrict__ text, const int *__restrict__ tile_base,
                                                          int *__restrict__ nl_pos)
{
  const int64_t off = (int64_t)blockIdx.x * DEC_TILE + threadIdx.x * DEC_BYTES_PER_THREAD;
  const uint4 *p = (const uint4 *)(text + off);
  const uint4 a = p[0], b = p[1];
  const int c = count_nl_32(a, b);
  int total;
  int at = block_exclusive_scan(c, &total) + tile_base[blockIdx.x];
  if(c == 0) return;
  const unsigned w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for(int k = 0; k < 8; k++) {
    unsigned m = eq_bytes(w[k], 0x0a0a0a0au);
    while(m) {
      const int bit = __ffs(m) - 1;    /* 7, 15, 23 or 31 */
      nl_pos[at++] = (int)(off + 4 * k + (bit >> 3));
      m &= m - 1;
    }
  }
}

/* ---- per line ----------------------------------------------------------- */
struct DecArgs {
  const uint8_t *text;
  int64_t n;              /* bytes of text (from the first non-blank character on) */
  const int *nl_pos;
  int64_t n_nl;           /* newlines */
  int64_t nlines;         /* n_nl (+1 when the text does not end with a newline) */
  int fmt;                /* SEQALIGN_FMT_* */
  int split;              /* records alternate side 0, 1 */
  int final;
  int64_t fq_lines;       /* FASTQ: lines of the complete records (multiple of 4) */
  uint8_t *kind;          /* per line */
  int *cs, *ce;           /* per line: content span in the text (data lines), name span (header lines) */
  int *recflag;           /* per line: 1 = starts a record */
  const int *rec_excl;    /* exclusive scan of recflag */
  int *contrib0, *contrib1;
  const int *at0, *at1;   /* exclusive scans of the contributions */
  uint8_t *out0, *out1;
  int64_t *off0, *off1;   /* per side: offset of each record of the side */
  int *rec_pos, *name_pos, *name_len;
  int *irregular;         /* first line the grammar declines (INT_MAX = none) */
};

__device__ __forceinline__ bool dec_isspace(unsigned c) { return c == ' ' || (c >= 9 && c <= 13); }

__global__ void __launch_bounds__(DEC_TPB) classify_kernel(const DecArgs A)
{
  const int64_t li = (int64_t)blockIdx.x * DEC_TPB + threadIdx.x;
  if(li >= A.nlines) return;
  const uint8_t *T = A.text;
  const int s = li == 0 ? 0 : A.nl_pos[li - 1] + 1;
  const int e = li < A.n_nl ? A.nl_pos[li] : (int)A.n;
  int kind = LK_NONE, cs = s, ce = e, rec = 0;
  bool bad = false;
  while(ce > s && T[ce - 1] == '\r') ce--;      /* chomp (stream_buffer.h:56-61) */
  if(A.fmt == SEQALIGN_FMT_PLAIN) {
    while(cs < e && dec_isspace(T[cs])) cs++;   /* seq_file.h:303, 315: white space in front of a record */
    if(cs < e) {
      if(T[cs] == '>' || T[cs] == '@') bad = true;
      kind = LK_REC | LK_DATA; rec = 1;
      if(ce < cs + 1) ce = cs + 1;
    } else { ce = cs; }
  } else if(A.fmt == SEQALIGN_FMT_FASTA) {
    while(cs < e && T[cs] == '\r') cs++;        /* seq_file.h:284: '\r' and '\n' between lines are skipped */
    if(cs < e) {
      if(ce < cs + 1) ce = cs + 1;
      if(T[cs] == '>') { kind = LK_REC; rec = 1; cs++; }   /* name = rest of the line, chomped (:280-281) */
      else kind = LK_DATA;
    } else { ce = cs; }
  } else {
    const int k = (int)(li & 3);
    if(li >= A.fq_lines) {
      /* behind the last complete record: the tail carried to the next chunk, or (end of input) nothing but blank lines */
      if(A.final && ce > s) bad = true;
      ce = cs;
    } else if(k == 0) {
      if(T[s] != '@') bad = true;
      kind = LK_REC; rec = 1; cs = s + 1;
      if(ce < cs) ce = cs;
    } else if(k == 1) {
      if(ce == s || T[s] == '+' || T[s] == '\r') bad = true;
      kind = LK_DATA;
    } else if(k == 2) {
      if(T[s] != '+') bad = true;
      ce = cs;
    } else {
      /* quality: one line at least as long as the sequence ends the record (seq_file.h:264-267) */
      const int s2 = li - 2 == 0 ? 0 : A.nl_pos[li - 3] + 1;
      int e2 = A.nl_pos[li - 2];
      while(e2 > s2 && T[e2 - 1] == '\r') e2--;
      if(ce - s < e2 - s2) bad = true;
      ce = cs;
    }
  }
  A.kind[li] = (uint8_t)kind;
  A.cs[li] = cs;
  A.ce[li] = ce;
  A.recflag[li] = rec;
  if(bad) atomicMin(A.irregular, (int)(li < 0x7fffffff ? li : 0x7ffffffe));
}

__global__ void __launch_bounds__(DEC_TPB) contrib_kernel(const DecArgs A)
{
  const int64_t li = (int64_t)blockIdx.x * DEC_TPB + threadIdx.x;
  if(li >= A.nlines) return;
  const int kind = A.kind[li];
  int c0 = 0, c1 = 0;
  if(kind & LK_DATA) {
    const int r = A.rec_excl[li] + A.recflag[li] - 1;   /* the record this line belongs to */
    const int len = A.ce[li] - A.cs[li];
    if(A.split && (r & 1)) c1 = len; else c0 = len;
  }
  A.contrib0[li] = c0;
  if(A.split) A.contrib1[li] = c1;
}

constexpr int EMIT_WARPS = DEC_TPB / 32;

/* One line's content, text[cs .. cs+len) -> dst[0 .. len), by a group of GL lanes (lane = index in the
 * group).  Source and destination have unrelated byte alignments: the body moves destination-aligned
 * 32-bit words, each assembled from the two aligned source words it straddles (one load per lane, the
 * neighbour's word by shuffle, a funnel shift); at most 3 bytes at either end go one by one.  Aligned
 * source words may reach 3 bytes outside the line: inside the text buffer all the same (its base is
 * 16-byte aligned, its end padded). */
template <int GL>
__device__ __forceinline__ void emit_line(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int len, int lane)
{
  const int head = imin((int)((4 - ((uintptr_t)dst & 3)) & 3), len);
  if(lane < head) dst[lane] = src[lane];
  const int nwords = (len - head) >> 2;
  const uint8_t *sb = src + head;
  unsigned *dw = (unsigned *)(dst + head);
  const int k8 = 8 * (int)((uintptr_t)sb & 3);
  const unsigned *sw = (const unsigned *)(sb - ((uintptr_t)sb & 3));
  /* every group of the warp runs as many rounds as the longest line among them needs (the shuffle is warp-wide) */
  int rounds = (nwords + GL - 1) / GL;
#pragma unroll
  for(int o = GL; o < 32; o <<= 1) rounds = imax(rounds, __shfl_xor_sync(0xffffffffu, rounds, o));
  for(int r = 0; r < rounds; r++) {
    const int w = r * GL + lane;
    /* the words of this round are sw[r*GL .. r*GL+GL]: lane l loads sw[r*GL+l], the last lane also the one behind */
    const unsigned lo = w <= nwords ? sw[w] : 0u;
    unsigned hi = __shfl_down_sync(0xffffffffu, lo, 1, GL);
    if(lane == GL - 1 && w < nwords && k8) hi = sw[w + 1];
    if(w < nwords) dw[w] = k8 ? (unsigned)((((unsigned long long)hi << 32) | lo) >> k8) : lo;
  }
  const int done = head + 4 * nwords;
  if(lane < len - done) dst[done + lane] = src[done + lane];
}

/* GL lanes per line: short lines (reads) leave a whole warp mostly idle and, worse, one line at a time per
 * warp makes the kernel wait out a chain of dependent loads per line (index -> span -> text); with 8 lanes
 * per line four such chains are in flight per warp */
template <int GL>
__global__ void __launch_bounds__(DEC_TPB) emit_kernel(const DecArgs A)
{
  constexpr int PER_WARP = 32 / GL;
  const int lane = threadIdx.x & 31, gl = lane % GL, grp = lane / GL;
  const int64_t warp0 = (int64_t)blockIdx.x * EMIT_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * EMIT_WARPS;
  for(int64_t l0 = warp0 * PER_WARP; l0 < A.nlines; l0 += nwarps * PER_WARP) {
    const int64_t li = l0 + grp;
    int kind = LK_NONE, cs = 0, ce = 0, r = 0, side = 0, at = 0;
    if(li < A.nlines) {
      kind = A.kind[li];
      if(kind != LK_NONE) {
        cs = A.cs[li]; ce = A.ce[li];
        r = A.rec_excl[li] + A.recflag[li] - 1;
        side = A.split ? (r & 1) : 0;
        at = side ? A.at1[li] : A.at0[li];
      }
    }
    if((kind & LK_REC) && gl == 0) {
      const int s = li == 0 ? 0 : A.nl_pos[li - 1] + 1;
      A.rec_pos[r] = s;
      /* a plain record has no name; a header's span is its name */
      A.name_pos[r] = (kind & LK_DATA) ? s : cs;
      A.name_len[r] = (kind & LK_DATA) ? 0 : ce - cs;
      /* (pointer and index picked in separate statements: gcc 13 with -fsanitize=undefined lost the index of
       * `(side ? A.off1 : A.off0)[i]` for side 0 in the emulator build) */
      int64_t *offs = side ? A.off1 : A.off0;
      const int slot = A.split ? (r >> 1) : r;
      offs[slot] = (int64_t)at;
    }
    /* every lane takes part (warp-wide shuffles inside); lines without data copy nothing */
    const int len = (kind & LK_DATA) ? ce - cs : 0;
    emit_line<GL>(A.text + cs, (side ? A.out1 : A.out0) + at, len, gl);
  }
}

} // namespace sa

/* =========================================================================
 * host side + C-ABI
 */
using namespace sa;

namespace {
struct Buf {
  void *p = nullptr;
  size_t cap = 0;
  bool pinned = false;
  void release() { if(p) { if(pinned) cudaFreeHost(p); else cudaFree(p); } p = nullptr; cap = 0; }
};
}

struct seqalign_reads {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  std::string err;
  Buf d_text, d_tiles, d_nl, d_kind, d_cs, d_ce, d_recflag, d_recex, d_c0, d_c1, d_a0, d_a1, d_sums, d_out0, d_out1,
      d_off0, d_off1, d_recpos, d_npos, d_nlen, d_scalars;
  Buf h_scalars, h_off0, h_off1, h_recpos, h_npos, h_nlen;
  /* result of the last decode */
  int fmt = 0;
  int split = 0;
  size_t p0 = 0;              /* text offset of the first non-blank character */
  size_t bytes = 0;
  size_t total_records = 0;   /* records seen (the last one may be incomplete) */
  size_t records = 0;         /* complete records */
  int64_t side_bytes[2] = {0, 0};
  int irregular_line = -1;
  size_t tail_start = 0;      /* text offset where the held-back tail starts */
  double last_ms = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

int rfail(seqalign_reads *r, int code, const std::string &msg) { r->err = msg; return code; }

#define RCU(call)                                                                                      \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if(e_ != cudaSuccess) {                                                                            \
      cudaGetLastError();                                                                              \
      return rfail(r, SEQALIGN_ERR_CUDA, std::string("CUDA error (") + cudaGetErrorString(e_) + ") at " #call); \
    }                                                                                                  \
  } while(0)

int need(seqalign_reads *r, Buf &b, size_t bytes, bool pinned = false)
{
  bytes = (bytes + 255) / 256 * 256 + 256;
  if(b.cap >= bytes) return 0;
  b.release();
  const size_t want = bytes + bytes / 4;
  b.pinned = pinned;
  const cudaError_t e = pinned ? cudaMallocHost(&b.p, want) : cudaMalloc(&b.p, want);
  if(e != cudaSuccess) { b.p = nullptr; cudaGetLastError(); return rfail(r, SEQALIGN_ERR_NOMEM, pinned ? "out of pinned host memory" : "out of device memory"); }
  b.cap = want;
  return 0;
}
#define RTRY(x) do { const int rc_ = (x); if(rc_ != 0) return rc_; } while(0)

/* exclusive scan of in[0..n) into out, grand total to d_total (device) */
int scan_i32(seqalign_reads *r, const int *in, int64_t n, int *out, int64_t *d_total)
{
  const int64_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  RTRY(need(r, r->d_sums, (size_t)(nb + 1) * 4));
  int *sums = (int *)r->d_sums.p;
  if(nb == 0) { RCU(cudaMemsetAsync(d_total, 0, 8, r->stream)); return 0; }
  SA_LAUNCH(scan3_reduce, (unsigned)nb, DEC_TPB, 0, r->stream, in, n, sums);
  SA_LAUNCH(scan3_sums, 1, DEC_TPB, 0, r->stream, sums, nb, d_total);
  SA_LAUNCH(scan3_apply, (unsigned)nb, DEC_TPB, 0, r->stream, in, n, (const int *)sums, out);
  RCU(cudaGetLastError());
  return 0;
}

} // namespace

extern "C" {

seqalign_reads_t *seqalign_reads_create(int device)
{
  int n = 0;
  if(cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); return nullptr; }
  cudaDeviceProp p;
  if(cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&p, device) != cudaSuccess || p.major != 10) {
    cudaGetLastError();
    return nullptr;   /* sm_100a kernels only; no CPU path */
  }
  seqalign_reads *r = new seqalign_reads();
  r->device = device;
  r->num_sms = p.multiProcessorCount;
  if(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&r->ev0) != cudaSuccess ||
     cudaEventCreate(&r->ev1) != cudaSuccess) {
    cudaGetLastError();
    delete r;
    return nullptr;
  }
  return r;
}

void seqalign_reads_destroy(seqalign_reads_t *r)
{
  if(!r) return;
  cudaSetDevice(r->device);
  Buf *all[] = {&r->d_text, &r->d_tiles, &r->d_nl, &r->d_kind, &r->d_cs, &r->d_ce, &r->d_recflag, &r->d_recex, &r->d_c0, &r->d_c1,
                &r->d_a0, &r->d_a1, &r->d_sums, &r->d_out0, &r->d_out1, &r->d_off0, &r->d_off1, &r->d_recpos, &r->d_npos,
                &r->d_nlen, &r->d_scalars, &r->h_scalars, &r->h_off0, &r->h_off1, &r->h_recpos, &r->h_npos, &r->h_nlen};
  for(Buf *b : all) b->release();
  if(r->ev0) cudaEventDestroy(r->ev0);
  if(r->ev1) cudaEventDestroy(r->ev1);
  if(r->stream) cudaStreamDestroy(r->stream);
  delete r;
}

const char *seqalign_reads_error(const seqalign_reads_t *r) { return r ? r->err.c_str() : "null reads object"; }

int seqalign_reads_decode(seqalign_reads_t *r, const char *text, size_t bytes, int final, int split)
{
  if(!r || (!text && bytes)) return SEQALIGN_ERR_ARG;
  r->err.clear();
  r->fmt = 0; r->split = split ? 1 : 0; r->records = r->total_records = 0; r->side_bytes[0] = r->side_bytes[1] = 0;
  r->bytes = bytes; r->irregular_line = -1; r->last_ms = 0; r->tail_start = bytes;
  if(bytes > ((size_t)1 << 30)) return rfail(r, SEQALIGN_ERR_ARG, "chunks of text are limited to 1 GiB");
  /* white space in front of the first record, then the format from its first character (seq_file.h:315-320) */
  size_t p0 = 0;
  while(p0 < bytes && isspace((unsigned char)text[p0])) p0++;
  r->p0 = p0;
  if(p0 == bytes) return SEQALIGN_OK;   /* nothing but white space */
  const int fmt = text[p0] == '@' ? SEQALIGN_FMT_FASTQ : text[p0] == '>' ? SEQALIGN_FMT_FASTA : SEQALIGN_FMT_PLAIN;
  r->fmt = fmt;
  const int64_t n = (int64_t)(bytes - p0);
  const bool tail_line = text[bytes - 1] != '\n';
  RCU(cudaSetDevice(r->device));
  cudaStream_t st = r->stream;

  const int64_t ntiles = (n + DEC_TILE - 1) / DEC_TILE;
  RTRY(need(r, r->d_text, (size_t)ntiles * DEC_TILE));
  RTRY(need(r, r->d_tiles, (size_t)(ntiles + 1) * 4));
  RTRY(need(r, r->d_scalars, 64));
  RTRY(need(r, r->h_scalars, 64, true));
  uint8_t *d_text = (uint8_t *)r->d_text.p;
  int64_t *d_sc = (int64_t *)r->d_scalars.p;          /* [0] newlines [1] records [2] bytes side 0 [3] bytes side 1 [4] irregular */
  volatile int64_t *h_sc = (volatile int64_t *)r->h_scalars.p;
  RCU(cudaEventRecord(r->ev0, st));
  RCU(cudaMemcpyAsync(d_text, text + p0, (size_t)n, cudaMemcpyHostToDevice, st));
  RCU(cudaMemsetAsync(d_text + n, 0, (size_t)(ntiles * DEC_TILE - n), st));
  /* newline index */
  SA_LAUNCH(nl_count_kernel, (unsigned)ntiles, DEC_TPB, 0, st, (const uint8_t *)d_text, (int *)r->d_tiles.p);
  SA_LAUNCH(scan3_sums, 1, DEC_TPB, 0, st, (int *)r->d_tiles.p, ntiles, d_sc);
  RCU(cudaGetLastError());
  RCU(cudaMemcpyAsync((void *)h_sc, d_sc, 8, cudaMemcpyDeviceToHost, st));
  RCU(cudaStreamSynchronize(st));
  const int64_t n_nl = h_sc[0];
  const int64_t nlines = n_nl + (tail_line ? 1 : 0);
  RTRY(need(r, r->d_nl, (size_t)(n_nl + 1) * 4));
  RTRY(need(r, r->d_kind, (size_t)nlines));
  Buf *per_line[] = {&r->d_cs, &r->d_ce, &r->d_recflag, &r->d_recex, &r->d_c0, &r->d_a0};
  for(Buf *b : per_line) RTRY(need(r, *b, (size_t)nlines * 4));
  if(split) { RTRY(need(r, r->d_c1, (size_t)nlines * 4)); RTRY(need(r, r->d_a1, (size_t)nlines * 4)); }
  if(n_nl) SA_LAUNCH(nl_fill_kernel, (unsigned)ntiles, DEC_TPB, 0, st, (const uint8_t *)d_text, (const int *)r->d_tiles.p, (int *)r->d_nl.p);

  DecArgs A;
  memset(&A, 0, sizeof(A));
  A.text = d_text; A.n = n; A.nl_pos = (const int *)r->d_nl.p; A.n_nl = n_nl; A.nlines = nlines;
  A.fmt = fmt; A.split = r->split; A.final = final ? 1 : 0;
  /* FASTQ: a record is complete when its four lines are (the last one may lack its newline at the end of the input) */
  A.fq_lines = ((final ? nlines : n_nl) / 4) * 4;
  A.kind = (uint8_t *)r->d_kind.p; A.cs = (int *)r->d_cs.p; A.ce = (int *)r->d_ce.p; A.recflag = (int *)r->d_recflag.p;
  A.rec_excl = (const int *)r->d_recex.p; A.contrib0 = (int *)r->d_c0.p; A.contrib1 = (int *)r->d_c1.p;
  A.at0 = (const int *)r->d_a0.p; A.at1 = (const int *)r->d_a1.p;
  int *d_irr = (int *)(d_sc + 4);
  A.irregular = d_irr;
  const int no_line = 0x7fffffff;
  RCU(cudaMemcpyAsync(d_irr, &no_line, 4, cudaMemcpyHostToDevice, st));
  const unsigned lgrid = (unsigned)((nlines + DEC_TPB - 1) / DEC_TPB);
  SA_LAUNCH(classify_kernel, lgrid, DEC_TPB, 0, st, A);
  RCU(cudaGetLastError());
  RTRY(scan_i32(r, A.recflag, nlines, (int *)r->d_recex.p, d_sc + 1));
  RCU(cudaMemcpyAsync((void *)(h_sc + 1), d_sc + 1, 8, cudaMemcpyDeviceToHost, st));
  RCU(cudaMemcpyAsync((void *)(h_sc + 4), d_sc + 4, 8, cudaMemcpyDeviceToHost, st));
  RCU(cudaStreamSynchronize(st));
  const int64_t R = h_sc[1];
  const int irr = (int)(h_sc[4] & 0xffffffff);
  if(irr != no_line) {
    r->irregular_line = irr;
    return rfail(r, SEQALIGN_ERR_IRREGULAR, "text outside the device decoder's grammar (line " + std::to_string(irr) +
                                               " of the chunk): read this input with the host reader");
  }
  r->total_records = (size_t)R;

  /* sides, offsets, bytes */
  const size_t side_cap = (size_t)(split ? (R + 1) / 2 : R) + 1;
  RTRY(need(r, r->d_out0, (size_t)n + 64));
  if(split) RTRY(need(r, r->d_out1, (size_t)n + 64));
  RTRY(need(r, r->d_off0, side_cap * 8));
  RTRY(need(r, r->d_off1, side_cap * 8));
  Buf *per_rec[] = {&r->d_recpos, &r->d_npos, &r->d_nlen};
  for(Buf *b : per_rec) RTRY(need(r, *b, (size_t)(R + 1) * 4));
  A.out0 = (uint8_t *)r->d_out0.p; A.out1 = (uint8_t *)r->d_out1.p;
  A.off0 = (int64_t *)r->d_off0.p; A.off1 = (int64_t *)r->d_off1.p;
  A.rec_pos = (int *)r->d_recpos.p; A.name_pos = (int *)r->d_npos.p; A.name_len = (int *)r->d_nlen.p;
  SA_LAUNCH(contrib_kernel, lgrid, DEC_TPB, 0, st, A);
  RCU(cudaGetLastError());
  RTRY(scan_i32(r, A.contrib0, nlines, (int *)r->d_a0.p, d_sc + 2));
  if(split) RTRY(scan_i32(r, A.contrib1, nlines, (int *)r->d_a1.p, d_sc + 3));
  else RCU(cudaMemsetAsync(d_sc + 3, 0, 8, st));
  /* lanes per line by the average line length: 8 for reads, a whole warp for long lines */
  const bool long_lines = n / (nlines > 0 ? nlines : 1) > 400;
  const int per_warp = long_lines ? 1 : 4;
  int egrid = (int)((nlines + (int64_t)EMIT_WARPS * per_warp - 1) / ((int64_t)EMIT_WARPS * per_warp));
  if(egrid > r->num_sms * 8) egrid = r->num_sms * 8;
  if(egrid < 1) egrid = 1;
  if(long_lines) SA_LAUNCH(emit_kernel<32>, egrid, DEC_TPB, 0, st, A);
  else SA_LAUNCH(emit_kernel<8>, egrid, DEC_TPB, 0, st, A);
  RCU(cudaGetLastError());
  RCU(cudaEventRecord(r->ev1, st));

  /* the record table and the offsets go back to the host */
  RTRY(need(r, r->h_off0, side_cap * 8, true));
  RTRY(need(r, r->h_off1, side_cap * 8, true));
  RTRY(need(r, r->h_recpos, (size_t)(R + 1) * 4, true));
  RTRY(need(r, r->h_npos, (size_t)(R + 1) * 4, true));
  RTRY(need(r, r->h_nlen, (size_t)(R + 1) * 4, true));
  RCU(cudaMemcpyAsync((void *)(h_sc + 2), d_sc + 2, 16, cudaMemcpyDeviceToHost, st));
  const size_t c0 = (size_t)(split ? (R + 1) / 2 : R), c1 = (size_t)(split ? R / 2 : 0);
  if(c0) RCU(cudaMemcpyAsync(r->h_off0.p, r->d_off0.p, c0 * 8, cudaMemcpyDeviceToHost, st));
  if(c1) RCU(cudaMemcpyAsync(r->h_off1.p, r->d_off1.p, c1 * 8, cudaMemcpyDeviceToHost, st));
  if(R) {
    RCU(cudaMemcpyAsync(r->h_recpos.p, r->d_recpos.p, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
    RCU(cudaMemcpyAsync(r->h_npos.p, r->d_npos.p, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
    RCU(cudaMemcpyAsync(r->h_nlen.p, r->d_nlen.p, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
  }
  RCU(cudaStreamSynchronize(st));
  float ms = 0;
  if(cudaEventElapsedTime(&ms, r->ev0, r->ev1) == cudaSuccess) r->last_ms = ms; else cudaGetLastError();
  r->side_bytes[0] = h_sc[2]; r->side_bytes[1] = h_sc[3];
  /* closing offsets, on both sides of the bus */
  ((int64_t *)r->h_off0.p)[c0] = r->side_bytes[0];
  ((int64_t *)r->h_off1.p)[c1] = r->side_bytes[1];
  RCU(cudaMemcpyAsync((int64_t *)r->d_off0.p + c0, (int64_t *)r->h_off0.p + c0, 8, cudaMemcpyHostToDevice, st));
  RCU(cudaMemcpyAsync((int64_t *)r->d_off1.p + c1, (int64_t *)r->h_off1.p + c1, 8, cudaMemcpyHostToDevice, st));
  RCU(cudaStreamSynchronize(st));

  /* which records are complete?  More text may still follow (final == 0):
   *   FASTA  the last record can grow by further sequence lines;
   *   plain  a last line without its newline can grow;
   *   FASTQ  complete = four newline-terminated lines.
   * At the end of the input every record counts, except a FASTA header that is a bare '>' as the very
   * last byte: the reference's name readline returns 0 there and the read fails (seq_file.h:280). */
  size_t complete = (size_t)R;
  const int *h_recpos = (const int *)r->h_recpos.p, *h_npos = (const int *)r->h_npos.p, *h_nlen = (const int *)r->h_nlen.p;
  if(fmt == SEQALIGN_FMT_FASTQ) complete = (size_t)(A.fq_lines / 4);
  else if(!final) {
    if(fmt == SEQALIGN_FMT_FASTA) complete = R ? (size_t)R - 1 : 0;
    else if(tail_line && R) {
      /* is the last record the unterminated line?  then its start lies behind the last newline */
      const char *last_nl = (const char *)memrchr(text + p0, '\n', bytes - p0);
      const size_t tail_start = last_nl ? (size_t)(last_nl - text) + 1 : p0;
      if((size_t)h_recpos[R - 1] + p0 >= tail_start) complete = (size_t)R - 1;
    }
  } else if(fmt == SEQALIGN_FMT_FASTA && R && text[bytes - 1] == '>' && h_nlen[R - 1] == 0 && (size_t)h_npos[R - 1] + p0 == bytes) {
    complete = (size_t)R - 1;
  }
  r->records = complete;
  /* where the held-back tail starts: at the first record that is not complete, or (FASTQ, whose
   * incomplete lines are not in the record table) at the first line behind the complete records */
  r->tail_start = bytes;
  if(complete < (size_t)R) r->tail_start = (size_t)h_recpos[complete] + p0;
  else if(fmt == SEQALIGN_FMT_FASTQ && !final && A.fq_lines < nlines) {
    int nl_before = -1;
    if(A.fq_lines > 0) {
      RCU(cudaMemcpyAsync((void *)(h_sc + 5), (const int *)r->d_nl.p + (A.fq_lines - 1), 4, cudaMemcpyDeviceToHost, st));
      RCU(cudaStreamSynchronize(st));
      nl_before = (int)(h_sc[5] & 0xffffffff);
    }
    r->tail_start = (size_t)(nl_before + 1) + p0;
  }
  return SEQALIGN_OK;
}

int seqalign_reads_format(const seqalign_reads_t *r) { return r ? r->fmt : 0; }
size_t seqalign_reads_records(const seqalign_reads_t *r) { return r ? r->records : 0; }
double seqalign_reads_last_ms(const seqalign_reads_t *r) { return r ? r->last_ms : 0; }

size_t seqalign_reads_count(const seqalign_reads_t *r, int side)
{
  if(!r || side < 0 || side > 1) return 0;
  if(!r->split) return side == 0 ? r->records : 0;
  return side == 0 ? (r->records + 1) / 2 : r->records / 2;
}

const int64_t *seqalign_reads_offsets(const seqalign_reads_t *r, int side)
{
  if(!r || side < 0 || side > 1 || r->total_records == 0) return nullptr;
  return (const int64_t *)(side ? r->h_off1.p : r->h_off0.p);
}

size_t seqalign_reads_record_start(const seqalign_reads_t *r, size_t i)
{
  if(!r) return 0;
  if(i >= r->records) return r->tail_start;
  return (size_t)((const int *)r->h_recpos.p)[i] + r->p0;
}

int seqalign_reads_name(const seqalign_reads_t *r, size_t i, size_t *pos, size_t *len)
{
  if(!r || i >= r->total_records || !pos || !len) return SEQALIGN_ERR_ARG;
  *pos = (size_t)((const int *)r->h_npos.p)[i] + r->p0;
  *len = (size_t)((const int *)r->h_nlen.p)[i];
  return SEQALIGN_OK;
}

const void *seqalign_reads_device_seq(const seqalign_reads_t *r, int side)
{
  if(!r || side < 0 || side > 1) return nullptr;
  return side ? r->d_out1.p : r->d_out0.p;
}

const void *seqalign_reads_device_offsets(const seqalign_reads_t *r, int side)
{
  if(!r || side < 0 || side > 1) return nullptr;
  return side ? r->d_off1.p : r->d_off0.p;
}

int seqalign_reads_device(const seqalign_reads_t *r) { return r ? r->device : -1; }

int seqalign_reads_fetch(seqalign_reads_t *r, int side, char *out)
{
  if(!r || side < 0 || side > 1 || !out) return SEQALIGN_ERR_ARG;
  const int64_t nbytes = r->side_bytes[side];
  if(nbytes == 0) return SEQALIGN_OK;
  RCU(cudaSetDevice(r->device));
  RCU(cudaMemcpyAsync(out, side ? r->d_out1.p : r->d_out0.p, (size_t)nbytes, cudaMemcpyDeviceToHost, r->stream));
  RCU(cudaStreamSynchronize(r->stream));
  return SEQALIGN_OK;
}

} // extern "C"
