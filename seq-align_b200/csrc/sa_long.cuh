/*
 * sa_long.cuh -- sm_100a kernels of the batch alignment engine (part 4):
 * the specialised fill for WIDE pairs (len_a beyond one warp-wide strip, up
 * to any length) and for Needleman-Wunsch with free end gaps.  This is the
 * kernel behind BASELINE config 3 (NW 10k x 10k, --freestartgap
 * --freeendgap, score + traceback).
 *
 * Same recurrence as alignment_fill_matrices (reference
 * src/alignment.c:89-167) in the H' = max(M,GA,GB)+open form of sa_fast.cuh
 * (one VIADDMNMX per matrix, H'+open on the FMA pipe, query profile in
 * shared memory).  What is new here is the work shape:
 *
 *   * a pair is cut into column strips of 32*K columns; a strip is swept top
 *     to bottom by one warp as an anti-diagonal wavefront (lane = K columns);
 *   * the LONG_WARPS warps of a CTA take the strips of the CTA's pairs in one
 *     continuous round-robin sequence (strip g of that sequence belongs to
 *     warp g % LONG_WARPS) -- a systolic pipeline that does not drain between
 *     pairs: while the last strips of one pair finish, the first strips of
 *     the next pair are already running;
 *   * the right edge of a strip (H', GB per row, 8 bytes) goes through a
 *     per-CTA boundary buffer in global memory (L2 resident), the consumer
 *     follows the producer's progress word in shared memory 32 rows at a
 *     time;
 *   * seq_b is streamed 32 rows at a time into a 64-entry ring per warp, so
 *     no sequence has to fit shared memory.
 *
 * Free end gaps (scoring->no_end_gap_penalty, alignment.c:122-155): in the
 * last column gap_a costs nothing, in the last row gap_b costs nothing.
 * Only the strip holding column len_a and the row len_b of every strip run
 * the "generic" row body that selects the penalties per column / per row;
 * everything else runs the lean body.
 *
 * Traceback comes in two shapes.  DIR writes one flag byte per cell (below).
 * CKPT writes almost nothing: the right edge of every strip (H', GB per row,
 * which the next strip needs anyway) is kept per pair instead of in a ring,
 * and every LONG_CK_ROWS rows each lane parks its row state (H', GA of its K
 * columns).  That is 8 bytes per 512 cells plus 8 bytes per 64 cells, 0.14
 * bytes per cell instead of 1, and the fill runs at the speed of the score
 * kernel.  walk_ckpt_kernel then recomputes, with this very row body, only the
 * tiles (one strip wide, LONG_CK_ROWS rows high) that the traceback path
 * crosses -- about 3 % of the cells of a 10k x 10k pair -- and takes the
 * reference's decisions (alignment.c:311-327) from the recomputed flags.
 *
 * Traceback flags are the five equality bits of sa_fast.cuh (same byte
 * layout, same walk): they are computed with the penalties that applied to
 * the cell, so alignment_reverse_move's zeroed penalties in the last column /
 * row (alignment.c:265-268) need nothing extra in the walk.
 */
#ifndef SA_LONG_CUH
#define SA_LONG_CUH

#include <vector>
#include "sa_platform.h"
#include "sa_flatten.h"
#include "sa_kernels.cuh"
#include "sa_fast.cuh"

namespace sa {

constexpr int LONG_WARPS = 8;
constexpr int LONG_NEG = -(1 << 29);   /* "minus infinity" of the NW borders: far from INT_MIN, below every real value */
constexpr int LONG_K = 16;             /* columns per lane */
constexpr int LONG_STRIP = 32 * LONG_K;
#ifndef SA_CK_ROWS
#define SA_CK_ROWS 64
#endif
constexpr int LONG_CK_ROWS = SA_CK_ROWS;   /* CKPT: a row checkpoint every this many rows (a power of two) */

/* CKPT: bytes of one pair's trace region, [strip edges | row checkpoints]:
 *   edge[s][y]  int2 (H', GB) of cell (x = (s+1)*STRIP, y), s < nstrips-1, y in [0, lb]
 *   ck[r][x-1]  int2 (H', GA) of cell (x, y = (r+1)*CK_ROWS), x in [1, roundup16(la)], r < lb/CK_ROWS */
__host__ __device__ __forceinline__ int64_t long_edge_ints2(int la, int lb)
{
  const int64_t nstrips = ((int64_t)la + LONG_STRIP - 1) / LONG_STRIP;
  /* even count: the checkpoints behind the edges are read and written 16 bytes at a time */
  return nstrips > 1 ? ((nstrips - 1) * ((int64_t)lb + 1) + 1) & ~(int64_t)1 : 0;
}
__host__ __device__ __forceinline__ int64_t long_ck_width(int la) { return ((int64_t)la + 15) & ~(int64_t)15; }
__host__ __device__ __forceinline__ int64_t long_trace_bytes(int la, int lb)
{
  const int64_t ints2 = long_edge_ints2(la, lb) + (int64_t)(lb / LONG_CK_ROWS) * long_ck_width(la);
  return (ints2 * 8 + 15) & ~(int64_t)15;
}

struct LongArgs {
  const uint8_t *seq_a, *seq_b;
  const int64_t *off_a, *off_b;   /* of the launch's first pair */
  int64_t npairs;
  ScoreParams sp;
  const int8_t *tab8;             /* [cb*(n+1) + ca] = sub - open, last column = padding code */
  const int32_t *tab32;
  const uint8_t *lut;
  int2 *bnd;                      /* [grid][2][LONG_WARPS][bnd_rows] strip edges */
  int64_t bnd_rows;
  uint8_t *dir;                   /* DIR: traceback flag bytes, row-major per pair; CKPT: the pairs' trace regions */
  const int64_t *dir_off;
  int32_t *score, *xend, *yend;   /* per pair of the launch */
  unsigned long long *swkey;      /* SW: per pair, best cell so far as (score << 40 | 0xFFFFF - x << 20 | 0xFFFFF - y) */
  int mul_one;
};

struct LongPlan {
  int K = 0;
  bool is_sw = false, prof32 = false, dir = false, noend = false;
  bool ckpt = false;   /* traceback through checkpoints + recomputed tiles instead of flag bytes */
  const char *name = "";
  std::vector<int8_t> tab8;
  std::vector<int32_t> tab32;
  size_t smem = 0;
};

/* one row of a lane's K columns.  GEN: penalties selected per column (bit j
 * of lastmask = this is column len_a) and per row (lastrow), for free end
 * gaps; otherwise the lean body. */
template <int K, bool IS_SW, bool PROF32, bool DIR, bool GEN>
__device__ __forceinline__ void long_row(int (&hp)[K], int (&ga)[K], int &hl, int &gb, int d,
                                         const unsigned *w, unsigned *dw,
                                         const int open, const int ext, const int mul_one,
                                         const unsigned lastmask, const bool lastrow, int *rowkey = nullptr)
{
  /* SW score kernels: the row's best match score with its column, key = M x 16 + (15 - j): a larger
   * key is a higher score or, at equal score, a smaller column (hit order: score desc, x asc) */
  int kbest = 0, kprev = 0;
#pragma unroll
  for(int j = 0; j < K; j++) {
    const int sub = PROF32 ? (int)w[j] : sext_byte_dyn(w[j / 4], j & 3);
    int eA = ext, eB = ext, hup = hp[j], hlf = hl;
    if(GEN) {
      if((lastmask >> j) & 1) { eA = 0; hup = hp[j] - open; }   /* gap_a is free in column len_a */
      if(lastrow) { eB = 0; hlf = hl - open; }                   /* gap_b is free in row len_b */
    }
    int uge = 0, lge = 0;
    if(DIR) {
      uge = ga[j] * mul_one + eA;   /* GA_up + ext, GB_left + ext (IMAD, FMA pipe) */
      lge = gb * mul_one + eB;
    }
    int m, h;
    if(IS_SW) {
      m = addmax(d, sub, 0);
      ga[j] = addmax_relu(ga[j], eA, hup);
      gb = addmax_relu(gb, eB, hlf);
    } else {
      /* no clamp at "min": every operand is a real value or LONG_NEG + one
       * penalty, which never wins a max (see long_plan) */
      m = d * mul_one + sub;
      ga[j] = addmax(ga[j], eA, hup);
      gb = addmax(gb, eB, hlf);
    }
    h = max3(m, ga[j], gb);
    if(IS_SW && !DIR) {
      const int kk = m * (mul_one * 16) + (15 - j);   /* IMAD: off the ALU pipe */
      if(j & 1) kbest = max3(kbest, kprev, kk);
      else if(j == K - 1) kbest = imax(kbest, kk);
      kprev = kk;
    }
    if(DIR) {
      /* five "not equal" bits (sa_fast.cuh), a >= b holds for every pair */
      const int f = imin(h - ga[j], 1) + 2 * imin(h - gb, 1) + 4 * imin(ga[j] - uge, 1) +
                    8 * imin(gb - lge, 1) + 16 * imin(gb - hlf, 1);
      dw[j / 4] += (unsigned)f << (8 * (j & 3));
    }
    d = hp[j];
    hl = h * mul_one + open;
    hp[j] = hl;
  }
  if(IS_SW && !DIR && rowkey) *rowkey = kbest;
}

template <int K, bool IS_SW, bool PROF32, bool DIR, bool NOEND, bool CKPT = false>
__global__ void __launch_bounds__(LONG_WARPS * 32, 2)
long_kernel(const LongArgs A)
{
  static_assert(!(DIR && CKPT), "flag bytes or checkpoints, not both");
  static_assert(!CKPT || K == LONG_K, "the walk recomputes tiles with LONG_K columns per lane");
  constexpr int W = LONG_WARPS;
  constexpr int STRIP = 32 * K;
  constexpr int KW = PROF32 ? K : (K + 3) / 4;
  constexpr int KS = PROF32 ? prof32_stride(K) : KW;
  constexpr int PSTRIDE = 32 * KS * 4;
  static_assert(K % 4 == 0, "flag bytes are stored as whole words");

  unsigned char *dsm = SA_DYN_SMEM();
  const ScoreParams &sp = A.sp;
  const int n = sp.ncodes, tw = n + 1;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;

  /* shared: [lut 256][table][per warp: profile | b ring 64 | edge chunk 32 x int2] */
  uint8_t *s_lut = dsm;
  int8_t *s_tab8 = (int8_t *)(dsm + 256);
  int32_t *s_tab32 = (int32_t *)(dsm + 256);
  const int tab_bytes = ((PROF32 ? 4 : 1) * n * tw + 15) & ~15;
  const int warp_bytes = n * PSTRIDE + 64 + 32 * 8;
  unsigned char *wbase = dsm + 256 + tab_bytes + wib * warp_bytes;
  unsigned char *s_prof = wbase;
  uint8_t *s_bring = wbase + n * PSTRIDE;
  int2 *s_chunk = (int2 *)(s_bring + 64);
  __shared__ volatile unsigned long long s_progress[W];   /* producer warp -> (strip << 32 | rows published) */
  __shared__ volatile unsigned s_fin[W];                   /* warp -> strips finished (index of the last one + 1) */

  for(int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = A.lut[i];
  if(PROF32) { for(int i = threadIdx.x; i < n * tw; i += blockDim.x) s_tab32[i] = A.tab32[i]; }
  else       { for(int i = threadIdx.x; i < n * tw; i += blockDim.x) s_tab8[i] = A.tab8[i]; }
  if(threadIdx.x < W) { s_progress[threadIdx.x] = 0; s_fin[threadIdx.x] = 0; }
  __syncthreads();

  const int open = sp.open, ext = sp.ext;
  const int mul_one = A.mul_one;
  int2 *cta_bnd = A.bnd + (int64_t)blockIdx.x * 2 * W * A.bnd_rows;
  const unsigned *prow = (const unsigned *)s_prof + lane * KS;

  unsigned gs_base = 0;   /* position of this pair's strip 0 in the CTA's strip sequence */
  for(int64_t p = blockIdx.x; p < A.npairs; p += gridDim.x) {
    const int64_t oa = A.off_a[p], ob = A.off_b[p];
    const int la = (int)(A.off_a[p + 1] - oa), lb = (int)(A.off_b[p + 1] - ob);
    const uint8_t *pa = A.seq_a + oa, *pb = A.seq_b + ob;
    const int nstrips = (la > 0 && lb > 0) ? (la + STRIP - 1) / STRIP : 0;
    if(nstrips == 0) {
      if(wib == 0 && lane == 0) {
        int sc = 0;
        if(!IS_SW && (la > 0 || lb > 0)) sc = sp.no_start ? 0 : addw(sp.gap_open, (la > 0 ? la : lb) * ext);
        A.score[p] = sc;
        if(A.xend) A.xend[p] = IS_SW ? 0 : la;
        if(A.yend) A.yend[p] = IS_SW ? 0 : lb;
      }
      continue;
    }
    uint8_t *dirp = nullptr;
    int dstride = 0;
    if(DIR) { dirp = A.dir + A.dir_off[p]; dstride = (int)dir_stride(la); }
    int2 *pair_edge = nullptr, *pair_ck = nullptr;
    int ckw = 0;
    if(CKPT) {
      pair_edge = (int2 *)(A.dir + A.dir_off[p]);
      pair_ck = pair_edge + long_edge_ints2(la, lb);
      ckw = (int)long_ck_width(la);
    }

    for(int s = (int)((wib + W - gs_base % W) % W); s < nstrips; s += W) {
      const unsigned gs = gs_base + (unsigned)s;
      const int x0 = s * STRIP;
      const int xf = x0 + lane * K + 1;   /* my first column, 1-based */
      const bool more = s + 1 < nstrips;
      int2 *out_bnd = cta_bnd + (int64_t)(((gs / W) & 1) * W + wib) * A.bnd_rows;
      const int2 *in_bnd = cta_bnd + (int64_t)((((gs - 1) / W) & 1) * W + (wib + W - 1) % W) * A.bnd_rows;
      if(CKPT) {
        /* the edges are kept per pair (the walk reads them later): no ring, no slot to wait for */
        out_bnd = pair_edge + (int64_t)s * (lb + 1);
        in_bnd = pair_edge + (int64_t)(s - 1) * (lb + 1);
      }
      const unsigned long long in_base = (unsigned long long)(gs - 1) << 32;
      /* my edge slot was last used two rounds ago; its reader must be done */
      if(!CKPT && more && gs >= 2 * W)
        while(s_fin[(wib + 1) % W] < gs - 2 * W + 2) SA_SPIN_HINT();

      /* query profile of my K columns: row c holds sub'(a[x], c) */
      unsigned lastmask = 0;
      {
        int acode[K];
#pragma unroll
        for(int j = 0; j < K; j++) {
          acode[j] = (xf + j <= la) ? s_lut[pa[xf + j - 1]] : n;   /* n = padding code */
          if(NOEND && xf + j == la) lastmask |= 1u << j;
        }
        for(int c = 0; c < n; c++) {
          unsigned *dst = (unsigned *)(s_prof + c * PSTRIDE) + lane * KS;
          if(PROF32) {
            const int32_t *trow = s_tab32 + c * tw;
#pragma unroll
            for(int j = 0; j < K; j++) dst[j] = (unsigned)trow[acode[j]];
          } else {
            const int8_t *trow = s_tab8 + c * tw;
#pragma unroll
            for(int q = 0; q < KW; q++) {
              unsigned word = 0;
#pragma unroll
              for(int b4 = 0; b4 < 4; b4++) {
                const int j = 4 * q + b4;
                if(j < K) word |= (unsigned)(uint8_t)trow[acode[j]] << (8 * b4);
              }
              dst[q] = word;
            }
          }
        }
      }

      /* row 0 (alignment.c:47-69) in H' form */
      int hp[K], ga[K];
#pragma unroll
      for(int j = 0; j < K; j++) {
        if(IS_SW) { hp[j] = open; ga[j] = 0; }
        else {
          hp[j] = (sp.no_start ? 0 : sp.gap_open + (xf + j) * ext) + open;
          ga[j] = LONG_NEG;
        }
      }
      int hd;   /* H'(xf-1, y-1) */
      if(IS_SW || xf == 1) hd = open;
      else hd = (sp.no_start ? 0 : sp.gap_open + (xf - 1) * ext) + open;

      int out_h = 0, out_gb = 0;
      int swbest = 0, swy = 0;   /* SW: this lane's best key in this strip and its row */
      const int nsteps = lb + 31;
      for(int st = 0; st < nsteps; st++) {
        const int y = st - lane + 1;
        const bool active = y >= 1 && y <= lb;

        if((st & 31) == 0) {
          /* next 32 rows: seq_b codes into the ring, the left strip's edge into the chunk */
          __syncwarp();
          const int row = st + 1 + lane;
          if(row <= lb) s_bring[(row - 1) & 63] = s_lut[pb[row - 1]];
          if(x0 > 0) {
            const unsigned long long need = in_base + (unsigned long long)(st + 32 < lb ? st + 32 : lb);
            while(s_progress[(wib + W - 1) % W] < need) SA_SPIN_HINT();
            __threadfence_block();
            int2 v = make_int2(0, 0);
            if(row <= lb) {
#if defined(__CUDA_ARCH__)
              v = __ldcg(&in_bnd[row]);   /* L2: written by another warp of this CTA */
#else
              v = in_bnd[row];
#endif
            }
            s_chunk[lane] = v;
          }
          __syncwarp();
        }

        /* left neighbour (xf-1, y) */
        int hl = __shfl_up_sync(FULL, out_h, 1);
        int gb = __shfl_up_sync(FULL, out_gb, 1);
        if(lane == 0) {
          if(x0 == 0) {
            /* column 0 (alignment.c:55-56, 72-80) */
            if(IS_SW) { hl = open; gb = 0; }
            else { hl = (sp.no_start ? 0 : sp.gap_open + y * ext) + open; gb = LONG_NEG; }
          } else {
            const int2 v = s_chunk[st & 31];
            hl = v.x; gb = v.y;
          }
        }
        const int hl_in = hl;

        if(active) {
          const int c = s_bring[(y - 1) & 63];
          const unsigned *pw = prow + c * (PSTRIDE / 4);
          unsigned w[KW];
          if(KW % 4 == 0 && KS % 4 == 0) {
#pragma unroll
            for(int q = 0; q < KW / 4; q++) {
              const uint4 v = ((const uint4 *)pw)[q];
              w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
            }
          } else {
#pragma unroll
            for(int q = 0; q < KW; q++) w[q] = pw[q];
          }
          unsigned dw[DIR ? K / 4 : 1];
          if(DIR) {
#pragma unroll
            for(int q = 0; q < K / 4; q++) dw[q] = 0;
          }
          int rowkey = 0;
          if(NOEND && (!more || y == lb))
            long_row<K, IS_SW, PROF32, DIR, true>(hp, ga, hl, gb, hd, w, dw, open, ext, mul_one, lastmask, y == lb);
          else
            long_row<K, IS_SW, PROF32, DIR, false>(hp, ga, hl, gb, hd, w, dw, open, ext, mul_one, 0u, false, &rowkey);
          /* a later row replaces the lane's best only with a larger key: at equal (score, x) the smaller y stays */
          if(IS_SW && !DIR && rowkey > swbest) { swbest = rowkey; swy = y; }
          if(DIR) {
            /* the row stride is a multiple of 16: a lane's 16 bytes are inside or outside as a whole */
            unsigned *drow = (unsigned *)(dirp + (int64_t)(y - 1) * dstride + (xf - 1));
            if(K == 16) {
              if(xf - 1 < dstride) *(uint4 *)drow = make_uint4(dw[0], dw[1], dw[2], dw[3]);
            } else {
#pragma unroll
              for(int q = 0; q < K / 4; q++)
                if(xf - 1 + 4 * q < dstride) drow[q] = dw[q];
            }
          }
          if(CKPT && (y & (LONG_CK_ROWS - 1)) == 0 && xf - 1 < ckw) {
            /* row checkpoint: this lane's (H', GA) of row y, 128 contiguous bytes */
            uint4 *ck = (uint4 *)(pair_ck + (int64_t)(y / LONG_CK_ROWS - 1) * ckw + (xf - 1));
#pragma unroll
            for(int q = 0; q < K / 2; q++)
              ck[q] = make_uint4((unsigned)hp[2 * q], (unsigned)ga[2 * q], (unsigned)hp[2 * q + 1], (unsigned)ga[2 * q + 1]);
          }
          if(more && lane == 31) {
            out_bnd[y] = make_int2(hl, gb);
            if((y & 31) == 0 || y == lb) {
              __threadfence_block();
              s_progress[wib] = ((unsigned long long)gs << 32) | (unsigned long long)y;
            }
          }
          out_h = hl;
          out_gb = gb;
          hd = hl_in;   /* next row's diagonal */
        }
      }

      if(!more) {
        /* the final cell (len_a, len_b) is in this strip */
        const int jf = (la - 1 - x0) % K, lf = (la - 1 - x0) / K;
        if(lane == lf) {
          int v = 0;
#pragma unroll
          for(int j = 0; j < K; j++) if(j == jf) v = hp[j];
          if(!IS_SW) {
            A.score[p] = v - open;
            if(A.xend) A.xend[p] = la;
            if(A.yend) A.yend[p] = lb;
          }
        }
      }
      if(IS_SW && !DIR) {
        /* the strip's best cell under the hit order (score desc, x asc, y asc: smith_waterman.c:71-86),
         * merged into the pair's by a 64-bit atomic max */
        int sv = swbest >> 4, sx = xf + 15 - (swbest & 15), sy = swy;
        if(sv <= 0) { sv = 0; sx = 0; sy = 0; }
#pragma unroll
        for(int o = 16; o > 0; o >>= 1) {
          const int v2 = __shfl_xor_sync(FULL, sv, o), x2 = __shfl_xor_sync(FULL, sx, o), y2 = __shfl_xor_sync(FULL, sy, o);
          if(hit_better(v2, x2, y2, sv, sx, sy)) { sv = v2; sx = x2; sy = y2; }
        }
        if(lane == 0 && sv > 0)
          atomicMax(A.swkey + p, ((unsigned long long)sv << 40) | ((unsigned long long)(0xFFFFF - sx) << 20) |
                                     (unsigned long long)(0xFFFFF - sy));
      }
      __syncwarp();
      if(lane == 0) s_fin[wib] = gs + 1;
    }
    gs_base += (unsigned)nstrips;
  }
}

/* ---------------------------------------------------------------------------
 * walk_ckpt_kernel: traceback over the checkpoints of a CKPT fill.  One warp
 * per pair.  The walk itself is walk_pair() of sa_kernels.cuh (the loops of
 * needleman_wunsch.c:79-132 / smith_waterman.c:187-255 over the five equality
 * flags); what differs is where the flags come from: when the walk asks for a
 * cell outside the window in shared memory, the warp recomputes the tile that
 * holds it -- one strip wide, up to LONG_CK_ROWS rows high, cut off at the
 * asking cell (the walk only moves up and left) -- from the row checkpoint
 * above it and the strip edge to its left, with the fill's own row body
 * (long_row<..., DIR>), so the flags are the ones the fill would have written.
 */
constexpr int WK_WARPS = 2;

inline size_t walk_ckpt_smem_bytes(int ncodes, bool prof32)
{
  const int KS = prof32 ? prof32_stride(LONG_K) : LONG_K / 4;
  const size_t tab = (((size_t)(prof32 ? 4 : 1) * ncodes * (ncodes + 1)) + 15) & ~(size_t)15;
  const size_t warp_bytes = (size_t)ncodes * 32 * KS * 4 + 64 + 32 * 8 + (size_t)LONG_CK_ROWS * LONG_STRIP;
  return 256 + tab + WK_WARPS * warp_bytes;
}

/* what a tile recompute needs to know about the pair and the warp's shared memory */
struct WalkCkCtx {
  const uint8_t *pa, *pb;
  int la, lb, nstrips, ckw;
  const int2 *pair_edge, *pair_ck;
  const uint8_t *s_lut;
  const void *s_tab;
  unsigned char *s_prof;
  uint8_t *s_bring;
  int2 *s_chunk;
  uint8_t *s_flags;
  int n, open, ext, gap_open, no_start, mul_one;
};

/* recompute the flags of the tile that holds cell (cx, cy) (0-based), cut off at that cell */
template <bool IS_SW, bool NOEND, bool PROF32>
__device__ __noinline__ void walk_ckpt_recompute(WalkCkCtx &C, const int cx, const int cy)
{
  constexpr int K = LONG_K, STRIP = LONG_STRIP;
  constexpr int KW = PROF32 ? K : K / 4;
  constexpr int KS = PROF32 ? prof32_stride(K) : KW;
  constexpr int PSTRIDE = 32 * KS * 4;
  const int lane = threadIdx.x & 31;
  const int n = C.n, tw = n + 1, open = C.open, ext = C.ext, la = C.la, lb = C.lb;
  const unsigned *prow = (const unsigned *)C.s_prof + lane * KS;
  /* H' of the border cell (x, 0) / (0, y) (alignment.c:47-81) */
  auto border_h = [&](int k) -> int { return IS_SW ? open : (C.no_start ? 0 : C.gap_open + k * ext) + open; };

  __syncwarp();
  const int s = cx / STRIP, x0c = s * STRIP;
  const int y0c = (cy / LONG_CK_ROWS) * LONG_CK_ROWS;
  const int nrows = cy - y0c + 1, nl = (cx - x0c) / K + 1;
  const bool more = s + 1 < C.nstrips;
  const int xf = x0c + lane * K + 1;
  const int2 *in_bnd = C.pair_edge + (int64_t)(s - 1) * (lb + 1);
  const int2 *ck = C.pair_ck + (int64_t)(y0c / LONG_CK_ROWS - 1) * C.ckw;

  unsigned lastmask = 0;
  {
    int acode[K];
#pragma unroll
    for(int j = 0; j < K; j++) {
      acode[j] = (xf + j <= la) ? C.s_lut[C.pa[xf + j - 1]] : n;
      if(NOEND && xf + j == la) lastmask |= 1u << j;
    }
    for(int c = 0; c < n; c++) {
      unsigned *dst = (unsigned *)(C.s_prof + c * PSTRIDE) + lane * KS;
      if(PROF32) {
        const int32_t *trow = (const int32_t *)C.s_tab + c * tw;
#pragma unroll
        for(int j = 0; j < K; j++) dst[j] = (unsigned)trow[acode[j]];
      } else {
        const int8_t *trow = (const int8_t *)C.s_tab + c * tw;
#pragma unroll
        for(int q = 0; q < KW; q++) {
          unsigned word = 0;
#pragma unroll
          for(int b4 = 0; b4 < 4; b4++) word |= (unsigned)(uint8_t)trow[acode[4 * q + b4]] << (8 * b4);
          dst[q] = word;
        }
      }
    }
  }

  /* the row above the tile: border row 0 or the checkpoint of row y0c */
  int hp[K], ga[K];
  int hd;
  if(y0c == 0) {
#pragma unroll
    for(int j = 0; j < K; j++) { hp[j] = border_h(xf + j); ga[j] = IS_SW ? 0 : LONG_NEG; }
    hd = (IS_SW || xf == 1) ? open : border_h(xf - 1);
  } else {
    const bool in_ck = xf - 1 < C.ckw && lane < nl;
#pragma unroll
    for(int q = 0; q < K / 2; q++) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if(in_ck) v = ((const uint4 *)(ck + (xf - 1)))[q];
      hp[2 * q] = (int)v.x; ga[2 * q] = (int)v.y; hp[2 * q + 1] = (int)v.z; ga[2 * q + 1] = (int)v.w;
    }
    if(lane > 0) hd = in_ck ? ck[xf - 2].x : 0;
    else if(x0c == 0) hd = border_h(y0c);
    else hd = in_bnd[y0c].x;
  }

  int out_h = 0, out_gb = 0;
  const int nsteps = nrows + nl - 1;
  for(int st = 0; st < nsteps; st++) {
    const int yl = st - lane + 1, y = y0c + yl;
    const bool active = yl >= 1 && yl <= nrows && lane < nl;
    if((st & 31) == 0) {
      __syncwarp();
      const int row = y0c + st + 1 + lane;
      if(row <= y0c + nrows) {
        C.s_bring[(row - 1) & 63] = C.s_lut[C.pb[row - 1]];
        if(x0c > 0) C.s_chunk[lane] = in_bnd[row];
      }
      __syncwarp();
    }
    int hl = __shfl_up_sync(FULL, out_h, 1);
    int gb = __shfl_up_sync(FULL, out_gb, 1);
    if(lane == 0) {
      if(x0c == 0) {
        if(IS_SW) { hl = open; gb = 0; }
        else { hl = border_h(y); gb = LONG_NEG; }
      } else {
        const int2 v = C.s_chunk[st & 31];
        hl = v.x; gb = v.y;
      }
    }
    const int hl_in = hl;
    if(active) {
      const int c = C.s_bring[(y - 1) & 63];
      const unsigned *pw = prow + c * (PSTRIDE / 4);
      unsigned w[KW];
#pragma unroll
      for(int q = 0; q < KW / 4; q++) {
        const uint4 v = ((const uint4 *)pw)[q];
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
      unsigned dw[K / 4];
#pragma unroll
      for(int q = 0; q < K / 4; q++) dw[q] = 0;
      if(NOEND && (!more || y == lb))
        long_row<K, IS_SW, PROF32, true, true>(hp, ga, hl, gb, hd, w, dw, open, ext, C.mul_one, lastmask, y == lb);
      else
        long_row<K, IS_SW, PROF32, true, false>(hp, ga, hl, gb, hd, w, dw, open, ext, C.mul_one, 0u, false);
      *(uint4 *)(C.s_flags + (yl - 1) * STRIP + lane * K) = make_uint4(dw[0], dw[1], dw[2], dw[3]);
      out_h = hl;
      out_gb = gb;
      hd = hl_in;
    }
  }
  __syncwarp();
}

template <bool IS_SW, bool NOEND, bool PROF32>
__global__ void __launch_bounds__(WK_WARPS * 32)
walk_ckpt_kernel(const WalkArgs A, const int8_t *__restrict__ tab8, const int32_t *__restrict__ tab32, const int mul_one)
{
  constexpr int K = LONG_K, STRIP = LONG_STRIP;
  constexpr int KS = PROF32 ? prof32_stride(K) : K / 4;
  constexpr int PSTRIDE = 32 * KS * 4;
  unsigned char *dsm = SA_DYN_SMEM();
  const ScoreParams &sp = A.sp;
  const int n = sp.ncodes, tw = n + 1;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;

  uint8_t *s_lut = dsm;
  int8_t *s_tab8 = (int8_t *)(dsm + 256);
  int32_t *s_tab32 = (int32_t *)(dsm + 256);
  const int tab_bytes = ((PROF32 ? 4 : 1) * n * tw + 15) & ~15;
  const int warp_bytes = n * PSTRIDE + 64 + 32 * 8 + LONG_CK_ROWS * STRIP;
  unsigned char *wbase = dsm + 256 + tab_bytes + wib * warp_bytes;

  for(int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = A.lut[i];
  if(PROF32) { for(int i = threadIdx.x; i < n * tw; i += blockDim.x) s_tab32[i] = tab32[i]; }
  else       { for(int i = threadIdx.x; i < n * tw; i += blockDim.x) s_tab8[i] = tab8[i]; }
  __syncthreads();

  WalkCkCtx C;
  C.s_lut = s_lut; C.s_tab = dsm + 256;
  C.s_prof = wbase;
  C.s_bring = wbase + n * PSTRIDE;
  C.s_chunk = (int2 *)(C.s_bring + 64);
  C.s_flags = (uint8_t *)(C.s_chunk + 32);
  C.n = n; C.open = sp.open; C.ext = sp.ext; C.gap_open = sp.gap_open; C.no_start = sp.no_start; C.mul_one = mul_one;

  for(int64_t r = (int64_t)blockIdx.x * WK_WARPS + wib; r < A.npairs; r += (int64_t)gridDim.x * WK_WARPS) {
    const int64_t p = A.pair0 + r;
    const int64_t oa = A.off_a[p], ob = A.off_b[p];
    C.la = (int)(A.off_a[p + 1] - oa); C.lb = (int)(A.off_b[p + 1] - ob);
    C.pa = A.seq_a + oa; C.pb = A.seq_b + ob;
    C.nstrips = (C.la + STRIP - 1) / STRIP;
    C.pair_edge = (const int2 *)(A.dir + A.dir_off[r]);
    C.pair_ck = C.pair_edge + long_edge_ints2(C.la, C.lb);
    C.ckw = (int)long_ck_width(C.la);
    int wx0 = 0, wy0 = 0, wx1 = -1, wy1 = -1;   /* window: cells (0-based) [wx0, wx1] x [wy0, wy1], kept in registers */
    const uint8_t *flags = C.s_flags;
    auto get = [&](int cx, int cy) -> unsigned {
      if(cx < wx0 || cx > wx1 || cy < wy0 || cy > wy1) {
        walk_ckpt_recompute<IS_SW, NOEND, PROF32>(C, cx, cy);
        wx0 = (cx / STRIP) * STRIP; wx1 = wx0 + ((cx - wx0) / K + 1) * K - 1;
        wy0 = (cy / LONG_CK_ROWS) * LONG_CK_ROWS; wy1 = cy;
      }
      return flags[(cy - wy0) * STRIP + (cx - wx0)];
    };
    walk_pair(A, r, get, lane == 0);
    __syncwarp();
  }
}

/* SW: the pairs' best-cell keys -> score, x_end, y_end */
__global__ void long_sw_finish_kernel(const unsigned long long *__restrict__ key, int64_t n, int32_t *__restrict__ score,
                                      int32_t *__restrict__ xend, int32_t *__restrict__ yend)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  const unsigned long long k = key[i];
  const int sv = (int)(k >> 40);
  score[i] = sv;
  if(xend) xend[i] = sv > 0 ? 0xFFFFF - (int)((k >> 20) & 0xFFFFF) : 0;
  if(yend) yend[i] = sv > 0 ? 0xFFFFF - (int)(k & 0xFFFFF) : 0;
}

/* ---- host side ---------------------------------------------------------- */

inline size_t long_smem_bytes(int K, int ncodes, bool prof32)
{
  const int KS = prof32 ? prof32_stride(K) : (K + 3) / 4;
  const size_t warp_bytes = (size_t)ncodes * 32 * KS * 4 + 64 + 32 * 8;
  const size_t tab = (((size_t)(prof32 ? 4 : 1) * ncodes * (ncodes + 1)) + 15) & ~(size_t)15;
  return 256 + tab + LONG_WARPS * warp_bytes;
}

/* can the wide-pair kernel take this batch?  Needleman-Wunsch with affine
 * gaps (gap_open, gap_extend <= 0), optional free start / end gaps, no gap or
 * mismatch restrictions, values far from the int32 limits. */
inline bool long_plan(const scoring_t *s, const FlatTable &ft, const ScoreParams &sp,
                      int64_t max_la, int64_t max_lb, bool want_dir, LongPlan *plan)
{
  /* Smith-Waterman: score and end cell, traceback only through checkpoints (the caller turns want_dir into
   * plan->ckpt; the flag-byte fill does not track the best cell); coordinates and score have to fit the
   * 20 + 20 + 24 bits of the best-cell key */
  if(sp.is_sw && (sp.no_end || max_la >= (1 << 20) || max_lb >= (1 << 20))) return false;
  if(sp.no_gaps_a || sp.no_gaps_b || sp.no_mismatches) return false;
  if(s->gap_open > 0 || s->gap_extend > 0) return false;   /* needs open <= ext <= 0 */
  if(ft.any_unknown) return false;
  for(size_t k = 0; k < ft.unknown.size(); k++) if(ft.unknown[k]) return false;
  if(max_la < 1 || max_lb < 1 || max_la > (1 << 24) || max_lb > (1 << 24)) return false;
  const long lo = (long)ft.min_sub - sp.open, hi = (long)ft.max_sub - sp.open;
  if(lo < -(1L << 20) || hi > (1L << 20)) return false;
  /* every real value stays inside (-2^28, 2^28): no clamp at the reference's
   * "min" (alignment.c:41) can ever bind, and LONG_NEG + penalty never wins */
  const long longest = (long)(max_la > max_lb ? max_la : max_lb);
  const long step = -(long)sp.open - (long)sp.ext - (ft.min_sub < 0 ? ft.min_sub : 0) + (ft.max_sub > 0 ? ft.max_sub : 0) + 1;
  if(step > (1L << 20) || 2 * longest * step + 2 * labs((long)sp.gap_open) > (1L << 28)) return false;
  const long room = labs((long)s->min_penalty);
  if(-(long)sp.open > room || -(long)sp.ext > room || -(long)ft.min_sub > room) return false;

  const int K = 16, n = ft.ncodes;
  bool prof32 = (size_t)n * 32 * prof32_stride(K) * 4 <= 13 * 1024;
  const int padsub = ft.min_sub < -1 ? ft.min_sub : -1;
  const bool fits8 = lo >= -127 && hi <= 127 && (long)padsub - sp.open >= -127 && (long)padsub - sp.open <= 127;
  if(!prof32 && !fits8) return false;
  if(sp.is_sw && (long)(max_la < max_lb ? max_la : max_lb) * (ft.max_sub > 0 ? ft.max_sub : 0) >= (1L << 23)) return false;
  plan->K = K; plan->is_sw = sp.is_sw != 0; plan->prof32 = prof32; plan->dir = want_dir; plan->noend = sp.no_end != 0;
  plan->smem = long_smem_bytes(K, n, prof32);
  if(plan->smem > 100 * 1024) return false;
  const int tw = n + 1;
  plan->tab8.assign(((size_t)tw * tw + 15) & ~(size_t)15, 0);   /* same geometry as the fast plans (padding row unused here) */
  plan->tab32.assign((size_t)tw * tw + 4, 0);
  for(int cb = 0; cb < tw; cb++)
    for(int ca = 0; ca < tw; ca++) {
      const int v = (ca < n && cb < n ? ft.sub[(size_t)cb * n + ca] : padsub) - sp.open;
      plan->tab32[(size_t)cb * tw + ca] = v;
      if(!prof32) plan->tab8[(size_t)cb * tw + ca] = (int8_t)v;
    }
  plan->name = sp.is_sw ? "long_sw_score_end" : want_dir ? "long_nw_dir" : "long_nw_score";
  plan->ckpt = false;
  return true;
}

template <int K, bool P32, bool DIR, bool NOEND, bool CKPT = false, bool IS_SW = false>
int long_launch_one(const LongPlan &plan, const LongArgs &L, int grid, cudaStream_t st)
{
  void (*kfn)(const LongArgs) = long_kernel<K, IS_SW, P32, DIR, NOEND, CKPT>;
  if(!smem_opt_in(kfn, plan.smem)) return -1;
  SA_LAUNCH(kfn, grid, LONG_WARPS * 32, plan.smem, st, L);
  return 0;
}

/* CTAs resident at once for this plan (the grid of a launch; the caller
 * sizes the boundary buffer with it) */
inline int long_grid(const LongPlan &plan, int num_sms, int64_t npairs)
{
  int64_t g = (int64_t)num_sms * 2;
  if(plan.smem > 110 * 1024) g = num_sms;
  if(g > npairs) g = npairs;
  return g < 1 ? 1 : (int)g;
}

inline int long_launch(const LongPlan &plan, LongArgs L, int grid, cudaStream_t st)
{
  L.mul_one = 1;
  if(plan.is_sw) {
    if(plan.dir) return -1;   /* SW traces back through checkpoints only */
    if(plan.ckpt) return plan.prof32 ? long_launch_one<16, true, false, false, true, true>(plan, L, grid, st)
                                     : long_launch_one<16, false, false, false, true, true>(plan, L, grid, st);
    return plan.prof32 ? long_launch_one<16, true, false, false, false, true>(plan, L, grid, st)
                       : long_launch_one<16, false, false, false, false, true>(plan, L, grid, st);
  }
  if(plan.ckpt) {
    if(plan.prof32) return plan.noend ? long_launch_one<16, true, false, true, true>(plan, L, grid, st)
                                      : long_launch_one<16, true, false, false, true>(plan, L, grid, st);
    return plan.noend ? long_launch_one<16, false, false, true, true>(plan, L, grid, st)
                      : long_launch_one<16, false, false, false, true>(plan, L, grid, st);
  }
#define SA_LONG_CASE(P32_, DIR_, NOEND_)                                          \
  if(plan.prof32 == P32_ && plan.dir == DIR_ && plan.noend == NOEND_)             \
    return long_launch_one<16, P32_, DIR_, NOEND_>(plan, L, grid, st)
  SA_LONG_CASE(true, true, true);   SA_LONG_CASE(true, true, false);
  SA_LONG_CASE(true, false, true);  SA_LONG_CASE(true, false, false);
  SA_LONG_CASE(false, true, true);  SA_LONG_CASE(false, true, false);
  SA_LONG_CASE(false, false, true); SA_LONG_CASE(false, false, false);
#undef SA_LONG_CASE
  return -1;
}

/* the recompute walk of a CKPT fill; tab8 = the plan's int8 table on the device */
inline int walk_ckpt_launch(const LongPlan &plan, const WalkArgs &W, const int8_t *d_tab8, const int32_t *d_tab32,
                            int grid, cudaStream_t st)
{
  const size_t smem = walk_ckpt_smem_bytes(W.sp.ncodes, plan.prof32);
  void (*kfn)(const WalkArgs, const int8_t *, const int32_t *, const int);
  if(plan.is_sw) kfn = plan.prof32 ? walk_ckpt_kernel<true, false, true> : walk_ckpt_kernel<true, false, false>;
  else if(plan.prof32) kfn = plan.noend ? walk_ckpt_kernel<false, true, true> : walk_ckpt_kernel<false, false, true>;
  else kfn = plan.noend ? walk_ckpt_kernel<false, true, false> : walk_ckpt_kernel<false, false, false>;
  if(!smem_opt_in(kfn, smem)) return -1;
  SA_LAUNCH(kfn, grid, WK_WARPS * 32, smem, st, W, d_tab8, d_tab32, 1);
  return 0;
}

/* CTAs of the recompute walk: 2 warps and their two flag tiles (LONG_CK_ROWS x 512 bytes each) in shared
 * memory; as many per SM as that leaves room for (64 rows: three) */
inline int walk_ckpt_grid(int num_sms, int64_t npairs, size_t smem = 70 * 1024)
{
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if(per_sm < 1) per_sm = 1;
  if(per_sm > 16) per_sm = 16;
  int64_t g = (npairs + WK_WARPS - 1) / WK_WARPS;
  if(g > (int64_t)num_sms * per_sm) g = (int64_t)num_sms * per_sm;
  return g < 1 ? 1 : (int)g;
}

} // namespace sa

#endif
