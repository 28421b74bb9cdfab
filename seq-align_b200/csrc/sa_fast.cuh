/*
 * sa_fast.cuh -- sm_100a kernels of the batch alignment engine (part 2):
 * the fill for the common scoring shape (affine gaps with gap_open <= 0, no
 * gap/mismatch restrictions, no free end gaps) and pairs up to 512 columns:
 * fast_score_kernel (int32: score, SW end cell, traceback flag bytes, int16
 * match scores for the multi-hit stage) and fast16_kernel (two pairs per
 * register in packed 16-bit lanes: score, optionally the SW end cell).
 * fast16_kernel is the kernel behind the headline metric (batched 150x150
 * DNA SW).
 *
 * Same recurrence as alignment_fill_matrices (reference
 * src/alignment.c:89-167) and the general kernel, restated for speed:
 *
 *   H = max(M, GA, GB) of a cell.  Because open = gap_open+gap_extend <=
 *   ext = gap_extend, max(M+open, GA+ext, GB+open) == max(H+open, GA+ext)
 *   (the dropped term GA+open is dominated by GA+ext), so a cell needs only
 *   H and GA from above, H and GB from the left and H from the diagonal.
 *   Registers hold H' = H+open, and the substitution scores are stored as
 *   sub' = sub-open, which removes one add per cell:
 *       M  = max(H'diag + sub', min)            VIADDMNMX
 *       GA = max(GA_up + ext, H'up, min)        VIADDMNMX(.RELU)
 *       GB = max(GB_left + ext, H'left, min)    VIADDMNMX(.RELU)
 *       H' = max3(M, GA, GB) + open             VIMNMX3 + IADD
 *
 * Work shape: a warp holds 32/G pairs at once; G lanes per pair, K columns
 * per lane (G*K >= len_a), lanes of a pair staggered one row apart
 * (anti-diagonal wavefront), two warp shuffles per row to pass the strip
 * edge.  Per pair the kernel builds a query profile in shared memory
 * (int8, one row per code of seq_b's alphabet, laid out so that the 32
 * lanes of a warp read conflict-free) -- one shared load per 4 cells and
 * one PRMT per cell replace the table lookup.  Raw sequences are brought
 * into shared memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier),
 * double-buffered so the next pairs load while the current ones compute.
 */
#ifndef SA_FAST_CUH
#define SA_FAST_CUH

#include <vector>
#include "sa_platform.h"
#include "sa_flatten.h"
#include "sa_kernels.cuh"

namespace sa {

constexpr int FAST_WARPS = 4;

struct FastArgs {
  const uint8_t *seq_a, *seq_b;
  const int64_t *off_a, *off_b;
  int64_t npairs;
  ScoreParams sp;
  const int8_t *tab8;   /* [cb*ncodes + ca] = sub - open (int8 profile) */
  const int32_t *tab32; /* same, int32 (small alphabets) */
  uint8_t *dir;         /* DIR kernels: traceback flag bytes, row-major per pair */
  const int64_t *dir_off;
  int16_t *m16;         /* HITS kernels: match scores, same layout/offsets (in elements) */
  int mul_one, mul_key; /* 1 and 32, as runtime values: see fast_score_kernel */
  const uint8_t *lut;
  unsigned long long *counter;
  int32_t *score, *xend, *yend;
  int max_lb;
  int a_stage, b_stage; /* bytes per pair per stage (multiples of 16) */
  int pad_row;          /* fast16: 1 = the profile has a padding row (couples of different shapes), 0 = uniform batch */
  const int *order;     /* fast16: work item i is pair order[i] (length buckets, see bucket_* below); null = identity */
  const int *range;     /* fast16 with order: the launch covers order[range[0] .. range[1]) -- read on the device, so
                           that the launches of all shape classes can be enqueued without waiting for the counts */
};

struct FastPlan {
  int G = 0, K = 0;
  bool is_sw = false;
  const char *name = "";
  std::vector<int8_t> tab8;
  std::vector<int32_t> tab32;
  int track = 0;
  bool prof32 = false;
  bool s16 = false;   /* two pairs per register (fast16_kernel) */
  bool s16_ends = false; /* ... with the SW end cell tracked in 16-bit keys */
  bool s16_rel = false;  /* ... those keys relative to the lane's row input (scores of 1024 - |open| and more) */
  bool dir = false;   /* also write traceback flag bytes */
  bool hits = false;  /* ... and int16 match scores (multi-hit stage) */
  size_t smem = 0;
  int warps = FAST_WARPS;  /* warps per CTA (the packed kernel takes fewer when that leaves more of them resident) */
  int a_stage = 0, b_stage = 0;
  bool pad_row = false; /* fast16: profile with a padding row, for batches whose pairs differ in shape */
};

template <int BYTE>
__device__ __forceinline__ int sext_byte(unsigned w)
{
#if defined(__CUDA_ARCH__)
  int r;
  /* PRMT with the replicate-sign flag: byte BYTE sign-extended to 32 bits */
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0u), "n"(BYTE + (BYTE | 8) * 0x1110));
  return r;
#else
  return (int)(int8_t)(w >> (8 * BYTE));
#endif
}

__device__ __forceinline__ int sext_byte_dyn(unsigned w, int k)
{
  switch(k & 3) {
    case 0: return sext_byte<0>(w);
    case 1: return sext_byte<1>(w);
    case 2: return sext_byte<2>(w);
    default: return sext_byte<3>(w);
  }
}

/* keep a value in its register: ptxas otherwise rematerialises lane indices
 * from SR_TID inside the step loop (an S2R + dependent ops every row) */
__device__ __forceinline__ int pin_reg(int v)
{
#if defined(__CUDA_ARCH__)
  asm volatile("" : "+r"(v));
#endif
  return v;
}

enum { TRACK_NONE = 0, TRACK_TREE = 1, TRACK_COLUMN = 2 };

/* DIR kernels with K not a multiple of 8 stage their rows in shared memory */
__host__ __device__ constexpr bool fast_dir_staged(int K) { return K % 8 != 0; }

/* lane stride (in 32-bit words) of an int32 profile row: K, bumped by 4 when
 * K/4 is even so that 128-bit loads of 8 consecutive lanes hit disjoint banks */
__host__ __device__ constexpr int prof32_stride(int K) { return (K % 4 == 0 && (K / 4) % 2 == 0) ? K + 4 : K; }

/*
 * TRACK  how the SW best cell is kept (ignored for NW):
 *   NONE    score only: running max of M (max3 tree over pairs of columns)
 *   TREE    one packed key per lane, (M<<16 | 31-j<<11 | 2047-y): the max3
 *           tree over a row picks the best column, one IMAD+max per row adds
 *           the row.  Needs len_b <= 2047 and score < 2^15.
 *   COLUMN  one key per column (M<<16 | 0xFFFF-y), len_b <= 65535.
 * PROF32 int32 query profile (small alphabets): no PRMT per cell.
 * DIR    also write one traceback byte per cell (align mode).  Unlike the
 *        general kernel's resolved 2-bit codes these are five raw equality
 *        flags (bit = 0 means "equal"), each one VIADDMNMX: min(a-b, 1) for
 *        operands known to satisfy a >= b:
 *          bit0  GA == H      bit1  GB == H       (of this cell)
 *          bit2  GA == GA_up + ext                (gap_a was extended)
 *          bit3  GB == GB_left + ext              (gap_b was extended)
 *          bit4  GB == H_left + open              (gap_b could have been opened)
 *        walk_kernel resolves them in the reference's priority order
 *        (alignment.c:311-327): the choices depend only on these equalities.
 * HITS   (with DIR) also store every cell's match score as int16, the input
 *        of the multi-hit stage (candidate sort + masked walks, sa_hits.cuh).
 * mul_one / mul_key are the constants 1 and 32 passed as kernel arguments:
 * a multiply-add with a runtime multiplier is a real IMAD (FMA pipe), which
 * takes "H+open" and the key packing off the saturated ALU pipe.
 */
template <int G, int K, bool IS_SW, int TRACK, bool PROF32, bool DIR, bool HITS = false>
__global__ void __launch_bounds__(FAST_WARPS * 32)
fast_score_kernel(const FastArgs A)
{
  constexpr int NG = 32 / G;              /* pairs per warp */
  constexpr int KW = PROF32 ? K : (K + 3) / 4;            /* profile words per lane */
  constexpr int KS = PROF32 ? prof32_stride(K) : KW;      /* lane stride in words */
  constexpr int PSTRIDE = 32 * KS * 4;    /* bytes per profile row */
  /* DIR: how a lane's K flag bytes of a row reach global memory.  K a multiple
   * of 8: one aligned 8/16-byte store per lane.  Otherwise (K = 12, 20) the
   * rows are staged in a shared-memory ring (G+1 rows per pair) and each
   * finished row leaves as full 16-byte vectors, coalesced over the group:
   * scattered 4-byte stores made the L1/L2 store path the bottleneck. */
  constexpr bool STAGED = DIR && fast_dir_staged(K);
  constexpr int RR = G + 1, RS = G * K;

  unsigned char *dsm = SA_DYN_SMEM();
  const ScoreParams &sp = A.sp;
  const int n = sp.ncodes;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int grp = lane / G, lig = pin_reg(lane % G);

  /* shared layout: [mbarriers][lut 256][table n*n (int8 or int32) pad16][per-warp: profile | a stages | b stages] */
  uint64_t *s_bar = (uint64_t *)dsm;                       /* FAST_WARPS*2 */
  uint8_t *s_lut = dsm + 64;
  int8_t *s_tab8 = (int8_t *)(dsm + 64 + 256);
  int32_t *s_tab32 = (int32_t *)(dsm + 64 + 256);
  const int tw = n + 1;   /* table row width: n codes + the padding code */
  const int tab_bytes = ((PROF32 ? 4 : 1) * n * tw + 15) & ~15;
  const int warp_bytes = n * PSTRIDE + 2 * NG * (A.a_stage + A.b_stage) + (STAGED ? NG * RR * RS : 0);
  unsigned char *wbase = dsm + 64 + 256 + tab_bytes + wib * warp_bytes;
  unsigned char *s_prof = wbase;
  unsigned char *s_a = wbase + n * PSTRIDE;                /* [stage][grp][a_stage] */
  unsigned char *s_b = s_a + 2 * NG * A.a_stage;           /* [stage][grp][b_stage] */
  unsigned char *s_ring = s_b + 2 * NG * A.b_stage + grp * (RR * RS);   /* [row slot][RS] of this group */
  uint64_t *bar = s_bar + wib * 2;

  for(int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = A.lut[i];
  if(PROF32) { for(int i = threadIdx.x; i < n * tw; i += blockDim.x) s_tab32[i] = A.tab32[i]; }
  else       { for(int i = threadIdx.x; i < n * tw; i += blockDim.x) s_tab8[i] = A.tab8[i]; }
  if(lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
  mbar_fence_init();
  __syncthreads();

  const int open = sp.open, ext = sp.ext, minv = sp.minv;
  const int mul_one = A.mul_one, mul_key = A.mul_key;
  const int64_t nsets = (A.npairs + NG - 1) / NG;

  /* issue the bulk loads of set t into stage st (one lane per warp) */
  auto issue = [&](int64_t t, int st) {
    if(lane == 0) {
      uint32_t bytes = 0;
      for(int g = 0; g < NG; g++) {
        const int64_t p = t * NG + g;
        if(p >= A.npairs) break;
        const int64_t oa = A.off_a[p], ob = A.off_b[p];
        const int64_t ea = A.off_a[p + 1], eb = A.off_b[p + 1];
        if(ea - oa > G * K || eb - ob > A.max_lb) continue;   /* does not fit the plan: skipped, see below */
        if(ea > oa) bytes += (uint32_t)(((ea + 15) & ~(int64_t)15) - (oa & ~(int64_t)15));
        if(eb > ob) bytes += (uint32_t)(((eb + 15) & ~(int64_t)15) - (ob & ~(int64_t)15));
      }
      mbar_expect_tx(&bar[st], bytes);
      for(int g = 0; g < NG; g++) {
        const int64_t p = t * NG + g;
        if(p >= A.npairs) break;
        const int64_t oa = A.off_a[p], ob = A.off_b[p];
        const int64_t ea = A.off_a[p + 1], eb = A.off_b[p + 1];
        if(ea - oa > G * K || eb - ob > A.max_lb) continue;
        const int64_t a0 = oa & ~(int64_t)15, b0 = ob & ~(int64_t)15;
        if(ea > oa)
          bulk_g2s(s_a + (st * NG + g) * A.a_stage, A.seq_a + a0, (uint32_t)(((ea + 15) & ~(int64_t)15) - a0), &bar[st]);
        if(eb > ob)
          bulk_g2s(s_b + (st * NG + g) * A.b_stage, A.seq_b + b0, (uint32_t)(((eb + 15) & ~(int64_t)15) - b0), &bar[st]);
      }
    }
  };
  auto next_set = [&]() -> int64_t {
    unsigned long long t = 0;
    if(lane == 0) t = atomicAdd(A.counter, 1ull);
    return (int64_t)__shfl_sync(FULL, t, 0);
  };

  int stage = 0;
  unsigned phase0 = 0, phase1 = 0;
  int64_t t = next_set();
  if(t < nsets) issue(t, 0);

  while(t < nsets) {
    /* prefetch the following set into the other stage */
    const int64_t tn = next_set();
    if(tn < nsets) issue(tn, stage ^ 1);

    /* this group's pair */
    const int64_t p = t * NG + grp;
    const bool have = p < A.npairs;
    int la = 0, lb = 0, sha = 0, shb = 0;
    if(have) {
      const int64_t oa = A.off_a[p], ob = A.off_b[p];
      la = (int)(A.off_a[p + 1] - oa); lb = (int)(A.off_b[p + 1] - ob);
      sha = (int)(oa & 15); shb = (int)(ob & 15);
      /* a pair larger than the launch was planned for (only possible when the
       * engine launched speculatively with the previous batch's plan) is
       * treated as empty; the engine notices from the scan and reruns */
      if(la > G * K || lb > A.max_lb) { la = 0; lb = 0; }
    }
    mbar_wait(&bar[stage], stage ? phase1 : phase0);
    if(stage) phase1 ^= 1; else phase0 ^= 1;

    unsigned char *ra = s_a + (stage * NG + grp) * A.a_stage + sha;
    unsigned char *rb = s_b + (stage * NG + grp) * A.b_stage + shb;
    uint8_t *dirp = nullptr;
    int dstride = 0;
    if(DIR && have) { dirp = A.dir + A.dir_off[p]; dstride = (int)dir_stride(la); }

    /* seq_b: raw bytes -> codes, in place */
    for(int i = lig; i < lb; i += G) rb[i] = s_lut[rb[i]];

    /* query profile of my K columns: row c holds sub'(a[x], c) */
    const int xf = lig * K + 1;
    {
      int acode[K];
#pragma unroll
      for(int j = 0; j < K; j++) acode[j] = (xf + j <= la) ? s_lut[ra[xf + j - 1]] : n;   /* n = padding code */
      for(int c = 0; c < n; c++) {
        unsigned *dst = (unsigned *)(s_prof + c * PSTRIDE) + lane * KS;
        if(PROF32) {
          const int32_t *trow = s_tab32 + c * tw;
#pragma unroll
          for(int j = 0; j < K; j++) dst[j] = (unsigned)trow[acode[j]];
        } else {
          const int8_t *trow = s_tab8 + c * tw;
#pragma unroll
          for(int w = 0; w < KW; w++) {
            unsigned word = 0;
#pragma unroll
            for(int q = 0; q < 4; q++) {
              const int j = 4 * w + q;
              if(j < K) word |= (unsigned)(uint8_t)trow[acode[j]] << (8 * q);
            }
            dst[w] = word;
          }
        }
      }
    }
    __syncwarp();

    /* borders (alignment.c:47-81), in H' = H+open form */
    int hp[K], ga[K];
    int colbest[TRACK == TRACK_COLUMN ? K : 1];
#pragma unroll
    for(int j = 0; j < K; j++) {
      const int x = xf + j;
      if(IS_SW) { hp[j] = open; ga[j] = 0; }
      else {
        const int gb0 = sp.no_start ? 0 : addw(sp.gap_open, x * ext);
        hp[j] = addw(imax(gb0, minv), open);
        /* DIR subtracts neighbouring values: keep the "minus infinity" of the
         * borders far from INT_MIN (no real value is below -2^28, see fast_plan) */
        ga[j] = DIR ? -(1 << 29) : minv;
      }
      if constexpr(TRACK == TRACK_COLUMN) colbest[j] = 0;
    }
    int best = 0;
    int hd; /* H'(xf-1, y-1) */
    if(IS_SW || xf == 1) hd = open;
    else hd = addw(imax(sp.no_start ? 0 : addw(sp.gap_open, (xf - 1) * ext), minv), open);

    int out_h = 0, out_gb = 0;
    int maxlb = lb;
#pragma unroll
    for(int o = 16; o >= G; o >>= 1) maxlb = imax(maxlb, __shfl_xor_sync(FULL, maxlb, o));
    const int nsteps = maxlb > 0 ? maxlb + G - 1 : 0;
    const unsigned *prow = (const unsigned *)s_prof + lane * KS;
    /* ring slots: the row this lane writes (y-1 mod RR) and the row the last lane completes */
    int wslot = (RR - lig) % RR, fslot = 2 % RR;

    for(int s = 0; s < nsteps; s++) {
      const int y = s - lig + 1;
      const bool active = y >= 1 && y <= lb;
      int hl = __shfl_up_sync(FULL, out_h, 1);
      int gb = __shfl_up_sync(FULL, out_gb, 1);
      if(lig == 0) {
        /* column 0 */
        if(IS_SW) { hl = open; gb = 0; }
        else {
          hl = addw(imax(sp.no_start ? 0 : addw(sp.gap_open, y * ext), minv), open);
          gb = DIR ? -(1 << 29) : minv;
        }
      }
      const int hl_in = hl;
      if(active) {
        const int c = rb[y - 1];
        const unsigned *pw = prow + c * (PSTRIDE / 4);
        unsigned w[KW];
        /* widest load the row layout allows; the lane strides are chosen so
         * that these loads are bank-conflict free across the warp */
        if(KW % 4 == 0 && KS % 4 == 0) {
#pragma unroll
          for(int q = 0; q < KW / 4; q++) {
            const uint4 v = ((const uint4 *)pw)[q];
            w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
          }
        } else if(KW % 2 == 0 && KS % 2 == 0) {
#pragma unroll
          for(int q = 0; q < KW / 2; q++) {
            const uint2 v = ((const uint2 *)pw)[q];
            w[2 * q] = v.x; w[2 * q + 1] = v.y;
          }
        } else {
#pragma unroll
          for(int q = 0; q < KW; q++) w[q] = pw[q];
        }
        const int ykey = 0xffff - y;
        int d = hd;
        int rowbest = (TRACK == TRACK_NONE) ? best : 0;
        int kprev = 0;
        unsigned dw[DIR ? (K + 3) / 4 : 1];
        unsigned mw[HITS ? (K + 1) / 2 : 1];
        if constexpr(DIR) {
#pragma unroll
          for(int q = 0; q < (K + 3) / 4; q++) dw[q] = 0;
        }
        if constexpr(HITS) {
#pragma unroll
          for(int q = 0; q < (K + 1) / 2; q++) mw[q] = 0;
        }
#pragma unroll
        for(int j = 0; j < K; j++) {
          const int sub = PROF32 ? (int)w[j] : sext_byte_dyn(w[j / 4], j & 3);
          int m, h;
          int uge = 0, lge = 0;
          const int hleft = hl;
          if constexpr(DIR) {
            uge = ga[j] * mul_one + ext;   /* GA_up + ext, GB_left + ext (IMAD) */
            lge = gb * mul_one + ext;
          }
          if(IS_SW) {
            m = addmax(d, sub, 0);
            ga[j] = addmax_relu(ga[j], ext, hp[j]);
            gb = addmax_relu(gb, ext, hl);
            h = max3(m, ga[j], gb);
            if constexpr(TRACK == TRACK_COLUMN) colbest[j] = imax(colbest[j], m * 65536 + ykey);
            else {
              /* padding columns need no mask: they score strictly below the
               * real best (see fast_plan).  TREE packs the column index (IMAD) */
              const int k = TRACK == TRACK_TREE ? m * mul_key + (31 - j) : m;
              if(j & 1) rowbest = max3(rowbest, kprev, k);
              else if(j == K - 1) rowbest = imax(rowbest, k);
              kprev = k;
            }
          } else {
            m = addmax(d, sub, minv);
            ga[j] = max3(addw(ga[j], ext), hp[j], minv);
            gb = max3(addw(gb, ext), hl, minv);
            h = max3(m, ga[j], gb);
          }
          if constexpr(DIR) {
            /* five "not equal" bits, one VIADDMNMX each (a >= b holds for every pair) */
            const int f = imin(h - ga[j], 1) + 2 * imin(h - gb, 1) + 4 * imin(ga[j] - uge, 1) +
                          8 * imin(gb - lge, 1) + 16 * imin(gb - hleft, 1);
            dw[j / 4] += (unsigned)f << (8 * (j & 3));
          }
          if constexpr(HITS) mw[j / 2] += (unsigned)m << (16 * (j & 1));   /* 0 <= m < 2^15 */
          d = hp[j];
          hl = h * mul_one + open;   /* IMAD (FMA pipe) */
          hp[j] = hl;
        }
        if constexpr(DIR) {
          if(have) {
            if constexpr(STAGED) {
              unsigned *rrow = (unsigned *)(s_ring + wslot * RS) + lig * (K / 4);
#pragma unroll
              for(int q = 0; q < K / 4; q++) rrow[q] = dw[q];
            } else if constexpr(K % 16 == 0) {
              uint4 *drow = (uint4 *)(dirp + (int64_t)(y - 1) * dstride + lig * K);
#pragma unroll
              for(int q = 0; q < K / 16; q++)
                if(lig * K + 16 * q < dstride) drow[q] = make_uint4(dw[4 * q], dw[4 * q + 1], dw[4 * q + 2], dw[4 * q + 3]);
            } else {
              uint2 *drow = (uint2 *)(dirp + (int64_t)(y - 1) * dstride + lig * K);
#pragma unroll
              for(int q = 0; q < K / 8; q++)
                if(lig * K + 8 * q < dstride) drow[q] = make_uint2(dw[2 * q], dw[2 * q + 1]);
            }
            if constexpr(HITS) {
              unsigned *mrow = (unsigned *)(A.m16 + A.dir_off[p] + (int64_t)(y - 1) * dstride) + lig * (K / 2);
#pragma unroll
              for(int q = 0; q < K / 2; q++)
                if(lig * K + 2 * q < dstride) mrow[q] = mw[q];
            }
          }
        }
        if(IS_SW && TRACK == TRACK_NONE) best = rowbest;
        if(IS_SW && TRACK == TRACK_TREE) best = imax(best, rowbest * 2048 + (2047 - y));
        out_h = hl;
        out_gb = gb;
        hd = hl_in; /* next row's diagonal */
      }
      if constexpr(STAGED) {
        /* the group's last lane has just finished row s-G+2: flush it */
        __syncwarp();
        const int yc = s - G + 2;
        if(have && yc >= 1 && yc <= lb) {
          const uint4 *src = (const uint4 *)(s_ring + fslot * RS);
          uint4 *dst = (uint4 *)(dirp + (int64_t)(yc - 1) * dstride);
          for(int i = lig; i < dstride / 16; i += G) dst[i] = src[i];
        }
        wslot = wslot + 1 == RR ? 0 : wslot + 1;
        fslot = fslot + 1 == RR ? 0 : fslot + 1;
      }
    }

    /* results */
    if(IS_SW) {
      int bv = 0, bx = 0, by = 0;
      if constexpr(TRACK == TRACK_COLUMN) {
#pragma unroll
        for(int j = 0; j < K; j++) {
          const int v = colbest[j] >> 16;
          if(xf + j <= la && v > bv) { bv = v; bx = xf + j; by = 0xffff - (colbest[j] & 0xffff); }
        }
      } else if(TRACK == TRACK_TREE) {
        bv = best >> 16;
        if(bv > 0) { bx = xf + 31 - ((best >> 11) & 31); by = 2047 - (best & 2047); }
      } else {
        bv = best;
      }
#pragma unroll
      for(int o = G / 2; o > 0; o >>= 1) {
        const int v2 = __shfl_xor_sync(FULL, bv, o);
        const int x2 = __shfl_xor_sync(FULL, bx, o);
        const int y2 = __shfl_xor_sync(FULL, by, o);
        if(hit_better(v2, x2, y2, bv, bx, by)) { bv = v2; bx = x2; by = y2; }
      }
      if(have && lig == 0) {
        A.score[p] = bv;
        if(A.xend) A.xend[p] = bx;
        if(A.yend) A.yend[p] = by;
      }
    } else {
      int sc = 0;
      const int jf = la > 0 ? (la - 1) % K : 0, lf = la > 0 ? (la - 1) / K : 0;
      int v = 0;
#pragma unroll
      for(int j = 0; j < K; j++) if(j == jf) v = hp[j];
      v = __shfl_sync(FULL, v, grp * G + lf);
      if(la > 0 && lb > 0) sc = (int)((unsigned)v - (unsigned)open);
      else if(la == 0 && lb == 0) sc = 0;
      else {
        const int len = la > 0 ? la : lb;
        sc = imax(sp.no_start ? 0 : addw(sp.gap_open, len * ext), minv);
      }
      if(have && lig == 0) {
        A.score[p] = sc;
        if(A.xend) A.xend[p] = la;
        if(A.yend) A.yend[p] = lb;
      }
    }

    /* this stage's buffers were written through the generic proxy (in-place
     * code conversion); order that before the async-proxy refill */
    fence_async_smem();
    __syncwarp();
    stage ^= 1;
    t = tn;
  }
}

/* ---------------------------------------------------------------------------
 * fast16_kernel<G,K>: Smith-Waterman score-only, TWO pairs per lane group,
 * packed in the low / high 16 bits of every register (VIADDMNMX.S16x2,
 * VIMNMX3.S16x2): each DPX instruction advances two alignments.  Used when
 * all pairs of the batch have the same shape and every score provably fits
 * (min(len)*max_sub + |open| < 2^15), so widening back to int32 is exact.
 *
 * All stored values carry a bias B = -open (>= 0), which keeps both halves
 * non-negative: H* = H+B, and the register copy H'* = H*+open.  Because the
 * low half of H* is always >= B, adding the packed constant
 * ((open-1)&0xffff)<<16 | (open&0xffff) with a plain 32-bit IMAD always
 * carries out of the low half, and the "-1" in the high half absorbs it: the
 * "+open" stays off the ALU pipe like in the int32 kernel.  GA/GB need no
 * clamp at zero (they can never drop below open, and M >= 0 dominates the
 * three-way max), M is clamped at B.
 */
template <int PAIR>
__device__ __forceinline__ unsigned pack_sub2(unsigned wlo, unsigned whi)
{
#if defined(__CUDA_ARCH__)
  unsigned r;
  /* byte PAIR of wlo -> low half, byte PAIR of whi -> high half, both sign-extended */
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(wlo), "r"(whi),
      "n"(PAIR | ((PAIR | 8) << 4) | ((4 + PAIR) << 8) | (((4 + PAIR) | 8) << 12)));
  return r;
#else
  const int lo = (int8_t)(wlo >> (8 * PAIR)), hi = (int8_t)(whi >> (8 * PAIR));
  return (unsigned)(unsigned short)lo | ((unsigned)(unsigned short)hi << 16);
#endif
}

__device__ __forceinline__ unsigned pack_sub2_dyn(unsigned wlo, unsigned whi, int k)
{
  switch(k & 3) {
    case 0: return pack_sub2<0>(wlo, whi);
    case 1: return pack_sub2<1>(wlo, whi);
    case 2: return pack_sub2<2>(wlo, whi);
    default: return pack_sub2<3>(wlo, whi);
  }
}

/* N loads of W words each from the profile rows at shared addresses pl / ph */
template <int N, int W, int Q = 0, int KWT>
__device__ __forceinline__ void fast16_load_rows(unsigned pl, unsigned ph, unsigned (&wl)[KWT], unsigned (&wh)[KWT],
                                                 const unsigned char *dsm)
{
  if constexpr(Q < N) {
    if constexpr(W == 4) {
      const uint4 u = lds_b128<16 * Q>(pl, dsm), v = lds_b128<16 * Q>(ph, dsm);
      wl[4 * Q] = u.x; wl[4 * Q + 1] = u.y; wl[4 * Q + 2] = u.z; wl[4 * Q + 3] = u.w;
      wh[4 * Q] = v.x; wh[4 * Q + 1] = v.y; wh[4 * Q + 2] = v.z; wh[4 * Q + 3] = v.w;
    } else if constexpr(W == 2) {
      const uint2 u = lds_b64<8 * Q>(pl, dsm), v = lds_b64<8 * Q>(ph, dsm);
      wl[2 * Q] = u.x; wl[2 * Q + 1] = u.y;
      wh[2 * Q] = v.x; wh[2 * Q + 1] = v.y;
    } else {
      wl[Q] = lds_b32<4 * Q>(pl, dsm);
      wh[Q] = lds_b32<4 * Q>(ph, dsm);
    }
    fast16_load_rows<N, W, Q + 1>(pl, ph, wl, wh, dsm);
  }
}

/* ENDS: also the SW end cell (x_end, y_end) under the reference's hit order.
 * Every cell's match score gets the column index appended, key = M* x 32 +
 * (31 - j), by one packed IMAD (scores below 1024 leave the room inside int16);
 * the row's best key comes out of the same max3 tree as before, and the ROW is
 * tracked once per row step (did the lane's best key grow?) instead of once
 * per cell.  A larger key is a higher score or, at equal score, a smaller
 * column; a later row only replaces an equal score if its column is smaller:
 * that is (score desc, x asc, y asc), smith_waterman.c:71-86. */
template <int G, int K, int ENDS, bool ORD = false>
__global__ void __launch_bounds__(FAST_WARPS * 32)
fast16_kernel(const FastArgs A)
{
  /* ORD: pair indices come through A.order / A.range (length buckets).  Its own instantiation: as a run-time
   * option the indirection cost the plain kernel 16 registers and 3 % of the headline number (measured). */
  constexpr int NG = 32 / G;              /* couples per warp */
  constexpr int NP = 2 * NG;              /* pairs per warp set */
  constexpr int KW = (K + 3) / 4;
  constexpr int PSTRIDE = 32 * KW * 4;    /* bytes per profile row */

  unsigned char *dsm = SA_DYN_SMEM();
  const ScoreParams &sp = A.sp;
  const int n = sp.ncodes, tw = n + 1;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int grp = lane / G, lig = pin_reg(lane % G);

  uint64_t *s_bar = (uint64_t *)dsm;
  uint8_t *s_lut = dsm + 64;
  int8_t *s_tab8 = (int8_t *)(dsm + 64 + 256);
  /* profile rows: the n codes of seq_b and, when the batch is not uniform, the padding row (a fifth row
   * for DNA costs the uniform 150 x 150 batch one resident CTA per SM, 8 % of its speed: measured) */
  const int nr = n + A.pad_row;
  const int padc = A.pad_row ? n : 0;                     /* code of the rows past a pair's end */
  const int tab_bytes = (tw * tw + 15) & ~15;
  const int warp_bytes = 2 * nr * PSTRIDE + 2 * NP * (A.a_stage + A.b_stage);
  unsigned char *wbase = dsm + 64 + 256 + tab_bytes + wib * warp_bytes;
  unsigned char *s_prof = wbase;                           /* [half][code 0..nr-1][PSTRIDE] */
  unsigned char *s_a = wbase + 2 * nr * PSTRIDE;           /* [stage][pair][a_stage] */
  unsigned char *s_b = s_a + 2 * NP * A.a_stage;
  uint64_t *bar = s_bar + wib * 2;

  for(int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = A.lut[i];
  for(int i = threadIdx.x; i < nr * tw; i += blockDim.x) s_tab8[i] = A.tab8[i];
  if(lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
  mbar_fence_init();
  __syncthreads();

  const int open = sp.open, ext = sp.ext;
  const unsigned B = (unsigned)(-open);
  const unsigned BB = B | (B << 16);
  const unsigned EXT2 = ((unsigned)ext & 0xffffu) | ((unsigned)ext << 16);
  /* open < 0: the low half always carries (H* >= B = -open), the -1 in the high half absorbs it.
   * open == 0 (gap_open = gap_extend = 0): nothing to add and no carry to absorb */
  const unsigned OPENC = open == 0 ? 0u : ((((unsigned)(open - 1) & 0xffffu) << 16) | ((unsigned)open & 0xffffu));
  const unsigned mul_one = (unsigned)A.mul_one;
  const int64_t npairs = (ORD && A.range) ? (int64_t)(A.range[1] - A.range[0]) : A.npairs;
  const int *order = !ORD ? nullptr : A.range ? A.order + A.range[0] : A.order;
  const int64_t nsets = (npairs + NP - 1) / NP;

  auto issue = [&](int64_t t, int st) {
    if(lane == 0) {
      uint32_t bytes = 0;
      for(int g = 0; g < NP; g++) {
        const int64_t wi = t * NP + g;
        if(wi >= npairs) break;
        const int64_t p = order ? (int64_t)order[wi] : wi;
        const int64_t oa = A.off_a[p], ob = A.off_b[p];
        const int64_t ea = A.off_a[p + 1], eb = A.off_b[p + 1];
        if(ea - oa > G * K || eb - ob > A.max_lb) continue;   /* does not fit the plan: skipped, see below */
        if(ea > oa) bytes += (uint32_t)(((ea + 15) & ~(int64_t)15) - (oa & ~(int64_t)15));
        if(eb > ob) bytes += (uint32_t)(((eb + 15) & ~(int64_t)15) - (ob & ~(int64_t)15));
      }
      mbar_expect_tx(&bar[st], bytes);
      for(int g = 0; g < NP; g++) {
        const int64_t wi = t * NP + g;
        if(wi >= npairs) break;
        const int64_t p = order ? (int64_t)order[wi] : wi;
        const int64_t oa = A.off_a[p], ob = A.off_b[p];
        const int64_t ea = A.off_a[p + 1], eb = A.off_b[p + 1];
        if(ea - oa > G * K || eb - ob > A.max_lb) continue;
        const int64_t a0 = oa & ~(int64_t)15, b0 = ob & ~(int64_t)15;
        if(ea > oa)
          bulk_g2s(s_a + (st * NP + g) * A.a_stage, A.seq_a + a0, (uint32_t)(((ea + 15) & ~(int64_t)15) - a0), &bar[st]);
        if(eb > ob)
          bulk_g2s(s_b + (st * NP + g) * A.b_stage, A.seq_b + b0, (uint32_t)(((eb + 15) & ~(int64_t)15) - b0), &bar[st]);
      }
    }
  };
  auto next_set = [&]() -> int64_t {
    unsigned long long t = 0;
    if(lane == 0) t = atomicAdd(A.counter, 1ull);
    return (int64_t)__shfl_sync(FULL, t, 0);
  };

  int stage = 0;
  unsigned phase0 = 0, phase1 = 0;
  int64_t t = next_set();
  if(t < nsets) issue(t, 0);

  while(t < nsets) {
    const int64_t tn = next_set();
    if(tn < nsets) issue(tn, stage ^ 1);

    /* this group's couple: pair plo in the low halves, phi in the high halves.  The two pairs
     * may differ in shape: columns past a pair's len_a and rows past its len_b see the padding
     * code n, which scores strictly negative against everything -- with open, ext <= 0 no such
     * cell can reach the pair's best real score (M there is below some earlier H), so neither
     * the running maximum nor the end-cell key needs a mask.  (Uniform batches carry no padding
     * row: only a missing or refused pair has rows past its end there, and its result is not kept.) */
    const int64_t wlo = t * NP + 2 * grp;
    const bool have_lo = wlo < npairs, have_hi = wlo + 1 < npairs;
    const int64_t plo = !have_lo ? 0 : order ? (int64_t)order[wlo] : wlo;
    const int64_t phi = !have_hi ? 0 : order ? (int64_t)order[wlo + 1] : wlo + 1;
    int la_lo = 0, lb_lo = 0, la_hi = 0, lb_hi = 0, sha_lo = 0, shb_lo = 0, sha_hi = 0, shb_hi = 0;
    if(have_lo) {
      const int64_t oa = A.off_a[plo], ob = A.off_b[plo];
      la_lo = (int)(A.off_a[plo + 1] - oa); lb_lo = (int)(A.off_b[plo + 1] - ob);
      sha_lo = (int)(oa & 15); shb_lo = (int)(ob & 15);
      /* speculative launches only: a pair the plan did not foresee becomes an empty one (as in issue()) */
      if(la_lo > G * K || lb_lo > A.max_lb) { la_lo = 0; lb_lo = 0; }
    }
    if(have_hi) {
      const int64_t oa = A.off_a[phi], ob = A.off_b[phi];
      la_hi = (int)(A.off_a[phi + 1] - oa); lb_hi = (int)(A.off_b[phi + 1] - ob);
      sha_hi = (int)(oa & 15); shb_hi = (int)(ob & 15);
      if(la_hi > G * K || lb_hi > A.max_lb) { la_hi = 0; lb_hi = 0; }
    }
    const int lb = imax(lb_lo, lb_hi);
    mbar_wait(&bar[stage], stage ? phase1 : phase0);
    if(stage) phase1 ^= 1; else phase0 ^= 1;

    const int slot_lo = stage * NP + 2 * grp, slot_hi = slot_lo + 1;
    unsigned char *ra_lo = s_a + slot_lo * A.a_stage + sha_lo;
    unsigned char *rb_lo = s_b + slot_lo * A.b_stage + shb_lo;
    unsigned char *ra_hi = s_a + slot_hi * A.a_stage + sha_hi;
    unsigned char *rb_hi = s_b + slot_hi * A.b_stage + shb_hi;

    /* seq_b of both pairs: raw bytes -> codes, in place; rows past the end of a pair get the padding code */
    for(int i = lig; i < lb; i += G) {
      rb_lo[i] = i < lb_lo ? s_lut[rb_lo[i]] : (unsigned char)padc;
      rb_hi[i] = i < lb_hi ? s_lut[rb_hi[i]] : (unsigned char)padc;
    }
    __syncwarp();
    const unsigned char *cb_hi = rb_hi;

    /* query profiles of my K columns, one per half; row n is the padding row */
    const int xf = lig * K + 1;
#pragma unroll
    for(int half = 0; half < 2; half++) {
      const unsigned char *ra = half ? ra_hi : ra_lo;   /* seq_a stays raw in shared memory */
      const int la = half ? la_hi : la_lo;
      int acode[K];
#pragma unroll
      for(int j = 0; j < K; j++) acode[j] = (xf + j <= la) ? s_lut[ra[xf + j - 1]] : n;
      for(int c = 0; c < nr; c++) {
        const int8_t *trow = s_tab8 + c * tw;
        unsigned *dst = (unsigned *)(s_prof + (half * nr + c) * PSTRIDE) + lane * KW;
#pragma unroll
        for(int w = 0; w < KW; w++) {
          unsigned word = 0;
#pragma unroll
          for(int q = 0; q < 4; q++) {
            const int j = 4 * w + q;
            if(j < K) word |= (unsigned)(uint8_t)trow[acode[j]] << (8 * q);
          }
          dst[w] = word;
        }
      }
    }
    __syncwarp();

    /* borders, biased: H = 0 -> H'* = B + open = 0 ; GA = GB = 0 -> B */
    unsigned hp[K], ga[K];
#pragma unroll
    for(int j = 0; j < K; j++) { hp[j] = 0; ga[j] = BB; }
    unsigned best = ENDS == 1 ? BB * 32u : BB, hd = 0, out_h = 0, out_gb = BB;
    unsigned bk_lo = B * 32u, bk_hi = B * 32u;   /* ENDS == 2: the lane's best (M* x 32 + 31 - j), one word per half */
    int ylo = 0, yhi = 0;
    const unsigned mul_key = (unsigned)A.mul_key;

    int maxlb = lb;
#pragma unroll
    for(int o = 16; o >= G; o >>= 1) maxlb = imax(maxlb, __shfl_xor_sync(FULL, maxlb, o));
    const int nsteps = maxlb > 0 ? maxlb + G - 1 : 0;
    const unsigned prow_lo = smem_addr(s_prof, dsm) + (unsigned)lane * (KW * 4);
    const unsigned prow_hi = prow_lo + (unsigned)nr * PSTRIDE;
    /* Everything of the row step that is not the recurrence stays off the ALU
     * pipe (it is the saturated one): the column-0 constants of a pair's first
     * lane come from multiply-adds with per-lane constants instead of selects,
     * the row's position in seq_b is a running shared-memory offset advanced
     * by an IMAD, and the codes of the next row are fetched one step ahead. */
    const unsigned nf = (unsigned)pin_reg(lig != 0), gbc = (unsigned)pin_reg(lig == 0 ? (int)BB : 0);
    unsigned ilo = smem_addr(rb_lo, dsm) - (unsigned)lig, ihi = smem_addr(cb_hi, dsm) - (unsigned)lig;
    unsigned clo = lds_u8(ilo, dsm), chi = lds_u8(ihi, dsm);   /* row 1 of lane 0; later lanes read slack bytes they do not use */

    for(int s = 0; s < nsteps; s++) {
      const bool active = (unsigned)(s - lig) < (unsigned)lb;   /* 1 <= y <= lb, y = s - lig + 1 */
      unsigned hl = __shfl_up_sync(FULL, out_h, 1) * nf;        /* column 0: H'* = 0 */
      unsigned gb = __shfl_up_sync(FULL, out_gb, 1) * nf + gbc; /*           GB* = B  */
      const unsigned hl_in = hl;
      const unsigned cl = clo, ch = chi;
      ilo = ilo * mul_one + 1u;
      ihi = ihi * mul_one + 1u;
      clo = lds_u8(ilo, dsm); chi = lds_u8(ihi, dsm);
      if(active) {
        /* profile rows of the two codes, through 32-bit shared addresses (row
         * offset by IMAD); widest loads the lane stride (KW words) allows --
         * scalar loads with an even stride would be 2- or 4-way bank conflicts */
        const unsigned pl = cl * (unsigned)PSTRIDE + prow_lo, ph = ch * (unsigned)PSTRIDE + prow_hi;
        unsigned wl[KW], wh[KW];
        if constexpr(KW % 4 == 0) {
          fast16_load_rows<KW / 4, 4>(pl, ph, wl, wh, dsm);
        } else if constexpr(KW % 2 == 0) {
          fast16_load_rows<KW / 2, 2>(pl, ph, wl, wh, dsm);
        } else {
          fast16_load_rows<KW, 1>(pl, ph, wl, wh, dsm);
        }
        unsigned d = hd, kprev = ENDS ? 0u : BB, rowkey = 0;
        /* ENDS == 2: keys relative to the value entering the lane on this row, (M* - H'*_in + 512) x 32 + (31 - j).
         * 32-bit arithmetic on the packed words is exact because every half of the result is in [0, 1023]
         * (fast_plan checks the bound) -- a negative half of relc borrows from its neighbour and the sum pays it back. */
        const unsigned relc = 0x02000200u - hl_in;
#pragma unroll
        for(int j = 0; j < K; j++) {
          const unsigned sub = pack_sub2_dyn(wl[j / 4], wh[j / 4], j & 3);
          const unsigned m = addmax_s16x2(d, sub, BB);
          ga[j] = addmax_s16x2(ga[j], EXT2, hp[j]);
          gb = addmax_s16x2(gb, EXT2, hl);
          const unsigned h = max3_s16x2(m, ga[j], gb);
          if constexpr(ENDS) {
            const unsigned mk = ENDS == 2 ? m * mul_one + relc : m;
#ifdef SA_EMU
            if(ENDS == 2 && ((mk & 0xffffu) > 1023u || (mk >> 16) > 1023u)) __builtin_trap();   /* fast_plan's bound */
#endif
            const unsigned kk = mk * mul_key + (unsigned)((31 - j) * 0x10001);   /* both halves: M* x 32 + (31 - j) */
            if(j & 1) rowkey = max3_s16x2(rowkey, kprev, kk);
            else if(j == K - 1) rowkey = max3_s16x2(rowkey, kk, kk);
            kprev = kk;
          } else {
            if(j & 1) best = max3_s16x2(best, kprev, m);
            else if(j == K - 1) best = max3_s16x2(best, m, m);
            kprev = m;
          }
          d = hp[j];
          hl = h * mul_one + OPENC;   /* packed H* + open, see header */
          hp[j] = hl;
        }
        if constexpr(ENDS == 2) {
          /* once per row: the row's best cell back to an absolute score (the key's order inside one row is
           * the order of the scores), then the same (score x 32 + 31 - j) comparison as below in full words */
          const unsigned rowm = ((rowkey >> 5) & 0x07ff07ffu) - relc;
          const unsigned k_lo = (rowm & 0xffffu) * mul_key + (rowkey & 31u);
          const unsigned k_hi = (rowm >> 16) * mul_key + ((rowkey >> 16) & 31u);
          const int y = s - lig + 1;
          if(k_lo > bk_lo) { bk_lo = k_lo; ylo = y; }
          if(k_hi > bk_hi) { bk_hi = k_hi; yhi = y; }
        } else if constexpr(ENDS == 1) {
          /* once per row: where the lane's best key grew, this row is its row */
          const unsigned nb = max3_s16x2(best, rowkey, rowkey);
          const unsigned grew = nb ^ best;
          const int y = s - lig + 1;
          if(grew & 0xffffu) ylo = y;
          if(grew >> 16) yhi = y;
          best = nb;
        }
        out_h = hl;
        out_gb = gb;
        hd = hl_in;
      }
    }

    if constexpr(ENDS) {
      /* per half: (score, x, y) of the lane, then the group's best under the hit order */
      int sv[2], sx[2], sy[2];
#pragma unroll
      for(int hsel = 0; hsel < 2; hsel++) {
        const int key = ENDS == 2 ? (int)(hsel ? bk_hi : bk_lo) : (int)((best >> (16 * hsel)) & 0xffffu);
        sv[hsel] = (key >> 5) - (int)B;
        sx[hsel] = sv[hsel] > 0 ? xf + 31 - (key & 31) : 0;
        sy[hsel] = sv[hsel] > 0 ? (hsel ? yhi : ylo) : 0;
#pragma unroll
        for(int o = G / 2; o > 0; o >>= 1) {
          const int v2 = __shfl_xor_sync(FULL, sv[hsel], o);
          const int x2 = __shfl_xor_sync(FULL, sx[hsel], o);
          const int y2 = __shfl_xor_sync(FULL, sy[hsel], o);
          if(hit_better(v2, x2, y2, sv[hsel], sx[hsel], sy[hsel])) { sv[hsel] = v2; sx[hsel] = x2; sy[hsel] = y2; }
        }
      }
      if(lig == 0) {
        if(have_lo) { A.score[plo] = sv[0]; if(A.xend) A.xend[plo] = sx[0]; if(A.yend) A.yend[plo] = sy[0]; }
        if(have_hi) { A.score[phi] = sv[1]; if(A.xend) A.xend[phi] = sx[1]; if(A.yend) A.yend[phi] = sy[1]; }
      }
    } else {
#pragma unroll
      for(int o = G / 2; o > 0; o >>= 1) {
        const unsigned v = __shfl_xor_sync(FULL, best, o);
        best = max3_s16x2(best, v, v);
      }
      if(lig == 0) {
        if(have_lo) {
          A.score[plo] = (int)(best & 0xffffu) - (int)B;
          if(A.xend) A.xend[plo] = 0;
          if(A.yend) A.yend[plo] = 0;
        }
        if(have_hi) {
          A.score[phi] = (int)(best >> 16) - (int)B;
          if(A.xend) A.xend[phi] = 0;
          if(A.yend) A.yend[phi] = 0;
        }
      }
    }

    fence_async_smem();
    __syncwarp();
    stage ^= 1;
    t = tn;
  }
}

/* ---- length buckets for the packed kernel ---------------------------------
 * A batch of reads of different lengths is shaped by its longest seq_a (columns) and every couple by
 * its longer seq_b (rows).  Bucketing gives each pair the narrowest kernel shape that holds its seq_a
 * and puts pairs of similar len_b next to each other: a counting sort on the key
 * (shape class, len_b >> shift) -- histogram, one-CTA scan, scatter -- whose output is the `order`
 * array the kernel reads its pair indices through.  The order inside a bin is whatever the atomics
 * make it; results are written by pair index, so they do not depend on it. */
constexpr int BUCKET_LB_BINS = 256;
constexpr int BUCKET_MAX_CLASSES = 16;
struct BucketArgs {
  const int64_t *off_a, *off_b;
  int64_t npairs;
  int nclasses;
  int width[BUCKET_MAX_CLASSES];   /* columns of class c, ascending */
  int shift;                       /* len_b >> shift < BUCKET_LB_BINS */
  int *bins;                       /* [nclasses * BUCKET_LB_BINS] counts, then starts */
  int *cursor;                     /* same size: running positions of the scatter */
  int *class_start;                /* [nclasses + 1] */
  int *order;
};

__device__ __forceinline__ int bucket_key(const BucketArgs &B, int64_t p)
{
  const int la = (int)(B.off_a[p + 1] - B.off_a[p]), lb = (int)(B.off_b[p + 1] - B.off_b[p]);
  int c = 0;
  while(c + 1 < B.nclasses && la > B.width[c]) c++;
  /* longest seq_b first inside a class: the persistent kernel hands out work in this order, and the long jobs
   * must not be the ones left for the tail */
  return c * BUCKET_LB_BINS + (BUCKET_LB_BINS - 1 - imin(lb >> B.shift, BUCKET_LB_BINS - 1));
}

__global__ void bucket_hist_kernel(const BucketArgs B)
{
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(p < B.npairs) atomicAdd(&B.bins[bucket_key(B, p)], 1);
}

/* one CTA of 256 threads: counts -> starts (in place), a copy for the scatter, the class boundaries */
__global__ void bucket_scan_kernel(const BucketArgs B)
{
  __shared__ int s_part[256];
  const int nb = B.nclasses * BUCKET_LB_BINS;
  const int per = (nb + 255) / 256, b0 = threadIdx.x * per;
  int sum = 0;
  for(int i = b0; i < b0 + per && i < nb; i++) sum += B.bins[i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if(threadIdx.x == 0) {
    int run = 0;
    for(int i = 0; i < 256; i++) { const int t = s_part[i]; s_part[i] = run; run += t; }
  }
  __syncthreads();
  int run = s_part[threadIdx.x];
  for(int i = b0; i < b0 + per && i < nb; i++) {
    const int t = B.bins[i];
    B.bins[i] = run; B.cursor[i] = run;
    if(i % BUCKET_LB_BINS == 0) B.class_start[i / BUCKET_LB_BINS] = run;
    run += t;
  }
  if(threadIdx.x == 0) B.class_start[B.nclasses] = (int)B.npairs;
}

__global__ void bucket_scatter_kernel(const BucketArgs B)
{
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(p < B.npairs) B.order[atomicAdd(&B.cursor[bucket_key(B, p)], 1)] = (int)p;
}

/* ---- host side ---------------------------------------------------------- */

struct FastShape { int G, K; };
static const FastShape kFastShapes[] = {
    {8, 8}, {8, 12}, {8, 16}, {8, 20}, {16, 12}, {16, 16}, {32, 10}, {32, 12}, {32, 16}};
static const FastShape kFast16Shapes[] = {
    {8, 8}, {8, 12}, {8, 13}, {8, 16}, {8, 19}, {8, 20}, {16, 12}, {16, 16}, {16, 19}, {32, 10}, {32, 12}, {32, 13}, {32, 16}};

inline size_t fast_smem_bytes(int G, int K, int ncodes, bool prof32, int a_stage, int b_stage, bool dir = false)
{
  const int NG = 32 / G;
  const int KS = prof32 ? prof32_stride(K) : (K + 3) / 4;
  const size_t ring = dir && fast_dir_staged(K) ? (size_t)NG * (G + 1) * G * K : 0;
  const size_t warp_bytes = (size_t)ncodes * 32 * KS * 4 + 2 * (size_t)NG * (a_stage + b_stage) + ring;
  const size_t tab = (((size_t)(prof32 ? 4 : 1) * ncodes * (ncodes + 1)) + 15) & ~(size_t)15;
  return 64 + 256 + tab + FAST_WARPS * warp_bytes;
}

/* can the fast kernel take this batch?  fills the plan if so.
 * want_ends: the caller needs the SW end cell (x_end, y_end); allow_s16: the packed 16-bit kernel may be chosen */
inline bool fast_plan(const scoring_t *s, const FlatTable &ft, const ScoreParams &sp,
                      int64_t max_la, int64_t max_lb, bool want_ends, bool allow_s16, FastPlan *plan,
                      bool want_dir = false, bool uniform = false)
{
  if(sp.no_end || sp.no_gaps_a || sp.no_gaps_b || sp.no_mismatches) return false;
  if(s->gap_open > 0 || s->gap_extend > 0) return false;   /* needs open <= ext <= 0 */
  if(ft.any_unknown) return false;
  if(max_la < 1 || max_lb < 1 || max_la > 512 || max_lb > 65535) return false;
  for(size_t k = 0; k < ft.unknown.size(); k++) if(ft.unknown[k]) return false;
  const long lo = (long)ft.min_sub - sp.open, hi = (long)ft.max_sub - sp.open;
  if(lo < -(1L << 24) || hi > (1L << 24)) return false;
  const long longest = (long)(max_la > max_lb ? max_la : max_lb), shortest = (long)(max_la < max_lb ? max_la : max_lb);
  if(sp.is_sw) {
    /* packed (score,position) keys: scores must stay below 2^15 */
    const long cap = shortest * (ft.max_sub > 0 ? ft.max_sub : 0);
    if(cap >= 32768) return false;
  } else {
    /* sentinel arithmetic must not wrap (alignment.c:41) */
    const long room = labs((long)s->min_penalty);
    if(-(long)sp.open > room || -(long)sp.ext > room || -(long)ft.min_sub > room) return false;
    if(sp.ext > 0 || sp.open > 0) return false;
    const long worst = (long)sp.gap_open + longest * sp.ext;
    if(worst < -(1L << 30)) return false;
    if(want_dir && (worst + longest * (ft.min_sub < 0 ? ft.min_sub : 0) < -(1L << 28) ||
                    shortest * (ft.max_sub > 0 ? ft.max_sub : 0) > (1L << 28)))
      return false;
  }
  const int n = ft.ncodes;
  const int padsub = ft.min_sub < -1 ? ft.min_sub : -1;
  const bool fits8 = lo >= -127 && hi <= 127 && (long)padsub - sp.open >= -127 && (long)padsub - sp.open <= 127;
  /* packed 16-bit kernel: SW score only, one shape for the whole batch, and
   * every biased value (score + |open|) inside int16 */
  const long best_biased = shortest * (ft.max_sub > 0 ? ft.max_sub : 0) - sp.open;   /* largest M + B */
  /* ... with the end cell: the 16-bit key is (M + B) x 32 + column, so M + B must stay below 1024; larger
   * scores take keys relative to the lane's input of the row (ends_mode 2), bounded further down */
  const int ends_mode = !want_ends || max_lb > 32767 ? 0 : best_biased < 1024 ? 1 : 2;
  /* (the packed kernel takes couples of different shapes: padding code for the shorter one) */
  bool s16 = sp.is_sw && (!want_ends || ends_mode) && !want_dir && allow_s16 && fits8 &&
             best_biased < 32000 && sp.open > -16000 && ft.min_sub > -16000;
  int G = 0, K = 0;
  if(s16) {
    /* the packed kernel has extra shapes that fit common read lengths tightly */
    /* relative keys (ends_mode 2): for a lane's columns x0 .. x0+K-1 on row y, D = M[x][y] - H[x0-1][y] obeys
     *   D <= (j + 1)(max_sub - open)                  (H rises by at most max_sub - open per step, any direction)
     *   D >= padsub - max_sub + 2 open + (j - 1) ext  (a gap from column x0-1 to x-1 on row y-1 bounds H from below)
     * and the key holds D - open + 512 in ten bits: the first shape whose K keeps both inside takes the batch */
    const bool rel = want_ends && ends_mode == 2;
    for(const FastShape &sh : kFast16Shapes) {
      if((int64_t)sh.G * sh.K < max_la) continue;
      if(rel) {
        const long up = (long)sh.K * ((long)ft.max_sub - sp.open) - sp.open;
        const long dn = -((long)padsub - ft.max_sub + sp.open + (long)(sh.K - 2) * sp.ext);
        if(up > 500 || dn > 500) continue;
      }
      G = sh.G; K = sh.K;
      break;
    }
    if(!G) s16 = false;
  }
  if(!s16) {
    for(const FastShape &sh : kFastShapes)
      if((int64_t)sh.G * sh.K >= max_la && (!want_dir || sh.K % 4 == 0)) { G = sh.G; K = sh.K; break; }
  }
  if(!G) return false;
  /* int32 profile when it stays small (DNA-sized alphabets); else int8, which
   * needs sub' = sub - open to fit a signed byte */
  bool prof32 = (size_t)n * 32 * prof32_stride(K) * 4 <= 16 * 1024;
  if(!prof32 && !fits8) return false;
  if(s16) prof32 = false;
  plan->s16 = s16;
  plan->s16_ends = s16 && want_ends;
  plan->s16_rel = s16 && want_ends && ends_mode == 2;
  plan->dir = want_dir;
  plan->G = G; plan->K = K; plan->is_sw = sp.is_sw != 0; plan->prof32 = prof32;
  plan->track = !sp.is_sw ? TRACK_NONE : (!want_ends ? TRACK_NONE : (max_lb <= 2047 ? TRACK_TREE : TRACK_COLUMN));
  plan->a_stage = (int)((G * K + 15 + 15) & ~15) + 16;
  plan->b_stage = (int)((max_lb + 15 + 15) & ~(int64_t)15) + 16;
  if(s16) {
    /* exact sizes (a fifth CTA per SM depends on them): a bulk copy brings at
     * most floor16(len + 30) bytes; the packed kernel also reads seq_b codes one
     * row ahead in every lane, up to G + 1 bytes past the end of the sequence */
    plan->a_stage = (int)((G * K + 15 + 15) & ~15);
    plan->b_stage = (int)((max_lb + 15 + 15 + G + 1) & ~(int64_t)15);
  }
  plan->smem = fast_smem_bytes(G, K, n, prof32, plan->a_stage, plan->b_stage, want_dir);
  if(s16) {
    plan->pad_row = !uniform;
    const size_t warp_bytes = 2 * (size_t)(n + (uniform ? 0 : 1)) * 32 * ((K + 3) / 4) * 4 + 2 * (size_t)(2 * (32 / G)) * (plan->a_stage + plan->b_stage);
    /* warps per CTA: four, unless the profile is so large (protein alphabets: 20+ rows of 512 bytes per half) that
     * few CTAs fit an SM and a smaller CTA leaves more warps resident (228 KB per SM, 1 KB reserved per CTA) */
    const size_t fixed = 64 + 256 + (((size_t)(n + 1) * (n + 1) + 15) & ~(size_t)15);
    int best_w = FAST_WARPS, best_res = 0;
    for(int w = FAST_WARPS; w >= 2; w--) {
      const int res = w * (int)(233472 / (fixed + (size_t)w * warp_bytes + 1024));
      if(w == FAST_WARPS) { best_res = res; if(res > 10) break; }
      else if(res > best_res) { best_res = res; best_w = w; }
    }
    if(const char *env = getenv("SEQALIGN_FAST16_WARPS")) {   /* experiment knob: the numbers in DESIGN.md 3 K1 */
      const int w = atoi(env);
      if(w >= 2 && w <= FAST_WARPS) best_w = w;
    }
    plan->warps = best_w;
    plan->smem = fixed + (size_t)best_w * warp_bytes;
  }
  if(plan->smem > 200 * 1024) return false;
  /* tables are n+1 rows (code of seq_b, last = padding row) x n+1 columns
   * (code of seq_a, last = padding).  A padding cell scores strictly negative
   * against everything, so (with open, ext <= 0) no cell in a padding column
   * or row can reach the best real score and the kernels need not mask it out
   * of the running maximum. */
  const int tw = n + 1;
  plan->tab8.assign(((size_t)tw * tw + 15) & ~(size_t)15, 0);
  plan->tab32.assign((size_t)tw * tw + 4, 0);
  for(int cb = 0; cb < tw; cb++)
    for(int ca = 0; ca < tw; ca++) {
      const int v = (ca < n && cb < n ? ft.sub[(size_t)cb * n + ca] : padsub) - sp.open;
      plan->tab32[(size_t)cb * tw + ca] = v;
      if(!prof32) plan->tab8[(size_t)cb * tw + ca] = (int8_t)v;
    }
  if(want_dir) {
    plan->name = sp.is_sw ? "fast_sw_dir" : "fast_nw_dir";
    return true;
  }
  plan->name = !sp.is_sw ? "fast_nw_score" : (s16 && want_ends) ? (ends_mode == 2 ? "fast16_sw_score_endrel" : "fast16_sw_score_end") : s16 ? "fast16_sw_score"
             : plan->track == TRACK_NONE ? "fast_sw_score" : plan->track == TRACK_TREE ? "fast_sw_score_end" : "fast_sw_score_endcol";
  return true;
}

/* Shared-memory opt-in and resident CTAs per SM of a kernel, asked from the
 * runtime once per (kernel, shared memory, device, host thread) and remembered:
 * the two runtime calls cost several microseconds on every launch otherwise. */
struct FastKernelInfo { const void *fn; size_t smem; int device; int per_sm; };
template <class KF>
int fast_per_sm(KF kfn, size_t smem, int threads = FAST_WARPS * 32)
{
  thread_local std::vector<FastKernelInfo> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  for(const FastKernelInfo &k : cache)
    if(k.fn == (const void *)kfn && k.device == dev && k.smem == smem) return k.per_sm;
  if(!smem_opt_in(kfn, smem)) return -1;
  int per_sm = 1;
  if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  cache.push_back({(const void *)kfn, smem, dev, per_sm});
  return per_sm;
}

/* persistent grid: exactly as many CTAs as are resident at once; -1 if the kernel cannot take this much shared memory */
template <class KF>
int fast_grid(KF kfn, size_t smem, int num_sms, int64_t need, int threads = FAST_WARPS * 32)
{
  const int per_sm = fast_per_sm(kfn, smem, threads);
  if(per_sm < 0) return -1;
  int64_t grid = (int64_t)num_sms * per_sm;
  if(grid > need) grid = need;
  return grid < 1 ? 1 : (int)grid;
}

template <int G, int K, bool P32>
int fast_launch_gkp(const FastPlan &plan, const FastArgs &F, int num_sms, int64_t need, cudaStream_t st)
{
  void (*kfn)(const FastArgs) = nullptr;
  if(plan.dir && plan.hits) {
    if constexpr(K % 4 == 0) kfn = fast_score_kernel<G, K, true, TRACK_NONE, P32, true, true>;
  }
  else if(plan.dir) {
    if constexpr(K % 4 == 0) {
      if(!plan.is_sw) kfn = fast_score_kernel<G, K, false, TRACK_NONE, P32, true>;
      else if(plan.track == TRACK_TREE) kfn = fast_score_kernel<G, K, true, TRACK_TREE, P32, true>;
      else kfn = fast_score_kernel<G, K, true, TRACK_COLUMN, P32, true>;
    }
  }
  else if(!plan.is_sw) kfn = fast_score_kernel<G, K, false, TRACK_NONE, P32, false>;
  else if(plan.track == TRACK_NONE) kfn = fast_score_kernel<G, K, true, TRACK_NONE, P32, false>;
  else if(plan.track == TRACK_TREE) kfn = fast_score_kernel<G, K, true, TRACK_TREE, P32, false>;
  else kfn = fast_score_kernel<G, K, true, TRACK_COLUMN, P32, false>;
  if(!kfn) return -1;
  const int grid = fast_grid(kfn, plan.smem, num_sms, need);
  if(grid < 0) return -1;
  SA_LAUNCH(kfn, grid, FAST_WARPS * 32, plan.smem, st, F);
  return 0;
}

inline int fast_launch(const FastPlan &plan, FastArgs F, int num_sms, size_t smem_optin, cudaStream_t st)
{
  F.a_stage = plan.a_stage; F.b_stage = plan.b_stage;
  F.pad_row = plan.pad_row ? 1 : 0;
  F.mul_one = 1; F.mul_key = 32;
  if(plan.smem > smem_optin) return -1;
  const int NG = (plan.s16 ? 2 : 1) * (32 / plan.G);
  const int64_t nsets = (F.npairs + NG - 1) / NG;
  const int warps = plan.s16 ? plan.warps : FAST_WARPS;
  const int64_t need = (nsets + warps - 1) / warps;
#define SA_FAST16_CASE(g, k)                                                                  \
  if(plan.G == g && plan.K == k && plan.s16) {                                                \
    void (*kfn)(const FastArgs) = F.order ? (plan.s16_rel ? fast16_kernel<g, k, 2, true> : plan.s16_ends ? fast16_kernel<g, k, 1, true> : fast16_kernel<g, k, 0, true>) \
                                          : (plan.s16_rel ? fast16_kernel<g, k, 2> : plan.s16_ends ? fast16_kernel<g, k, 1> : fast16_kernel<g, k, 0>); \
    /* SEQALIGN_FAST_PAD_SMEM: extra bytes of (unused) shared memory per CTA, an occupancy knob for experiments */ \
    const char *pad_env = getenv("SEQALIGN_FAST_PAD_SMEM");                                   \
    const size_t smem16 = plan.smem + (pad_env ? (size_t)atoi(pad_env) : 0);                  \
    const int grid16 = fast_grid(kfn, smem16, num_sms, need, warps * 32);                     \
    if(grid16 < 0) return -1;                                                                 \
    SA_LAUNCH(kfn, grid16, warps * 32, smem16, st, F);                                        \
    return 0;                                                                                 \
  }
#define SA_FAST_CASE(g, k)                                                                    \
  SA_FAST16_CASE(g, k)                                                                        \
  if(plan.G == g && plan.K == k)                                                              \
    return plan.prof32 ? fast_launch_gkp<g, k, true>(plan, F, num_sms, need, st)               \
                       : fast_launch_gkp<g, k, false>(plan, F, num_sms, need, st)
  SA_FAST16_CASE(8, 13) SA_FAST16_CASE(8, 19) SA_FAST16_CASE(16, 19) SA_FAST16_CASE(32, 13)
  SA_FAST_CASE(8, 8); SA_FAST_CASE(8, 12); SA_FAST_CASE(8, 16); SA_FAST_CASE(8, 20);
  SA_FAST_CASE(16, 12); SA_FAST_CASE(16, 16);
  SA_FAST_CASE(32, 10); SA_FAST_CASE(32, 12); SA_FAST_CASE(32, 16);
#undef SA_FAST_CASE
#undef SA_FAST16_CASE
  return -1;
}

} // namespace sa

#endif
