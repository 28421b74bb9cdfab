/*
 * sa_hits.cuh -- Smith-Waterman multi-hit iteration on the device.
 *
 * Takes over what smith_waterman_align2 + smith_waterman_fetch do after the
 * fill (reference src/smith_waterman.c:152-161 and 165-277): collect every
 * cell whose match score is positive, order the candidates by score
 * descending, then column (x) ascending, then row (y) ascending -- the
 * comparator of smith_waterman.c:71-86 made total by glibc's stable qsort_r
 * over indices generated in ascending order -- and walk them back one after
 * the other through a visited mask: a candidate that is already marked is
 * skipped, a walk that runs into a marked cell is dropped (its own marks
 * stay), every other walk is a hit.  In the reference this sort is the
 * dominant cost of SW (about 5x the fill).
 *
 * Inputs are what fast_score_kernel<..., DIR, HITS> left in HBM: one flag
 * byte and one int16 match score per cell.
 *
 *   hits_sort_kernel   one warp per pair.  Candidates >= min_score are
 *                      compacted in row-major order (y, then x ascending), so
 *                      a STABLE counting sort by x followed by a stable
 *                      counting sort by descending score yields the required
 *                      order; no comparison sort, all accesses coalesced.
 *                      Stability inside a 32-key tile comes from
 *                      __match_any_sync ranks, across tiles from the warp
 *                      processing them in order.
 *   hits_walk_kernel   one warp per pair: the sequential part (mask).  All
 *                      lanes follow the same walk (no divergence, lane 0
 *                      writes); together they skip marked candidates 32 at a
 *                      time.
 */
#ifndef SA_HITS_CUH
#define SA_HITS_CUH

#include "sa_platform.h"
#include "sa_kernels.cuh"

namespace sa {

constexpr int HITS_WARPS = 4;
constexpr int HITS_BINS = 2048;          /* counters per warp (shared memory: 4 x 8 KB) */
constexpr int HITS_DIGIT_BITS = 11;

struct HitsArgs {
  const uint8_t *seq_a, *seq_b;
  const int64_t *off_a, *off_b;          /* already shifted to the wave's first pair */
  int64_t npairs;
  ScoreParams sp;
  const int32_t *sub;                    /* [cb*ncodes+ca], for the penalty of a match step */
  const uint8_t *lut;
  const uint8_t *dir;                    /* flag bytes   [pair][lb][stride]            */
  const int16_t *m16;                    /* match scores, same offsets (elements)      */
  const int64_t *dir_off;
  unsigned long long *keys0, *keys1;     /* candidate keys, ping-pong, dir_off elements */
  int32_t *ncand;                        /* per pair                                   */
  int32_t *which;                        /* per pair: 0/1 = buffer holding the sorted keys */
  unsigned *mask;                        /* visited bits, dir_off/32 words per pair    */
  int32_t min_score, max_hits;
  /* outputs */
  int32_t *nhits;                        /* per pair                                   */
  int32_t *rec;                          /* [pair][max_hits][8]: score,pos_a,pos_b,len_a,len_b,aln_len,aln_start,- */
  uint8_t *out_a, *out_b;                /* [pair][max_hits][la+lb], offsets out_off   */
  const int64_t *out_off;
  unsigned long long *counter;
};

/* key: score << 32 | x << 16 | y  (x, y 1-based cell coordinates) */
__host__ __device__ __forceinline__ unsigned long long hit_key(int score, int x, int y)
{
  return ((unsigned long long)(unsigned)score << 32) | ((unsigned long long)(unsigned)x << 16) | (unsigned)y;
}

/* one stable counting-sort pass of n keys (src -> dst) by digit(key), done by
 * one warp; cnt has HITS_BINS counters; nbins <= HITS_BINS */
template <class DigitFn>
__device__ void warp_counting_pass(const unsigned long long *src, unsigned long long *dst, int n,
                                   unsigned *cnt, int nbins, int lane, DigitFn digit)
{
  for(int i = lane; i < nbins; i += 32) cnt[i] = 0;
  __syncwarp();
  for(int i = lane; i < n; i += 32) atomicAdd(&cnt[digit(src[i])], 1u);
  __syncwarp();
  /* exclusive prefix over the bins */
  unsigned run = 0;
  for(int b0 = 0; b0 < nbins; b0 += 32) {
    const int b = b0 + lane;
    const unsigned c = b < nbins ? cnt[b] : 0;
    unsigned inc = c;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(FULL, inc, o);
      if(lane >= o) inc += t;
    }
    if(b < nbins) cnt[b] = run + inc - c;
    run += __shfl_sync(FULL, inc, 31);
  }
  __syncwarp();
  /* stable scatter: tiles in order, lanes in order inside a tile */
  for(int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const bool valid = i < n;
    const unsigned long long k = valid ? src[i] : 0ull;
    const unsigned d = valid ? (unsigned)digit(k) : 0xffffffffu;
    const unsigned peers = __match_any_sync(FULL, d);
    const int leader = __ffs(peers) - 1;
    const unsigned rank = (unsigned)__popc(peers & ((1u << lane) - 1u));
    unsigned base = 0;
    if(valid && lane == leader) { base = cnt[d]; cnt[d] = base + (unsigned)__popc(peers); }
    base = __shfl_sync(FULL, base, leader);
    if(valid) dst[base + rank] = k;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(HITS_WARPS * 32)
hits_sort_kernel(const HitsArgs A)
{
  __shared__ unsigned s_cnt[HITS_WARPS][HITS_BINS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned *cnt = s_cnt[wib];
  for(;;) {
    unsigned long long t = 0;
    if(lane == 0) t = atomicAdd(A.counter, 1ull);
    t = __shfl_sync(FULL, t, 0);
    if(t >= (unsigned long long)A.npairs) break;
    const int64_t p = (int64_t)t;
    const int la = (int)(A.off_a[p + 1] - A.off_a[p]), lb = (int)(A.off_b[p + 1] - A.off_b[p]);
    const int stride = (int)dir_stride(la);
    const int16_t *m = A.m16 + A.dir_off[p];
    unsigned long long *k0 = A.keys0 + A.dir_off[p], *k1 = A.keys1 + A.dir_off[p];

    /* 1. compact the candidates, row-major (y asc, then x asc) */
    int n = 0, maxs = 0;
    const int64_t total = (int64_t)stride * lb;
    /* (four tiles of 32 cells per iteration: their loads are in flight together) */
    for(int64_t i0 = 0; i0 < total; i0 += 128) {
      int scs[4];
#pragma unroll
      for(int q = 0; q < 4; q++) {
        const int64_t i = i0 + 32 * q + lane;
        scs[q] = i < total ? (int)m[i] : 0;
      }
#pragma unroll
      for(int q = 0; q < 4; q++) {
        const int64_t i = i0 + 32 * q + lane;
        const int x = (int)(i % stride), y = (int)(i / stride);
        const int sc = (i < total && x < la) ? scs[q] : 0;
        const bool take = sc >= A.min_score && sc > 0;
        const unsigned b = __ballot_sync(FULL, take);
        if(take) k0[n + __popc(b & ((1u << lane) - 1u))] = hit_key(sc, x + 1, y + 1);
        n += __popc(b);
        maxs = imax(maxs, sc);
      }
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) maxs = imax(maxs, __shfl_xor_sync(FULL, maxs, o));
    __syncwarp();

    /* 2. stable by x ascending: (x asc, y asc) */
    unsigned long long *src = k0, *dst = k1;
    int passes = 0;
    if(n > 1) {
      if(la + 1 <= HITS_BINS) {
        warp_counting_pass(src, dst, n, cnt, la + 1, lane,
                           [](unsigned long long k) { return (unsigned)((k >> 16) & 0xffffu); });
        { unsigned long long *tmp = src; src = dst; dst = tmp; } passes++;
      } else {
        for(int sh = 0; sh < 16; sh += HITS_DIGIT_BITS) {
          warp_counting_pass(src, dst, n, cnt, HITS_BINS, lane,
                             [sh](unsigned long long k) { return (unsigned)((k >> (16 + sh)) & (HITS_BINS - 1)); });
          { unsigned long long *tmp = src; src = dst; dst = tmp; } passes++;
        }
      }
      /* 3. stable by score descending: digit = maxs - score, low digit first */
      const int range = maxs - (A.min_score > 1 ? A.min_score : 1) + 1;
      for(int sh = 0; (range - 1) >> sh; sh += HITS_DIGIT_BITS) {
        const int nb = imin(HITS_BINS, ((range - 1) >> sh) + 1);
        warp_counting_pass(src, dst, n, cnt, nb, lane,
                           [sh, maxs](unsigned long long k) {
                             return (unsigned)(((unsigned)maxs - (unsigned)(k >> 32)) >> sh) & (unsigned)(HITS_BINS - 1);
                           });
        { unsigned long long *tmp = src; src = dst; dst = tmp; } passes++;
        if(sh + HITS_DIGIT_BITS >= 31) break;
      }
    }
    if(lane == 0) { A.ncand[p] = n; A.which[p] = passes & 1; }
  }
}

/* One WARP per pair.  The walks of a pair are sequential by nature (each one
 * reads the marks its predecessors left), so the parallelism is across pairs;
 * a thread per pair made 32 diverging walks share one instruction stream and
 * left the SMs with a handful of warps, every dependent load exposed (62 ms for
 * 20k pairs).  Here all lanes of a warp run the SAME walk on the same addresses
 * (loads coalesce to one transaction, no divergence; lane 0 alone writes), the
 * machine holds thousands of such warps, and the lanes split the one thing
 * that is parallel: skipping candidates that are already marked, 32 at a time. */
constexpr int HITS_WALK_WARPS = 4;

__global__ void __launch_bounds__(HITS_WALK_WARPS * 32)
hits_walk_kernel(const HitsArgs A)
{
  const ScoreParams &sp = A.sp;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for(int64_t p = warp0; p < A.npairs; p += nwarps) {
    const int64_t oa = A.off_a[p], ob = A.off_b[p];
    const int la = (int)(A.off_a[p + 1] - oa), lb = (int)(A.off_b[p + 1] - ob);
    const uint8_t *a = A.seq_a + oa, *b = A.seq_b + ob;
    const int64_t stride = dir_stride(la);
    const uint8_t *dirp = A.dir + A.dir_off[p];
    unsigned *mask = A.mask + A.dir_off[p] / 32;
    const unsigned long long *keys = (A.which[p] ? A.keys1 : A.keys0) + A.dir_off[p];
    const int n = A.ncand[p];
    const int cap = la + lb;
    int nh = 0;

    int c = 0;
    while(c < n && nh < A.max_hits) {
      /* next candidate whose end cell is not marked yet (smith_waterman.c:270):
       * the lanes look at candidates c .. c+31 together */
      unsigned long long key = 0;
      bool open_cand = false;
      if(c + lane < n) {
        key = keys[c + lane];
        const int64_t cell = (int64_t)((int)(key & 0xffffu) - 1) * stride + ((int)((key >> 16) & 0xffffu) - 1);
        open_cand = !((mask[cell >> 5] >> (cell & 31)) & 1u);
      }
      const unsigned cand = __ballot_sync(FULL, open_cand);
      if(cand == 0) { c += 32; continue; }
      const int j = __ffs(cand) - 1;
      key = __shfl_sync(FULL, key, j);
      c += j + 1;   /* the candidates after it are looked at again: this walk may mark them */

      const int xe = (int)((key >> 16) & 0xffffu), ye = (int)(key & 0xffffu);
      const int score = (int)(key >> 32);
      uint8_t *ra = A.out_a + A.out_off[p] + (int64_t)nh * cap;
      uint8_t *rb = A.out_b + A.out_off[p] + (int64_t)nh * cap;
      int x = xe, y = ye, st = ST_M, cs = score, len = 0;
      bool ok = true;
      /* smith_waterman.c:187-199 and 217-244 in one pass: the strings are
       * written speculatively and only kept if the walk completes */
      for(;;) {
        const bool border = x == 0 || y == 0;
        int64_t cell = 0;
        if(!border) {
          cell = (int64_t)(y - 1) * stride + (x - 1);
          const unsigned word = mask[cell >> 5];
          if((word >> (cell & 31)) & 1u) { ok = false; break; }
          __syncwarp();                     /* every lane has read the word before lane 0 changes it */
          if(lane == 0) mask[cell >> 5] = word | (1u << (cell & 31));
          __syncwarp();
        }
        if(cs == 0) break;
        len++;
        /* everything this step may need is requested up front (independent loads) */
        const unsigned f = dirp[cell];
        const unsigned g_diag = (x > 1 && y > 1) ? dirp[cell - stride - 1] : 0u;
        const unsigned g_up = y > 1 ? dirp[cell - stride] : 0u;
        const unsigned g_left = x > 1 ? dirp[cell - 1] : 0u;
        const unsigned ca = a[x - 1], cb = b[y - 1];
        if(lane == 0) {
          ra[cap - len] = st == ST_GA ? '-' : (uint8_t)ca;
          rb[cap - len] = st == ST_GB ? '-' : (uint8_t)cb;
        }
        /* predecessor state from the equality flags (as walk_kernel, fmt 1) */
        int code;
        if(st == ST_M) {
          if(x == 1 || y == 1) code = ST_M;
          else code = !(g_diag & 1) ? ST_GA : !(g_diag & 2) ? ST_GB : ST_M;
        } else if(st == ST_GA) {
          if(!(f & 4)) code = ST_GA;
          else if(y == 1) code = ST_M;
          else code = !(g_up & 2) ? ST_GB : ST_M;
        } else {
          if(x == 1) code = ST_M;
          else code = (!(g_left & 1) && !(f & 16)) ? ST_GA : !(f & 8) ? ST_GB : ST_M;
        }
        int pen;
        if(st == ST_M) { pen = A.sub[A.lut[cb] * sp.ncodes + A.lut[ca]]; x--; y--; }
        else if(st == ST_GA) { pen = code == ST_GA ? sp.ext : sp.open; y--; }
        else { pen = code == ST_GB ? sp.ext : sp.open; x--; }
        cs -= pen;
        if(x == 0 || y == 0) cs = 0;
        st = code;
      }
      if(!ok) continue;
      if(lane == 0) {
        int32_t *r = A.rec + ((int64_t)p * A.max_hits + nh) * 8;
        r[0] = score; r[1] = x; r[2] = y; r[3] = xe - x; r[4] = ye - y; r[5] = len; r[6] = cap - len; r[7] = 0;
      }
      nh++;
    }
    if(lane == 0) A.nhits[p] = nh;
    __syncwarp();
  }
}

/* The same iteration with one THREAD per pair, written as a flat state machine: every pass of the one loop
 * does a scan step (look at the next candidate) or a walk step (one cell of the current walk) for the
 * thread's pair, so the 32 pairs of a warp share every instruction they issue instead of one pair owning the
 * warp.  ncu on the warp-per-pair kernel above showed what that costs: 69 % ALU-pipe and 72 % issue-slot
 * utilisation spent on 32 lanes computing one walk (profiles/ncu_r02p_all_kernels.csv).  The first
 * thread-per-pair version of round 1 lost to divergence because its nested loops (candidates / walk) let the
 * lanes of a warp wait for each other; here a lane that finishes a walk goes straight on with its next
 * candidate in the same loop.  Marked candidates are skipped one per pass instead of 32 at a time -- a
 * few instructions each.  Same rules, same outputs as hits_walk_kernel (smith_waterman.c:165-277). */
constexpr int HITS_FLAT_THREADS = 64;   /* small CTAs: the pairs of a batch spread over every SM */

__global__ void __launch_bounds__(HITS_FLAT_THREADS)
hits_walk_flat_kernel(const HitsArgs A)
{
  const ScoreParams &sp = A.sp;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  enum { PH_LOAD, PH_SCAN, PH_WALK };
  int phase = PH_LOAD;
  /* the pair */
  const uint8_t *a = nullptr, *b = nullptr, *dirp = nullptr;
  unsigned *mask = nullptr;
  const unsigned long long *keys = nullptr;
  uint8_t *ra = nullptr, *rb = nullptr;
  int64_t stride = 0;
  int n = 0, cap = 0, c = 0, nh = 0;
  /* the walk */
  int xe = 0, ye = 0, score = 0, x = 0, y = 0, st = ST_M, cs = 0, len = 0;

  while(p < A.npairs) {
    if(phase == PH_LOAD) {
      const int64_t oa = A.off_a[p], ob = A.off_b[p];
      const int la = (int)(A.off_a[p + 1] - oa), lb = (int)(A.off_b[p + 1] - ob);
      a = A.seq_a + oa; b = A.seq_b + ob;
      stride = dir_stride(la);
      dirp = A.dir + A.dir_off[p];
      mask = A.mask + A.dir_off[p] / 32;
      keys = (A.which[p] ? A.keys1 : A.keys0) + A.dir_off[p];
      n = A.ncand[p];
      cap = la + lb;
      c = 0; nh = 0;
      phase = PH_SCAN;
    }
    if(phase == PH_SCAN) {
      if(c >= n || nh >= A.max_hits) {
        A.nhits[p] = nh;
        p += nthreads;
        phase = PH_LOAD;
        continue;
      }
      /* next candidate whose end cell is not marked yet (smith_waterman.c:270) */
      const unsigned long long key = keys[c++];
      xe = (int)((key >> 16) & 0xffffu); ye = (int)(key & 0xffffu);
      const int64_t cell = (int64_t)(ye - 1) * stride + (xe - 1);
      if((mask[cell >> 5] >> (cell & 31)) & 1u) continue;
      score = (int)(key >> 32);
      ra = A.out_a + A.out_off[p] + (int64_t)nh * cap;
      rb = A.out_b + A.out_off[p] + (int64_t)nh * cap;
      x = xe; y = ye; st = ST_M; cs = score; len = 0;
      phase = PH_WALK;
    }
    /* one cell of the walk (smith_waterman.c:187-199 and 217-244 in one pass: the strings are written
     * speculatively and only kept if the walk completes) */
    int64_t cell = 0;
    if(x != 0 && y != 0) {
      cell = (int64_t)(y - 1) * stride + (x - 1);
      const unsigned word = mask[cell >> 5];
      if((word >> (cell & 31)) & 1u) { phase = PH_SCAN; continue; }   /* ran into a used cell: no hit, its marks stay */
      mask[cell >> 5] = word | (1u << (cell & 31));
    }
    if(cs == 0) {
      int32_t *r = A.rec + ((int64_t)p * A.max_hits + nh) * 8;
      r[0] = score; r[1] = x; r[2] = y; r[3] = xe - x; r[4] = ye - y; r[5] = len; r[6] = cap - len; r[7] = 0;
      nh++;
      phase = PH_SCAN;
      continue;
    }
    len++;
    const unsigned f = dirp[cell];
    const unsigned g_diag = (x > 1 && y > 1) ? dirp[cell - stride - 1] : 0u;
    const unsigned g_up = y > 1 ? dirp[cell - stride] : 0u;
    const unsigned g_left = x > 1 ? dirp[cell - 1] : 0u;
    const unsigned ca = a[x - 1], cb = b[y - 1];
    ra[cap - len] = st == ST_GA ? '-' : (uint8_t)ca;
    rb[cap - len] = st == ST_GB ? '-' : (uint8_t)cb;
    /* predecessor state from the equality flags (as walk_kernel, fmt 1) */
    int code;
    if(st == ST_M) {
      if(x == 1 || y == 1) code = ST_M;
      else code = !(g_diag & 1) ? ST_GA : !(g_diag & 2) ? ST_GB : ST_M;
    } else if(st == ST_GA) {
      if(!(f & 4)) code = ST_GA;
      else if(y == 1) code = ST_M;
      else code = !(g_up & 2) ? ST_GB : ST_M;
    } else {
      if(x == 1) code = ST_M;
      else code = (!(g_left & 1) && !(f & 16)) ? ST_GA : !(f & 8) ? ST_GB : ST_M;
    }
    int pen;
    if(st == ST_M) { pen = A.sub[A.lut[cb] * sp.ncodes + A.lut[ca]]; x--; y--; }
    else if(st == ST_GA) { pen = code == ST_GA ? sp.ext : sp.open; y--; }
    else { pen = code == ST_GB ? sp.ext : sp.open; x--; }
    cs -= pen;
    if(x == 0 || y == 0) cs = 0;
    st = code;
  }
}

} // namespace sa

#endif
