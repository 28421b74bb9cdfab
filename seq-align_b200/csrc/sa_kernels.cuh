/*
 * sa_kernels.cuh -- sm_100a kernels of the batch alignment engine (part 1):
 * alphabet scan, the general three-matrix fill (all scoring flags, all
 * modes) and the traceback walk.
 *
 * Conventions (same as the reference, src/alignment_macros.h:11,
 * src/alignment.c:178): x in [0,len_a] indexes seq_a = columns, y in
 * [0,len_b] indexes seq_b = rows; cell (x,y) depends on (x-1,y-1) [match],
 * (x,y-1) [gap_a] and (x-1,y) [gap_b].
 *
 * Parallel shape: a warp sweeps a strip of 32*GK columns.  Lane l owns GK
 * consecutive columns and, at step s, computes them for row y = s-l+1: the
 * lanes form an anti-diagonal wavefront, the right-most column of lane l-1
 * reaches lane l through one warp shuffle per matrix per step, everything
 * else lives in registers.  Pairs wider than a strip are strip-mined; the
 * right edge of a strip is parked in a per-warp boundary buffer.
 */
#ifndef SA_KERNELS_CUH
#define SA_KERNELS_CUH

#include "sa_platform.h"
#include <mutex>
#include <vector>

namespace sa {

/* Opt-in to more than 48 KB of dynamic shared memory.  The attribute belongs to the FUNCTION (per
 * device), not to the calling thread: engines on several host threads launch the same kernels with
 * different sizes, so the limit is only ever raised, under one process-wide lock.  (A per-thread
 * record of what had been opted in let one thread lower the limit under another thread's launch:
 * "invalid argument" on hardware, invisible in the emulator.) */
template <class KF>
inline bool smem_opt_in(KF kfn, size_t smem)
{
  struct Opted { const void *fn; int device; size_t smem; };
  static std::mutex mu;
  static std::vector<Opted> opted;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for(Opted &o : opted) {
    if(o.fn != (const void *)kfn || o.device != dev) continue;
    if(o.smem >= smem) return true;
    if(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return false; }
    o.smem = smem;
    return true;
  }
  if(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return false; }
  opted.push_back({(const void *)kfn, dev, smem});
  return true;
}


constexpr unsigned FULL = 0xffffffffu;
constexpr int GK = 8;            /* columns per lane, general kernel */
constexpr int GSTRIP = 32 * GK;  /* columns per strip */
constexpr int GEN_WARPS = 4;     /* warps per CTA, general kernel */
constexpr int SMEM_TABLE_MAX_CODES = 64;

enum { ST_M = 0, ST_GA = 1, ST_GB = 2, ST_FAIL = 3 };
enum { MODE_SCORE = 0, MODE_DIR = 1, MODE_MATS = 2 };
enum { WALK_OK = 0, WALK_NOHIT = 1, WALK_FAIL = 2 };

/* scoring model as the kernels see it (host flattens scoring_t into this
 * plus a dense ncodes x ncodes table) */
struct ScoreParams {
  int open;      /* gap_open + gap_extend: first gap position   (alignment.c:38) */
  int ext;       /* gap_extend                                  (alignment.c:39) */
  int gap_open;  /* raw gap_open, for the borders            (alignment.c:67-78) */
  int minv;      /* 0 for SW, INT_MIN+|min_penalty| for NW      (alignment.c:41) */
  int is_sw;
  int no_start, no_end, no_gaps_a, no_gaps_b, no_mismatches;
  int ncodes;
};

/* meta block written by scan_kernel, read back by the host (16 x u64) */
enum { META_PRES_A = 0, META_PRES_B = 4, META_MAX_LA = 8, META_MAX_LB = 9,
       META_CELLS = 10, META_MAX_CELLS = 11, META_MIN_LA = 12, META_MIN_LB = 13, META_WORDS = 16 };

/* offsets of a batch whose sequences all have the same length: off[i] = i*len.
 * Such a batch (every BASELINE config) does not ship its offset arrays over
 * PCIe -- 16 bytes per pair, 5 % of a 150 bp batch -- they are made here. */
__global__ void __launch_bounds__(256)
uniform_offsets_kernel(int64_t *__restrict__ off_a, int64_t *__restrict__ off_b, int64_t n, int64_t la, int64_t lb)
{
  for(int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
    off_a[i] = i * la;
    off_b[i] = i * lb;
  }
}

/* ---------------------------------------------------------------------------
 * scan_kernel: which byte values occur in seq_a / seq_b (256-bit sets), the
 * longest sequences and the cell count.  The host needs the alphabet to turn
 * scoring_t (256x256) into a dense table small enough for shared memory.
 * HBM-bound streaming read.
 */
__global__ void __launch_bounds__(256)
scan_kernel(const uint8_t *__restrict__ seq_a_base, const uint8_t *__restrict__ seq_b_base,
            const int64_t *__restrict__ off_a, const int64_t *__restrict__ off_b,
            int64_t npairs, unsigned long long *__restrict__ meta)
{
  /* the pairs' bytes are seq_*_base[off[0] .. off[npairs]) */
  const uint8_t *seq_a = seq_a_base + off_a[0], *seq_b = seq_b_base + off_b[0];
  const int64_t total_a = off_a[npairs] - off_a[0], total_b = off_b[npairs] - off_b[0];
  __shared__ unsigned s_seen[2][256];
  __shared__ unsigned long long s_red[6];
  for(int i = threadIdx.x; i < 512; i += blockDim.x) (&s_seen[0][0])[i] = 0;
  if(threadIdx.x < 4) s_red[threadIdx.x] = 0;
  if(threadIdx.x >= 4 && threadIdx.x < 6) s_red[threadIdx.x] = ~0ull;
  __syncthreads();

  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for(int which = 0; which < 2; which++) {
    const uint8_t *seq = which ? seq_b : seq_a;
    const int64_t total = which ? total_b : total_a;
    /* scalar head up to 16-byte alignment, 16-byte vector body, scalar tail */
    int64_t head = (int64_t)((16 - ((uintptr_t)seq & 15)) & 15);
    if(head > total) head = total;
    for(int64_t i = tid; i < head; i += nthreads) s_seen[which][seq[i]] = 1;
    const int64_t nvec = (total - head) / 16;
    const uint4 *v = (const uint4 *)(seq + head);
    for(int64_t i = tid; i < nvec; i += nthreads) {
      uint4 w = v[i];
      unsigned words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for(int k = 0; k < 4; k++) {
        s_seen[which][words[k] & 0xff] = 1;
        s_seen[which][(words[k] >> 8) & 0xff] = 1;
        s_seen[which][(words[k] >> 16) & 0xff] = 1;
        s_seen[which][words[k] >> 24] = 1;
      }
    }
    for(int64_t i = head + nvec * 16 + tid; i < total; i += nthreads) s_seen[which][seq[i]] = 1;
  }

  unsigned long long max_la = 0, max_lb = 0, cells = 0, max_cells = 0, min_la = ~0ull, min_lb = ~0ull;
  for(int64_t p = tid; p < npairs; p += nthreads) {
    unsigned long long la = (unsigned long long)(off_a[p + 1] - off_a[p]);
    unsigned long long lb = (unsigned long long)(off_b[p + 1] - off_b[p]);
    max_la = la > max_la ? la : max_la;
    max_lb = lb > max_lb ? lb : max_lb;
    min_la = la < min_la ? la : min_la;
    min_lb = lb < min_lb ? lb : min_lb;
    cells += la * lb;
    max_cells = la * lb > max_cells ? la * lb : max_cells;
  }
  atomicMax(&s_red[0], max_la);
  atomicMax(&s_red[1], max_lb);
  atomicAdd(&s_red[2], cells);
  atomicMax(&s_red[3], max_cells);
  atomicMin(&s_red[4], min_la);
  atomicMin(&s_red[5], min_lb);
  __syncthreads();

  if(threadIdx.x < 8) {
    /* 256 flags -> 4 x u64 per sequence set */
    const int which = threadIdx.x >> 2, word = threadIdx.x & 3;
    unsigned long long bits = 0;
    for(int i = 0; i < 64; i++)
      if(s_seen[which][word * 64 + i]) bits |= 1ull << i;
    if(bits) atomicOr(&meta[(which ? META_PRES_B : META_PRES_A) + word], bits);
  }
  if(threadIdx.x == 0) {
    atomicMax(&meta[META_MAX_LA], s_red[0]);
    atomicMax(&meta[META_MAX_LB], s_red[1]);
    atomicAdd(&meta[META_CELLS], s_red[2]);
    atomicMax(&meta[META_MAX_CELLS], s_red[3]);
    atomicMin(&meta[META_MIN_LA], s_red[4]);
    atomicMin(&meta[META_MIN_LB], s_red[5]);
  }
}

/* ---------------------------------------------------------------------------
 * general fill kernel
 */
struct GenArgs {
  const uint8_t *seq_a, *seq_b;
  const int64_t *off_a, *off_b;
  int64_t pair0, npairs;       /* this launch covers pairs [pair0, pair0+npairs) */
  ScoreParams sp;
  const int32_t *sub;          /* [cb*ncodes + ca] substitution score       */
  const uint8_t *forbid;       /* same shape: no_mismatches && !is_match    */
  const uint8_t *lut;          /* byte -> dense code                        */
  int table_in_smem;
  unsigned long long *counter; /* dynamic pair scheduler                    */
  int4 *bnd;                   /* strip boundary rows, per warp             */
  int64_t bnd_rows;
  int32_t *score, *xend, *yend, *state; /* indexed by absolute pair id      */
  uint8_t *dir;                /* MODE_DIR: 1 byte/cell direction codes     */
  const int64_t *dir_off;      /* per pair (relative to pair0), bytes       */
  int32_t *mat_m, *mat_ga, *mat_gb;     /* MODE_MATS: one pair, pitched     */
  int64_t pitch;               /* ints per row; element (x,y) at y*pitch+3+x */
};

__host__ __device__ __forceinline__ int64_t dir_stride(int la) { return ((int64_t)la + 15) & ~(int64_t)15; }

struct Cell3 { int m, ga, gb; };

/* better-hit comparator of the reference's hit order
 * (smith_waterman.c:71-86 + stable sort): score desc, x asc, y asc */
__host__ __device__ __forceinline__ bool hit_better(int v, int x, int y, int v2, int x2, int y2)
{
  return v > v2 || (v == v2 && (x < x2 || (x == x2 && y < y2)));
}

/* per-warp partial result of a pair (a CTA's warps are merged in shared memory) */
struct PairPart { int bestV, bestX, bestY, has_fin, fm, fga, fgb; };

__device__ __forceinline__ int4 ld_cg_int4(const int4 *p)
{
#if defined(__CUDA_ARCH__)
  return __ldcg(p);   /* L2: the row may have been written by another warp of this CTA */
#else
  return *p;
#endif
}

/*
 * One pair, swept by W warps of a CTA (W = 1: a warp on its own).  Warp w
 * takes the column strips w, w+W, ...; strip s reads the right edge of strip
 * s-1 from the boundary slot (s-1)%W and writes its own into slot s%W.  With
 * W > 1 the warps form a systolic pipeline: the consumer of a strip waits on
 * a progress word in shared memory that the producer bumps every 32 rows.
 */
template <int MODE>
__device__ void general_pair(const GenArgs &A, const int64_t p, const int lane,
                             const int32_t *T, const uint8_t *F, const uint8_t *L,
                             int4 *my_bnd, int4 *s_chunk,
                             const int w, const int W, volatile unsigned long long *progress,
                             PairPart *part)
{
  const ScoreParams &sp = A.sp;
  const int64_t oa = A.off_a[p], ob = A.off_b[p];
  const int la = (int)(A.off_a[p + 1] - oa), lb = (int)(A.off_b[p + 1] - ob);
  const uint8_t *pa = A.seq_a + oa, *pb = A.seq_b + ob;
  const int n = sp.ncodes;
  const int minv = sp.minv;

  uint8_t *dirp = nullptr;
  int64_t dstride = 0;
  if(MODE == MODE_DIR) {
    dirp = A.dir + A.dir_off[p - A.pair0];
    dstride = dir_stride(la);
  }

  if(MODE == MODE_MATS && w == 0) {
    /* borders: row 0 and column 0 (alignment.c:47-81) */
    for(int x = lane; x <= la; x += 32) {
      int m = minv, ga = minv, gb = sp.no_start ? 0 : addw(sp.gap_open, x * sp.ext);
      if(sp.is_sw || x == 0) m = ga = gb = 0;
      A.mat_m[3 + x] = m; A.mat_ga[3 + x] = ga; A.mat_gb[3 + x] = gb;
    }
    for(int y = 1 + lane; y <= lb; y += 32) {
      int m = minv, gb = minv, ga = sp.no_start ? 0 : addw(sp.gap_open, y * sp.ext);
      if(sp.is_sw) m = ga = gb = minv;
      A.mat_m[y * A.pitch + 3] = m; A.mat_ga[y * A.pitch + 3] = ga; A.mat_gb[y * A.pitch + 3] = gb;
    }
  }

  /* running best hit of this lane (SW) and the final cell (NW) */
  int bestV = 0, bestX = 0, bestY = 0, has_fin = 0;
  Cell3 fin = {0, 0, 0};

  const int nstrips = (la > 0 && lb > 0) ? (la + GSTRIP - 1) / GSTRIP : 0;
  for(int strip = w; strip < nstrips; strip += W) {
    const int x0 = strip * GSTRIP;
    const int xf = x0 + lane * GK + 1; /* my first column, 1-based */
    const bool more = strip + 1 < nstrips;
    int4 *out_bnd = my_bnd + (int64_t)(strip % W) * A.bnd_rows;
    const int4 *in_bnd = my_bnd + (int64_t)((strip + W - 1) % W) * A.bnd_rows;
    const unsigned long long in_base = (unsigned long long)(strip - 1) * (unsigned long long)(lb + 1);

    int ac[GK], uM[GK], uGA[GK], uGB[GK], colV[GK], colY[GK];
#pragma unroll
    for(int j = 0; j < GK; j++) {
      const int x = xf + j;
      ac[j] = x <= la ? L[pa[x - 1]] : 0;
      /* row 0 (alignment.c:51-69) */
      uM[j] = uGA[j] = sp.is_sw ? 0 : minv;
      uGB[j] = sp.is_sw ? 0 : (sp.no_start ? 0 : addw(sp.gap_open, x * sp.ext));
      colV[j] = 0; colY[j] = 0;
    }
    /* diagonal predecessor of my first column in row 1: cell (xf-1, 0) */
    Cell3 dg;
    if(xf - 1 == 0 || sp.is_sw) { dg.m = dg.ga = dg.gb = 0; }
    else { dg.m = dg.ga = minv; dg.gb = sp.no_start ? 0 : addw(sp.gap_open, (xf - 1) * sp.ext); }

    Cell3 out = {0, 0, 0};
    const int nsteps = lb + 31;
    for(int s = 0; s < nsteps; s++) {
      const int y = s - lane + 1;
      const bool active = y >= 1 && y <= lb;

      if(x0 > 0 && (s & 31) == 0) {
        /* next 32 rows of the previous strip's right edge -> shared */
        __syncwarp();
        const int row = s + 1 + lane;
        if(W > 1) {
          /* wait until the producer strip has published these 32 rows */
          const int need = s + 32 < lb ? s + 32 : lb;
          while(progress[(strip + W - 1) % W] < in_base + (unsigned long long)need) SA_SPIN_HINT();
          __threadfence_block();
        }
        int4 v = make_int4(0, 0, 0, 0);
        if(row <= lb) v = W > 1 ? ld_cg_int4(&in_bnd[row]) : in_bnd[row];
        s_chunk[lane] = v;
        __syncwarp();
      }

      /* left neighbour (xf-1, y) */
      Cell3 lf;
      lf.m = __shfl_up_sync(FULL, out.m, 1);
      lf.ga = __shfl_up_sync(FULL, out.ga, 1);
      lf.gb = __shfl_up_sync(FULL, out.gb, 1);
      if(lane == 0) {
        if(x0 == 0) {
          /* column 0 (alignment.c:55-56, 72-80) */
          lf.m = lf.gb = minv;
          lf.ga = sp.is_sw ? minv : (sp.no_start ? 0 : addw(sp.gap_open, y * sp.ext));
        } else {
          const int4 v = s_chunk[s & 31];
          lf.m = v.x; lf.ga = v.y; lf.gb = v.z;
        }
      }
      const Cell3 lf_in = lf;

      if(active) {
        const int cb = L[pb[y - 1]];
        const int32_t *Trow = T + cb * n;
        const uint8_t *Frow = F + cb * n;
        /* gap_b penalties of this row (alignment.c:140-155) and of the
         * reverse move out of GAP_B (alignment.c:265-268) */
        const bool lastrow = y == lb;
        const bool freeB = lastrow && sp.no_end;
        const int oB = freeB ? 0 : sp.open, eB = freeB ? 0 : sp.ext;
        const bool disB = sp.no_gaps_b && !lastrow;
        Cell3 d = dg;
        unsigned codes_lo = 0, codes_hi = 0;
        int4 sm[GK / 4], sga[GK / 4], sgb[GK / 4];
#pragma unroll
        for(int j = 0; j < GK; j++) {
          const int x = xf + j;
          const bool lastcol = x == la;
          const bool freeA = lastcol && sp.no_end;
          const int oA = freeA ? 0 : sp.open, eA = freeA ? 0 : sp.ext;
          const bool disA = sp.no_gaps_a && !lastcol;
          const int sub = Trow[ac[j]];
          const bool fb = sp.no_mismatches && Frow[ac[j]];

          /* match (alignment.c:101-116) */
          const int m = fb ? minv : imax(addw(max3(d.m, d.ga, d.gb), sub), minv);
          /* gap in a, predecessor above (alignment.c:122-137) */
          const int ga = disA ? minv
                       : imax(imax(addw(imax(uM[j], uGB[j]), oA), addw(uGA[j], eA)), minv);
          /* gap in b, predecessor left (alignment.c:140-155) */
          const int gb = disB ? minv
                       : imax(imax(addw(imax(lf.m, lf.ga), oB), addw(lf.gb, eB)), minv);

          if(MODE == MODE_DIR) {
            /* the choice alignment_reverse_move (alignment.c:311-327) would
             * make when it leaves this cell in each of the three states,
             * taken now while all operands are in registers */
            const bool okA_x1 = !sp.no_gaps_a || x == 1;    /* x' = x-1 */
            const bool okA_x = !sp.no_gaps_a || lastcol;    /* x' = x   */
            const bool okB_y1 = !sp.no_gaps_b || y == 1;    /* y' = y-1 */
            const bool okB_y = !sp.no_gaps_b || lastrow;    /* y' = y   */
            unsigned cm = ST_FAIL, cga = ST_FAIL, cgb = ST_FAIL;
            if(okA_x1 && addw(d.ga, sub) == m) cm = ST_GA;
            else if(okB_y1 && addw(d.gb, sub) == m) cm = ST_GB;
            else if(addw(d.m, sub) == m) cm = ST_M;
            if(okA_x && addw(uGA[j], eA) == ga) cga = ST_GA;
            else if(okB_y1 && addw(uGB[j], oA) == ga) cga = ST_GB;
            else if(addw(uM[j], oA) == ga) cga = ST_M;
            if(okA_x1 && addw(lf.ga, oB) == gb) cgb = ST_GA;
            else if(okB_y && addw(lf.gb, eB) == gb) cgb = ST_GB;
            else if(addw(lf.m, oB) == gb) cgb = ST_M;
            const unsigned code = cm | (cga << 2) | (cgb << 4);
            if(j < 4) codes_lo |= code << (8 * j);
            else codes_hi |= code << (8 * (j - 4));
          }
          if(MODE == MODE_MATS) {
            int *qm = &sm[j / 4].x, *qa = &sga[j / 4].x, *qb = &sgb[j / 4].x;
            qm[j & 3] = m; qa[j & 3] = ga; qb[j & 3] = gb;
          }
          if(MODE == MODE_SCORE) {
            if(sp.is_sw && x <= la && m > colV[j]) { colV[j] = m; colY[j] = y; }
          }
          d.m = uM[j]; d.ga = uGA[j]; d.gb = uGB[j];
          uM[j] = m; uGA[j] = ga; uGB[j] = gb;
          lf.m = m; lf.ga = ga; lf.gb = gb;
        }
        out = lf;

        if(MODE == MODE_DIR) {
          if(xf - 1 < dstride)
            *(uint2 *)(dirp + (int64_t)(y - 1) * dstride + (xf - 1)) = make_uint2(codes_lo, codes_hi);
        }
        if(MODE == MODE_MATS) {
          const int64_t base = (int64_t)y * A.pitch + 3 + xf;
          if(xf + GK - 1 <= la) {
#pragma unroll
            for(int q = 0; q < GK / 4; q++) {
              *(int4 *)(A.mat_m + base + 4 * q) = sm[q];
              *(int4 *)(A.mat_ga + base + 4 * q) = sga[q];
              *(int4 *)(A.mat_gb + base + 4 * q) = sgb[q];
            }
          } else {
#pragma unroll
            for(int j = 0; j < GK; j++)
              if(xf + j <= la) {
                A.mat_m[base + j] = (&sm[j / 4].x)[j & 3];
                A.mat_ga[base + j] = (&sga[j / 4].x)[j & 3];
                A.mat_gb[base + j] = (&sgb[j / 4].x)[j & 3];
              }
          }
        }
        if(more && lane == 31) {
          out_bnd[y] = make_int4(out.m, out.ga, out.gb, 0);
          if(W > 1 && ((y & 31) == 0 || y == lb)) {
            __threadfence_block();
            progress[strip % W] = (unsigned long long)strip * (unsigned long long)(lb + 1) + (unsigned long long)y;
          }
        }
        /* next row's diagonal predecessor is this row's left neighbour */
        dg = lf_in;
      }
    }

    /* fold this strip's columns into the lane's running result */
    if(MODE == MODE_SCORE && sp.is_sw) {
#pragma unroll
      for(int j = 0; j < GK; j++)
        if(hit_better(colV[j], xf + j, colY[j], bestV, bestX, bestY) && colV[j] > 0) {
          bestV = colV[j]; bestX = xf + j; bestY = colY[j];
        }
    }
    if(!more) {
      const int jf = (la - 1 - x0) % GK;
#pragma unroll
      for(int j = 0; j < GK; j++)
        if(j == jf) { fin.m = uM[j]; fin.ga = uGA[j]; fin.gb = uGB[j]; }
      has_fin = 1;
    }
    __syncwarp();
  }

  /* ---- this warp's share of the result ---- */
  if(MODE == MODE_SCORE && sp.is_sw) {
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) {
      const int v2 = __shfl_xor_sync(FULL, bestV, o);
      const int x2 = __shfl_xor_sync(FULL, bestX, o);
      const int y2 = __shfl_xor_sync(FULL, bestY, o);
      if(hit_better(v2, x2, y2, bestV, bestX, bestY)) { bestV = v2; bestX = x2; bestY = y2; }
    }
  }
  {
    /* the final cell (la, lb) lives in the lane that owns column la of the last strip */
    const int x0 = nstrips > 0 ? (nstrips - 1) * GSTRIP : 0;
    const int lf_lane = nstrips > 0 ? (la - 1 - x0) / GK : 0;
    fin.m = __shfl_sync(FULL, fin.m, lf_lane);
    fin.ga = __shfl_sync(FULL, fin.ga, lf_lane);
    fin.gb = __shfl_sync(FULL, fin.gb, lf_lane);
  }
  part->bestV = bestV; part->bestX = bestX; part->bestY = bestY;
  part->has_fin = has_fin; part->fm = fin.m; part->fga = fin.ga; part->fgb = fin.gb;
}

/* merge the warps' partial results and write the pair's outputs (one lane) */
template <int MODE>
__device__ void general_finish(const GenArgs &A, const int64_t p, const PairPart *parts, const int W)
{
  const ScoreParams &sp = A.sp;
  if(MODE == MODE_MATS || (sp.is_sw && MODE == MODE_DIR)) return;
  const int la = (int)(A.off_a[p + 1] - A.off_a[p]), lb = (int)(A.off_b[p + 1] - A.off_b[p]);
  const int minv = sp.minv;
  int score = 0, xe = 0, ye = 0, st = ST_M;
  if(sp.is_sw) {
    int bv = 0, bx = 0, by = 0;
    for(int i = 0; i < W; i++)
      if(hit_better(parts[i].bestV, parts[i].bestX, parts[i].bestY, bv, bx, by)) {
        bv = parts[i].bestV; bx = parts[i].bestX; by = parts[i].bestY;
      }
    score = bv; xe = bx; ye = by;
  } else {
    Cell3 fin = {0, 0, 0};
    if(la > 0 && lb > 0) {
      for(int i = 0; i < W; i++)
        if(parts[i].has_fin) { fin.m = parts[i].fm; fin.ga = parts[i].fga; fin.gb = parts[i].fgb; }
    } else if(la == 0 && lb == 0) {
      fin.m = fin.ga = fin.gb = 0;
    } else if(lb == 0) {
      fin.m = fin.ga = minv;
      fin.gb = sp.no_start ? 0 : addw(sp.gap_open, la * sp.ext);
    } else {
      fin.m = fin.gb = minv;
      fin.ga = sp.no_start ? 0 : addw(sp.gap_open, lb * sp.ext);
    }
    /* end state: GAP_A over GAP_B over MATCH on ties (needleman_wunsch.c:53-66) */
    score = fin.m; st = ST_M;
    if(fin.gb >= score) { score = fin.gb; st = ST_GB; }
    if(fin.ga >= score) { score = fin.ga; st = ST_GA; }
    xe = la; ye = lb;
  }
  A.score[p] = score;
  if(A.xend) A.xend[p] = xe;
  if(A.yend) A.yend[p] = ye;
  if(A.state) A.state[p] = st;
}

template <int MODE>
__global__ void __launch_bounds__(GEN_WARPS * 32)
general_kernel(const GenArgs A)
{
  unsigned char *dsm = SA_DYN_SMEM();
  const int n = A.sp.ncodes;
  uint8_t *s_lut = dsm;
  int4 *s_chunks = (int4 *)(dsm + 256);
  int32_t *s_sub = (int32_t *)(dsm + 256 + GEN_WARPS * 32 * sizeof(int4));
  uint8_t *s_forbid = (uint8_t *)(s_sub + (A.table_in_smem ? n * n : 0));

  for(int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = A.lut[i];
  if(A.table_in_smem)
    for(int i = threadIdx.x; i < n * n; i += blockDim.x) { s_sub[i] = A.sub[i]; s_forbid[i] = A.forbid[i]; }
  __syncthreads();

  const int32_t *T = A.table_in_smem ? s_sub : A.sub;
  const uint8_t *F = A.table_in_smem ? s_forbid : A.forbid;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t slot = (int64_t)blockIdx.x * GEN_WARPS + wib;
  int4 *my_bnd = A.bnd ? A.bnd + slot * A.bnd_rows : nullptr;

  for(;;) {
    unsigned long long t = 0;
    if(lane == 0) t = atomicAdd(A.counter, 1ull);
    t = __shfl_sync(FULL, t, 0);
    if(t >= (unsigned long long)A.npairs) break;
    PairPart part;
    general_pair<MODE>(A, A.pair0 + (int64_t)t, lane, T, F, s_lut, my_bnd, s_chunks + wib * 32, 0, 1, nullptr, &part);
    if(lane == 0) general_finish<MODE>(A, A.pair0 + (int64_t)t, &part, 1);
  }
}

/* CTA-per-pair variant for wide pairs: COOP_WARPS warps sweep one pair as a
 * systolic pipeline of column strips (see general_pair) */
constexpr int COOP_WARPS = 8;

template <int MODE>
__global__ void __launch_bounds__(COOP_WARPS * 32)
general_coop_kernel(const GenArgs A)
{
  unsigned char *dsm = SA_DYN_SMEM();
  const int n = A.sp.ncodes;
  uint8_t *s_lut = dsm;
  int4 *s_chunks = (int4 *)(dsm + 256);
  int32_t *s_sub = (int32_t *)(dsm + 256 + COOP_WARPS * 32 * sizeof(int4));
  uint8_t *s_forbid = (uint8_t *)(s_sub + (A.table_in_smem ? n * n : 0));
  __shared__ unsigned long long s_progress[COOP_WARPS];
  __shared__ unsigned long long s_next;
  __shared__ PairPart s_parts[COOP_WARPS];

  for(int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = A.lut[i];
  if(A.table_in_smem)
    for(int i = threadIdx.x; i < n * n; i += blockDim.x) { s_sub[i] = A.sub[i]; s_forbid[i] = A.forbid[i]; }
  __syncthreads();

  const int32_t *T = A.table_in_smem ? s_sub : A.sub;
  const uint8_t *F = A.table_in_smem ? s_forbid : A.forbid;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int4 *cta_bnd = A.bnd + (int64_t)blockIdx.x * COOP_WARPS * A.bnd_rows;

  for(;;) {
    if(threadIdx.x == 0) s_next = atomicAdd(A.counter, 1ull);
    if(threadIdx.x < COOP_WARPS) s_progress[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long t = s_next;
    if(t >= (unsigned long long)A.npairs) break;
    general_pair<MODE>(A, A.pair0 + (int64_t)t, lane, T, F, s_lut, cta_bnd, s_chunks + wib * 32,
                       wib, COOP_WARPS, s_progress, &s_parts[wib]);
    __syncthreads();
    if(threadIdx.x == 0) general_finish<MODE>(A, A.pair0 + (int64_t)t, s_parts, COOP_WARPS);
    __syncthreads();
  }
}

inline size_t general_smem_bytes(int ncodes, bool table_in_smem, int warps = GEN_WARPS)
{
  size_t b = 256 + warps * 32 * sizeof(int4);
  if(table_in_smem) b += (size_t)ncodes * ncodes * 5;
  return b + 16;
}

/* ---------------------------------------------------------------------------
 * SW best cell from direction-mode scores is not available (MODE_DIR stores
 * no scores), so SW alignment runs MODE_SCORE first (score + end cell) and
 * MODE_DIR second; walk_kernel then follows the direction bytes.
 *
 * walk_kernel: one thread per pair.  Restates the loops of
 * needleman_wunsch.c:79-132 (NW) and smith_waterman.c:187-255 (first hit on
 * a fresh mask) over the direction bytes.  Strings are written right-aligned
 * into per-pair buffers of capacity la+lb; aln_start/aln_len say where.
 */
struct WalkArgs {
  const uint8_t *seq_a, *seq_b;
  const int64_t *off_a, *off_b;
  int64_t pair0, npairs;
  ScoreParams sp;
  const int32_t *sub;
  const uint8_t *lut;
  const int32_t *score, *xend, *yend, *state;
  const uint8_t *dir;
  const int64_t *dir_off;
  uint8_t *out_a, *out_b;
  const int64_t *out_off;      /* per pair (relative to pair0), la+lb each */
  int32_t *aln_start, *aln_len, *pos_a, *pos_b, *len_a, *len_b, *status; /* relative */
  int fmt;                     /* 0: resolved 2-bit codes (general kernel), 1: equality flags (fast kernel) */
};

/* the walk of one pair; get(cx, cy) returns the flag byte of cell (cx+1, cy+1);
 * only the `writer` thread stores results (the tiled kernel runs the walk
 * warp-uniformly so that all lanes can refill the window together) */
template <class Get>
__device__ __forceinline__ void walk_pair(const WalkArgs &A, const int64_t r, Get get, const bool writer)
{
  const ScoreParams &sp = A.sp;
  const int64_t p = A.pair0 + r;
  const int64_t oa = A.off_a[p], ob = A.off_b[p];
  const int la = (int)(A.off_a[p + 1] - oa), lb = (int)(A.off_b[p + 1] - ob);
  const uint8_t *a = A.seq_a + oa, *b = A.seq_b + ob;
  uint8_t *ra = A.out_a + A.out_off[r], *rb = A.out_b + A.out_off[r];
  const int cap = la + lb;
  int n = 0, status = WALK_OK;
  int x, y, st;

  /* predecessor state when leaving cell (x,y) in state st */
  auto next_state = [&](int x, int y, int st) -> int {
    const unsigned f = get(x - 1, y - 1);
    if(A.fmt == 0) return (f >> (2 * st)) & 3;
    /* equality flags (bit clear = equal), resolved in the order GAP_A, GAP_B,
     * MATCH of alignment.c:311-327.  A predecessor on the border ends the
     * walk whatever its state, so it needs no flags. */
    if(st == ST_M) {
      if(x == 1 || y == 1) return ST_M;
      const unsigned g = get(x - 2, y - 2);
      return !(g & 1) ? ST_GA : !(g & 2) ? ST_GB : ST_M;
    }
    if(st == ST_GA) {
      if(!(f & 4)) return ST_GA;
      if(y == 1) return ST_M;
      const unsigned g = get(x - 1, y - 2);
      return !(g & 2) ? ST_GB : ST_M;
    }
    if(x == 1) return ST_M;
    const unsigned g = get(x - 2, y - 1);
    if(!(g & 1) && !(f & 16)) return ST_GA;
    return !(f & 8) ? ST_GB : ST_M;
  };

  if(!sp.is_sw) {
    x = la; y = lb;
    if(A.fmt == 0) st = A.state[p];
    else if(la > 0 && lb > 0) {
      /* end state GAP_A over GAP_B over MATCH on ties (needleman_wunsch.c:53-66) */
      const unsigned f = get(la - 1, lb - 1);
      st = !(f & 1) ? ST_GA : !(f & 2) ? ST_GB : ST_M;
    } else st = ST_M;
    while(x > 0 && y > 0) {
      n++;
      if(writer) {
        ra[cap - n] = st == ST_GA ? '-' : a[x - 1];
        rb[cap - n] = st == ST_GB ? '-' : b[y - 1];
      }
      const int code = next_state(x, y, st);
      if(code == ST_FAIL) { status = WALK_FAIL; break; }
      if(st == ST_M) { x--; y--; } else if(st == ST_GA) y--; else x--;
      st = code;
    }
    if(status == WALK_OK) {
      for(; y > 0; y--) { n++; if(writer) { ra[cap - n] = '-'; rb[cap - n] = b[y - 1]; } }
      for(; x > 0; x--) { n++; if(writer) { ra[cap - n] = a[x - 1]; rb[cap - n] = '-'; } }
    }
    if(writer) { A.pos_a[r] = 0; A.pos_b[r] = 0; A.len_a[r] = la; A.len_b[r] = lb; }
  } else {
    const int xe = A.xend[p], ye = A.yend[p];
    int cs = A.score[p];
    x = xe; y = ye; st = ST_M;
    if(cs <= 0) status = WALK_NOHIT;
    while(cs > 0) {
      n++;
      if(writer) {
        ra[cap - n] = st == ST_GA ? '-' : a[x - 1];
        rb[cap - n] = st == ST_GB ? '-' : b[y - 1];
      }
      const int code = next_state(x, y, st);
      if(code == ST_FAIL) { status = WALK_FAIL; break; }
      /* the penalty the reference subtracts implicitly: it continues
       * with the predecessor's stored value (alignment.c:311-327) */
      int pen;
      if(st == ST_M) {
        pen = A.sub[A.lut[b[y - 1]] * sp.ncodes + A.lut[a[x - 1]]];
        x--; y--;
      } else if(st == ST_GA) {
        const bool fr = sp.no_end && x == la;
        pen = fr ? 0 : (code == ST_GA ? sp.ext : sp.open);
        y--;
      } else {
        const bool fr = sp.no_end && y == lb;
        pen = fr ? 0 : (code == ST_GB ? sp.ext : sp.open);
        x--;
      }
      cs = (int)((unsigned)cs - (unsigned)pen);
      if(x == 0 || y == 0) cs = 0;   /* the SW borders are all zero: the hit starts here */
      st = code;
    }
    if(writer) { A.pos_a[r] = x; A.pos_b[r] = y; A.len_a[r] = xe - x; A.len_b[r] = ye - y; }
  }
  if(writer) {
    A.aln_start[r] = cap - n;
    A.aln_len[r] = n;
    A.status[r] = status;
  }
}

/* one thread per pair, flag bytes read straight from global memory: many
 * small pairs */
__global__ void __launch_bounds__(128)
walk_kernel(const WalkArgs A)
{
  for(int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < A.npairs;
      r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = A.pair0 + r;
    const uint8_t *dirp = A.dir + A.dir_off[r];
    const int64_t dstride = dir_stride((int)(A.off_a[p + 1] - A.off_a[p]));
    walk_pair(A, r, [&](int cx, int cy) -> unsigned { return dirp[(int64_t)cy * dstride + cx]; }, true);
  }
}

/* one warp per pair, for long walks: the flag bytes up and left of the
 * current cell are fetched a 32 x 128 byte window at a time into shared
 * memory by all lanes (one round trip to HBM per ~32 steps instead of two
 * dependent ones per step); the walk itself runs warp-uniformly */
constexpr int WT_ROWS = 32, WT_COLS = 128, WT_WARPS = 4;

__global__ void __launch_bounds__(WT_WARPS * 32)
walk_tiled_kernel(const WalkArgs A)
{
  __shared__ uint4 s_win[WT_WARPS][WT_ROWS * WT_COLS / 16];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint8_t *win = (uint8_t *)s_win[wib];
  for(int64_t r = (int64_t)blockIdx.x * WT_WARPS + wib; r < A.npairs; r += (int64_t)gridDim.x * WT_WARPS) {
    const int64_t p = A.pair0 + r;
    const uint8_t *dirp = A.dir + A.dir_off[r];
    const int64_t dstride = dir_stride((int)(A.off_a[p + 1] - A.off_a[p]));
    int r0 = 0, c0 = 0, r1 = -1, c1 = -1;   /* window = rows [r0, r1] x columns [c0, c1] */
    auto get = [&](int cx, int cy) -> unsigned {
      if(cx < c0 || cx > c1 || cy < r0 || cy > r1) {
        /* refill with (cx, cy) in the bottom-right corner block: the walk only moves up and left */
        __syncwarp();
        r1 = cy; r0 = cy - (WT_ROWS - 1); if(r0 < 0) r0 = 0;
        c0 = ((cx >> 4) << 4) + 16 - WT_COLS; if(c0 < 0) c0 = 0;
        c1 = c0 + WT_COLS - 1;
#pragma unroll
        for(int i = 0; i < WT_ROWS / 4; i++) {
          const int row = r0 + i * 4 + (lane >> 3), col = c0 + 16 * (lane & 7);
          uint4 v = make_uint4(0, 0, 0, 0);
          if(row <= r1 && col < dstride) v = *(const uint4 *)(dirp + (int64_t)row * dstride + col);
          *(uint4 *)(win + (i * 4 + (lane >> 3)) * WT_COLS + 16 * (lane & 7)) = v;
        }
        __syncwarp();
      }
      return win[(cy - r0) * WT_COLS + (cx - c0)];
    };
    walk_pair(A, r, get, lane == 0);
    __syncwarp();
  }
}

} // namespace sa

#endif
