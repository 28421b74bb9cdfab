/*
 * sa_mats.cuh -- sm_100a kernels of the batch alignment engine (part 5):
 * Smith-Waterman in MATERIALISE mode for whole batches -- the literal
 * contract of aligner_align() (reference src/alignment.c:170-193): the three
 * int32 matrices match / gap_a / gap_b of every pair, (len_a+1) x (len_b+1)
 * each, borders included, row-major with index = y*(len_a+1) + x (reference
 * src/alignment_macros.h:11).  12 bytes per cell leave the SM: this is the one
 * mode of the path that the HBM roofline bounds (mode M:
 * 0.546 Tcells/s at the measured copy bandwidth), so the kernel is built
 * around its stores, not around its arithmetic.
 *
 * Work shape: one warp per pair, one ROW per step.  Lane l owns the columns
 * x = 32*j + l (j = 0..NB-1, x = 0 is the border column), so that every store
 * instruction of the warp writes 32 consecutive ints of one matrix row -- a
 * full 128-byte line straight from registers, no staging.  The price is the
 * dependency inside a row (gap_b runs along it); it is resolved exactly by a
 * max-plus prefix scan:
 *
 *   T[x]    = max(M[x], GA[x])                      (both need only row y-1)
 *   GB[x]   = max(0, max over x' < x of  T[x'] + open + (x-1-x')*ext)
 *           = max(0, open + (x-1)*ext + prefixmax_{x' < x}( T[x'] - x'*ext ))
 *
 * which equals the reference's recurrence GB[x] = max(M[x-1]+open,
 * GA[x-1]+open, GB[x-1]+ext, 0) (alignment.c:140-155 with min = 0): a run of
 * gap_b always starts from a match or gap_a value because open <= ext, and
 * the clamps at 0 inside the run are absorbed by the final one (ext <= 0).
 * The scan is 5 shuffle+max steps per block of 32 columns plus a carry between
 * blocks; at 12 B/cell the arithmetic has an order of magnitude of slack.
 *
 * Smith-Waterman with the scoring shapes of sa_fast.cuh (affine gaps with
 * gap_open <= 0, no gap / mismatch restrictions), len_a <= 511.
 *
 * Needleman-Wunsch (template flag NW) runs the same rows without the clamps:
 * with no restriction flags every interior cell has a real predecessor, so the
 * INT_MIN-based sentinel `min` of alignment.c:41 only ever shows up in the
 * border cells (alignment.c:62-80), which are written with its exact value;
 * inside the kernel the sentinel is MATS_NEG, far below any real score.  The
 * host admits a batch only if none of the reference's `min + penalty` sums
 * could wrap (run_mats), so leaving the sentinel out of the maxima is exact.
 * gap_b is the same scan, seeded by the border column: T[0] = gap_a[y][0].
 * Free start gaps only change the border values (0); free end gaps (template
 * flag FREE) make the last column's gap_a the plain best of the cell above and
 * the last row's gap_b the same scan with open = ext = 0
 * (alignment.c:117-122,136-141).
 *
 * Everything else goes pair by pair through general_kernel<MODE_MATS>.
 */
#ifndef SA_MATS_CUH
#define SA_MATS_CUH

#include "sa_platform.h"
#include "sa_kernels.cuh"

namespace sa {

constexpr int MATS_WARPS = 4;
constexpr int MATS_NEG = -(1 << 29);

__device__ __forceinline__ void st_stream(int32_t *p, int v)
{
#if defined(__CUDA_ARCH__)
  __stcs(p, v);
#else
  *p = v;
#endif
}

/* two-input max as a three-input DPX op: VIMNMX3 issues at full rate on the
 * ALU pipe, the two-input IMNMX only at half rate (tools/microbench.cu) */
__device__ __forceinline__ int fmax2(int a, int b)
{
  int c = b;
#if defined(__CUDA_ARCH__)
  asm volatile("" : "+r"(c));   /* keeps ptxas from folding max3(a,b,b) back into the two-input form */
#endif
  return max3(a, b, c);
}

struct MatsArgs {
  const uint8_t *seq_a, *seq_b;
  const int64_t *off_a, *off_b;
  int64_t npairs;
  ScoreParams sp;
  const int32_t *sub;          /* [cb*ncodes + ca] plain substitution scores */
  const uint8_t *lut;
  int32_t *mats;               /* per pair: match | gap_a | gap_b planes, dense */
  const int64_t *mat_off;      /* first int of pair p's match plane */
  int32_t *score;              /* best match score per pair (as score mode), may be null */
  unsigned long long *counter;
};

/* one block of a row, everything that needs only row y-1: match, gap_a and
 * the scan input u = max(M, GA) - x*ext.  diag: H[y-1][x-1] comes from the
 * lane before; lane 0 takes it from lane 31 of the block before, which that
 * lane sends instead of its own value (one rotating shuffle, no second one) */
template <int NB, bool NW, bool FREE = false>
__device__ __forceinline__ void mats_block_head(int j, int lane, int la, const int16_t *prow, const int (&hp)[NB],
                                                const int (&gap)[NB], int &prev_old, int open, int ext,
                                                int bord, int minv, int ext_r, int &m, int &ga, int &u)
{
  const int x = 32 * j + lane;
  const bool cell = x >= 1 && x <= la;
  const int send = lane == 31 ? prev_old : hp[j];
  const int diag = __shfl_sync(FULL, send, (lane + 31) & 31);
  prev_old = hp[j];
  const int sub = prow[x];
  if constexpr(NW) {
    /* column 0 is the border: match = min, gap_a = bord (alignment.c:71-79) */
    m = cell ? diag + sub : minv;
    /* free end gaps: in the last column gap_a takes the best of the cell above as it is (alignment.c:117-122) */
    if constexpr(FREE) {
      ga = cell ? (x == la ? hp[j] : fmax2(gap[j] + ext, hp[j] + open)) : (x == 0 ? bord : minv);
      u = x == 0 ? bord : (cell ? fmax2(m, ga) - x * ext_r : MATS_NEG);
    } else {
      ga = cell ? fmax2(gap[j] + ext, hp[j] + open) : (x == 0 ? bord : minv);
      u = x == 0 ? bord : (cell ? fmax2(m, ga) - x * ext : MATS_NEG);
    }
    return;
  }
  m = cell ? addmax(diag, sub, 0) : 0;
  ga = cell ? max3(gap[j] + ext, hp[j] + open, 0) : 0;
  u = x <= la ? fmax2(m, ga) - x * ext : 0;          /* x = 0: the border's 0; past len_a: never read by a real cell */
}

/* NB blocks of 32 columns cover x = 0 .. len_a.
 * CS: streaming stores (st.global.cs) -- an experiment, no measurable difference.
 * PACK: the prefix scans of two neighbouring blocks share their shuffles, the two
 * values packed in the halves of a register (needs every scan value in int16).
 * NW + PACK is the default wherever every scan value fits (round 2: timed on a B200,
 * 69 vs 67 % of the HBM roofline at 150x150 DNA, 73 vs 58 % at 400x400 protein;
 * bit-exact in the GPU fuzz, profiles/fuzz_r02l.json); SEQALIGN_MATS_NOPACK=1 = int32 scans.
 * FREE (NW): free end gaps; a template flag because open / ext as per-row values
 * cost the common rows their loop-invariant x*ext terms (measured: 60 / 43 % of the
 * HBM roofline instead of 67 / 60 %). */
template <int NB, bool CS, bool PACK, bool NW = false, bool FREE = false>
__global__ void __launch_bounds__(MATS_WARPS * 32)
mats_kernel(const MatsArgs A)
{
  constexpr int PW = NB * 32;                         /* profile row width (columns) */
  unsigned char *dsm = SA_DYN_SMEM();
  const ScoreParams &sp = A.sp;
  const int n = sp.ncodes;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  /* shared layout: [lut 256][sub n*n int32][per warp: profile n*PW int16] */
  uint8_t *s_lut = dsm;
  int32_t *s_sub = (int32_t *)(dsm + 256);
  int16_t *s_prof = (int16_t *)(dsm + 256 + ((n * n * 4 + 15) & ~15)) + (size_t)wib * n * PW;
  for(int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = A.lut[i];
  for(int i = threadIdx.x; i < n * n; i += blockDim.x) s_sub[i] = A.sub[i];
  __syncthreads();
  const int open = sp.open, ext = sp.ext;
  const int minv = sp.minv;                           /* NW: the value the border cells carry */

  for(;;) {
    unsigned long long t = 0;
    if(lane == 0) t = atomicAdd(A.counter, 1ull);
    t = __shfl_sync(FULL, t, 0);
    if(t >= (unsigned long long)A.npairs) break;
    const int64_t p = (int64_t)t;
    const int64_t oa = A.off_a[p], ob = A.off_b[p];
    const int la = (int)(A.off_a[p + 1] - oa), lb = (int)(A.off_b[p + 1] - ob);
    const int W = la + 1;
    const int64_t cells = (int64_t)W * (lb + 1);
    int32_t *pm = A.mats + A.mat_off[p], *pga = pm + cells, *pgb = pga + cells;

    /* query profile: s_prof[c][x] = sub(a[x-1], c) for 1 <= x <= len_a */
    __syncwarp();
#pragma unroll
    for(int j = 0; j < NB; j++) {
      const int x = 32 * j + lane;
      const int ca = (x >= 1 && x <= la) ? s_lut[A.seq_a[oa + x - 1]] : -1;
      for(int c = 0; c < n; c++) s_prof[c * PW + x] = ca >= 0 ? (int16_t)s_sub[c * n + ca] : (int16_t)0;
    }
    __syncwarp();

    /* row 0: borders, all three matrices 0 (alignment.c:47-57) */
    int hp[NB], gap[NB];
    int endv = 0;                  /* NW: H of the newest row at column len_a (its owner lane only) */
#pragma unroll
    for(int j = 0; j < NB; j++) {
      hp[j] = 0; gap[j] = 0;
      const int x = 32 * j + lane;
      if constexpr(NW) {
        /* NW row 0: match = gap_a = min, gap_b = gap_open + x*ext (alignment.c:62-69) */
        if(x >= 1 && x <= la) {
          hp[j] = sp.no_start ? 0 : sp.gap_open + x * ext; gap[j] = MATS_NEG;
          pm[x] = minv; pga[x] = minv; pgb[x] = hp[j];
        } else if(x == 0) { pm[0] = 0; pga[0] = 0; pgb[0] = 0; }
      } else if(x <= la) { pm[x] = 0; pga[x] = 0; pgb[x] = 0; }
      if(NW && x == la) endv = hp[j];
    }
    int best = 0;
    int codes = 0;
    for(int y = 1; y <= lb; y++) {
      /* codes of 32 rows of seq_b at a time, one per lane */
      if(((y - 1) & 31) == 0) codes = (y + lane <= lb) ? s_lut[A.seq_b[ob + y - 1 + lane]] : 0;
      const int c = __shfl_sync(FULL, codes, (y - 1) & 31);
      const int16_t *prow = s_prof + c * PW;
      const int64_t row = (int64_t)y * W;
      const int bord = (NW && !sp.no_start) ? sp.gap_open + y * ext : 0;   /* NW: gap_a[y][0] (alignment.c:76-77) */
      /* free end gaps: along the last row gap_b costs nothing (alignment.c:136-141), the same scan with 0 / 0 */
      const bool free_row = FREE && y == lb;
      const int open_r = free_row ? 0 : open, ext_r = free_row ? 0 : ext;
      int prev_old = 0;            /* this lane's H[y-1] in the block before (lane 31's is the one that travels) */
      int run = MATS_NEG;          /* prefix maximum of u over the blocks before */

      /* the tail of a block: gap_b from the exclusive prefix maximum, H, stores */
      auto finish = [&](int j, int m, int ga, int excl_in_block, int total) {
        const int x = 32 * j + lane;
        const bool cell = x >= 1 && x <= la;
        const int excl = lane == 0 ? run : fmax2(excl_in_block, run);
        run = fmax2(run, total);
        int gb;
        if constexpr(NW && FREE) gb = cell ? excl + open_r + (x - 1) * ext_r : minv;
        else if constexpr(NW) gb = cell ? excl + open + (x - 1) * ext : minv;
        else gb = cell ? fmax2(excl + open + (x - 1) * ext, 0) : 0;
        const int h = max3(m, ga, gb);
        if(x <= la) {
          if(CS) { st_stream(pm + row + x, m); st_stream(pga + row + x, ga); st_stream(pgb + row + x, gb); }
          else { pm[row + x] = m; pga[row + x] = ga; pgb[row + x] = gb; }
        }
        best = fmax2(best, m);
        hp[j] = (NW ? x <= la : cell) ? h : 0;           /* NW column 0: h = gap_a[y][0] */
        if(NW && x == la) endv = h;
        gap[j] = ga;
      };

      if constexpr(PACK) {
#pragma unroll
        for(int j = 0; j < NB; j += 2) {
          int m0, ga0, u0, m1 = 0, ga1 = 0, u1 = 0;
          /* NW: columns past len_a carry MATS_NEG, whose low half is 0 -- like SW's 0 there it only reaches columns that are not cells */
          mats_block_head<NB, NW, FREE>(j, lane, la, prow, hp, gap, prev_old, open, ext, bord, minv, ext_r, m0, ga0, u0);
          if(j + 1 < NB) mats_block_head<NB, NW, FREE>(j + 1, lane, la, prow, hp, gap, prev_old, open, ext, bord, minv, ext_r, m1, ga1, u1);
          unsigned w = ((unsigned)u0 & 0xffffu) | ((unsigned)u1 << 16);
#pragma unroll
          for(int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(FULL, w, o);
            if(lane >= o) w = max3_s16x2(w, v, v);
          }
          const unsigned e = __shfl_up_sync(FULL, w, 1), tot = __shfl_sync(FULL, w, 31);
          finish(j, m0, ga0, (int)(short)(e & 0xffffu), (int)(short)(tot & 0xffffu));
          if(j + 1 < NB) finish(j + 1, m1, ga1, (int)e >> 16, (int)tot >> 16);
        }
      } else {
#pragma unroll
        for(int j = 0; j < NB; j++) {
          int m, ga, incl;
          mats_block_head<NB, NW, FREE>(j, lane, la, prow, hp, gap, prev_old, open, ext, bord, minv, ext_r, m, ga, incl);
#pragma unroll
          for(int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if(lane >= o) incl = fmax2(incl, v);
          }
          const int e = __shfl_up_sync(FULL, incl, 1), tot = __shfl_sync(FULL, incl, 31);
          finish(j, m, ga, e, tot);
        }
      }
    }
    if(A.score) {
      if constexpr(NW) {
        /* NW score: max of the three matrices at [len_a][len_b] (needleman_wunsch.c:52-55) = H there */
        const int v = __shfl_sync(FULL, endv, la & 31);
        if(lane == 0) A.score[p] = v;
      } else {
#pragma unroll
        for(int o = 16; o > 0; o >>= 1) best = fmax2(best, __shfl_xor_sync(FULL, best, o));
        if(lane == 0) A.score[p] = best;
      }
    }
  }
}

inline size_t mats_smem_bytes(int NB, int ncodes)
{
  return 256 + (((size_t)ncodes * ncodes * 4 + 15) & ~(size_t)15) + (size_t)MATS_WARPS * ncodes * NB * 32 * 2;
}

/* smallest instantiated block count covering max_la + 1 columns; 0 if none */
inline int mats_blocks(int64_t max_la)
{
  static const int kNB[] = {1, 2, 3, 4, 5, 6, 8, 10, 13, 16};
  for(int nb : kNB) if((int64_t)nb * 32 >= max_la + 1) return nb;
  return 0;
}

template <int NB>
int mats_launch_nb(const MatsArgs &M, bool pack, bool nw, size_t smem, int num_sms, cudaStream_t st)
{
  const char *cs_env = getenv("SEQALIGN_MATS_STORES");   /* "stream" = st.global.cs (experiments: no measurable difference) */
  const bool cs = cs_env && cs_env[0] == 's';
  void (*kfn)(const MatsArgs) = pack ? (cs ? mats_kernel<NB, true, true> : mats_kernel<NB, false, true>)
                                     : (cs ? mats_kernel<NB, true, false> : mats_kernel<NB, false, false>);
  /* NW: free end gaps (FREE) are their own instantiations, so that the common rows keep open / ext loop-invariant */
  if(nw) kfn = M.sp.no_end ? (pack ? mats_kernel<NB, false, true, true, true> : mats_kernel<NB, false, false, true, true>)
                           : (pack ? mats_kernel<NB, false, true, true, false> : mats_kernel<NB, false, false, true, false>);
  if(!smem_opt_in(kfn, smem)) return -1;
  int per_sm = 1;
  if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, MATS_WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)num_sms * per_sm;
  const int64_t need = (M.npairs + MATS_WARPS - 1) / MATS_WARPS;
  if(grid > need) grid = need;
  SA_LAUNCH(kfn, (int)(grid < 1 ? 1 : grid), MATS_WARPS * 32, smem, st, M);
  return 0;
}

inline int mats_launch(int NB, bool pack, bool nw, const MatsArgs &M, int ncodes, int num_sms, size_t smem_optin, cudaStream_t st)
{
  const size_t smem = mats_smem_bytes(NB, ncodes);
  if(smem > smem_optin) return -1;
  switch(NB) {
    case 1: return mats_launch_nb<1>(M, pack, nw, smem, num_sms, st);
    case 2: return mats_launch_nb<2>(M, pack, nw, smem, num_sms, st);
    case 3: return mats_launch_nb<3>(M, pack, nw, smem, num_sms, st);
    case 4: return mats_launch_nb<4>(M, pack, nw, smem, num_sms, st);
    case 5: return mats_launch_nb<5>(M, pack, nw, smem, num_sms, st);
    case 6: return mats_launch_nb<6>(M, pack, nw, smem, num_sms, st);
    case 8: return mats_launch_nb<8>(M, pack, nw, smem, num_sms, st);
    case 10: return mats_launch_nb<10>(M, pack, nw, smem, num_sms, st);
    case 13: return mats_launch_nb<13>(M, pack, nw, smem, num_sms, st);
    case 16: return mats_launch_nb<16>(M, pack, nw, smem, num_sms, st);
  }
  return -1;
}

} // namespace sa

#endif
