/*
 * sa_fast.cuh -- sm_100a kernels of the batch alignment engine (part 2):
 * the score-only fill for the common scoring shape (affine gaps with
 * gap_open <= 0, no gap/mismatch restrictions, no free end gaps).  This is
 * the kernel behind the headline metric (batched 150x150 DNA SW).
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
  const int8_t *tab8;   /* [cb*ncodes + ca] = sub - open */
  const uint8_t *lut;
  unsigned long long *counter;
  int32_t *score, *xend, *yend;
  int max_lb;
  int a_stage, b_stage; /* bytes per pair per stage (multiples of 16) */
};

struct FastPlan {
  int G = 0, K = 0;
  bool is_sw = false;
  const char *name = "";
  std::vector<int8_t> tab8;
  size_t smem = 0;
  int a_stage = 0, b_stage = 0;
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

template <int G, int K, bool IS_SW>
__global__ void __launch_bounds__(FAST_WARPS * 32)
fast_score_kernel(const FastArgs A)
{
  constexpr int NG = 32 / G;              /* pairs per warp */
  constexpr int KW = (K + 3) / 4;         /* profile words per lane */
  constexpr int KPAD = KW * 4;
  constexpr int PSTRIDE = 32 * KPAD;      /* bytes per profile row */

  unsigned char *dsm = SA_DYN_SMEM();
  const ScoreParams &sp = A.sp;
  const int n = sp.ncodes;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int grp = lane / G, lig = lane % G;

  /* shared layout: [mbarriers][lut 256][tab8 n*n pad16][per-warp: profile | a stages | b stages] */
  uint64_t *s_bar = (uint64_t *)dsm;                       /* FAST_WARPS*2 */
  uint8_t *s_lut = dsm + 64;
  int8_t *s_tab = (int8_t *)(dsm + 64 + 256);
  const int tab_bytes = (n * n + 15) & ~15;
  const int warp_bytes = n * PSTRIDE + 2 * NG * (A.a_stage + A.b_stage);
  unsigned char *wbase = dsm + 64 + 256 + tab_bytes + wib * warp_bytes;
  unsigned char *s_prof = wbase;
  unsigned char *s_a = wbase + n * PSTRIDE;                /* [stage][grp][a_stage] */
  unsigned char *s_b = s_a + 2 * NG * A.a_stage;           /* [stage][grp][b_stage] */
  uint64_t *bar = s_bar + wib * 2;

  for(int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = A.lut[i];
  for(int i = threadIdx.x; i < n * n; i += blockDim.x) s_tab[i] = A.tab8[i];
  if(lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
  mbar_fence_init();
  __syncthreads();

  const int open = sp.open, ext = sp.ext, minv = sp.minv;
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
        if(ea > oa) bytes += (uint32_t)(((ea + 15) & ~(int64_t)15) - (oa & ~(int64_t)15));
        if(eb > ob) bytes += (uint32_t)(((eb + 15) & ~(int64_t)15) - (ob & ~(int64_t)15));
      }
      mbar_expect_tx(&bar[st], bytes);
      for(int g = 0; g < NG; g++) {
        const int64_t p = t * NG + g;
        if(p >= A.npairs) break;
        const int64_t oa = A.off_a[p], ob = A.off_b[p];
        const int64_t ea = A.off_a[p + 1], eb = A.off_b[p + 1];
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
    }
    mbar_wait(&bar[stage], stage ? phase1 : phase0);
    if(stage) phase1 ^= 1; else phase0 ^= 1;

    unsigned char *ra = s_a + (stage * NG + grp) * A.a_stage + sha;
    unsigned char *rb = s_b + (stage * NG + grp) * A.b_stage + shb;

    /* seq_b: raw bytes -> codes, in place */
    for(int i = lig; i < lb; i += G) rb[i] = s_lut[rb[i]];

    /* query profile of my K columns: row c holds sub'(a[x], c) */
    const int xf = lig * K + 1;
    int acode[K];
#pragma unroll
    for(int j = 0; j < K; j++) acode[j] = (xf + j <= la) ? s_lut[ra[xf + j - 1]] : 0;
    for(int c = 0; c < n; c++) {
      const int8_t *trow = s_tab + c * n;
      unsigned *dst = (unsigned *)(s_prof + c * PSTRIDE + lane * KPAD);
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
    __syncwarp();

    /* borders (alignment.c:47-81), in H' = H+open form */
    int hp[K], ga[K], best[K];
#pragma unroll
    for(int j = 0; j < K; j++) {
      const int x = xf + j;
      if(IS_SW) { hp[j] = open; ga[j] = 0; }
      else {
        const int gb0 = sp.no_start ? 0 : addw(sp.gap_open, x * ext);
        hp[j] = addw(imax(gb0, minv), open);
        ga[j] = minv;
      }
      best[j] = 0;
    }
    int hd; /* H'(xf-1, y-1) */
    if(IS_SW || xf == 1) hd = open;
    else hd = addw(imax(sp.no_start ? 0 : addw(sp.gap_open, (xf - 1) * ext), minv), open);

    int out_h = 0, out_gb = 0;
    int maxlb = lb;
#pragma unroll
    for(int o = 16; o >= G; o >>= 1) maxlb = imax(maxlb, __shfl_xor_sync(FULL, maxlb, o));
    const int nsteps = maxlb > 0 ? maxlb + G - 1 : 0;
    const unsigned char *prow = s_prof + lane * KPAD;

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
          gb = minv;
        }
      }
      const int hl_in = hl;
      if(active) {
        const int c = rb[y - 1];
        const unsigned *pw = (const unsigned *)(prow + c * PSTRIDE);
        unsigned w[KW];
        /* widest load the row layout allows: lane*KPAD is 16/8/4-byte
         * aligned for KW%4==0 / KW%2==0 / odd KW, and those are exactly
         * the widths that keep the 32 lanes on distinct banks */
        if(KW % 4 == 0) {
#pragma unroll
          for(int q = 0; q < KW / 4; q++) {
            const uint4 v = ((const uint4 *)pw)[q];
            w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
          }
        } else if(KW % 2 == 0) {
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
#pragma unroll
        for(int j = 0; j < K; j++) {
          const int sub = sext_byte_dyn(w[j / 4], j & 3);
          int m, h;
          if(IS_SW) {
            m = addmax(d, sub, 0);
            ga[j] = addmax_relu(ga[j], ext, hp[j]);
            gb = addmax_relu(gb, ext, hl);
            h = max3(m, ga[j], gb);
            best[j] = imax(best[j], m * 65536 + ykey);
          } else {
            m = addmax(d, sub, minv);
            ga[j] = max3(addw(ga[j], ext), hp[j], minv);
            gb = max3(addw(gb, ext), hl, minv);
            h = max3(m, ga[j], gb);
          }
          d = hp[j];
          hl = addw(h, open);
          hp[j] = hl;
        }
        out_h = hl;
        out_gb = gb;
        hd = hl_in; /* next row's diagonal */
      }
    }

    /* results */
    if(IS_SW) {
      int bv = 0, bx = 0, by = 0;
#pragma unroll
      for(int j = 0; j < K; j++) {
        const int v = best[j] >> 16;
        if(xf + j <= la && v > bv) { bv = v; bx = xf + j; by = 0xffff - (best[j] & 0xffff); }
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

/* ---- host side ---------------------------------------------------------- */

struct FastShape { int G, K; };
static const FastShape kFastShapes[] = {
    {8, 8}, {8, 12}, {8, 16}, {8, 20}, {16, 12}, {16, 16}, {32, 10}, {32, 12}, {32, 16}};

inline size_t fast_smem_bytes(int G, int K, int ncodes, int a_stage, int b_stage)
{
  const int NG = 32 / G, KPAD = (K + 3) / 4 * 4;
  const size_t warp_bytes = (size_t)ncodes * 32 * KPAD + 2 * (size_t)NG * (a_stage + b_stage);
  return 64 + 256 + (((size_t)ncodes * ncodes + 15) & ~(size_t)15) + FAST_WARPS * warp_bytes;
}

/* can the fast kernel take this batch?  fills the plan if so */
inline bool fast_plan(const scoring_t *s, const FlatTable &ft, const ScoreParams &sp,
                      int64_t max_la, int64_t max_lb, FastPlan *plan)
{
  if(sp.no_end || sp.no_gaps_a || sp.no_gaps_b || sp.no_mismatches) return false;
  if(s->gap_open > 0) return false;                 /* needs open <= ext */
  if(ft.any_unknown) return false;
  if(max_la < 1 || max_lb < 1 || max_la > 512 || max_lb > 65535) return false;
  /* sub' = sub - open must fit int8 */
  const long lo = (long)ft.min_sub - sp.open, hi = (long)ft.max_sub - sp.open;
  if(lo < -127 || hi > 127) return false;
  for(size_t k = 0; k < ft.unknown.size(); k++) if(ft.unknown[k]) return false;
  if(sp.is_sw) {
    /* packed (score,row) key: scores must stay below 2^15 */
    const long cap = (long)(max_la < max_lb ? max_la : max_lb) * (ft.max_sub > 0 ? ft.max_sub : 0);
    if(cap >= 32768) return false;
  } else {
    /* sentinel arithmetic must not wrap (alignment.c:41) */
    const long room = labs((long)s->min_penalty);
    if(-(long)sp.open > room || -(long)sp.ext > room || -(long)ft.min_sub > room) return false;
    if(sp.ext > 0 || sp.open > 0) return false;
    const long worst = (long)sp.gap_open + (long)(max_la > max_lb ? max_la : max_lb) * sp.ext;
    if(worst < -(1L << 30)) return false;
  }
  int G = 0, K = 0;
  for(const FastShape &sh : kFastShapes)
    if((int64_t)sh.G * sh.K >= max_la) { G = sh.G; K = sh.K; break; }
  if(!G) return false;
  plan->G = G; plan->K = K; plan->is_sw = sp.is_sw != 0;
  plan->a_stage = (int)((G * K + 15 + 15) & ~15) + 16;
  plan->b_stage = (int)((max_lb + 15 + 15) & ~(int64_t)15) + 16;
  plan->smem = fast_smem_bytes(G, K, ft.ncodes, plan->a_stage, plan->b_stage);
  if(plan->smem > 200 * 1024) return false;
  const int n = ft.ncodes;
  plan->tab8.assign(((size_t)n * n + 15) & ~(size_t)15, 0);
  for(int i = 0; i < n * n; i++) plan->tab8[i] = (int8_t)(ft.sub[i] - sp.open);
  plan->name = sp.is_sw ? "fast_sw_score" : "fast_nw_score";
  return true;
}

template <int G, int K>
int fast_launch_gk(const FastPlan &plan, const FastArgs &F, int grid, cudaStream_t st)
{
  void (*kfn)(const FastArgs) = plan.is_sw ? fast_score_kernel<G, K, true> : fast_score_kernel<G, K, false>;
  if(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem) != cudaSuccess) return -1;
  SA_LAUNCH(kfn, grid, FAST_WARPS * 32, plan.smem, st, F);
  return 0;
}

inline int fast_launch(const FastPlan &plan, FastArgs F, int num_sms, size_t smem_optin, cudaStream_t st)
{
  F.a_stage = plan.a_stage; F.b_stage = plan.b_stage;
  if(plan.smem > smem_optin) return -1;
  /* persistent grid: as many CTAs per SM as shared memory allows (<= 4) */
  int per_sm = (int)((smem_optin + 1024) / (plan.smem + 1024));
  if(per_sm > 4) per_sm = 4;
  if(per_sm < 1) per_sm = 1;
  const int NG = 32 / plan.G;
  int64_t nsets = (F.npairs + NG - 1) / NG;
  int64_t grid = (int64_t)num_sms * per_sm;
  const int64_t need = (nsets + FAST_WARPS - 1) / FAST_WARPS;
  if(grid > need) grid = need;
  if(grid < 1) grid = 1;
#define SA_FAST_CASE(g, k) if(plan.G == g && plan.K == k) return fast_launch_gk<g, k>(plan, F, (int)grid, st)
  SA_FAST_CASE(8, 8); SA_FAST_CASE(8, 12); SA_FAST_CASE(8, 16); SA_FAST_CASE(8, 20);
  SA_FAST_CASE(16, 12); SA_FAST_CASE(16, 16);
  SA_FAST_CASE(32, 10); SA_FAST_CASE(32, 12); SA_FAST_CASE(32, 16);
#undef SA_FAST_CASE
  return -1;
}

} // namespace sa

#endif
