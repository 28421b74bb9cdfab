// microbench.cu -- issue rates of the integer ops the DP cell is made of, on
// the box's B200.  Output feeds DESIGN.md's "INT/DPX roof" (SURVEY.md 8d):
// lane-ops per clock per SM for each op, alone and in the cell's mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

enum Op { IADD, IMAX, VIMAX3, VIADDMAX, VIADDMAX_RELU, IMAD, PRMT, SHFL, LDS32, CELL_SW, CELL_SW_KEY, VIMAX3_S16X2, VIADDMAX_S16X2_RELU, CELL_SW_S16 };
static const char *names[] = {"IADD3", "IMNMX(max)", "VIMNMX3", "VIADDMNMX", "VIADDMNMX.RELU", "IMAD", "PRMT(sext)", "SHFL.UP", "LDS.32", "SW cell int32 score-only mix (fast_score_kernel)", "SW cell int32 end-cell mix (fast_score_kernel TREE)", "VIMNMX3.S16x2", "VIADDMNMX.S16x2.RELU", "SW cell s16x2 mix (fast16_kernel, 2 cells/op)"};
static const int ops_per_iter[] = {8, 8, 8, 8, 8, 8, 8, 8, 8, 7 * 4, 9 * 4, 8, 8, 7 * 4};
// DP cells advanced per loop iteration (0 = not a cell mix)
static const int cells_per_iter[] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 0, 0, 8};

template <int OP>
__global__ void __launch_bounds__(256) bench(int iters, int seed, int one, int one32, long long *cyc, int *sink)
{
  __shared__ int sm[1024];
  for(int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * seed;
  __syncthreads();
  int v[8];
#pragma unroll
  for(int k = 0; k < 8; k++) v[k] = threadIdx.x * (k + 1) + seed;
  int a = seed * 3 + 1, b = seed - 7, c = seed ^ 0x55;
  unsigned w = seed * 0x01010101u;
  long long t0 = clock64();
  for(int i = 0; i < iters; i++) {
    if(OP == IADD) {
#pragma unroll
      for(int k = 0; k < 8; k++) v[k] = v[k] + a + b;
    } else if(OP == IMAX) {
#pragma unroll
      for(int k = 0; k < 8; k++) v[k] = max(v[k], a + k) ^ 1;
    } else if(OP == VIMAX3) {
#pragma unroll
      for(int k = 0; k < 8; k++) v[k] = __vimax3_s32(v[k], a, v[(k + 1) & 7]);
    } else if(OP == VIADDMAX) {
#pragma unroll
      for(int k = 0; k < 8; k++) v[k] = __viaddmax_s32(v[k], b, v[(k + 3) & 7]);
    } else if(OP == VIADDMAX_RELU) {
#pragma unroll
      for(int k = 0; k < 8; k++) v[k] = __viaddmax_s32_relu(v[k], b, v[(k + 3) & 7]);
    } else if(OP == IMAD) {
#pragma unroll
      for(int k = 0; k < 8; k++) v[k] = v[k] * a + b;
    } else if(OP == PRMT) {
#pragma unroll
      for(int k = 0; k < 8; k++) asm volatile("prmt.b32 %0, %1, %2, 0x9991;" : "=r"(v[k]) : "r"(v[k]), "r"(0));
    } else if(OP == SHFL) {
#pragma unroll
      for(int k = 0; k < 8; k++) v[k] = __shfl_up_sync(0xffffffffu, v[k], 1);
    } else if(OP == LDS32) {
#pragma unroll
      for(int k = 0; k < 8; k++) v[k] = sm[(v[k] + threadIdx.x) & 1023];
    } else if(OP == CELL_SW || OP == CELL_SW_KEY) {
      // 4 cells of the int32 recurrence as the fast kernel issues them:
      // 3 x VIADDMNMX + VIMNMX3 per cell, max3 tree over pairs of columns,
      // H+open (and the key packing) as IMAD with a runtime multiplier
      int hl = v[0], gb = v[1], d = v[2], kprev = 0;
#pragma unroll
      for(int k = 0; k < 4; k++) {
        int sub = sm[(k + i) & 1023];
        int m = __viaddmax_s32(d, sub, 0);
        v[4 + k] = __viaddmax_s32_relu(v[4 + k], b, v[k]);
        gb = __viaddmax_s32_relu(gb, b, hl);
        int h = __vimax3_s32(m, v[4 + k], gb);
        int key = OP == CELL_SW_KEY ? m * one32 + (31 - k) : m;
        if(k & 1) c = __vimax3_s32(c, kprev, key);
        kprev = key;
        d = v[k];
        hl = h * one + a;
        v[k] = hl;
      }
      v[1] ^= gb;
    } else if(OP == VIMAX3_S16X2) {
#pragma unroll
      for(int k = 0; k < 8; k++) v[k] = __vimax3_s16x2(v[k], a, v[(k + 1) & 7]);
    } else if(OP == VIADDMAX_S16X2_RELU) {
#pragma unroll
      for(int k = 0; k < 8; k++) v[k] = __viaddmax_s16x2_relu(v[k], b, v[(k + 3) & 7]);
    } else if(OP == CELL_SW_S16) {
      // 4 packed cell pairs as fast16_kernel issues them: PRMT combine,
      // 3 x VIADDMNMX.S16x2, VIMNMX3.S16x2, max3 tree, packed add as IMAD
      unsigned hl = v[0], gb = v[1], d = v[2], kprev = 0, cc = c;
      // profile words of the two pairs: one shared-memory word each per 4 cells, as in the kernel
      // (an earlier version derived the PRMT inputs with two extra ALU ops per cell and so
      // understated this ceiling by about a fifth)
      const unsigned wl = (unsigned)sm[(threadIdx.x * 5 + (i & 3)) & 1023], wh = (unsigned)sm[(threadIdx.x * 5 + 640 + (i & 3)) & 1023];
#pragma unroll
      for(int k = 0; k < 4; k++) {
        unsigned sub;
        if(k == 0) asm volatile("prmt.b32 %0, %1, %2, 0xc480;" : "=r"(sub) : "r"(wl), "r"(wh));
        else if(k == 1) asm volatile("prmt.b32 %0, %1, %2, 0xd591;" : "=r"(sub) : "r"(wl), "r"(wh));
        else if(k == 2) asm volatile("prmt.b32 %0, %1, %2, 0xe6a2;" : "=r"(sub) : "r"(wl), "r"(wh));
        else asm volatile("prmt.b32 %0, %1, %2, 0xf7b3;" : "=r"(sub) : "r"(wl), "r"(wh));
        unsigned m = __viaddmax_s16x2(d, sub, 0x00020002u);
        v[4 + k] = __viaddmax_s16x2(v[4 + k], b, v[k]);
        gb = __viaddmax_s16x2(gb, b, hl);
        unsigned h = __vimax3_s16x2(m, v[4 + k], gb);
        if(k & 1) cc = __vimax3_s16x2(cc, kprev, m);
        kprev = m;
        d = v[k];
        hl = h * (unsigned)one + (unsigned)a;
        v[k] = hl;
      }
      v[1] ^= gb;
      c = cc;
    }
  }
  long long t1 = clock64();
  int s = c;
#pragma unroll
  for(int k = 0; k < 8; k++) s ^= v[k];
  if(s == 0x7fffffff) sink[0] = s;
  if(threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int sms, int clock_khz)
{
  const int blocks = sms * 4, threads = 256, iters = 4096;
  long long *cyc; int *sink;
  cudaMalloc(&cyc, blocks * 8); cudaMalloc(&sink, 4);
  bench<OP><<<blocks, threads>>>(16, 1, 1, 32, cyc, sink);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<OP><<<blocks, threads>>>(iters, 3, 1, 32, cyc, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long *h = new long long[blocks];
  cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for(int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
  // 4 blocks x 256 threads resident per SM for the whole run
  const double lane_ops_per_sm = 4.0 * threads * (double)iters * ops_per_iter[OP];
  // event-timed rates (the clock64 figure is kept for reference only: blocks
  // do not all start together, so it over-counts)
  const double total_ops = (double)blocks * threads * iters * ops_per_iter[OP];
  const double gops = total_ops / ms / 1e6;
  const double cells_gcups = (double)blocks * threads * iters * cells_per_iter[OP] / ms / 1e6;
  printf("{\"op\": \"%s\", \"gops\": %.1f, \"lane_ops_per_clk_per_sm_at_max_clock\": %.1f, \"cells_gcups\": %.1f, \"ms\": %.3f, \"clock64_lane_ops_per_clk_per_sm\": %.1f}\n",
         names[OP], gops, gops * 1e9 / (sms * clock_khz * 1e3), cells_gcups, ms, lane_ops_per_sm / avg);
  cudaFree(cyc); cudaFree(sink); delete[] h;
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
  const int s = p.multiProcessorCount;
#define run_(OP) run<OP>(s, p.clockRate)
  run_(IADD); run_(IMAX); run_(VIMAX3); run_(VIADDMAX); run_(VIADDMAX_RELU); run_(IMAD);
  run_(PRMT); run_(SHFL); run_(LDS32); run_(CELL_SW); run_(CELL_SW_KEY);
  run_(VIMAX3_S16X2); run_(VIADDMAX_S16X2_RELU); run_(CELL_SW_S16);
  return 0;
}
