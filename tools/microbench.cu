// microbench.cu -- issue rates of the integer ops the DP cell is made of, on
// the box's B200.  Output feeds DESIGN.md's "INT/DPX roof" (SURVEY.md 8d):
// lane-ops per clock per SM for each op, alone and in the cell's mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

enum Op { IADD, IMAX, VIMAX3, VIADDMAX, VIADDMAX_RELU, IMAD, PRMT, SHFL, LDS32, CELL_SW, CELL_SW_KEY, VIMAX3_S16X2, VIADDMAX_S16X2_RELU, CELL_SW_S16 };
static const char *names[] = {"IADD3", "IMNMX(max)", "VIMNMX3", "VIADDMNMX", "VIADDMNMX.RELU", "IMAD", "PRMT(sext)", "SHFL.UP", "LDS.32", "SW cell (7 ops, no key)", "SW cell (9 ops, packed key)", "VIMNMX3.S16x2", "VIADDMNMX.S16x2.RELU", "SW cell s16x2 (2 cells)"};
static const int ops_per_iter[] = {8, 8, 8, 8, 8, 8, 8, 8, 8, 7 * 4, 9 * 4, 8, 8, 7 * 4};

template <int OP>
__global__ void __launch_bounds__(256) bench(int iters, int seed, long long *cyc, int *sink)
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
      // 4 cells of the fast SW recurrence chained left to right
      int hl = v[0], gb = v[1], d = v[2];
#pragma unroll
      for(int k = 0; k < 4; k++) {
        int sub;
        asm volatile("prmt.b32 %0, %1, %2, 0x9991;" : "=r"(sub) : "r"(w + k), "r"(0));
        int m = __viaddmax_s32(d, sub, 0);
        v[4 + k] = __viaddmax_s32_relu(v[4 + k], b, v[k]);
        gb = __viaddmax_s32_relu(gb, b, hl);
        int h = __vimax3_s32(m, v[4 + k], gb);
        if(OP == CELL_SW_KEY) c = max(c, m * 65536 + i);
        else c = max(c, m);
        d = v[k];
        hl = h + a;
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
      unsigned hl = v[0], gb = v[1], d = v[2];
#pragma unroll
      for(int k = 0; k < 4; k++) {
        unsigned sub;
        asm volatile("prmt.b32 %0, %1, %2, 0x9180;" : "=r"(sub) : "r"(w + k), "r"(0));
        unsigned m = __viaddmax_s16x2(d, sub, 0);
        v[4 + k] = __viaddmax_s16x2_relu(v[4 + k], b, v[k]);
        gb = __viaddmax_s16x2_relu(gb, b, hl);
        unsigned h = __vimax3_s16x2(m, v[4 + k], gb);
        c = __vmaxs2(c, m);
        d = v[k];
        hl = __vadd2(h, a);
        v[k] = hl;
      }
      v[1] ^= gb;
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
void run(int sms)
{
  const int blocks = sms * 4, threads = 256, iters = 4096;
  long long *cyc; int *sink;
  cudaMalloc(&cyc, blocks * 8); cudaMalloc(&sink, 4);
  bench<OP><<<blocks, threads>>>(16, 1, cyc, sink);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<OP><<<blocks, threads>>>(iters, 3, cyc, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long *h = new long long[blocks];
  cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for(int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
  // 4 blocks x 256 threads resident per SM for the whole run
  const double lane_ops_per_sm = 4.0 * threads * (double)iters * ops_per_iter[OP];
  printf("{\"op\": \"%s\", \"lane_ops_per_clk_per_sm\": %.1f, \"ms\": %.3f, \"gops\": %.1f}\n", names[OP],
         lane_ops_per_sm / avg, ms, (double)blocks * threads * iters * ops_per_iter[OP] / ms / 1e6);
  cudaFree(cyc); cudaFree(sink); delete[] h;
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
  const int s = p.multiProcessorCount;
  run<IADD>(s); run<IMAX>(s); run<VIMAX3>(s); run<VIADDMAX>(s); run<VIADDMAX_RELU>(s); run<IMAD>(s);
  run<PRMT>(s); run<SHFL>(s); run<LDS32>(s); run<CELL_SW>(s); run<CELL_SW_KEY>(s);
  run<VIMAX3_S16X2>(s); run<VIADDMAX_S16X2_RELU>(s); run<CELL_SW_S16>(s);
  return 0;
}
