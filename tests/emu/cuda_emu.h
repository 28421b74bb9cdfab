/*
 * cuda_emu.h -- a small warp-synchronous CUDA emulator (TEST INFRASTRUCTURE).
 *
 * The build container has nvcc but no GPU.  To check the kernels' recurrence,
 * border handling and strip logic against the oracle before spending GPU
 * time, tests/emu compiles the *unchanged* kernel + engine sources with g++
 * (-DSA_EMU) against this header.  Every CUDA thread of a block becomes a
 * fiber (ucontext) on one OS thread; __syncthreads / __syncwarp / the
 * __shfl_*_sync family are rendezvous points between fibers; blocks of a
 * grid run one after another; the runtime API (cudaMalloc, cudaMemcpyAsync,
 * streams, events) degenerates to malloc / memcpy / no-ops.
 *
 * It is slow (a context switch per lane per shuffle) and is never linked
 * into the product library: libseqalign_b200.so is nvcc-built and has no CPU
 * path.  Only tests/test_emu_*.py use the emulated build.
 */
#ifndef CUDA_EMU_H
#define CUDA_EMU_H

#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <ucontext.h>
#include <algorithm>
#include <functional>
#include <vector>
#include <chrono>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static thread_local
#define __constant__ static

using std::max;
using std::min;

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };

struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(4) char4 { signed char x, y, z, w; };
static inline int2 make_int2(int x, int y) { int2 r = {x, y}; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r = {x, y}; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r = {x, y, z, w}; return r; }

namespace emu {

struct Warp {
  int live = 0, arrived = 0;
  unsigned gen = 0;
  uint64_t slot[2][32];
};

struct Block;

struct Fiber {
  ucontext_t ctx;
  void *stack = nullptr;
  bool done = false;
  uint3_emu tid;
  int lane = 0, warp = 0;
  unsigned shfl_count = 0;
  Block *blk = nullptr;
};

struct Block {
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  int live = 0, arrived = 0;
  unsigned gen = 0;
  unsigned char *dyn_smem = nullptr;
  uint3_emu bid;
  ucontext_t sched;
};

extern thread_local Fiber *cur;
extern thread_local dim3 g_blockDim, g_gridDim;
extern thread_local std::function<void()> g_entry;

void yield();
void block_barrier();
void warp_barrier();
uint64_t warp_exchange(uint64_t v, int src_lane);
unsigned warp_ballot(int pred);
void run_grid(dim3 grid, dim3 block, size_t smem);

template <class K, class... Args>
void launch(K kernel, dim3 grid, dim3 block, size_t smem, Args... args)
{
  g_entry = [=]() { kernel(args...); };
  run_grid(grid, block, smem);
}

} // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::cur->blk->bid)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define warpSize 32

#define SA_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch(kernel, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)
#define SA_SPIN_HINT() emu::yield()
#define SA_DYN_SMEM() (emu::cur->blk->dyn_smem)

static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) { emu::yield(); }

template <class T>
static inline T emu_shfl(T v, int src)
{
  static_assert(sizeof(T) <= 8, "shuffle of >8 bytes");
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  bits = emu::warp_exchange(bits, src);
  T r;
  memcpy(&r, &bits, sizeof(T));
  return r;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
  int lane = emu::cur->lane;
  int base = lane & ~(width - 1);
  return emu_shfl(v, base + (src & (width - 1)));
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32)
{
  int lane = emu::cur->lane;
  int base = lane & ~(width - 1);
  int src = lane - (int)d;
  return emu_shfl(v, src < base ? lane : src);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32)
{
  int lane = emu::cur->lane;
  int base = lane & ~(width - 1);
  int src = lane + (int)d;
  return emu_shfl(v, src >= base + width ? lane : src);
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32)
{
  (void)width;
  return emu_shfl(v, emu::cur->lane ^ m);
}
static inline unsigned __ballot_sync(unsigned, int pred) { return emu::warp_ballot(pred); }
template <class T>
static inline unsigned __match_any_sync(unsigned, T v)
{
  unsigned m = 0;
  for(int src = 0; src < 32; src++) {
    T o = emu_shfl(v, src);
    if(o == v) m |= 1u << src;
  }
  return m;
}
static inline int __any_sync(unsigned, int pred) { return emu::warp_ballot(pred) != 0; }
static inline int __all_sync(unsigned, int pred) { return emu::warp_ballot(!pred) == 0; }
static inline unsigned __activemask() { return 0xffffffffu; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }

template <class T> static inline T __ldg(const T *p) { return *p; }

template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if(v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; if(v < o) *p = v; return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if(o == c) *p = v; return o; }

static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
  uint64_t both = ((uint64_t)b << 32) | a;
  unsigned r = 0;
  for(int i = 0; i < 4; i++) {
    unsigned s = (sel >> (4 * i)) & 0xf;
    unsigned byte = (unsigned)(both >> (8 * (s & 7))) & 0xff;
    if(s & 8) byte = (byte & 0x80) ? 0xff : 0x00;
    r |= byte << (8 * i);
  }
  return r;
}

/* ---- runtime API subset -------------------------------------------------- */
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef struct emu_event { std::chrono::steady_clock::time_point t; } *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
struct cudaDeviceProp {
  char name[256];
  int major, minor, multiProcessorCount;
  size_t totalGlobalMem, sharedMemPerBlockOptin;
};

static inline const char *cudaGetErrorString(cudaError_t e) { return e ? "emulated CUDA error" : "no error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
  memset(p, 0, sizeof(*p));
  strcpy(p->name, "emulated sm_100");
  p->major = 10; p->minor = 0;
  const char *e = getenv("SA_EMU_SMS");
  p->multiProcessorCount = e ? atoi(e) : 2;
  p->totalGlobalMem = (size_t)4 << 30;
  p->sharedMemPerBlockOptin = 227 * 1024;
  return cudaSuccess;
}
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = (size_t)2 << 30; *t = (size_t)4 << 30; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = 0)
{
  for(size_t r = 0; r < h; r++) memcpy((char *)d + r * dp, (const char *)s + r * sp, w);
  return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event(); return cudaSuccess; }
static const unsigned cudaEventDisableTiming = 2;
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new emu_event(); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
struct cudaIpcMemHandle_t { char reserved[64]; };
static const unsigned cudaIpcMemLazyEnablePeerAccess = 1;
/* single process: a "handle" is the pointer itself */
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof(*h)); memcpy(h, &p, sizeof(p)); return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, &h, sizeof(*p)); return cudaSuccess; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
static const cudaError_t cudaErrorPeerAccessAlreadyEnabled = (cudaError_t)704;
static inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
template <class K>
static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <class K>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t) { *n = 1; return cudaSuccess; }

#endif
