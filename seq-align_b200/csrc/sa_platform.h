/*
 * sa_platform.h -- the one place that knows whether the kernels are being
 * compiled by nvcc for sm_100a (the product) or by g++ against the lane
 * emulator in tests/emu/ (a development aid for the GPU-less build container:
 * it runs the *same kernel source* warp-synchronously on fibers so the
 * recurrence, border and strip logic can be checked against the oracle
 * before a GPU call is spent).  The emulator is test infrastructure: it is
 * never linked into libseqalign_b200.so.
 */
#ifndef SA_PLATFORM_H
#define SA_PLATFORM_H

#include <stdint.h>
#include <stddef.h>

#ifdef SA_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define SA_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define SA_SPIN_HINT() __nanosleep(20)
extern __shared__ __align__(1024) unsigned char sa_dyn_smem_[];
#define SA_DYN_SMEM() (sa_dyn_smem_)
#endif

namespace sa {

/* wrapping int32 add: the reference adds penalties to INT_MIN-based
 * sentinels with plain ints (src/alignment.c:41,110-155); doing it in
 * unsigned keeps the compiler from exploiting signed-overflow UB */
__host__ __device__ __forceinline__ int addw(int a, int b)
{
  return (int)((unsigned)a + (unsigned)b);
}

#if defined(__CUDA_ARCH__)
/* DPX: single-instruction three-way max / add-max (VIMNMX3 / VIADDMNMX) */
__device__ __forceinline__ int max3(int a, int b, int c) { return __vimax3_s32(a, b, c); }
__device__ __forceinline__ int addmax(int a, int b, int c) { return __viaddmax_s32(a, b, c); }
__device__ __forceinline__ int addmax_relu(int a, int b, int c) { return __viaddmax_s32_relu(a, b, c); }
#else
__host__ __device__ __forceinline__ int max3(int a, int b, int c)
{
  int m = a > b ? a : b;
  return m > c ? m : c;
}
__host__ __device__ __forceinline__ int addmax(int a, int b, int c)
{
  int s = addw(a, b);
  return s > c ? s : c;
}
__host__ __device__ __forceinline__ int addmax_relu(int a, int b, int c)
{
  int s = addmax(a, b, c);
  return s > 0 ? s : 0;
}
#endif

__host__ __device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }
__host__ __device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }

/* packed 2 x int16 DPX (VIADDMNMX.S16x2 / VIMNMX3.S16x2): two alignments per
 * instruction when the scores fit 16 bits */
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned addmax_s16x2(unsigned a, unsigned b, unsigned c) { return __viaddmax_s16x2(a, b, c); }
__device__ __forceinline__ unsigned max3_s16x2(unsigned a, unsigned b, unsigned c) { return __vimax3_s16x2(a, b, c); }
#else
__host__ __device__ inline unsigned addmax_s16x2(unsigned a, unsigned b, unsigned c)
{
  unsigned r = 0;
  for(int h = 0; h < 2; h++) {
    const int x = (short)(a >> (16 * h)), y = (short)(b >> (16 * h)), z = (short)(c >> (16 * h));
    const int s = (short)(x + y);
    r |= (unsigned)(unsigned short)(s > z ? s : z) << (16 * h);
  }
  return r;
}
__host__ __device__ inline unsigned max3_s16x2(unsigned a, unsigned b, unsigned c)
{
  unsigned r = 0;
  for(int h = 0; h < 2; h++) {
    int x = (short)(a >> (16 * h)), y = (short)(b >> (16 * h)), z = (short)(c >> (16 * h));
    int m = x > y ? x : y;
    m = m > z ? m : z;
    r |= (unsigned)(unsigned short)m << (16 * h);
  }
  return r;
}
#endif

/* ---- async bulk copy (TMA, 1-D) + mbarrier ------------------------------
 * cp.async.bulk moves a 16-byte-aligned span global -> shared without
 * touching registers and signals an mbarrier with the byte count
 * (SASS: UBLKCP + SYNCS).  Used to prefetch the next pair's sequences
 * while the current pair is in the DP loop. */
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(phase)
      : "memory");
}
/* byte loads through a 32-bit shared-memory address kept in a register (the
 * hot loops advance it with an IMAD, off the ALU pipe); `base` is the start
 * of dynamic shared memory and only used by the host-side emulation */
__device__ __forceinline__ unsigned smem_addr(const void *p, const unsigned char *) { return smem_u32(p); }
__device__ __forceinline__ unsigned lds_u8(unsigned addr, const unsigned char *)
{
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
template <int OFF>
__device__ __forceinline__ unsigned lds_b32(unsigned addr, const unsigned char *)
{
  unsigned v;
  asm volatile("ld.shared.b32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ uint2 lds_b64(unsigned addr, const unsigned char *)
{
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(addr), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ uint4 lds_b128(unsigned addr, const unsigned char *)
{
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(OFF));
  return v;
}
/* make generic-proxy smem writes/reads ordered against the async proxy */
__device__ __forceinline__ void fence_async_smem()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
#else
__host__ __device__ inline void mbar_init(uint64_t *bar, int) { *bar = 0; }
__host__ __device__ inline void mbar_fence_init() {}
__host__ __device__ inline void mbar_expect_tx(uint64_t *, uint32_t) {}
__host__ __device__ inline void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *)
{
  const unsigned char *s = (const unsigned char *)src;
  unsigned char *d = (unsigned char *)dst;
  for(uint32_t i = 0; i < bytes; i++) d[i] = s[i];
}
__host__ __device__ inline void mbar_wait(uint64_t *, uint32_t) {}
__host__ __device__ inline unsigned smem_addr(const void *p, const unsigned char *base) { return (unsigned)((const unsigned char *)p - base); }
__host__ __device__ inline unsigned lds_u8(unsigned addr, const unsigned char *base) { return base[addr]; }
template <int OFF> inline unsigned lds_b32(unsigned addr, const unsigned char *base) { return *(const unsigned *)(base + addr + OFF); }
template <int OFF> inline uint2 lds_b64(unsigned addr, const unsigned char *base) { return *(const uint2 *)(base + addr + OFF); }
template <int OFF> inline uint4 lds_b128(unsigned addr, const unsigned char *base) { return *(const uint4 *)(base + addr + OFF); }
__host__ __device__ inline void fence_async_smem() {}
#endif

} // namespace sa

#endif
