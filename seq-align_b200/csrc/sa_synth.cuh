/*
 * sa_synth.cuh -- counter-based synthetic pair generator.
 *
 * The reference has no generator (it reads FASTA/FASTQ files,
 * src/alignment_cmdline.c:578-640); BASELINE's configs are "synthetic DNA /
 * protein pairs", so the benchmark and the multi-GPU job need one that any
 * rank can evaluate for any pair range without talking to anybody:
 *
 *   u(seed, pair, stream, pos) = splitmix64(seed ^ pair*0x9E3779B97F4A7C15
 *                                           ^ stream*0xBF58476D1CE4E5B9 ^ pos)
 *
 *   seq_a[pos]  = alphabet[u(.., 0, pos) % k]
 *   seq_b       = seq_a read through a mutation channel, position by position
 *                 (read pointer i starts at 0; r = u(.., 1, j) for output j):
 *                   t = r & 0xffff
 *                   t <  T_indel            insertion: emit alphabet[(r>>40) % k], i stays
 *                   t <  2*T_indel          deletion:  i += 1, then copy as below
 *                   copy: c = src(i); i += 1;
 *                         if ((r>>16) & 0xffff) < T_sub: c = (c + 1 + ((r>>32) & 0xff) % (k-1)) % k
 *                 src(i) = seq_a code for i < len_a, else the fresh letter
 *                 u(.., 2, i) % k (padding when seq_b outruns seq_a)
 *   DNA:     alphabet ACGT,                  T_sub = 0.05, T_indel = 0.01 (x 65536, rounded)
 *   protein: alphabet ARNDCQEGHILKMFPSTWYV,  T_sub = 0.15, T_indel = 0.02
 *
 * seqalign/synth.py holds the same function in numpy (the CPU arm of the
 * benchmark and the tests use it); tests/test_synth.py pins the two to each
 * other.  One thread per pair (seq_b is a sequential channel); set-up code,
 * not on any timed path.  All `% k` are exact on the full 64-bit value
 * (folded through 2^32 mod k so the device only divides 32-bit numbers).
 */
#ifndef SA_SYNTH_CUH
#define SA_SYNTH_CUH

#include "sa_platform.h"

namespace sa {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ uint64_t synth_u(uint64_t seed, uint64_t pair, uint64_t stream, uint64_t pos)
{
  return splitmix64(seed ^ (pair * 0x9E3779B97F4A7C15ull) ^ (stream * 0xBF58476D1CE4E5B9ull) ^ pos);
}

/* u % k for k < 2^16 with 32-bit divisions only: u = hi*2^32 + lo */
__host__ __device__ __forceinline__ unsigned synth_mod(uint64_t u, unsigned k)
{
  const unsigned p32 = (unsigned)(0x100000000ull % k);
  const unsigned hi = (unsigned)(u >> 32) % k, lo = (unsigned)u % k;
  return (hi * p32 + lo) % k;
}

struct SynthArgs {
  uint64_t seed;
  int64_t first_pair, npairs;
  int len_a, len_b;
  int k;                 /* alphabet size: 4 or 20 */
  unsigned t_sub, t_indel;
  uint8_t *seq_a, *seq_b;   /* npairs*len_a / npairs*len_b bytes, pair i at i*len */
};

__device__ __forceinline__ uint8_t synth_letter(int k, unsigned code)
{
  /* "ACGT" / "ARNDCQEGHILKMFPSTWYV" */
  const unsigned long long dna = 0x54474341ull;   /* 'A','C','G','T' little endian */
  if(k == 4) return (uint8_t)(dna >> (8 * code));
  const char *prot = "ARNDCQEGHILKMFPSTWYV";
  return (uint8_t)prot[code];
}

__global__ void __launch_bounds__(128)
synth_kernel(const SynthArgs A)
{
  const int k = A.k;
  for(int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < A.npairs; p += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t pair = (uint64_t)(A.first_pair + p);
    uint8_t *oa = A.seq_a + p * A.len_a, *ob = A.seq_b + p * A.len_b;
    for(int x = 0; x < A.len_a; x++) oa[x] = synth_letter(k, synth_mod(synth_u(A.seed, pair, 0, (uint64_t)x), (unsigned)k));
    int64_t i = 0;
    for(int j = 0; j < A.len_b; j++) {
      const uint64_t r = synth_u(A.seed, pair, 1, (uint64_t)j);
      const unsigned t = (unsigned)(r & 0xffff);
      unsigned c;
      if(t < A.t_indel) {
        c = (unsigned)(r >> 40) % (unsigned)k;
      } else {
        if(t < 2 * A.t_indel) i++;
        c = synth_mod(synth_u(A.seed, pair, i < A.len_a ? 0 : 2, (uint64_t)i), (unsigned)k);
        i++;
        if((unsigned)((r >> 16) & 0xffff) < A.t_sub) c = (c + 1 + ((unsigned)(r >> 32) & 0xffu) % (unsigned)(k - 1)) % (unsigned)k;
      }
      ob[j] = synth_letter(k, c);
    }
  }
}

} // namespace sa

#endif
