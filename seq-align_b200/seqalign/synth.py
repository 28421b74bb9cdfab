"""Counter-based synthetic pairs -- numpy mirror of csrc/sa_synth.cuh.

u(seed, pair, stream, pos) = splitmix64(seed ^ pair*0x9E3779B97F4A7C15 ^ stream*0xBF58476D1CE4E5B9 ^ pos)

seq_a[pos] = alphabet[u(.., 0, pos) % k]; seq_b is seq_a read through a
mutation channel (substitutions, insertions, deletions; fresh letters from
stream 2 once the read pointer runs past seq_a), see the header of
sa_synth.cuh for the exact rule.  A pair depends only on (seed, pair index):
any rank can make any shard of a job, on the device
(``seqalign.synth_device``, the CUDA kernel) or on the host (this module:
the CPU arm of bench.py, the oracle side of the tests).  The reference has
no generator (it reads files, src/alignment_cmdline.c:578-640).
"""
import numpy as np

DNA = np.frombuffer(b"ACGT", dtype=np.uint8)
PROTEIN = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
KINDS = {"dna": (0, DNA, 3277, 655), "protein": (1, PROTEIN, 9830, 1311)}

_G = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(x):
    with np.errstate(over="ignore"):
        z = x + _G
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def synth_u(seed, pair, stream, pos):
    """pair, pos: uint64 arrays (broadcast against each other)"""
    with np.errstate(over="ignore"):
        return splitmix64(np.uint64(seed) ^ (pair * _G) ^ (np.uint64(stream) * _M1) ^ pos)


def synth_batch(seed, first_pair, npairs, len_a, len_b, kind="dna", block=8192):
    """pairs [first_pair, first_pair+npairs) as a packed batch (seq_a, off_a, seq_b, off_b)"""
    _, alphabet, t_sub, t_indel = KINDS[kind]
    k = np.uint64(len(alphabet))
    A = np.empty((npairs, len_a), dtype=np.uint8)
    B = np.empty((npairs, len_b), dtype=np.uint8)
    for b0 in range(0, npairs, block):
        m = min(block, npairs - b0)
        pair = (np.arange(m, dtype=np.uint64) + np.uint64(first_pair + b0))[:, None]
        ca = (synth_u(seed, pair, 0, np.arange(len_a, dtype=np.uint64)[None, :]) % k).astype(np.int64)
        A[b0:b0 + m] = alphabet[ca]
        i = np.zeros(m, dtype=np.int64)
        pcol = pair[:, 0]
        rows = np.arange(m)
        cb = np.empty((m, len_b), dtype=np.int64)
        for j in range(len_b):
            r = synth_u(seed, pcol, 1, np.uint64(j))
            t = (r & np.uint64(0xffff)).astype(np.int64)
            ins = t < t_indel
            dele = (~ins) & (t < 2 * t_indel)
            i = i + dele
            inside = i < len_a
            src = np.where(inside, ca[rows, np.minimum(i, max(len_a - 1, 0))] if len_a else 0, 0)
            if not inside.all():
                fresh = (synth_u(seed, pcol, 2, i.astype(np.uint64)) % k).astype(np.int64)
                src = np.where(inside, src, fresh)
            sub = ((r >> np.uint64(16)) & np.uint64(0xffff)).astype(np.int64) < t_sub
            shift = (((r >> np.uint64(32)) & np.uint64(0xff)) % (k - np.uint64(1))).astype(np.int64)
            copy = np.where(sub, (src + 1 + shift) % int(k), src)
            fresh_ins = ((r >> np.uint64(40)) % k).astype(np.int64)
            cb[:, j] = np.where(ins, fresh_ins, copy)
            i = i + (~ins)
        B[b0:b0 + m] = alphabet[cb]
    off_a = np.arange(npairs + 1, dtype=np.int64) * len_a
    off_b = np.arange(npairs + 1, dtype=np.int64) * len_b
    return A.reshape(-1), off_a, B.reshape(-1), off_b
