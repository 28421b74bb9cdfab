#!/usr/bin/env python3
"""One launch of every kernel family on its BASELINE shape, for a single ncu capture
(ncu -k regex:... --set full python tools/gpu_prof_all.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import *
specs = scoring_specs()
eng = seqalign.BatchAligner(0)
A, OA, B, OB = synthetic_batch(2, 100000, 150, 150)
eng.set_scoring(specs["sw_cli"]())
eng.force_general(3); eng.submit_packed(seqalign.SW, seqalign.MODE_SCORE, A, OA, B, OB); print(eng.last_kernel, eng.last_kernel_ms)
eng.force_general(0); eng.submit_packed(seqalign.SW, seqalign.MODE_SCORE, A, OA, B, OB); print(eng.last_kernel, eng.last_kernel_ms)
eng.force_general(5); eng.submit_packed(seqalign.SW, seqalign.MODE_SCORE, A, OA, B, OB); print(eng.last_kernel, eng.last_kernel_ms)
eng.force_general(0)
n = 20000
eng.submit_packed(seqalign.SW, seqalign.MODE_MATS, A[:150 * n], OA[:n + 1], B[:150 * n], OB[:n + 1]); print(eng.last_kernel, eng.last_kernel_ms)
eng.submit_packed(seqalign.SW, seqalign.MODE_ALIGN, A[:150 * n], OA[:n + 1], B[:150 * n], OB[:n + 1]); print(eng.last_kernel, eng.last_kernel_ms, eng.last_walk_ms)
eng.set_hit_limits(8, 60)
eng.submit_packed(seqalign.SW, seqalign.MODE_HITS, A[:150 * n], OA[:n + 1], B[:150 * n], OB[:n + 1]); print(eng.last_kernel, eng.last_kernel_ms)
PA, POA, PB, POB = synthetic_batch(4, 20000, 400, 400, kind="protein")
eng.set_scoring(specs["blosum62"]())
eng.submit_packed(seqalign.SW, seqalign.MODE_ALIGN, PA, POA, PB, POB); print(eng.last_kernel, eng.last_kernel_ms, eng.last_walk_ms)
eng.force_general(3); eng.submit_packed(seqalign.SW, seqalign.MODE_SCORE, PA, POA, PB, POB); print(eng.last_kernel, eng.last_kernel_ms)
eng.force_general(0)
L = 10000; n = 296
eng.set_scoring(scoring_from_spec(SPECS["free_ends"]))
LA, LOA, LB, LOB = synthetic_batch(3, n, L, L, block=16)
eng.submit_packed(seqalign.NW, seqalign.MODE_SCORE, LA, LOA, LB, LOB); print(eng.last_kernel, eng.last_kernel_ms)
eng.submit_packed(seqalign.NW, seqalign.MODE_ALIGN, LA, LOA, LB, LOB); print(eng.last_kernel, eng.last_kernel_ms, eng.last_walk_ms)
# ---- round 2 ----
# SW beyond 512 columns: score + end cell, and the first hit through checkpoints + recompute walk
WA, WOA, WB, WOB = synthetic_batch(2040, 500, 2000, 2000)
eng.set_scoring(specs["sw_cli"]())
eng.submit_packed(seqalign.SW, seqalign.MODE_SCORE, WA, WOA, WB, WOB); print(eng.last_kernel, eng.last_kernel_ms)
eng.submit_packed(seqalign.SW, seqalign.MODE_ALIGN, WA, WOA, WB, WOB); print(eng.last_kernel, eng.last_kernel_ms, eng.last_walk_ms)
# NW materialise (packed scans are the default now)
eng.set_scoring(specs["nw_default"]())
n = 20000
eng.submit_packed(seqalign.NW, seqalign.MODE_MATS, A[:150 * n], OA[:n + 1], B[:150 * n], OB[:n + 1]); print(eng.last_kernel, eng.last_kernel_ms)
# a FASTA file decoded on the device (66 MB of text) and aligned in place
import ctypes
m = 200000
names = np.frombuffer(b"".join(b">read%09d\n" % i for i in range(2 * m)), np.uint8).reshape(2 * m, 15)
seqs = np.frombuffer(b"ACGT", np.uint8)[np.random.default_rng(1).integers(0, 4, size=(2 * m, 150), dtype=np.uint8)]
text = np.concatenate([names, seqs, np.full((2 * m, 1), 10, np.uint8)], axis=1).tobytes()
lib = seqalign.load()
hp = lib.seqalign_host_alloc(len(text) + 64)
ctypes.memmove(hp, text, len(text))
rd = seqalign.Reads(0)
assert lib.seqalign_reads_decode(rd._h, hp, len(text), 1, 1) == 0
print("decode", rd.records, rd.last_ms)
eng.set_scoring(specs["sw_cli"]())
eng.submit_reads(seqalign.SW, seqalign.MODE_SCORE, rd, 0, rd, 1, m); print(eng.last_kernel, eng.last_kernel_ms)
