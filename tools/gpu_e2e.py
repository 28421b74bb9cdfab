#!/usr/bin/env python3
"""Where does end-to-end time go?  pinned H2D bandwidth vs submit() wall time per chunk count."""
import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import *
A, OA, B, OB = synthetic_batch(2, 100000, 150, 150)
pa, pb = torch.from_numpy(A).pin_memory(), torch.from_numpy(B).pin_memory()
poa, pob = torch.from_numpy(OA).pin_memory(), torch.from_numpy(OB).pin_memory()
d = torch.empty(len(A), dtype=torch.uint8, device="cuda")
for _ in range(3): d.copy_(pa, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(20): d.copy_(pa, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 20
print("pinned H2D 15 MB: %.3f ms  (%.1f GB/s)" % (dt * 1e3, len(A) / dt / 1e9))
if len(sys.argv) > 1:
    os.environ["SEQALIGN_CHUNKS"] = sys.argv[1]
eng = seqalign.BatchAligner(0, seqalign.Scoring.sw_cli_default())
for mode in (seqalign.MODE_SCORE_ONLY, seqalign.MODE_SCORE):
    for _ in range(3): eng.submit_ptrs(seqalign.SW, mode, pa.data_ptr(), poa.data_ptr(), pb.data_ptr(), pob.data_ptr(), 100000)
    t = time.perf_counter()
    for _ in range(20): eng.submit_ptrs(seqalign.SW, mode, pa.data_ptr(), poa.data_ptr(), pb.data_ptr(), pob.data_ptr(), 100000)
    dt = (time.perf_counter() - t) / 20
    print("chunks=%s mode=%d submit: %.3f ms -> %.0f GCUPS  (kernel sum %.3f ms, %s)" % (os.environ.get("SEQALIGN_CHUNKS", "auto"), mode, dt * 1e3, 2250 / dt / 1e3, eng.last_kernel_ms, eng.last_kernel))
