#!/usr/bin/env python3
"""Where does end-to-end time go?  pinned H2D bandwidth vs submit() wall time per chunk count,
and PipelinedAligner throughput per depth."""
import os, sys, time
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
print("pinned H2D 15 MB: %.3f ms  (%.1f GB/s) -> a 31.6 MB step needs %.3f ms of PCIe" % (dt * 1e3, len(A) / dt / 1e9, dt * 1e3 * 31.6 / 15), flush=True)
ptrs = (pa.data_ptr(), poa.data_ptr(), pb.data_ptr(), pob.data_ptr())
eng = seqalign.BatchAligner(0, seqalign.Scoring.sw_cli_default())
for chunks in ("1", "2", "3", "4", "6", "8"):
    os.environ["SEQALIGN_CHUNKS"] = chunks
    for mode in (seqalign.MODE_SCORE_ONLY,):
        for _ in range(3): eng.submit_ptrs(seqalign.SW, mode, *ptrs, 100000)
        t = time.perf_counter()
        for _ in range(20): eng.submit_ptrs(seqalign.SW, mode, *ptrs, 100000)
        dt = (time.perf_counter() - t) / 20
        print("chunks=%s mode=%d submit: %.3f ms -> %.0f GCUPS  (kernel sum %.3f ms, %s)" % (chunks, mode, dt * 1e3, 2250 / dt / 1e3, eng.last_kernel_ms, eng.last_kernel), flush=True)
eng.close()
for chunks in ("2", "4", "8"):
    os.environ["SEQALIGN_CHUNKS"] = chunks
    for depth in (1, 2, 3, 4):
        pipe = seqalign.PipelinedAligner(0, seqalign.Scoring.sw_cli_default(), depth=depth)
        pend = []
        for k in range(depth + 2): pend.append(pipe.submit_ptrs(seqalign.SW, seqalign.MODE_SCORE_ONLY, *ptrs, 100000))
        for f in pend: f.result()
        t = time.perf_counter(); pend = []
        for k in range(40):
            pend.append(pipe.submit_ptrs(seqalign.SW, seqalign.MODE_SCORE_ONLY, *ptrs, 100000))
            if len(pend) > depth: pend.pop(0).result()
        for f in pend: f.result()
        dt = (time.perf_counter() - t) / 40
        print("pipelined chunks=%s depth=%d: %.3f ms/step -> %.0f GCUPS" % (chunks, depth, dt * 1e3, 2250 / dt / 1e3), flush=True)
        pipe.close()
