#!/usr/bin/env python3
"""Multi-hit mode throughput next to the reference's smith_waterman_align2 + fetch loop."""
import os, sys, time, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import *
eng = seqalign.BatchAligner(0, seqalign.Scoring.sw_cli_default())
n = 50000
A, OA, B, OB = synthetic_batch(2, n, 150, 150)
cells = n * 22500.0
for max_hits, min_score in ((1, 60), (8, 60), (8, 1), (64, 1)):
    eng.set_hit_limits(max_hits, min_score)
    for r in range(2):
        t = time.time(); eng.submit_packed(seqalign.SW, seqalign.MODE_HITS, A, OA, B, OB); dt = time.time() - t
    nh = sum(len(eng.hits(i)) for i in range(0, n, 500))
    print(json.dumps(dict(what="SW hits 50k x 150x150", max_hits=max_hits, min_score=min_score, kernel=eng.last_kernel,
                          kernel_ms=round(eng.last_kernel_ms, 3), gcups_kernel=round(cells / eng.last_kernel_ms / 1e6, 1),
                          e2e_ms=round(dt * 1e3, 1), gcups_e2e=round(cells / dt / 1e9, 1), hits_in_sample_of_100=nh)), flush=True)
