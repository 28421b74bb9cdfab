#!/usr/bin/env python3
"""Occupancy sweep of the headline kernel: pad shared memory per CTA, time the 100k-pair launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import *
os.environ["SEQALIGN_CHUNKS"] = "1"
eng = seqalign.BatchAligner(0, seqalign.Scoring.sw_cli_default())
A, OA, B, OB = synthetic_batch(2, 100000, 150, 150)
for pad in (0, 2000, 12000, 31000, 70000):
    os.environ["SEQALIGN_FAST_PAD_SMEM"] = str(pad)
    best = 1e9
    for r in range(5):
        eng.submit_packed(seqalign.SW, seqalign.MODE_SCORE_ONLY, A, OA, B, OB)
        best = min(best, eng.last_kernel_ms)
    print("pad %6d B: kernel %.4f ms -> %.0f GCUPS (%s)" % (pad, best, 2250 / best, eng.last_kernel), flush=True)
