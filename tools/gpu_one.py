#!/usr/bin/env python3
"""Run the headline batch a few times (for ncu captures): python tools/gpu_one.py [force] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import *
force = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = seqalign.BatchAligner(0, seqalign.Scoring.sw_cli_default())
eng.force_general(force)
A, OA, B, OB = synthetic_batch(2, 100000, 150, 150)
for r in range(reps):
    eng.submit_packed(seqalign.SW, seqalign.MODE_SCORE, A, OA, B, OB)
    print(eng.last_kernel, eng.last_kernel_ms, 2250.0 / eng.last_kernel_ms, "TCUPS")
