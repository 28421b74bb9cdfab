#!/usr/bin/env python3
"""one launch of the batch materialise kernel (for ncu): 20k DNA pairs 150x150"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import *
eng = seqalign.BatchAligner(0, seqalign.Scoring.sw_cli_default())
a, oa, b, ob = synthetic_batch(2, 20000, 150, 150)
for r in range(2):
    eng.submit_packed(seqalign.SW, seqalign.MODE_MATS, a, oa, b, ob)
    print(eng.last_kernel, eng.last_kernel_ms)
