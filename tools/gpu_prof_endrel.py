#!/usr/bin/env python3
"""Two launches of the packed kernel with relative end-cell keys on BASELINE config 4's shape (protein 400x400, BLOSUM62),
for an ncu capture: ncu -k regex:fast16_kernel --set full python tools/gpu_prof_endrel.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import *
eng = seqalign.BatchAligner(0)
PA, POA, PB, POB = synthetic_batch(4, 20000, 400, 400, kind="protein")
eng.set_scoring(scoring_specs()["blosum62"]())
for _ in range(2):
    eng.submit_packed(seqalign.SW, seqalign.MODE_SCORE, PA, POA, PB, POB); print(eng.last_kernel, eng.last_kernel_ms)
