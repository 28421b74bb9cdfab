#!/usr/bin/env python3
"""Short run of the wide-pair kernel for ncu: NW score, then NW score+traceback, n pairs of L x L."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import *
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
L = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
algo = seqalign.SW if len(sys.argv) > 3 and sys.argv[3] == "sw" else seqalign.NW
sc = scoring_from_spec(SPECS["free_ends" if algo == seqalign.NW else "sw_cli"])
eng = seqalign.BatchAligner(0, sc)
A, OA, B, OB = synthetic_batch(3, n, L, L, block=16)
for mode in (seqalign.MODE_SCORE, seqalign.MODE_ALIGN):
    eng.submit_packed(algo, mode, A, OA, B, OB)
    print(eng.last_kernel, eng.last_kernel_ms, n * L * L / eng.last_kernel_ms / 1e6, "GCUPS")
