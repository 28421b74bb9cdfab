#!/usr/bin/env python3
"""Ragged-batch throughput (VERDICT r1 next #4a): 100k DNA pairs with lengths uniform in [100,150], SW 2/-2/-2/-1,
score mode; which kernel the plan takes and GCUPS over the real cells (sum la*lb)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import *

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rng = np.random.default_rng(5)
la = rng.integers(100, 151, size=n); lb = rng.integers(100, 151, size=n)
oa = np.zeros(n + 1, np.int64); ob = np.zeros(n + 1, np.int64)
np.cumsum(la, out=oa[1:]); np.cumsum(lb, out=ob[1:])
A = rng.integers(0, 4, size=int(oa[-1]), dtype=np.uint8); A = np.frombuffer(b"ACGT", np.uint8)[A]
B = rng.integers(0, 4, size=int(ob[-1]), dtype=np.uint8); B = np.frombuffer(b"ACGT", np.uint8)[B]
cells = float((la * lb).sum())
eng = seqalign.BatchAligner(0, scoring_from_spec(SPECS["sw_cli"]))
for force, tag in ((0, "auto"), (5, "int32 tree"), (3, "score-only s16x2"), (4, "score-only int32")):
    eng.force_general(force)
    best = 1e9
    for r in range(4):
        eng.submit_packed(seqalign.SW, seqalign.MODE_SCORE, A, oa, B, ob)
        best = min(best, eng.last_kernel_ms)
    s = eng.scores().copy()
    if force == 0: ref = s
    print(json.dumps(dict(what="ragged 100-150 bp SW score", pairs=n, variant=tag, kernel=eng.last_kernel, kernel_ms=round(best, 4),
                          gcups_real_cells=round(cells / best / 1e6, 1), same_scores=bool(np.array_equal(s, ref)))), flush=True)
# oracle on a sample
o = orc_from_scoring(scoring_from_spec(SPECS["sw_cli"]))
m = 2000
es = orc_batch_sw(o, np.ascontiguousarray(A[:oa[m]]), np.ascontiguousarray(oa[:m + 1]), np.ascontiguousarray(B[:ob[m]]), np.ascontiguousarray(ob[:m + 1]))
print(json.dumps(dict(what="ragged oracle sample", pairs=m, equal=bool(np.array_equal(ref[:m], es[0])))), flush=True)
