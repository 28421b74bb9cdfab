#!/usr/bin/env python3
"""Kernel-time survey on the GPU box: headline shapes x kernel variants."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import *

specs = scoring_specs()
eng = seqalign.BatchAligner(0)
rows = []

def timeit(tag, name, algo, force, a, oa, b, ob, cells, reps=4, check=None):
    sc = specs[name](); eng.set_scoring(sc); eng.force_general(force)
    best = 1e9; e2e = 1e9
    for r in range(reps):
        t = time.time()
        eng.submit_packed(algo, seqalign.MODE_SCORE, a, oa, b, ob)
        e2e = min(e2e, time.time() - t)
        best = min(best, eng.last_kernel_ms)
    s = eng.scores()
    ok = None if check is None else bool(np.array_equal(s, check))
    res = dict(tag=tag, scoring=name, algo="SW" if algo else "NW", kernel=eng.last_kernel, kernel_ms=round(best, 4),
               gcups_kernel=round(cells / best / 1e6, 1), e2e_ms=round(e2e * 1e3, 3), gcups_e2e=round(cells / e2e / 1e9, 1), same_scores=ok)
    print(json.dumps(res), flush=True)
    rows.append(res)
    eng.force_general(0)
    return s

A, OA, B, OB = synthetic_batch(2, 100000, 150, 150)
c = 100000 * 22500.0
ref = timeit("dna150 auto (end cell, packed 16-bit keys)", "sw_cli", seqalign.SW, 0, A, OA, B, OB, c)
timeit("dna150 end cell, int32 tree", "sw_cli", seqalign.SW, 5, A, OA, B, OB, c, check=ref)
timeit("dna150 end,column", "sw_cli", seqalign.SW, 2, A, OA, B, OB, c, check=ref)
timeit("dna150 score-only s16x2", "sw_cli", seqalign.SW, 3, A, OA, B, OB, c, check=ref)
timeit("dna150 score-only int32", "sw_cli", seqalign.SW, 4, A, OA, B, OB, c, check=ref)
timeit("dna150 general", "sw_cli", seqalign.SW, 1, A, OA, B, OB, c, reps=2, check=ref)
timeit("dna150 nw", "nw_default", seqalign.NW, 0, A, OA, B, OB, c)
timeit("dna150 libdefault sw", "nw_default", seqalign.SW, 0, A, OA, B, OB, c)
PA, POA, PB, POB = synthetic_batch(4, 50000, 400, 400, kind="protein")
c4 = 50000 * 160000.0
ref = timeit("prot400 auto", "blosum62", seqalign.SW, 0, PA, POA, PB, POB, c4)
timeit("prot400 end cell, int32 tree", "blosum62", seqalign.SW, 5, PA, POA, PB, POB, c4, check=ref)
timeit("prot400 score-only s16x2", "blosum62", seqalign.SW, 3, PA, POA, PB, POB, c4, check=ref)
timeit("prot400 score-only int32", "blosum62", seqalign.SW, 4, PA, POA, PB, POB, c4, check=ref)
for L in (64, 100, 128, 250, 300, 512):
    n = int(2.0e9 / (L * L))
    a, oa, b, ob = synthetic_batch(20 + L, n, L, L)
    timeit("dna%d score-only" % L, "sw_cli", seqalign.SW, 3, a, oa, b, ob, n * L * L)
# Smith-Waterman beyond 512 columns: strip-pipelined kernel (was the general kernel in round 1)
for L, n in ((2000, 500), (5000, 100)):
    a, oa, b, ob = synthetic_batch(40 + L, n, L, L)
    ref = timeit("dna%d SW score + end cell (wide)" % L, "sw_cli", seqalign.SW, 0, a, oa, b, ob, n * L * L, reps=3)
    if L == 2000:
        timeit("dna%d SW general kernel" % L, "sw_cli", seqalign.SW, 1, a, oa, b, ob, n * L * L, reps=1, check=ref)
# align mode (score + end cell + traceback strings on the host)
def align(tag, name, algo, a, oa, b, ob, cells, walk=None):
    if walk: os.environ["SEQALIGN_WALK"] = walk
    else: os.environ.pop("SEQALIGN_WALK", None)
    sc = specs[name](); eng.set_scoring(sc)
    best = None
    for r in range(2):
        t = time.time(); eng.submit_packed(algo, seqalign.MODE_ALIGN, a, oa, b, ob); dt = time.time() - t
        if best is None or dt < best[0]: best = (dt, eng.last_kernel_ms, eng.last_walk_ms)
    res = dict(tag=tag, scoring=name, kernel=eng.last_kernel, kernel_ms=round(best[1], 3), walk_ms=round(best[2], 3),
               gcups_kernel=round(cells / best[1] / 1e6, 1), e2e_ms=round(best[0] * 1e3, 2), gcups_e2e=round(cells / best[0] / 1e9, 1))
    print(json.dumps(res), flush=True)
    rows.append(res)
n = 20000
align("dna150 SW align 20k", "sw_cli", seqalign.SW, A[:150 * n], OA[:n + 1], B[:150 * n], OB[:n + 1], n * 22500.0)
align("dna150 NW align 20k", "nw_default", seqalign.NW, A[:150 * n], OA[:n + 1], B[:150 * n], OB[:n + 1], n * 22500.0)
align("dna150 NW align 20k tiled walk", "nw_default", seqalign.NW, A[:150 * n], OA[:n + 1], B[:150 * n], OB[:n + 1], n * 22500.0, walk="tiled")
align("prot400 SW align 50k (config 4), thread walk", "blosum62", seqalign.SW, PA, POA, PB, POB, c4, walk="thread")
align("prot400 SW align 50k (config 4), tiled walk", "blosum62", seqalign.SW, PA, POA, PB, POB, c4, walk="tiled")
# wide SW pairs with traceback (round 2: checkpoints + recompute walk; was the general kernel)
a, oa, b, ob = synthetic_batch(2040, 500, 2000, 2000)
align("dna2000 SW align 500 (wide)", "sw_cli", seqalign.SW, a, oa, b, ob, 500 * 2000.0 * 2000.0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "gpu_perf.json"), "w"), indent=1)
