#!/usr/bin/env python3
"""First GPU contact: microbench, a parity sweep against the oracle, first timings."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import *

out = {}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
try:
    mb = subprocess.run([os.path.join(ROOT, "bin", "microbench")], capture_output=True, text=True, timeout=120)
    print(mb.stdout); print(mb.stderr)
    open(os.path.join(ROOT, "gpurun_out", "microbench.jsonl"), "w").write(mb.stdout)
except Exception as e:
    print("microbench failed", e)

specs = scoring_specs()
eng = seqalign.BatchAligner(0)

def parity_scores(name, algo, fg, a, oa, b, ob):
    sc = specs[name](); o = orc_from_scoring(sc)
    eng.set_scoring(sc); eng.force_general(fg)
    eng.submit_packed(algo, seqalign.MODE_SCORE, a, oa, b, ob)
    s, x, y = eng.ends()
    if algo == seqalign.SW:
        es, ex, ey = orc_batch_sw(o, a, oa, b, ob)
        ok = bool((s == es).all() and (x == ex).all() and (y == ey).all())
    else:
        es = orc_batch_nw(o, a, oa, b, ob); ok = bool((s == es).all())
    print("parity", name, "SW" if algo else "NW", eng.last_kernel, "OK" if ok else "MISMATCH", flush=True)
    return ok

allok = True
a, oa, b, ob = synthetic_batch(2, 2000, 150, 150)
for name in ("sw_cli", "nw_default"):
    for algo in (seqalign.SW, seqalign.NW):
        for fg in (False, True):
            allok &= parity_scores(name, algo, fg, a, oa, b, ob)
pa, poa, pb, pob = synthetic_batch(4, 300, 400, 400, kind="protein")
for algo in (seqalign.SW, seqalign.NW):
    for fg in (False, True):
        allok &= parity_scores("blosum62", algo, fg, pa, poa, pb, pob)
sa, sb = ragged_batch(7, 300, 300, 300, alphabet=b"ACGTNacgtn")
ra, roa = seqalign.pack(sa); rb, rob = seqalign.pack(sb)
for name in ("wild_n", "free_ends", "no_mismatch_wild", "mutations", "big_scores", "linear_gap"):
    for algo in (seqalign.SW, seqalign.NW):
        allok &= parity_scores(name, algo, False, ra, roa, rb, rob)

# align mode
for name, algo in (("sw_cli", seqalign.SW), ("nw_default", seqalign.NW), ("free_ends", seqalign.NW)):
    sc = specs[name](); o = orc_from_scoring(sc); eng.set_scoring(sc); eng.force_general(False)
    sa2, sb2 = sa[:100], sb[:100]
    eng.submit(algo, seqalign.MODE_ALIGN, sa2, sb2)
    bad = 0
    for i, (x, y) in enumerate(zip(sa2, sb2)):
        al = eng.alignment(i)
        if algo == seqalign.NW:
            rc, es, ea, eb = orc_nw(o, x, y)
            ok = al is not None and (al.score, al.result_a, al.result_b) == (es, ea, eb)
        else:
            n, hits = orc_sw_hits(o, x, y, 1)
            ok = (al is None) if n == 0 else (al is not None and (al.score, al.result_a, al.result_b, al.pos_a, al.pos_b) ==
                                              (hits[0]["score"], hits[0]["result_a"], hits[0]["result_b"], hits[0]["pos_a"], hits[0]["pos_b"]))
        bad += not ok
    print("align parity", name, "bad=%d" % bad, eng.last_kernel, flush=True)
    allok &= bad == 0
# matrices
sc = specs["nw_default"](); o = orc_from_scoring(sc); eng.set_scoring(sc)
for is_sw in (0, 1):
    for x, y in list(zip(sa, sb))[:10] + [(b"A" * 700, b"ACGT" * 100)]:
        m, ga, gb = eng.fill_matrices(x, y, is_sw)
        rc, em, ega, egb = orc_fill(o, x, y, is_sw)
        ok = bool((m == em).all() and (ga == ega).all() and (gb == egb).all())
        allok &= ok
        if not ok: print("mats MISMATCH", is_sw, len(x), len(y))
print("PARITY", "ALL OK" if allok else "FAILURES", flush=True)
out["parity_ok"] = allok

# timings
def timeit(name, algo, fg, a, oa, b, ob, cells, reps=5):
    sc = specs[name](); eng.set_scoring(sc); eng.force_general(fg)
    best = 1e9; e2e = 1e9
    for r in range(reps):
        t = time.time()
        eng.submit_packed(algo, seqalign.MODE_SCORE, a, oa, b, ob)
        e2e = min(e2e, time.time() - t)
        best = min(best, eng.last_kernel_ms)
    res = dict(workload=name, algo="SW" if algo else "NW", kernel=eng.last_kernel, kernel_ms=best,
               gcups_kernel=cells / best / 1e6, e2e_ms=e2e * 1e3, gcups_e2e=cells / e2e / 1e9)
    print(json.dumps(res), flush=True)
    return res

t = time.time()
A, OA, B, OB = synthetic_batch(2, 100000, 150, 150)
print("gen 100k pairs: %.1fs" % (time.time() - t))
out["timings"] = []
out["timings"].append(timeit("sw_cli", seqalign.SW, False, A, OA, B, OB, 100000 * 22500.0))
out["timings"].append(timeit("sw_cli", seqalign.SW, True, A, OA, B, OB, 100000 * 22500.0, reps=2))
out["timings"].append(timeit("nw_default", seqalign.NW, False, A, OA, B, OB, 100000 * 22500.0))
PA, POA, PB, POB = synthetic_batch(4, 20000, 400, 400, kind="protein")
out["timings"].append(timeit("blosum62", seqalign.SW, False, PA, POA, PB, POB, 20000 * 160000.0))
# align-mode timing (general dir + walk)
sc = specs["sw_cli"](); eng.set_scoring(sc); eng.force_general(False)
t = time.time(); eng.submit_packed(seqalign.SW, seqalign.MODE_ALIGN, A[:150 * 20000], OA[:20001], B[:150 * 20000], OB[:20001]); dt = time.time() - t
print(json.dumps(dict(workload="sw_cli align 20k", e2e_ms=dt * 1e3, kernel_ms=eng.last_kernel_ms, gcups_kernel=20000 * 22500 / eng.last_kernel_ms / 1e6)))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gpu_first.json"), "w"), indent=1)
