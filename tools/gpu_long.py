#!/usr/bin/env python3
"""BASELINE config 3 shape (NW 10k x 10k, free start/end gaps): score and score+traceback, a few pairs."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import *
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
sc = scoring_from_spec(SPECS["free_ends"])
o = orc_from_scoring(sc)
eng = seqalign.BatchAligner(0, sc)
t = time.time(); A, OA, B, OB = synthetic_batch(3, n, L, L, block=16); print("gen %.1fs" % (time.time() - t), flush=True)
cells = n * L * L
for rep in range(2):
    t = time.time(); eng.submit_packed(seqalign.NW, seqalign.MODE_SCORE, A, OA, B, OB); dt = time.time() - t
    print(json.dumps(dict(what="NW score %dx%d x%d" % (L, L, n), kernel=eng.last_kernel, kernel_ms=eng.last_kernel_ms, gcups_kernel=cells / eng.last_kernel_ms / 1e6, e2e_s=dt)), flush=True)
s_score = eng.scores().copy()
for rep in range(2):
    t = time.time(); eng.submit_packed(seqalign.NW, seqalign.MODE_ALIGN, A, OA, B, OB); dt = time.time() - t
    print(json.dumps(dict(what="NW align %dx%d x%d" % (L, L, n), kernel=eng.last_kernel, kernel_ms=eng.last_kernel_ms, walk_ms=eng.last_walk_ms, gcups_kernel=cells / eng.last_kernel_ms / 1e6, gcups_e2e=cells / dt / 1e9, e2e_s=dt)), flush=True)
assert (eng.scores() == s_score).all()
# parity of two pairs against the oracle (score + strings); the oracle needs 3 x 400 MB per pair
for i in range(2):
    a = A[i * L:(i + 1) * L].tobytes(); b = B[i * L:(i + 1) * L].tobytes()
    t = time.time(); rc, es, ea, eb = orc_nw(o, a, b); dt = time.time() - t
    al = eng.alignment(i)
    print("pair %d oracle %.1fs: score %d vs %d, strings equal: %s" % (i, dt, es, al.score, (al.result_a, al.result_b) == (ea, eb)), flush=True)
# a single big pair through the classic API
t = time.time(); al = seqalign.needleman_wunsch(A[:L].tobytes(), B[:L].tobytes(), sc); dt = time.time() - t
print("classic needleman_wunsch_align %dx%d (matrices materialised + traceback): %.2fs score %d" % (L, L, dt, al.score))
os.environ["SEQALIGN_SKIP_MATRICES"] = "1"
t = time.time(); al = seqalign.needleman_wunsch(A[:L].tobytes(), B[:L].tobytes(), sc); dt = time.time() - t
print("classic needleman_wunsch_align, SEQALIGN_SKIP_MATRICES=1: %.3fs score %d" % (dt, al.score))
