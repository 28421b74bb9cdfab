#!/usr/bin/env python3
"""The classic single-pair API on the GPU box: one 10k x 10k pair through
needleman_wunsch_align (deferred matrices: one fill), the same with the reference's eager matrices
(SEQALIGN_EAGER_MATRICES=1), and smith_waterman_align + a loop of fetches on a 2k x 2k pair (every hit
from the device's list), each against the oracle.   python tools/gpu_classic.py >> profiles/classic_api_r02.jsonl"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import *

if len(sys.argv) > 1 and sys.argv[1] == "eager":
    os.environ["SEQALIGN_EAGER_MATRICES"] = "1"
L = 10000
sc = scoring_from_spec(SPECS["free_ends"])
A, OA, B, OB = synthetic_batch(3, 2, L, L, block=16)
a, b = A[:L].tobytes(), B[:L].tobytes()
seqalign.needleman_wunsch(a[:500], b[:500], sc)      # context, module load
best = 1e9
for rep in range(3):
    t = time.time(); al = seqalign.needleman_wunsch(a, b, sc); best = min(best, time.time() - t)
o = orc_from_scoring(sc)
rc, es, ea, eb = orc_nw(o, a, b)
print(json.dumps(dict(what="needleman_wunsch_align, one pair %dx%d, free end gaps" % (L, L), eager_matrices=os.environ.get("SEQALIGN_EAGER_MATRICES") == "1",
                      seconds=round(best, 4), equals_oracle=bool((al.score, al.result_a, al.result_b) == (es, ea, eb)))), flush=True)
if os.environ.get("SEQALIGN_EAGER_MATRICES") != "1":
    # SW: all hits of a 2k x 2k related pair, hit by hit
    sw = scoring_from_spec(SPECS["sw_cli"])
    a2, b2 = a[:2000], b[:2000]
    t = time.time(); hits = seqalign.smith_waterman(a2, b2, sw, max_hits=200); dt = time.time() - t
    nw_, want = orc_sw_hits(orc_from_scoring(sw), a2, b2, 200)
    same = len(hits) == len(want) and all((h.score, h.result_a, h.result_b, h.pos_a, h.pos_b) == (w["score"], w["result_a"], w["result_b"], w["pos_a"], w["pos_b"]) for h, w in zip(hits, want))
    print(json.dumps(dict(what="smith_waterman_align + %d fetches, one pair 2000x2000" % len(hits), seconds=round(dt, 4), hits=len(hits),
                          equals_oracle=bool(same))), flush=True)
    p = subprocess.run([sys.executable, __file__, "eager"], capture_output=True, text=True)
    sys.stdout.write(p.stdout)
