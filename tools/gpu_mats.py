#!/usr/bin/env python3
"""Materialise mode (SURVEY 8d mode M, 12 B/cell): the batch kernel against the HBM roofline, and
the single-pair aligner_align() path for comparison."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import *
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6650.0
eng = seqalign.BatchAligner(0)
rows = []
def batch(tag, scoring, n, L, kind):
    a, oa, b, ob = synthetic_batch(2, n, L, L, kind=kind)
    eng.set_scoring(scoring)
    best = 1e9
    for r in range(3):
        eng.submit_packed(seqalign.SW, seqalign.MODE_MATS, a, oa, b, ob)
        best = min(best, eng.last_kernel_ms)
    cells = n * L * L
    bytes_written = 12 * n * (L + 1) * (L + 1)
    # spot check of one pair against the oracle
    o = orc_from_scoring(scoring)
    i = n // 2
    m, ga, gb = eng.matrices(i, L, L)
    rc, em, ega, egb = orc_fill(o, a[i * L:(i + 1) * L].tobytes(), b[i * L:(i + 1) * L].tobytes(), True)
    ok = bool(np.array_equal(m, em) and np.array_equal(ga, ega) and np.array_equal(gb, egb))
    row = dict(what=tag, pairs=n, kernel=eng.last_kernel, kernel_ms=round(best, 3), gcups=round(cells / best / 1e6, 1),
               gb_written=round(bytes_written / 1e9, 2), gbs=round(bytes_written / best / 1e6, 1), hbm_peak_gbs=HBM,
               hbm_frac=round(bytes_written / best / 1e6 / HBM, 3), pair_checked_vs_oracle=ok)
    print(json.dumps(row), flush=True); rows.append(row)
batch("SW DNA 150x150, 2/-2/-2/-1", seqalign.Scoring.sw_cli_default(), 100000, 150, "dna")
batch("SW DNA 150x150, 2/-2/-2/-1 (20k pairs)", seqalign.Scoring.sw_cli_default(), 20000, 150, "dna")
batch("SW protein 400x400, BLOSUM62", scoring_specs()["blosum62"](), 20000, 400, "protein")
batch("SW DNA 500x500", seqalign.Scoring.sw_cli_default(), 10000, 500, "dna")
# single pair through the literal aligner_align contract (general kernel, matrices to the host)
for L in (2000, 10000):
    a, oa, b, ob = synthetic_batch(3, 1, L, L, block=16)
    eng.set_scoring(seqalign.Scoring.sw_cli_default())
    t = time.time(); eng.fill_matrices(a.tobytes(), b.tobytes(), True); dt = time.time() - t
    t = time.time(); eng.fill_matrices(a.tobytes(), b.tobytes(), True); dt = time.time() - t
    row = dict(what="single pair %dx%d, seqalign_fill_matrices" % (L, L), kernel=eng.last_kernel, kernel_ms=round(eng.last_kernel_ms, 2),
               wall_s=round(dt, 3), gcups_kernel=round(L * L / eng.last_kernel_ms / 1e6, 2), gbs_kernel=round(12 * L * L / eng.last_kernel_ms / 1e6, 1),
               hbm_frac=round(12 * L * L / eng.last_kernel_ms / 1e6 / HBM, 4))
    print(json.dumps(row), flush=True); rows.append(row)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "gpu_mats.json"), "w"), indent=1)
