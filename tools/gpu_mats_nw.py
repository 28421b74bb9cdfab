#!/usr/bin/env python3
"""Needleman-Wunsch in batch materialise mode (mats_kernel<NW>): parity against the oracle's fill over every
block count of the kernel, then the kernel against the HBM roofline (12 B/cell written)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import *
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6650.0
eng = seqalign.BatchAligner(0)
rows = []
out = os.path.join(ROOT, "gpurun_out", "gpu_mats_nw.json")
os.makedirs(os.path.dirname(out), exist_ok=True)

def parity(name, alphabet):
    sa, sb = ragged_batch(77, 40, 300, 300, alphabet=alphabet)
    for w in (1, 31, 32, 33, 63, 64, 65, 96, 127, 128, 160, 200, 255, 256, 300, 320, 416, 511):
        x, y = ragged_batch(w, 1, w, w, alphabet=alphabet, min_len=w)
        sa.append(x[0]); sb.append(y[0][:50])
    sa += [b"", sa[0]]; sb += [sb[0], b""]
    sc = scoring_from_spec(SPECS[name]); o = orc_from_scoring(sc); eng.set_scoring(sc)
    eng.submit(seqalign.NW, seqalign.MODE_MATS, sa, sb)
    scores = eng.scores()
    bad = 0
    for i, (a, b) in enumerate(zip(sa, sb)):
        m, ga, gb = eng.matrices(i, len(a), len(b))
        rc, em, ega, egb = orc_fill(o, a, b, False)
        if not (np.array_equal(m, em) and np.array_equal(ga, ega) and np.array_equal(gb, egb)
                and scores[i] == max(em[-1, -1], ega[-1, -1], egb[-1, -1])):
            bad += 1
    row = dict(what="parity NW MODE_MATS vs oracle fill, " + name, kernel=eng.last_kernel, pairs=len(sa), mismatching_pairs=bad)
    print(json.dumps(row), flush=True); rows.append(row)

def batch(tag, scoring, n, L, kind):
    n = max(4, n // int(os.environ.get("MATS_NW_SHRINK", "1")))   # emulator dry runs
    a, oa, b, ob = synthetic_batch(2, n, L, L, kind=kind)
    eng.set_scoring(scoring)
    o = orc_from_scoring(scoring)
    for algo, is_sw, nw_pack in ((seqalign.NW, False, ""), (seqalign.NW, False, "1"), (seqalign.SW, True, "")):
        # the opt-in packed 16-bit scans of the NW rows beside the default int32 ones
        if nw_pack: os.environ.pop("SEQALIGN_MATS_NOPACK", None)
        else: os.environ["SEQALIGN_MATS_NOPACK"] = "1"
        best = 1e9
        for r in range(3):
            eng.submit_packed(algo, seqalign.MODE_MATS, a, oa, b, ob)
            best = min(best, eng.last_kernel_ms)
        bytes_written = 12 * n * (L + 1) * (L + 1)
        i = n // 2
        m, ga, gb = eng.matrices(i, L, L)
        rc, em, ega, egb = orc_fill(o, a[i * L:(i + 1) * L].tobytes(), b[i * L:(i + 1) * L].tobytes(), is_sw)
        ok = bool(np.array_equal(m, em) and np.array_equal(ga, ega) and np.array_equal(gb, egb))
        row = dict(what=("SW " if is_sw else "NW ") + tag, pairs=n, kernel=eng.last_kernel, kernel_ms=round(best, 3),
                   gcups=round(n * L * L / best / 1e6, 1), gb_written=round(bytes_written / 1e9, 2),
                   gbs=round(bytes_written / best / 1e6, 1), hbm_peak_gbs=HBM,
                   hbm_frac=round(bytes_written / best / 1e6 / HBM, 3), pair_checked_vs_oracle=ok)
        print(json.dumps(row), flush=True); rows.append(row)
        json.dump(rows, open(out, "w"), indent=1)

parity("nw_default", b"ACGT")
parity("blosum62", b"ARNDCQEGHILKMFPSTWYV")
parity("big_scores", b"ACGT")
json.dump(rows, open(out, "w"), indent=1)
batch("DNA 150x150, default scoring", scoring_from_spec(SPECS["nw_default"]), 50000, 150, "dna")
batch("protein 400x400, BLOSUM62", scoring_specs()["blosum62"](), 10000, 400, "protein")
