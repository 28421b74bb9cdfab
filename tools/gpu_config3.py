#!/usr/bin/env python3
"""BASELINE config 3 at its stated size: NW, 10,000 synthetic DNA pairs of 10,000 x 10,000, 1/-2/-4/-1,
--freestartgap --freeendgap, score + traceback (both gapped strings of every pair), 1 x B200.

    python tools/gpu_config3.py [pairs] [length] [oracle_pairs]  >> profiles/config3_r02.jsonl

1e12 cells.  The flag bytes (1 B/cell, 100 MB per pair) do not fit HBM for the whole batch, so
SEQALIGN_MODE_ALIGN runs it in waves (fill + walk + strings to the host per wave).  Reported: kernel
and end-to-end TCUPS, walk time; parity: every score against the score-only kernel, and
`oracle_pairs` pairs (default 16) against the oracle -- score and both strings."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from helpers import *

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
n_orc = int(sys.argv[3]) if len(sys.argv) > 3 else 16
sc = scoring_from_spec(SPECS["free_ends"])
eng = seqalign.BatchAligner(0, sc)
t = time.time()
da = torch.empty(n * L, dtype=torch.uint8, device="cuda:0"); db = torch.empty(n * L, dtype=torch.uint8, device="cuda:0")
seqalign.synth_device(0, "dna", 3, 0, n, L, L, da.data_ptr(), db.data_ptr())
A, B = da.cpu().numpy(), db.cpu().numpy()
del da, db
OA = np.arange(n + 1, dtype=np.int64) * L; OB = OA.copy()
gen_s = time.time() - t
cells = n * L * L
out = dict(what="config 3: NW %d pairs %dx%d, free start/end gaps, score + traceback" % (n, L, L), cells=cells, gen_s=gen_s)
t = time.time(); eng.submit_packed(seqalign.NW, seqalign.MODE_SCORE, A, OA, B, OB); dt = time.time() - t
s_score = eng.scores().copy()
out.update(score_kernel=eng.last_kernel, score_kernel_ms=eng.last_kernel_ms, score_tcups_kernel=cells / eng.last_kernel_ms / 1e9, score_tcups_e2e=cells / dt / 1e12)
t = time.time(); eng.submit_packed(seqalign.NW, seqalign.MODE_ALIGN, A, OA, B, OB); dt = time.time() - t
out.update(align_kernel=eng.last_kernel, align_kernel_ms=eng.last_kernel_ms, walk_ms=eng.last_walk_ms, launches=eng.last_launches,
           align_tcups_kernel=cells / eng.last_kernel_ms / 1e9, align_tcups_kernel_plus_walk=cells / (eng.last_kernel_ms + eng.last_walk_ms) / 1e9,
           align_tcups_e2e=cells / dt / 1e12, align_e2e_s=dt)
s_align = eng.scores()
out["all_scores_equal_score_only_kernel"] = bool(np.array_equal(s_align, s_score))
out["score_checksum"] = int(s_align.astype(np.int64).sum())
# every string pair spells its inputs (cheap, all pairs)
spell = 0
for i in range(0, n, max(1, n // 200)):
    al = eng.alignment(i)
    spell += int(al.result_a.replace(b"-", b"") == A[i * L:(i + 1) * L].tobytes() and al.result_b.replace(b"-", b"") == B[i * L:(i + 1) * L].tobytes() and al.score == s_score[i])
out["strings_spell_inputs_sampled"] = "%d / %d" % (spell, len(range(0, n, max(1, n // 200))))
# oracle on a fixed sample spread over the batch
o = orc_from_scoring(sc)
ok, t = 0, time.time()
idx = [int(v) for v in np.linspace(0, n - 1, n_orc).astype(int)] if n_orc else []
for i in idx:
    rc, es, ea, eb = orc_nw(o, A[i * L:(i + 1) * L].tobytes(), B[i * L:(i + 1) * L].tobytes())
    al = eng.alignment(i)
    ok += int((al.score, al.result_a, al.result_b) == (es, ea, eb))
out.update(oracle_pairs=len(idx), oracle_pairs_equal=ok, oracle_s=time.time() - t)
print(json.dumps(out), flush=True)
