#!/usr/bin/env python3
"""Differential fuzzing of the SW end cell (score, x_end, y_end) against the oracle under scoring models whose
scores pass 1024, the range where the packed kernel's 16-bit keys turn relative (fast16_kernel ENDS == 2):
random match / mismatch / gap values around fast_plan's bound, uniform and ragged batches, unrelated pairs,
pairs with indels, homopolymers, copies with one substitution.
    python tools/gpu_fuzz_ends.py [seed] [rounds] [max length] [pairs per round]
Runs on the GPU library or, with SEQALIGN_LIB=tests/emu/libseqalign_emu.so, in the lane emulator (which also
traps on a key outside its ten bits)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "seq-align_b200"))
import numpy as np, seqalign
from seqalign import SW, MODE_SCORE
from helpers import orc_batch_sw, orc_from_scoring, scoring_from_spec
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
eng = seqalign.BatchAligner(0)
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
maxlen = int(sys.argv[3]) if len(sys.argv) > 3 else 200
maxpairs = int(sys.argv[4]) if len(sys.argv) > 4 else 8
pairs = 0
used = {}
L = np.frombuffer(b"ACGT", dtype=np.uint8)
def mutate(a, lb):
    out = []
    i = 0
    while len(out) < lb and i < len(a):
        r = rng.random()
        if r < 0.04: i += int(rng.integers(1, 6)); continue
        if r < 0.08: out.extend(L[rng.integers(0, 4, size=int(rng.integers(1, 6)))]); continue
        out.append(a[i] if rng.random() > 0.08 else L[rng.integers(0, 4)]); i += 1
    out = np.array(out[:lb], dtype=np.uint8)
    if len(out) < lb: out = np.concatenate([out, L[rng.integers(0, 4, size=lb - len(out))]])
    return out
for it in range(iters):
    match = int(rng.integers(8, 40)); mism = -int(rng.integers(0, 100))
    go = -int(rng.integers(0, 14)); ge = -int(rng.integers(0, 8))
    sc = scoring_from_spec(dict(init=[match, mism, go, ge, 0, 0, 0, 0, 0, 0]))
    n = int(rng.integers(1, maxpairs + 1))
    uniform = rng.random() < 0.5
    ma, mb = int(rng.integers(20, maxlen)), int(rng.integers(20, maxlen))
    sa, sb = [], []
    for p in range(n):
        la = ma if uniform else int(rng.integers(0, ma + 1)); lb = mb if uniform else int(rng.integers(0, mb + 1))
        kind = rng.integers(0, 4)
        if kind == 0: a = L[rng.integers(0, 4, size=la)]; b = L[rng.integers(0, 4, size=lb)]
        elif kind == 1: a = L[rng.integers(0, 4, size=la)]; b = mutate(a, lb)
        elif kind == 2: a = np.full(la, 65, np.uint8); b = np.full(lb, 65, np.uint8)
        else:
            a = L[rng.integers(0, 4, size=la)]; b = np.resize(a, lb) if la else L[rng.integers(0, 4, size=lb)]
            if lb > 10: b = b.copy(); b[int(rng.integers(0, lb))] = 67
        sa.append(a.astype(np.uint8).tobytes()); sb.append(np.asarray(b, np.uint8).tobytes())
    a, oa = seqalign.pack(sa); b, ob = seqalign.pack(sb)
    eng.set_scoring(sc); eng.force_general(0)
    eng.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
    k = eng.last_kernel; used[k] = used.get(k, 0) + 1; pairs += n
    s, x, y = eng.ends()
    es, ex, ey = orc_batch_sw(orc_from_scoring(sc), a, oa, b, ob)
    if not (np.array_equal(s, es) and np.array_equal(x, ex) and np.array_equal(y, ey)):
        print("MISMATCH", it, k, (match, mism, go, ge), n, uniform, ma, mb, s, es, x, ex, y, ey); sys.exit(1)
print(json.dumps(dict(what="end-cell fuzz", rounds=iters, pairs=pairs, max_length=maxlen, kernels=used, mismatches=0)))
