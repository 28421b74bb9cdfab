#!/usr/bin/env python3
"""Differential fuzzing of every mode of the engine against the oracle: random scoring models
(init values, built-in systems with poked gap penalties, flags), random ragged batches up to the
kernels' shape limits, all modes.  python tools/gpu_fuzz.py [seconds] [seed]
FUZZ_MODES=4 (comma list of mode numbers), FUZZ_NW=1 (NW only) and FUZZ_SMALL=1 (short batches) narrow a run, e.g.
for a dry run in the lane emulator (SEQALIGN_LIB=tests/emu/libseqalign_emu.so)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import *
from seqalign import NW, SW, MODE_SCORE, MODE_ALIGN, MODE_HITS, MODE_MATS, MODE_SCORE_ONLY
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
eng = seqalign.BatchAligner(0)
t_end = time.time() + budget
stats = dict(rounds=0, pairs=0, kernels={}, failures=[])
PROT = b"ARNDCQEGHILKMFPSTWYV"
MODES = [int(x) for x in os.environ.get("FUZZ_MODES", "0,1,3,4,2").split(",")]
ONLY_NW = bool(os.environ.get("FUZZ_NW"))
SMALL = bool(os.environ.get("FUZZ_SMALL"))

def model():
    r = rng.random()
    if r < 0.55:
        match = int(rng.integers(1, 7)); mism = int(rng.integers(-6, 1))
        go = int(rng.integers(-8, 1)); ge = int(rng.integers(-4, 1))
        flags = [False] * 6
        f = rng.random()
        if f < 0.10: flags[0] = True                      # free start gap
        elif f < 0.20: flags[1] = True                    # free end gap
        elif f < 0.30: flags[0] = flags[1] = True
        elif f < 0.34: flags[2] = True                    # no gaps in a
        elif f < 0.38: flags[3] = True
        elif f < 0.42: flags[4] = True                    # no mismatches
        sc = seqalign.Scoring(match, mism, go, ge, *flags)
        desc = dict(init=[match, mism, go, ge] + [int(x) for x in flags])
        if rng.random() < 0.15: sc.add_wildcard("N", int(rng.integers(-2, 2))); desc["wild"] = True
        return sc, desc, b"ACGTN" if desc.get("wild") else b"ACGT"
    name = ["BLOSUM62", "PAM30", "PAM70", "BLOSUM80"][int(rng.integers(0, 4))]
    sc = seqalign.Scoring.system(name)
    desc = dict(system=name)
    if rng.random() < 0.5:
        go, ge = int(rng.integers(-14, 1)), int(rng.integers(-3, 1))
        sc.poke(gap_open=go, gap_extend=ge); desc["poke"] = [go, ge]
    return sc, desc, PROT

def unsafe(sc, algo):
    s = sc.s
    # upstream signed overflow: NW with both no_gaps flags, or penalties beyond the stale min_penalty
    if algo == NW and (s.no_gaps_in_a or s.no_gaps_in_b or s.no_mismatches): return True
    room = abs(s.min_penalty)
    return algo == NW and (-(s.gap_open + s.gap_extend) > room or -s.gap_extend > room)

while time.time() < t_end:
    sc, desc, alpha = model()
    maxlen = int(([10, 20, 40, 70, 100, 140] if SMALL else [20, 60, 150, 300, 512, 700, 1500])[int(rng.integers(0, 6 if SMALL else 7))])
    n = int(rng.integers(3, 12)) if SMALL else int(rng.integers(3, 400 if maxlen <= 150 else 60 if maxlen <= 700 else 12))
    uniform = rng.random() < 0.15 and maxlen <= 512     # one shape for the whole batch: the packed 16-bit kernel's case
    if uniform:
        la_u, lb_u = int(rng.integers(1, maxlen + 1)), int(rng.integers(1, maxlen + 1))
        ua, uoa, ub, uob = synthetic_batch(int(rng.integers(0, 1 << 30)), n, la_u, lb_u, kind="protein" if alpha == PROT else "dna")
        sa = [ua[i * la_u:(i + 1) * la_u].tobytes() for i in range(n)]; sb = [ub[i * lb_u:(i + 1) * lb_u].tobytes() for i in range(n)]
    else:
        sa, sb = ragged_batch(int(rng.integers(0, 1 << 30)), n, maxlen, maxlen, alphabet=alpha, min_len=int(rng.integers(0, 2)))
    a, oa = seqalign.pack(sa); b, ob = seqalign.pack(sb)
    o = orc_from_scoring(sc)
    eng.set_scoring(sc)
    algo = NW if ONLY_NW or rng.random() >= 0.6 else SW
    mode = MODES[int(rng.integers(0, len(MODES)))]
    # NW batch matrices: the engine itself has to refuse what it cannot reproduce (stale min_penalty, restriction flags)
    if unsafe(sc, algo) and not (algo == NW and mode == MODE_MATS): algo = SW
    if uniform and rng.random() < 0.9 and not os.environ.get("FUZZ_MODES"): algo, mode = SW, (MODE_SCORE_ONLY if rng.random() < 0.5 else MODE_SCORE)
    if algo == NW and mode == MODE_HITS: mode = MODE_ALIGN
    eng.force_general(1 if rng.random() < 0.1 and mode in (MODE_SCORE, MODE_ALIGN) else 0)
    eng.set_hit_limits(8, 1)
    case = dict(model=desc, algo="SW" if algo == SW else "NW", mode=mode, n=n, maxlen=maxlen)
    try:
        try:
            eng.submit_packed(algo, mode, a, oa, b, ob)
        except seqalign.SeqAlignError as e:
            if e.code == seqalign.ERR_ARG and mode in (MODE_HITS, MODE_MATS):
                continue          # shapes outside the specialised kernels are refused by design
            raise
        k = eng.last_kernel
        stats["kernels"][k] = stats["kernels"].get(k, 0) + 1
        if algo == SW: es, ex, ey = orc_batch_sw(o, a, oa, b, ob)
        else: es, ex, ey = orc_batch_nw(o, a, oa, b, ob), None, None
        if mode in (MODE_SCORE, MODE_SCORE_ONLY, MODE_ALIGN, MODE_MATS):
            assert np.array_equal(eng.scores(), es), "scores"
        if mode == MODE_SCORE and algo == SW:
            s_, x_, y_ = eng.ends()
            assert np.array_equal(x_, ex) and np.array_equal(y_, ey), "end cells"
        pick = rng.choice(n, size=min(n, 12), replace=False)
        for i in pick:
            i = int(i)
            if mode == MODE_ALIGN:
                al = eng.alignment(i)
                if algo == NW:
                    rc, s1, ra, rb = orc_nw(o, sa[i], sb[i])
                    assert rc == 0 and (al.score, al.result_a, al.result_b) == (s1, ra, rb), "NW alignment %d" % i
                else:
                    nh, hits = orc_sw_hits(o, sa[i], sb[i], 1)
                    if nh == 0: assert al is None, "SW alignment %d should be empty" % i
                    else:
                        h = hits[0]
                        assert (al.score, al.result_a, al.result_b, al.pos_a, al.pos_b) == (h["score"], h["result_a"], h["result_b"], h["pos_a"], h["pos_b"]), "SW alignment %d" % i
            elif mode == MODE_HITS:
                nh, hits = orc_sw_hits(o, sa[i], sb[i], 8)
                got = eng.hits(i)
                assert [(h.score, h.result_a, h.result_b, h.pos_a, h.pos_b) for h in got] == [(h["score"], h["result_a"], h["result_b"], h["pos_a"], h["pos_b"]) for h in hits], "hits %d" % i
            elif mode == MODE_MATS:
                m, ga, gb = eng.matrices(i, len(sa[i]), len(sb[i]))
                rc, em, ega, egb = orc_fill(o, sa[i], sb[i], algo == SW)
                assert np.array_equal(m, em) and np.array_equal(ga, ega) and np.array_equal(gb, egb), "matrices %d" % i
        stats["rounds"] += 1; stats["pairs"] += n
    except AssertionError as e:
        case["what"] = str(e); case["kernel"] = eng.last_kernel
        stats["failures"].append(case)
        print("FAIL", json.dumps(case), flush=True)
        if len(stats["failures"]) > 10: break
print(json.dumps(stats))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(stats, open(os.path.join(ROOT, "gpurun_out", "fuzz_seed%d.json" % seed), "w"), indent=1)
sys.exit(1 if stats["failures"] else 0)
