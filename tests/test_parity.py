"""Parity of the CUDA path with the oracle, through the C-ABI.

Every test here goes product library -> oracle comparison, bit-exact (int32
scores, end cells, alignment strings, matrices).  On the GPU box these run
on cuda:0 against libseqalign_b200.so at full sizes (marked `gpu` by
conftest); in the build container the same tests drive the same kernel
sources through the lane emulator at small sizes.
"""
import json
import os
import zlib

import numpy as np
import pytest

import seqalign
from seqalign import NW, SW, MODE_SCORE, MODE_ALIGN, MODE_HITS, MODE_MATS
from helpers import (ROOT, SPECS, orc_batch_nw, orc_batch_sw, orc_fill, orc_from_scoring, orc_nw,
                     orc_sw_hits, ragged_batch, scoring_from_spec, synthetic_batch)

pytestmark = pytest.mark.parity

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))
PROTEIN_SPECS = ("blosum62", "pam30", "pam70", "blosum80")
# no_gaps_* with unequal lengths are fine one flag at a time; both flags
# together overflow in the reference itself and are excluded
SWEEP = [n for n in SPECS]


def _h(name):
    return zlib.crc32(name.encode())


def _alphabet(name):
    if name in PROTEIN_SPECS:
        return b"ARNDCQEGHILKMFPSTWYVBZX"
    if name == "dna_hyb":
        return b"ACGTacgt"
    return b"ACGTNacgtn"


def _score_check(engine, sc, algo, a, oa, b, ob, general):
    o = orc_from_scoring(sc)
    engine.set_scoring(sc)
    engine.force_general(general)
    engine.submit_packed(algo, MODE_SCORE, a, oa, b, ob)
    s, x, y = engine.ends()
    if algo == SW:
        es, ex, ey = orc_batch_sw(o, a, oa, b, ob)
        assert np.array_equal(s, es), np.nonzero(s != es)[0][:8]
        assert np.array_equal(x, ex) and np.array_equal(y, ey)
    else:
        es = orc_batch_nw(o, a, oa, b, ob)
        assert np.array_equal(s, es), np.nonzero(s != es)[0][:8]
        assert np.array_equal(x, np.diff(oa)) and np.array_equal(y, np.diff(ob))
    assert engine.last_launches >= 1
    kernel = engine.last_kernel
    if algo == SW and not general:
        # the other two end-cell strategies of the specialised kernel
        engine.force_general(2)
        engine.submit_packed(algo, MODE_SCORE, a, oa, b, ob)
        s2, x2, y2 = engine.ends()
        assert np.array_equal(s2, es) and np.array_equal(x2, ex) and np.array_equal(y2, ey)
        for mode in (3, 4):   # score only: packed 16-bit kernel if the batch qualifies / int32 only
            engine.force_general(mode)
            engine.submit_packed(algo, MODE_SCORE, a, oa, b, ob)
            assert np.array_equal(engine.scores(), es), (mode, engine.last_kernel)
        engine.force_general(0)
    return kernel


@pytest.mark.parametrize("algo", [SW, NW], ids=["sw", "nw"])
@pytest.mark.parametrize("name", SWEEP)
def test_scores_ragged(engine, big, name, algo):
    """ragged lengths (including empty sequences), every scoring variant,
    specialised and general kernels"""
    n, maxlen = (300, 220) if big else (20, 44)
    sa, sb = ragged_batch(_h(name) % 1000, n, maxlen, maxlen, alphabet=_alphabet(name))
    a, oa = seqalign.pack(sa)
    b, ob = seqalign.pack(sb)
    sc = scoring_from_spec(SPECS[name])
    _score_check(engine, sc, algo, a, oa, b, ob, general=False)
    _score_check(engine, sc, algo, a, oa, b, ob, general=True)


def test_headline_config_sample(engine, big):
    """BASELINE config 2 shape: SW, DNA 150x150, smith_waterman CLI scoring"""
    n = 4000 if big else 6
    a, oa, b, ob = synthetic_batch(2, n, 150, 150)
    sc = scoring_from_spec(SPECS["sw_cli"])
    k = _score_check(engine, sc, SW, a, oa, b, ob, general=False)
    assert k == "fast16_sw_score_end"     # uniform batch, scores below 1024: end cells from the packed kernel
    engine.force_general(5)               # the int32 tree kernel on the same batch
    engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
    assert engine.last_kernel == "fast_sw_score_end"
    es, ex, ey = orc_batch_sw(orc_from_scoring(sc), a, oa, b, ob)
    s5, x5, y5 = engine.ends()
    assert np.array_equal(s5, es) and np.array_equal(x5, ex) and np.array_equal(y5, ey)
    engine.force_general(3)
    engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
    assert engine.last_kernel == "fast16_sw_score"
    engine.force_general(0)
    # unrelated pairs too (throughput must not depend on the data; scores do)
    a, oa, b, ob = synthetic_batch(12, n // 2, 150, 150, related=False)
    _score_check(engine, sc, SW, a, oa, b, ob, general=False)
    # library-default scoring on the same shape
    _score_check(engine, scoring_from_spec(SPECS["nw_default"]), SW, a, oa, b, ob, general=False)


@pytest.mark.parametrize("la", [97, 100, 104, 150, 152, 153, 300, 304])
def test_fast16_tight_shapes(engine, big, la):
    """packed 16-bit kernel on the shapes cut for common read lengths (8x13, 8x19,
    16x19 lanes x columns): odd pair counts leave the high half of a register empty"""
    n, lb = (333, la) if big else (5, 36)
    a, oa, b, ob = synthetic_batch(900 + la, n, la, lb)
    sc = scoring_from_spec(SPECS["sw_cli"])
    es, _, _ = orc_batch_sw(orc_from_scoring(sc), a, oa, b, ob)
    engine.set_scoring(sc)
    engine.force_general(3)
    engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
    assert engine.last_kernel == "fast16_sw_score"
    assert np.array_equal(engine.scores(), es)
    engine.force_general(0)
    # with end cells: the same shapes through the 16-bit keys (ties between equal scores are
    # frequent on unrelated pairs: x ascending first, then y)
    for related in (True, False):
        a, oa, b, ob = synthetic_batch(1900 + la, n, la, lb, related=related)
        es, ex, ey = orc_batch_sw(orc_from_scoring(sc), a, oa, b, ob)
        engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
        assert engine.last_kernel == "fast16_sw_score_end"
        s, x, y = engine.ends()
        assert np.array_equal(s, es) and np.array_equal(x, ex) and np.array_equal(y, ey), (la, related)


@pytest.mark.parametrize("spec,kind", [
    (dict(init=[30, -20, -4, -1, 0, 0, 0, 0, 0, 0]), "dna"),          # identical stretches of 36+ letters pass 1024
    (dict(init=[25, -3, -11, -1, 0, 0, 0, 0, 0, 0]), "dna"),          # cheap mismatches, dear gaps
    (dict(init=[22, -17, 0, 0, 0, 0, 0, 0, 0, 0]), "dna"),            # free gaps: B = 0
    (dict(system="BLOSUM62", poke=dict(gap_open=-11, gap_extend=-1)), "protein"),
])
def test_packed_end_cells_relative_keys(engine, big, spec, kind):
    """SW end cells from the packed kernel when (score + |open|) passes 1024: the 16-bit keys become relative
    to the value entering the lane on the row (fast16_kernel ENDS == 2).  Uniform and ragged batches,
    related pairs (the best grows row after row) and unrelated ones (ties: x ascending, then y)."""
    sc = scoring_from_spec(spec)
    o = orc_from_scoring(sc)
    engine.set_scoring(sc)
    engine.force_general(0)
    shapes = ((400, 400), (150, 150), (247, 300), (512, 90)) if big else ((120, 100), (97, 96))
    n = 600 if big else 5
    seen = set()
    for la, lb in shapes:
        for related in (True, False):
            a, oa, b, ob = synthetic_batch(4100 + la, n, la, lb, kind=kind, related=related)
            es, ex, ey = orc_batch_sw(o, a, oa, b, ob)
            engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
            seen.add(engine.last_kernel)
            if kind == "protein" and (la, lb) == shapes[0]:
                assert engine.last_kernel == "fast16_sw_score_endrel"     # BASELINE config 4's shape
            s, x, y = engine.ends()
            assert np.array_equal(s, es), (la, lb, related, np.nonzero(s != es)[0][:8])
            assert np.array_equal(x, ex) and np.array_equal(y, ey), (la, lb, related)
    # (a shape whose K breaks fast_plan's bound, or a batch that stays below 1024, goes to the kernels it went to before)
    assert "fast16_sw_score_endrel" in seen and seen <= {"fast16_sw_score_endrel", "fast16_sw_score_end", "fast_sw_score_end"}, seen
    # ragged: couples of different shapes, the padding code on both axes
    alphabet = b"ACGT" if kind == "dna" else b"ARNDCQEGHILKMFPSTWYV"
    sa, sb = ragged_batch(4200, 400 if big else 9, 400 if big else 110, 380 if big else 100, alphabet=alphabet)
    # one long identical pair (tryptophan scores 11 against itself, the largest BLOSUM62 entry)
    sa[0] = sb[0] = bytes((ord("W") if kind == "protein" else alphabet[i % 4]) for i in range(380 if big else 100))
    a, oa = seqalign.pack(sa)
    b, ob = seqalign.pack(sb)
    es, ex, ey = orc_batch_sw(o, a, oa, b, ob)
    engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
    assert engine.last_kernel == "fast16_sw_score_endrel", engine.last_kernel
    s, x, y = engine.ends()
    assert np.array_equal(s, es) and np.array_equal(x, ex) and np.array_equal(y, ey)
    assert es.max() >= 1024, es.max()      # the case this test is about


@pytest.mark.parametrize("name", ["free_gaps", "linear_gap", "nw_default", "blosum62"])
def test_fast16_gap_models(engine, big, name):
    """packed kernel under gap models at the edges of its "+open" trick: gap_open = gap_extend = 0
    (nothing to add, no carry between the halves -- the fuzzer's find), linear gaps, protein tables;
    scores alone and with end cells"""
    n, la, lb = (301, 150, 140) if big else (7, 30, 26)
    a, oa, b, ob = synthetic_batch(77, n, la, lb, kind="protein" if name == "blosum62" else "dna")
    sc = scoring_from_spec(SPECS[name])
    es, ex, ey = orc_batch_sw(orc_from_scoring(sc), a, oa, b, ob)
    engine.set_scoring(sc)
    engine.force_general(3)
    engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
    assert engine.last_kernel == "fast16_sw_score"
    assert np.array_equal(engine.scores(), es)
    engine.force_general(0)
    engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
    # absolute 16-bit keys need (score + |open|) < 1024: 140 x 11 with BLOSUM62 is past that, the keys turn relative
    assert engine.last_kernel == ("fast16_sw_score_endrel" if name == "blosum62" and big else "fast16_sw_score_end")
    s, x, y = engine.ends()
    assert np.array_equal(s, es) and np.array_equal(x, ex) and np.array_equal(y, ey)


def test_uniform_batch_offsets_made_on_device(engine, big):
    """a host batch of >= 4096 equal-shaped pairs does not ship its offset arrays: the engine
    generates them on the device; results must equal the ragged path's (one pair trimmed)"""
    n, la, lb = 4200, (150 if big else 7), (150 if big else 6)
    a, oa, b, ob = synthetic_batch(31, n, la, lb)
    sc = scoring_from_spec(SPECS["sw_cli"])
    es, ex, ey = orc_batch_sw(orc_from_scoring(sc), a, oa, b, ob)
    engine.set_scoring(sc)
    engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
    s, x, y = engine.ends()
    assert np.array_equal(s, es) and np.array_equal(x, ex) and np.array_equal(y, ey)
    # same sequences, last pair one base shorter: not uniform any more, offsets travel as before
    ob2 = ob.copy(); ob2[-1] -= 1
    engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob2)
    s2, _, _ = engine.ends()
    assert np.array_equal(s2[:-1], es[:-1])


def test_fast16_ragged_batches(engine, big):
    """the packed 16-bit kernel on NON-uniform batches: couples of pairs of different shapes (the
    shorter one sees padding columns / rows), empty sequences, odd pair counts; scores and end cells"""
    sc = scoring_from_spec(SPECS["sw_cli"])
    o = orc_from_scoring(sc)
    engine.set_scoring(sc)
    for seed, n, lo, hi in ((1, 2001 if big else 9, 100, 150), (2, 777 if big else 7, 0, 60), (3, 301 if big else 5, 140, 300)):
        sa, sb = ragged_batch(seed, n, hi, hi, min_len=lo)
        a, oa = seqalign.pack(sa)
        b, ob = seqalign.pack(sb)
        es, ex, ey = orc_batch_sw(o, a, oa, b, ob)
        engine.submit_packed(SW, seqalign.MODE_SCORE_ONLY, a, oa, b, ob)
        assert engine.last_kernel == "fast16_sw_score", engine.last_kernel
        assert np.array_equal(engine.scores(), es)
        engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
        assert engine.last_kernel == "fast16_sw_score_end", engine.last_kernel
        s, x, y = engine.ends()
        assert np.array_equal(s, es) and np.array_equal(x, ex) and np.array_equal(y, ey)


def test_uniform_submit_and_result_sink(engine, big):
    """seqalign_batch_submit_uniform (no offset arrays) + seqalign_batch_set_result_sink (scores written
    straight into the caller's array): same numbers as the packed submit, for SW and NW, with and without
    end cells; the engine's own accessors refuse after a submit into a sink"""
    from seqalign.synth import synth_batch
    n = 5000 if big else 9
    a, oa, b, ob = synth_batch(7, 100, n, 150 if big else 40, 140 if big else 33)
    la, lb = int(oa[1]), int(ob[1])
    for algo, spec in ((SW, "sw_cli"), (NW, "nw_default")):
        engine.set_scoring(scoring_from_spec(SPECS[spec]))
        engine.submit_packed(algo, MODE_SCORE, a, oa, b, ob)
        es, ex, ey = engine.ends()
        engine.submit_uniform_ptrs(algo, MODE_SCORE, a.ctypes.data, la, b.ctypes.data, lb, n)
        s, x, y = engine.ends()
        assert np.array_equal(s, es) and np.array_equal(x, ex) and np.array_equal(y, ey)
        sink = [np.full(n, -77, dtype=np.int32) for _ in range(3)]
        engine.set_result_sink(*[v.ctypes.data for v in sink])
        try:
            engine.submit_uniform_ptrs(algo, MODE_SCORE, a.ctypes.data, la, b.ctypes.data, lb, n)
            assert np.array_equal(sink[0], es) and np.array_equal(sink[1], ex) and np.array_equal(sink[2], ey)
            with pytest.raises(seqalign.SeqAlignError):
                engine.scores()
            sink[0][:] = -77
            engine.set_result_sink(sink[0].ctypes.data)          # scores only
            engine.submit_packed(algo, seqalign.MODE_SCORE_ONLY, a, oa, b, ob)
            assert np.array_equal(sink[0], es)
        finally:
            engine.set_result_sink(0)
        engine.submit_uniform_ptrs(algo, MODE_ALIGN, a.ctypes.data, la, b.ctypes.data, lb, min(n, 50))
        assert engine.alignment(0).score == es[0]
    assert np.array_equal(engine.scores()[: min(n, 50)], es[: min(n, 50)])


def test_protein_config_sample(engine, big):
    """BASELINE config 4 shape: SW, protein 400x400, BLOSUM62"""
    n, L = (400, 400) if big else (2, 70)
    a, oa, b, ob = synthetic_batch(4, n, L, L, kind="protein")
    sc = scoring_from_spec(SPECS["blosum62"])
    _score_check(engine, sc, SW, a, oa, b, ob, general=False)
    _score_check(engine, sc, NW, a, oa, b, ob, general=False)


@pytest.mark.parametrize("shape", [(1, 1), (1, 37), (37, 1), (255, 20), (256, 9), (257, 9), (513, 30), (700, 3), (3, 700)])
def test_shapes(engine, big, shape):
    """strip boundaries of both kernels (8/256/512 columns), degenerate shapes"""
    la, lb = shape
    if not big and la * lb > 16000:
        lb = max(1, 16000 // la)
    a, oa, b, ob = synthetic_batch(la * 1000 + lb, 5 if big else 2, la, lb)
    for name in ("sw_cli", "free_ends"):
        sc = scoring_from_spec(SPECS[name])
        for algo in (SW, NW):
            _score_check(engine, sc, algo, a, oa, b, ob, general=False)
            _score_check(engine, sc, algo, a, oa, b, ob, general=True)


@pytest.mark.parametrize("shape", [(1100, 70), (2600, 40), (40, 2600)])
def test_wide_pairs_cooperative(engine, big, shape):
    """pairs of >= 4 strips are swept by a whole CTA (warps pipelined over the
    column strips, more strips than warps): scores, tracebacks, matrices"""
    la, lb = shape
    if big:
        la, lb = la * 2, lb * 4
    a, oa, b, ob = synthetic_batch(la + lb, 3 if big else 2, la, lb)
    sa = [a[i * la:(i + 1) * la].tobytes() for i in range(len(oa) - 1)]
    sb = [b[i * lb:(i + 1) * lb].tobytes() for i in range(len(ob) - 1)]
    for name in ("free_ends", "sw_cli", "nw_default"):
        sc = scoring_from_spec(SPECS[name])
        o = orc_from_scoring(sc)
        for algo in (SW, NW):
            _score_check(engine, sc, algo, a, oa, b, ob, general=True)
            engine.submit(algo, MODE_ALIGN, sa, sb)
            assert "general_dir" in engine.last_kernel
            for i in range(len(sa)):
                _check_alignment(engine.alignment(i), algo, o, sa[i], sb[i])
        m, ga, gb = engine.fill_matrices(sa[0], sb[0], 0)
        rc, em, ega, egb = orc_fill(o, sa[0], sb[0], 0)
        assert np.array_equal(m, em) and np.array_equal(ga, ega) and np.array_equal(gb, egb)


@pytest.mark.parametrize("case", ["ragged_wide", "many_strips", "short_free_ends", "protein_wide", "tall_checkpoints"])
def test_wide_pairs_strip_pipeline(engine, big, case, monkeypatch):
    """NW beyond one strip / with free end gaps: the strip-pipelined kernel
    (sa_long.cuh).  Ragged widths (strips of several pairs interleave in a
    CTA), more than 2 x LONG_WARPS strips (edge slots are reused), the
    last-column / last-row rules of free end gaps; scores and strings.  The traceback runs both
    ways: through checkpoints + recomputed tiles (default; tall pairs cross several row
    checkpoints) and through flag bytes (SEQALIGN_LONG_FLAGS=1)."""
    alphabet = b"ACGT"
    names = ("free_ends", "nw_default", "free_end", "free_start", "linear_gap")
    if case == "ragged_wide":
        sa, sb = ragged_batch(8, 40 if big else 8, 3000 if big else 1300, 700 if big else 24, min_len=1)
    elif case == "many_strips":
        sa, sb = ragged_batch(9, 6 if big else 3, 512 * 19, 300 if big else 12, min_len=5)
        sa[0] = (sa[0] * 40)[:512 * 18 + 7]
    elif case == "short_free_ends":
        sa, sb = ragged_batch(7, 300 if big else 24, 200 if big else 60, 200 if big else 60)
        names = ("free_ends", "free_end")
    elif case == "tall_checkpoints":
        sa, sb = ragged_batch(11, 24 if big else 3, 2100 if big else 1100, 1500 if big else 200, min_len=70)
        sa[1], sb[1] = sa[1][:64 * 3], sb[1][:64 * 2]          # one strip, last row ON a checkpoint row
        # two strips and an even number of rows: an odd count of 8-byte strip-edge records sits in front of
        # the 16-byte checkpoint records (their base was once misaligned: CUDA error 716 on the device)
        xa, xb = ragged_batch(12, 1, 600, 128, min_len=128)
        sa.append((xa[0] * 8)[:600]); sb.append((xb[0] * 2)[:128])
        names = ("free_ends", "nw_default", "linear_gap")
    else:
        alphabet = b"ARNDCQEGHILKMFPSTWYVBZX"
        sa, sb = ragged_batch(10, 20 if big else 5, 2000 if big else 700, 500 if big else 30, alphabet=alphabet, min_len=1)
        names = ("blosum62", "pam30")
    a, oa = seqalign.pack(sa)
    b, ob = seqalign.pack(sb)
    engine.force_general(0)
    for name in names:
        sc = scoring_from_spec(SPECS[name])
        o = orc_from_scoring(sc)
        engine.set_scoring(sc)
        engine.submit_packed(NW, MODE_SCORE, a, oa, b, ob)
        assert engine.last_kernel == "long_nw_score", engine.last_kernel
        es = orc_batch_nw(o, a, oa, b, ob)
        s = engine.scores()
        assert np.array_equal(s, es), (name, np.nonzero(s != es)[0][:8])
        for flags, kernel in (("0", "long_nw_ckpt+walk_recompute"), ("1", "long_nw_dir+walk")):
            monkeypatch.setenv("SEQALIGN_LONG_FLAGS", flags)
            engine.submit_packed(NW, MODE_ALIGN, a, oa, b, ob)
            assert engine.last_kernel == kernel, engine.last_kernel
            assert np.array_equal(engine.scores(), es)
            for i in range(len(sa)):
                _check_alignment(engine.alignment(i), NW, o, sa[i], sb[i])


@pytest.mark.parametrize("case", ["dna_wide", "protein_wide", "ties"])
def test_wide_pairs_sw_score(engine, big, case):
    """Smith-Waterman beyond 512 columns: the strip-pipelined kernel (sa_long.cuh, IS_SW) -- score and the end
    cell under the reference's hit order (score desc, x asc, y asc), merged over the strips of a pair"""
    if case == "dna_wide":
        sa, sb = ragged_batch(31, 40 if big else 6, 3000 if big else 1200, 2500 if big else 90, min_len=1)
        names = ("sw_cli", "nw_default", "linear_gap")
    elif case == "protein_wide":
        sa, sb = ragged_batch(32, 24 if big else 4, 2200 if big else 700, 900 if big else 40,
                              alphabet=b"ARNDCQEGHILKMFPSTWYVBZX", min_len=1)
        names = ("blosum62", "pam30")
    else:
        # the same motif many times over: equal best scores in different strips, lanes and rows
        motif = b"ACGTTGCA"
        sa = [motif * (150 if big else 80), b"T" * 700 + motif + b"T" * 700, motif * 70 + b"G" + motif * 70]
        sb = [motif * 3, motif, motif * 2]
        names = ("sw_cli",)
    a, oa = seqalign.pack(sa)
    b, ob = seqalign.pack(sb)
    engine.force_general(0)
    for name in names:
        sc = scoring_from_spec(SPECS[name])
        o = orc_from_scoring(sc)
        engine.set_scoring(sc)
        engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
        assert engine.last_kernel == "long_sw_score_end", engine.last_kernel
        es, ex, ey = orc_batch_sw(o, a, oa, b, ob)
        s, x, y = engine.ends()
        assert np.array_equal(s, es), (name, np.nonzero(s != es)[0][:8])
        assert np.array_equal(x, ex) and np.array_equal(y, ey), (name, np.nonzero((x != ex) | (y != ey))[0][:8])
        # the first hit with its traceback: fill with checkpoints + best-cell tracking, recompute walk from the end cell
        engine.submit_packed(SW, MODE_ALIGN, a, oa, b, ob)
        assert engine.last_kernel == "long_sw_ckpt+walk_recompute", engine.last_kernel
        assert np.array_equal(engine.scores(), es)
        for i in range(len(sa)):
            nh, hits = orc_sw_hits(o, sa[i], sb[i], 1)
            al = engine.alignment(i)
            if nh == 0:
                assert al is None, (name, i)
            else:
                h = hits[0]
                assert (al.score, al.result_a, al.result_b, al.pos_a, al.pos_b, al.len_a, al.len_b) == \
                       (h["score"], h["result_a"], h["result_b"], h["pos_a"], h["pos_b"], h["len_a"], h["len_b"]), (name, i)


def test_length_buckets(engine, big, monkeypatch):
    """reads of different lengths through the packed kernel: pairs are counting-sorted by (shape class of
    len_a, len_b) on the device and every class gets the narrowest kernel shape that holds it
    (sa_fast.cuh "length buckets"); scores and end cells must not depend on it"""
    rng = np.random.default_rng(17)
    n = 20000 if big else 700
    if not big:
        monkeypatch.setenv("SEQALIGN_BUCKET_MIN", "64")     # the emulator's batch is small
    la = rng.integers(1, 151, size=n); lb = rng.integers(1, 151, size=n)
    la[:7] = (150, 104, 105, 64, 65, 1, 128); lb[:7] = (1, 150, 149, 150, 2, 1, 77)
    oa = np.zeros(n + 1, np.int64); ob = np.zeros(n + 1, np.int64)
    np.cumsum(la, out=oa[1:]); np.cumsum(lb, out=ob[1:])
    letters = np.frombuffer(b"ACGT", np.uint8)
    A = letters[rng.integers(0, 4, size=int(oa[-1]))]
    B = letters[rng.integers(0, 4, size=int(ob[-1]))]
    # half of the pairs related: b starts as a copy of a
    for i in range(0, n, 2):
        m = min(la[i], lb[i])
        B[ob[i]:ob[i] + m] = A[oa[i]:oa[i] + m]
    sc = scoring_from_spec(SPECS["sw_cli"])
    engine.set_scoring(sc)
    engine.force_general(0)
    o = orc_from_scoring(sc)
    m_orc = n if not big else 4000
    es, ex, ey = orc_batch_sw(o, np.ascontiguousarray(A[:oa[m_orc]]), np.ascontiguousarray(oa[:m_orc + 1]),
                              np.ascontiguousarray(B[:ob[m_orc]]), np.ascontiguousarray(ob[:m_orc + 1]))
    got = {}
    for buckets in ("1", "rows", "0"):
        monkeypatch.setenv("SEQALIGN_BUCKETS", {"1": "shapes", "rows": "rows", "0": "0"}[buckets])
        for mode in (MODE_SCORE, seqalign.MODE_SCORE_ONLY):
            engine.submit_packed(SW, mode, A, oa, B, ob)
            assert engine.last_kernel.startswith("fast16_sw_score"), engine.last_kernel
            got[(buckets, mode)] = (engine.ends(), engine.last_launches)
    (s1, x1, y1), l1 = got[("1", MODE_SCORE)]
    (s0, x0, y0), l0 = got[("0", MODE_SCORE)]
    assert l1 > l0 + 3        # several shape classes were launched
    assert np.array_equal(s1, s0) and np.array_equal(x1, x0) and np.array_equal(y1, y0)
    assert np.array_equal(s1[:m_orc], es) and np.array_equal(x1[:m_orc], ex) and np.array_equal(y1[:m_orc], ey)
    assert np.array_equal(got[("1", seqalign.MODE_SCORE_ONLY)][0][0], s0)
    (sr, xr, yr), lr = got[("rows", MODE_SCORE)]
    assert lr == l0 + 3 and np.array_equal(sr, s0) and np.array_equal(xr, x0) and np.array_equal(yr, y0)


def test_classic_sw_fetch_loop_from_device_list(engine):
    """smith_waterman_align + fetch until exhausted: the first hit from align mode, the rest from the device's
    multi-hit list, rebuilt eightfold larger each time the caller reads past its end (8 -> 64 -> 512);
    restriction shapes iterate on the host.  Every hit equals the oracle's, in order."""
    rng = np.random.default_rng(23)
    letters = np.frombuffer(b"ACGT", np.uint8)
    a = bytes(letters[rng.integers(0, 4, size=90)]); b = bytes(letters[rng.integers(0, 4, size=80)])
    for name in ("sw_cli", "no_gaps_a", "no_mismatch"):
        sc = scoring_from_spec(SPECS[name])
        hits = seqalign.smith_waterman(a, b, sc)
        n, want = orc_sw_hits(orc_from_scoring(sc), a, b, 4000)
        assert len(hits) == n and (name != "sw_cli" or n > 64), (name, len(hits), n)
        assert [(h.score, h.result_a, h.result_b, h.pos_a, h.pos_b, h.len_a, h.len_b) for h in hits] == \
               [(w["score"], w["result_a"], w["result_b"], w["pos_a"], w["pos_b"], w["len_a"], w["len_b"]) for w in want], name


def test_empty_inputs(engine):
    sc = scoring_from_spec(SPECS["nw_default"])
    engine.set_scoring(sc)
    engine.submit(NW, MODE_SCORE, [], [])
    assert len(engine.scores()) == 0
    engine.submit(NW, MODE_SCORE, [b"", b"", b"ACGT"], [b"", b"ACG", b""])
    assert engine.scores().tolist() == [0, -7, -8]
    engine.submit(SW, MODE_SCORE, [b"", b"", b"ACGT"], [b"", b"ACG", b""])
    s, x, y = engine.ends()
    assert s.tolist() == [0, 0, 0] and x.tolist() == [0, 0, 0] and y.tolist() == [0, 0, 0]
    engine.submit(NW, MODE_ALIGN, [b"", b"ACGT", b""], [b"ACG", b"", b""])
    al = [engine.alignment(i) for i in range(3)]
    assert (al[0].result_a, al[0].result_b, al[0].score) == (b"---", b"ACG", -7)
    assert (al[1].result_a, al[1].result_b, al[1].score) == (b"ACGT", b"----", -8)
    assert (al[2].result_a, al[2].result_b, al[2].score) == (b"", b"", 0)


def _check_alignment(al, algo, o, a, b):
    if algo == NW:
        rc, es, ea, eb = orc_nw(o, a, b)
        assert rc == 0
        assert al is not None
        assert (al.score, al.result_a, al.result_b) == (es, ea, eb), (a, b)
    else:
        n, hits = orc_sw_hits(o, a, b, 1)
        if n == 0:
            assert al is None, (a, b)
        else:
            h = hits[0]
            assert al is not None, (a, b)
            assert (al.score, al.result_a, al.result_b, al.pos_a, al.pos_b, al.len_a, al.len_b) == (
                h["score"], h["result_a"], h["result_b"], h["pos_a"], h["pos_b"], h["len_a"], h["len_b"]), (a, b)


@pytest.mark.parametrize("algo", [SW, NW], ids=["sw", "nw"])
@pytest.mark.parametrize("name", SWEEP)
def test_alignments_ragged(engine, big, name, algo):
    """score + traceback: alignment strings byte for byte"""
    n, maxlen = (120, 180) if big else (10, 40)
    sa, sb = ragged_batch(7000 + _h(name) % 1000, n, maxlen, maxlen, alphabet=_alphabet(name))
    sc = scoring_from_spec(SPECS[name])
    o = orc_from_scoring(sc)
    engine.set_scoring(sc)
    kernels = set()
    for mode in (0, 1, 2):   # specialised fill (flag bytes), general fill (2-bit codes), per-column end keys
        engine.force_general(mode)
        engine.submit(algo, MODE_ALIGN, sa, sb)
        kernels.add(engine.last_kernel)
        for i, (a, b) in enumerate(zip(sa, sb)):
            _check_alignment(engine.alignment(i), algo, o, a, b)
    engine.force_general(0)
    if name in ("sw_cli", "nw_default", "blosum62", "wild_n", "mutations"):
        assert any(k.startswith("fast_") for k in kernels) and any("general_dir" in k for k in kernels), kernels


@pytest.mark.parametrize("la", [60, 90, 120, 150, 180, 250, 300, 400])
def test_alignments_every_fill_shape(engine, big, la):
    """flag-byte fill over each lanes x columns shape: rows staged through the shared-memory
    ring (12 / 20 columns per lane) and direct 8 / 16-byte stores, ragged len_b per group"""
    n, maxb = (24, la) if big else (5, 30)
    rng = np.random.default_rng(la)
    a_all, _, b_all, _ = synthetic_batch(300 + la, n, la, maxb)
    sa = [a_all[i * la:(i + 1) * la].tobytes() for i in range(n)]
    sb = [b_all[i * maxb:i * maxb + int(rng.integers(1, maxb + 1))].tobytes() for i in range(n)]
    sa[1] = sa[1][:la - 7]          # one narrower pair: dstride below the shape's width
    for name, algo in (("sw_cli", SW), ("nw_default", NW)):
        sc = scoring_from_spec(SPECS[name])
        o = orc_from_scoring(sc)
        engine.set_scoring(sc)
        engine.submit(algo, MODE_ALIGN, sa, sb)
        assert engine.last_kernel.startswith("fast_")
        for i, (a, b) in enumerate(zip(sa, sb)):
            _check_alignment(engine.alignment(i), algo, o, a, b)


def test_alignments_wide_and_waves(engine, big, monkeypatch):
    """pairs wider than one strip; direction bytes processed in several waves"""
    monkeypatch.setenv("SEQALIGN_DIR_BUDGET", "60000")
    sa, sb = ragged_batch(77, 6 if big else 3, 640 if big else 300, 200 if big else 60, min_len=40)
    for name, algo in (("nw_default", NW), ("free_ends", NW), ("sw_cli", SW)):
        sc = scoring_from_spec(SPECS[name])
        o = orc_from_scoring(sc)
        engine.set_scoring(sc)
        engine.submit(algo, MODE_ALIGN, sa, sb)
        for i, (a, b) in enumerate(zip(sa, sb)):
            _check_alignment(engine.alignment(i), algo, o, a, b)


@pytest.mark.parametrize("name", ["nw_default", "sw_cli", "free_ends", "no_gaps_a", "no_gaps_b",
                                  "no_mismatch_wild", "blosum62", "big_scores"])
def test_matrices(engine, big, name):
    """materialise mode: all three int32 matrices, borders and sentinels included"""
    sa, sb = ragged_batch(900 + _h(name) % 100, 8 if big else 4, 300 if big else 40, 120 if big else 40,
                          alphabet=_alphabet(name))
    sa.append(b"ACGT" * (70 if big else 66))   # > 256 columns: two strips
    sb.append(b"AGGT" * 5)
    sc = scoring_from_spec(SPECS[name])
    o = orc_from_scoring(sc)
    engine.set_scoring(sc)
    for a, b in zip(sa, sb):
        for is_sw in (0, 1):
            m, ga, gb = engine.fill_matrices(a, b, is_sw)
            rc, em, ega, egb = orc_fill(o, a, b, is_sw)
            assert rc == 0
            assert np.array_equal(m, em) and np.array_equal(ga, ega) and np.array_equal(gb, egb), (a, b, is_sw)


@pytest.mark.parametrize("name", ["sw_cli", "nw_default", "linear_gap", "wild_n", "mutations", "blosum62", "pam30",
                                  "free_ends", "free_start", "free_end"])
def test_batch_matrices(engine, big, name, monkeypatch):
    """MODE_MATS: aligner_align() for a whole batch (row-per-step kernel with the gap_b prefix scan):
    all three matrices of every pair, element for element, against the oracle's fill; widths
    around the 32-column blocks, empty sequences.  SW (packed and plain scans) and NW (borders
    carrying the reference's INT_MIN-based sentinel, alignment.c:41,62-80; free start / end gaps)"""
    n, maxlen = (60, 200) if big else (10, 45)
    sa, sb = ragged_batch(5200 + _h(name) % 100, n, maxlen, maxlen, alphabet=_alphabet(name))
    widths = [31, 32, 33, 63, 64, 65] + ([127, 160, 300, 511] if big else [])
    for w in widths:
        x, y = ragged_batch(w, 1, w, w, alphabet=_alphabet(name), min_len=w)
        sa += [x[0]]; sb += [y[0][: 40 if big else 12]]
    sa += [b"", sa[0]]; sb += [sb[0], b""]
    sc = scoring_from_spec(SPECS[name])
    o = orc_from_scoring(sc)
    engine.set_scoring(sc)
    engine.force_general(0)
    # free end gaps: NW only (for SW the flag also changes the fill and stays with the general kernel)
    for nopack in () if SPECS[name].get("init", [0] * 6)[5] else ("", "1"):     # packed 16-bit prefix scans / plain int32 scans
        if nopack:
            monkeypatch.setenv("SEQALIGN_MATS_NOPACK", "1")
        engine.submit(SW, MODE_MATS, sa, sb)
        assert engine.last_kernel == ("mats_sw" if nopack else "mats_sw_packed")
        scores = engine.scores()
        for i, (a, b) in enumerate(zip(sa, sb)):
            m, ga, gb = engine.matrices(i, len(a), len(b))
            rc, em, ega, egb = orc_fill(o, a, b, True)
            assert rc == 0
            assert np.array_equal(m, em) and np.array_equal(ga, ega) and np.array_equal(gb, egb), (name, nopack, i, a, b)
            assert scores[i] == em.max()
    for nw_pack in ("1", ""):    # packed 16-bit scans (the default where every value fits) / int32 scans
        if nw_pack:
            monkeypatch.delenv("SEQALIGN_MATS_NOPACK", raising=False)
        else:
            monkeypatch.setenv("SEQALIGN_MATS_NOPACK", "1")
        engine.submit(NW, MODE_MATS, sa, sb)
        assert engine.last_kernel == ("mats_nw_packed" if nw_pack else "mats_nw")
        scores = engine.scores()
        for i, (a, b) in enumerate(zip(sa, sb)):
            m, ga, gb = engine.matrices(i, len(a), len(b))
            rc, em, ega, egb = orc_fill(o, a, b, False)
            assert rc == 0
            assert np.array_equal(m, em) and np.array_equal(ga, ega) and np.array_equal(gb, egb), (name, "nw", nw_pack, i, a, b)
            assert scores[i] == max(em[-1, -1], ega[-1, -1], egb[-1, -1])


def test_batch_matrices_in_waves(engine, big, monkeypatch):
    """MODE_MATS on a batch whose matrices do not fit the device block at once (SEQALIGN_MATS_BUDGET makes
    that a few KB here): the kernel runs wave by wave, one wave stays resident, seqalign_batch_matrices()
    re-runs the wave it is asked about -- in order, backwards and at random the matrices are the oracle's"""
    n = 40 if big else 14
    sa, sb = ragged_batch(77, n, 60, 60, min_len=1)
    sc = scoring_from_spec(SPECS["sw_cli"])
    o = orc_from_scoring(sc)
    engine.set_scoring(sc)
    engine.force_general(0)
    monkeypatch.setenv("SEQALIGN_MATS_BUDGET", str(12 * 61 * 61 * 3))     # about three pairs per wave
    for algo, is_sw in ((SW, True), (NW, False)):
        engine.submit(algo, MODE_MATS, sa, sb)
        launches = engine.last_launches
        assert launches >= 4, launches               # several waves (+ the first one again)
        scores = engine.scores()
        order = list(range(n)) + list(range(n - 1, -1, -1)) + [int(v) for v in np.random.default_rng(3).integers(0, n, size=n)]
        for i in order:
            m, ga, gb = engine.matrices(i, len(sa[i]), len(sb[i]))
            rc, em, ega, egb = orc_fill(o, sa[i], sb[i], is_sw)
            assert rc == 0
            assert np.array_equal(m, em) and np.array_equal(ga, ega) and np.array_equal(gb, egb), (algo, i)
            assert scores[i] == (em.max() if is_sw else max(em[-1, -1], ega[-1, -1], egb[-1, -1]))
    monkeypatch.setenv("SEQALIGN_MATS_BUDGET", "64")
    with pytest.raises(seqalign.SeqAlignError):                            # one pair alone does not fit
        engine.submit(SW, MODE_MATS, sa, sb)


def test_batch_matrices_rejects_other_shapes(engine):
    """scoring shapes outside the specialised kernel are refused, not approximated"""
    sc = scoring_from_spec(SPECS["no_gaps_a"])
    engine.set_scoring(sc)
    with pytest.raises(seqalign.SeqAlignError) as e:
        engine.submit(SW, MODE_MATS, [b"ACGT"], [b"ACGT"])
    assert e.value.code == seqalign.ERR_ARG
    with pytest.raises(seqalign.SeqAlignError):
        engine.submit(NW, MODE_MATS, [b"ACGT"], [b"ACGT"])
    for name in ("no_gaps_a", "no_gaps_b", "no_mismatch"):   # NW shapes the row kernel does not cover
        engine.set_scoring(scoring_from_spec(SPECS[name]))
        with pytest.raises(seqalign.SeqAlignError) as e:
            engine.submit(NW, MODE_MATS, [b"ACGT"], [b"ACGT"])
        assert e.value.code == seqalign.ERR_ARG


@pytest.mark.parametrize("chunk", range(4))
def test_reference_golden_vectors(engine, chunk):
    """tests/golden/reference_vectors.json (made by the unmodified reference)
    straight against the CUDA path: batch API and classic single-pair API"""
    cases = GOLD["cases"][chunk::4]
    by_spec = {}
    for c in cases:
        by_spec.setdefault(c["spec"], []).append(c)
    for spec, cs in by_spec.items():
        sc = scoring_from_spec(GOLD["specs"][spec])
        engine.set_scoring(sc)
        engine.force_general(False)
        sa = [c["a"].encode() for c in cs]
        sb = [c["b"].encode() for c in cs]
        engine.submit(NW, MODE_ALIGN, sa, sb)
        for i, c in enumerate(cs):
            al = engine.alignment(i)
            assert (al.score, al.result_a.decode(), al.result_b.decode()) == (
                c["nw"]["score"], c["nw"]["result_a"], c["nw"]["result_b"]), c
        engine.submit(SW, MODE_ALIGN, sa, sb)
        for i, c in enumerate(cs):
            al = engine.alignment(i)
            if not c["sw"]:
                assert al is None
                continue
            e = c["sw"][0]
            assert (al.score, al.result_a.decode(), al.result_b.decode(), al.pos_a, al.pos_b, al.len_a, al.len_b) == (
                e["score"], e["result_a"], e["result_b"], e["pos_a"], e["pos_b"], e["len_a"], e["len_b"]), c
        for c in cs:
            for key, is_sw in (("nw_mats", 0), ("sw_mats", 1)):
                if key in c:
                    m, ga, gb = engine.fill_matrices(c["a"], c["b"], is_sw)
                    assert m.ravel().tolist() == c[key][0]
                    assert ga.ravel().tolist() == c[key][1]
                    assert gb.ravel().tolist() == c[key][2]


def test_reference_golden_matrices_batch_mode(engine):
    """the matrices recorded from the unmodified reference (tests/golden/reference_vectors.json) against
    MODE_MATS for whole batches, NW and SW; scoring shapes the row kernel refuses (SEQALIGN_ERR_ARG) are
    counted, the rest must match element for element"""
    by_spec = {}
    for c in GOLD["cases"]:
        if "nw_mats" in c or "sw_mats" in c:
            by_spec.setdefault(c["spec"], []).append(c)
    checked = {NW: 0, SW: 0}
    for spec, cs in by_spec.items():
        engine.set_scoring(scoring_from_spec(GOLD["specs"][spec]))
        engine.force_general(False)
        for algo, key in ((NW, "nw_mats"), (SW, "sw_mats")):
            sub = [c for c in cs if key in c]
            if not sub:
                continue
            try:
                engine.submit(algo, MODE_MATS, [c["a"].encode() for c in sub], [c["b"].encode() for c in sub])
            except seqalign.SeqAlignError as e:
                assert e.code in (seqalign.ERR_ARG, seqalign.ERR_UNKNOWN_PAIR), (spec, e)
                continue
            for i, c in enumerate(sub):
                m, ga, gb = engine.matrices(i, len(c["a"]), len(c["b"]))
                assert m.ravel().tolist() == c[key][0], (spec, key, c["a"], c["b"])
                assert ga.ravel().tolist() == c[key][1], (spec, key, c["a"], c["b"])
                assert gb.ravel().tolist() == c[key][2], (spec, key, c["a"], c["b"])
                checked[algo] += 1
    print("pairs checked against the reference\x27s matrices:", checked)
    assert checked[NW] >= 70 and checked[SW] >= 60, checked


def test_classic_api_single_pair(engine):
    """needleman_wunsch_align / smith_waterman_align + fetch (all hits) via the
    reference's own function names; hit order and visited-mask semantics"""
    # BASELINE config 1 / README.md:71-74
    al = seqalign.needleman_wunsch("CAGACGT", "CGATA", seqalign.Scoring.nw_default())
    assert (al.result_a, al.result_b, al.score) == (b"C-AGACGT", b"CGATA---", -11)
    for c in GOLD["cases"][::5]:
        sc = scoring_from_spec(GOLD["specs"][c["spec"]])
        al = seqalign.needleman_wunsch(c["a"], c["b"], sc)
        assert (al.score, al.result_a.decode(), al.result_b.decode()) == (
            c["nw"]["score"], c["nw"]["result_a"], c["nw"]["result_b"])
        hits = seqalign.smith_waterman(c["a"], c["b"], sc, max_hits=6)
        assert len(hits) == len(c["sw"]), c
        for h, e in zip(hits, c["sw"]):
            assert (h.score, h.result_a.decode(), h.result_b.decode(), h.pos_a, h.pos_b, h.len_a, h.len_b) == (
                e["score"], e["result_a"], e["result_b"], e["pos_a"], e["pos_b"], e["len_a"], e["len_b"]), c


def test_classic_api_reused_aligner(engine):
    """one sw_aligner_t across pairs must behave like a fresh one per pair
    (the reference's reused mask is stale: SURVEY.md 8c H1)"""
    L = seqalign.load()
    import ctypes
    sc = seqalign.Scoring.sw_cli_default()
    o = orc_from_scoring(sc)
    sw = L.smith_waterman_new()
    res = L.alignment_create(256)
    sa, sb = ragged_batch(5, 6, 60, 60)
    for a, b in list(zip(sa, sb)) * 2:
        L.smith_waterman_align2(a, b, len(a), len(b), sc.ptr, sw)
        got = []
        while len(got) < 4 and L.smith_waterman_fetch(sw, res):
            r = res.contents
            got.append((r.score, ctypes.string_at(r.result_a, r.length), ctypes.string_at(r.result_b, r.length),
                        r.pos_a, r.pos_b))
        n, hits = orc_sw_hits(o, a, b, 4)
        assert got == [(h["score"], h["result_a"], h["result_b"], h["pos_a"], h["pos_b"]) for h in hits]
    L.alignment_free(res)
    L.smith_waterman_free(sw)


def test_unknown_character_pair(engine):
    """DNA_hybridization has no match/mismatch fallback: the reference exits on
    the first unknown pair (alignment_scoring.c:179-181); the batch API reports it"""
    sc = seqalign.Scoring.system("DNA_hybridization")
    engine.set_scoring(sc)
    engine.submit(SW, MODE_SCORE, [b"ACGT", b"ACCA"], [b"ACGT", b"TTGA"])   # fine
    with pytest.raises(seqalign.SeqAlignError) as e:
        engine.submit(SW, MODE_SCORE, [b"ACGT", b"ACNA"], [b"ACGT", b"TTGA"])
    assert e.value.code == seqalign.ERR_UNKNOWN_PAIR
    assert "Unknown character pair (n,t)" in str(e.value)
    # 'N' present in the batch but never paired with an unknown partner is fine:
    # a-side N only meets an empty b
    engine.submit(SW, MODE_SCORE, [b"ACGT", b"NN"], [b"ACGT", b""])
    assert engine.scores().tolist()[1] == 0


def test_full_size_invariants(engine, big):
    """BASELINE full size (100k pairs, 150x150): properties that need no oracle,
    plus specialised-vs-general kernel agreement and an oracle-checked sample"""
    if not big:
        pytest.skip("full size runs on the GPU only")
    n = 100000
    a, oa, b, ob = synthetic_batch(2, n, 150, 150)
    sc = scoring_from_spec(SPECS["sw_cli"])
    engine.set_scoring(sc)
    engine.force_general(False)
    engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
    s, x, y = engine.ends()
    assert engine.last_kernel == "fast16_sw_score_end"
    engine.force_general(5)
    engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
    s5, x5, y5 = engine.ends()
    assert engine.last_kernel == "fast_sw_score_end"
    assert np.array_equal(s, s5) and np.array_equal(x, x5) and np.array_equal(y, y5)
    engine.force_general(True)
    engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
    s2, x2, y2 = engine.ends()
    assert np.array_equal(s, s2) and np.array_equal(x, x2) and np.array_equal(y, y2)
    engine.force_general(False)
    # symmetry of the score under swapping the sequences (symmetric scoring)
    engine.submit_packed(SW, MODE_SCORE, b, ob, a, oa)
    assert np.array_equal(engine.scores(), s)
    # self alignment: 150 matches
    engine.submit_packed(SW, MODE_SCORE, a, oa, a, oa)
    s3, x3, y3 = engine.ends()
    assert (s3 == 300).all() and (x3 == 150).all() and (y3 == 150).all()
    assert s.min() >= 0 and s.max() <= 300 and (x >= 0).all() and (x <= 150).all()
    # oracle on a strided sample
    idx = np.arange(0, n, 97)
    o = orc_from_scoring(sc)
    A, B = a.reshape(n, 150)[idx].reshape(-1), b.reshape(n, 150)[idx].reshape(-1)
    off = np.arange(len(idx) + 1, dtype=np.int64) * 150
    es, ex, ey = orc_batch_sw(o, A, off, B, off)
    assert np.array_equal(s[idx], es) and np.array_equal(x[idx], ex) and np.array_equal(y[idx], ey)


def _rescore(ra, rb, sub, gap_open, gap_extend, free_ends=False):
    """score of a gapped alignment under affine gaps (a gap of length N costs open + N*extend),
    recomputed column by column; free_ends: leading / trailing gap runs cost nothing"""
    n = len(ra)
    cols = list(zip(ra, rb))
    lead, trail = 0, n
    if free_ends and n:
        # one leading run of gaps in ONE row lies on the matrix border (free), likewise one trailing run
        row = 0 if cols[0][0] == 45 else 1 if cols[0][1] == 45 else None
        while row is not None and lead < n and cols[lead][row] == 45:
            lead += 1
        row = 0 if cols[-1][0] == 45 else 1 if cols[-1][1] == 45 else None
        while row is not None and trail > lead and cols[trail - 1][row] == 45:
            trail -= 1
    total, prev = 0, 0   # prev: 0 = match column, 1 = gap in a, 2 = gap in b
    for x, y in cols[lead:trail]:
        if x == 45:
            total += gap_extend + (gap_open if prev != 1 else 0); prev = 1
        elif y == 45:
            total += gap_extend + (gap_open if prev != 2 else 0); prev = 2
        else:
            total += sub(x, y); prev = 0
    return total


def _cached_lookup(sc):
    memo = {}

    def look(x, y):
        if (x, y) not in memo:
            memo[(x, y)] = sc.lookup(bytes([x]), bytes([y]))[0]
        return memo[(x, y)]
    return look


def test_full_size_protein_alignments_rescore(engine, big):
    """BASELINE config 4 at full width (SW, protein 400x400, BLOSUM62, first hit): every
    reported hit must be a substring pair of its inputs whose column-by-column score equals the
    reported score, and the score must equal the score-only kernel's; oracle on a sample"""
    if not big:
        pytest.skip("full size runs on the GPU only")
    n, L = 5000, 400
    a, oa, b, ob = synthetic_batch(4, n, L, L, kind="protein")
    sc = scoring_from_spec(SPECS["blosum62"])
    engine.set_scoring(sc)
    engine.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
    s_score, xe, ye = [v.copy() for v in engine.ends()]
    engine.submit_packed(SW, MODE_ALIGN, a, oa, b, ob)
    assert engine.last_kernel.startswith("fast_sw_dir")
    look = _cached_lookup(sc)
    A, B = a.reshape(n, L), b.reshape(n, L)
    for i in range(0, n, 7):
        al = engine.alignment(i)
        if s_score[i] == 0:
            assert al is None
            continue
        assert al.score == s_score[i]
        assert al.pos_a + al.len_a == xe[i] and al.pos_b + al.len_b == ye[i]
        assert al.result_a.replace(b"-", b"") == A[i, al.pos_a: al.pos_a + al.len_a].tobytes()
        assert al.result_b.replace(b"-", b"") == B[i, al.pos_b: al.pos_b + al.len_b].tobytes()
        assert _rescore(al.result_a, al.result_b, look, sc.s.gap_open, sc.s.gap_extend) == al.score
    o = orc_from_scoring(sc)
    for i in range(0, n, 500):
        _check_alignment(engine.alignment(i), SW, o, A[i].tobytes(), B[i].tobytes())


def test_full_size_long_pairs_rescore(engine, big):
    """BASELINE config 3 at full width (NW 10k x 10k, free start and end gaps, score +
    traceback): the gapped strings spell the inputs, and their column score with free end
    gaps equals both the reported score and the score-only kernel's"""
    if not big:
        pytest.skip("full size runs on the GPU only")
    n, L = 8, 10000
    a, oa, b, ob = synthetic_batch(3, n, L, L, block=16)
    sc = scoring_from_spec(SPECS["free_ends"])
    engine.set_scoring(sc)
    engine.submit_packed(NW, MODE_SCORE, a, oa, b, ob)
    s_score = engine.scores().copy()
    engine.submit_packed(NW, MODE_ALIGN, a, oa, b, ob)
    assert engine.last_kernel.startswith("long_nw_ckpt")   # wide pairs trace back through checkpoints
    look = _cached_lookup(sc)
    A, B = a.reshape(n, L), b.reshape(n, L)
    for i in range(n):
        al = engine.alignment(i)
        assert al.score == s_score[i]
        assert al.result_a.replace(b"-", b"") == A[i].tobytes() and al.result_b.replace(b"-", b"") == B[i].tobytes()
        assert len(al.result_a) == len(al.result_b)
        assert _rescore(al.result_a, al.result_b, look, sc.s.gap_open, sc.s.gap_extend, free_ends=True) == al.score


def _on_gpu():
    import conftest
    return conftest.BACKEND == "gpu"


def _device_arrays(big, arrays):
    """CUDA copies of numpy arrays (GPU) / the arrays themselves (the emulator's
    device memory is host memory); returns (keepalive, pointers).  `big` only sizes the
    callers' batches; what counts here is which backend runs."""
    if _on_gpu():
        import torch
        ts = [torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0") for a in arrays]
        torch.cuda.synchronize()
        return ts, [t.data_ptr() for t in ts]
    # pad: the kernels stage sequences with 16-byte bulk copies
    ts = [np.concatenate([a, np.zeros(32, dtype=a.dtype)]) for a in arrays]
    return ts, [t.ctypes.data for t in ts]


def _device_result(big, t, n):
    return t.cpu().numpy()[:n] if _on_gpu() else t[:n]


def test_full_size_long_pairs_oracle(engine, big):
    """BASELINE config 3 shape against the oracle itself: 16 pairs of 10k x 10k (SURVEY.md 8d asks for
    a fixed 16-pair sample), NW with free start and end gaps: score AND both gapped strings, i.e. the
    reference's GA > GB > M tie-breaking along 20,000 traceback steps per pair.  The oracle fills three
    400 MB matrices per pair (about a second each)."""
    if not big:
        pytest.skip("full size runs on the GPU only")
    import torch
    n, L = 16, 10000
    da = torch.empty(n * L, dtype=torch.uint8, device="cuda:0")
    db = torch.empty(n * L, dtype=torch.uint8, device="cuda:0")
    seqalign.synth_device(0, "dna", 3, 4000, n, L, L, da.data_ptr(), db.data_ptr())   # pairs 4000.. of the config-3 stream
    a, b = da.cpu().numpy(), db.cpu().numpy()
    oa = np.arange(n + 1, dtype=np.int64) * L
    sc = scoring_from_spec(SPECS["free_ends"])
    engine.set_scoring(sc)
    engine.submit_packed(NW, MODE_ALIGN, a, oa, b, oa)
    assert engine.last_kernel.startswith("long_nw")
    o = orc_from_scoring(sc)
    for i in range(n):
        rc, es, ea, eb = orc_nw(o, a[i * L:(i + 1) * L].tobytes(), b[i * L:(i + 1) * L].tobytes())
        al = engine.alignment(i)
        assert rc == 0 and al.score == es, i
        assert al.result_a == ea and al.result_b == eb, i


def test_device_resident_api(big):
    """seqalign_batch_run_device: inputs and outputs stay in device memory.  A stream of
    batches exercises the speculative launch (previous plan reused, verified against the
    scan afterwards): hits, and misses on a new alphabet, a new shape, ragged lengths."""
    sc = scoring_from_spec(SPECS["sw_cli"])
    o = orc_from_scoring(sc)
    eng = seqalign.BatchAligner(0, sc)
    n, L = (3000, 150) if big else (6, 30)
    ragged = seqalign.pack(ragged_batch(3, n, L + 20, L + 20)[0]) + seqalign.pack(ragged_batch(3, n, L + 20, L + 20)[1])
    withn = ragged_batch(8, n, L, L, alphabet=b"ACGTN")
    batches = [
        synthetic_batch(9, n, L, L), synthetic_batch(10, n, L, L), synthetic_batch(11, n, L, L),   # miss, hit, hit
        seqalign.pack(withn[0]) + seqalign.pack(withn[1]),                                          # new alphabet + ragged
        synthetic_batch(12, n, L, L),                                                              # back to uniform
        synthetic_batch(13, n, L + 40, L + 33),                                                     # larger shape
        synthetic_batch(14, n, L, L), synthetic_batch(15, n, L - 7, L - 3),                         # smaller shape: hit
        ragged,
    ]
    for want_ends in (False, True):
        for k, (a, oa, b, ob) in enumerate(batches):
            m = len(oa) - 1
            es, ex, ey = orc_batch_sw(o, a, oa, b, ob)
            out = [np.zeros(m, dtype=np.int32) for _ in range(3)]
            keep, (pa, poa, pb, pob, ps, px, py) = _device_arrays(big, [a, oa, b, ob] + out)
            eng.run_device(SW, pa, poa, pb, pob, m, ps, px if want_ends else 0, py if want_ends else 0)
            assert np.array_equal(_device_result(big, keep[4], m), es), (want_ends, k, eng.last_kernel)
            if want_ends:
                assert np.array_equal(_device_result(big, keep[5], m), ex)
                assert np.array_equal(_device_result(big, keep[6], m), ey)
    hits, misses = eng.speculation_stats()
    assert hits >= 6 and misses >= 4, (hits, misses)
    eng.close()


def test_device_resident_async(big):
    """seqalign_batch_run_device_async / _wait: several runs enqueued before any is verified.
    The stream mixes batches that fit the guessed plan with ones that do not (new alphabet, larger
    shape, ragged): a wrong guess redoes that run and everything enqueued after it; every result
    must equal the oracle's whatever the order of events"""
    sc = scoring_from_spec(SPECS["sw_cli"])
    o = orc_from_scoring(sc)
    eng = seqalign.BatchAligner(0, sc)
    n, L = (2000, 150) if big else (6, 30)
    withn = ragged_batch(8, n, L, L, alphabet=b"ACGTN")
    batches = [synthetic_batch(20 + k, n, L, L) for k in range(5)]
    batches += [seqalign.pack(withn[0]) + seqalign.pack(withn[1])]            # wrong guess in the middle of a queue
    batches += [synthetic_batch(30 + k, n, L, L) for k in range(3)]
    batches += [synthetic_batch(40, n, L + 30, L + 11)]                       # larger shape
    batches += [synthetic_batch(41 + k, n, L - 5, L - 9) for k in range(4)]   # smaller: fits the larger plan
    for want_ends in (False, True):
        for depth in (1, 2, 4, 6):
            pending, checked = [], 0
            for k, (a, oa, b, ob) in enumerate(batches):
                m = len(oa) - 1
                out = [np.zeros(m, dtype=np.int32) for _ in range(3)]
                keep, (pa, poa, pb, pob, ps, px, py) = _device_arrays(big, [a, oa, b, ob] + out)
                eng.run_device_async(SW, pa, poa, pb, pob, m, ps, px if want_ends else 0, py if want_ends else 0)
                pending.append((k, keep, m))
                while len(pending) > depth or (k == len(batches) - 1 and pending):
                    eng.run_device_wait()
                    kk, kp, mm = pending.pop(0)
                    es, ex, ey = orc_batch_sw(o, *batches[kk])
                    assert np.array_equal(_device_result(big, kp[4], mm), es), (want_ends, depth, kk, eng.last_kernel)
                    if want_ends:
                        assert np.array_equal(_device_result(big, kp[5], mm), ex) and np.array_equal(_device_result(big, kp[6], mm), ey)
                    checked += 1
            assert checked == len(batches)
            eng.run_device_wait()      # nothing outstanding: a no-op
    hits, misses = eng.speculation_stats()
    assert hits >= 40 and misses >= 8, (hits, misses)
    # a blocking entry point with runs outstanding waits for them first
    a, oa, b, ob = batches[0]
    m = len(oa) - 1
    out = [np.zeros(m, dtype=np.int32) for _ in range(3)]
    keep, (pa, poa, pb, pob, ps, px, py) = _device_arrays(big, [a, oa, b, ob] + out)
    eng.run_device(SW, pa, poa, pb, pob, m, ps, 0, 0)
    eng.run_device_async(SW, pa, poa, pb, pob, m, ps, 0, 0)
    eng.submit_packed(SW, MODE_SCORE, a, oa, b, ob)
    es, _, _ = orc_batch_sw(o, a, oa, b, ob)
    assert np.array_equal(eng.scores(), es) and np.array_equal(_device_result(big, keep[4], m), es)
    eng.close()


def test_pipelined_aligner(big):
    """several batches in flight (one engine per worker thread), results in order"""
    sc = scoring_from_spec(SPECS["sw_cli"])
    o = orc_from_scoring(sc)
    pipe = seqalign.PipelinedAligner(0, sc, depth=3)
    batches = [synthetic_batch(40 + k, 3000 if big else 5, 150 if big else 30, 150 if big else 30) for k in range(5)]
    got = list(pipe.map_scores(SW, batches))
    full = [pipe.submit_packed(SW, MODE_SCORE, *b) for b in batches]
    for b, s, f in zip(batches, got, full):
        es, ex, ey = orc_batch_sw(o, *b)
        assert np.array_equal(s, es)
        fs, fx, fy = f.result()
        assert np.array_equal(fs, es) and np.array_equal(fx, ex) and np.array_equal(fy, ey)
    pipe.close()


def _hit_tuple(h):
    return (h.score, h.result_a, h.result_b, h.pos_a, h.pos_b, h.len_a, h.len_b)


@pytest.mark.parametrize("name", ["sw_cli", "nw_default", "linear_gap", "wild_n", "mutations", "blosum62", "pam30"])
def test_multi_hit_on_device(engine, big, name):
    """MODE_HITS: candidate sort + masked walks on the device vs the oracle's restatement of
    smith_waterman_align2 + fetch loop (fresh mask), for several hit limits"""
    n, maxlen = (150, 160) if big else (8, 36)
    sa, sb = ragged_batch(4100 + _h(name) % 100, n, maxlen, maxlen, alphabet=_alphabet(name), min_len=1)
    # self-similar pairs: many equal-score candidates and colliding walks (the lcs use case)
    sa += [b"abcabcdabcdexabcd".upper() if name not in PROTEIN_SPECS else b"ARNDARNDCQARNDCQE", sa[0]]
    sb += [sa[-2], sa[0]]
    sc = scoring_from_spec(SPECS[name])
    o = orc_from_scoring(sc)
    engine.set_scoring(sc)
    engine.force_general(0)
    for max_hits, min_score in ((1, 1), (5, 1), (40, 1), (6, 4)):
        engine.set_hit_limits(max_hits, min_score)
        engine.submit(SW, MODE_HITS, sa, sb)
        assert "hits" in engine.last_kernel
        for i, (a, b) in enumerate(zip(sa, sb)):
            nref, ref = orc_sw_hits(o, a, b, max_hits)
            ref = [h for h in ref if h["score"] >= min_score]
            got = engine.hits(i)
            assert [_hit_tuple(h) for h in got] == [
                (h["score"], h["result_a"], h["result_b"], h["pos_a"], h["pos_b"], h["len_a"], h["len_b"]) for h in ref], (
                name, max_hits, min_score, a, b)
    engine.set_hit_limits(8, 1)


def test_multi_hit_in_waves(engine, big, monkeypatch):
    """MODE_HITS with the flag-byte budget cut down: the batch runs in several waves and the hit strings are
    gathered wave by wave (a one-wave batch is read straight from the landing buffers instead)"""
    n, maxlen = (60, 120) if big else (7, 30)
    sa, sb = ragged_batch(4477, n, maxlen, maxlen, min_len=1)
    sc = scoring_from_spec(SPECS["sw_cli"])
    o = orc_from_scoring(sc)
    engine.set_scoring(sc)
    engine.force_general(0)
    engine.set_hit_limits(5, 1)
    lists = []
    for budget in (None, str(2 * (maxlen + 16) * (maxlen + 16) * 4)):
        if budget:
            monkeypatch.setenv("SEQALIGN_DIR_BUDGET", budget)
        engine.submit(SW, MODE_HITS, sa, sb)
        lists.append([[_hit_tuple(h) for h in engine.hits(i)] for i in range(len(sa))])
    assert lists[0] == lists[1]
    for i, (a, b) in enumerate(zip(sa, sb)):
        nref, ref = orc_sw_hits(o, a, b, 5)
        assert lists[1][i] == [(h["score"], h["result_a"], h["result_b"], h["pos_a"], h["pos_b"], h["len_a"], h["len_b"])
                               for h in ref if h["score"] >= 1], (a, b)
    engine.set_hit_limits(8, 1)


def test_multi_hit_golden(engine):
    """the reference's own hit lists (tests/golden/reference_vectors.json, up to six hits per pair)"""
    engine.set_hit_limits(6, 1)
    by_spec = {}
    for c in GOLD["cases"]:
        by_spec.setdefault(c["spec"], []).append(c)
    checked = 0
    for spec, cs in by_spec.items():
        sc = scoring_from_spec(GOLD["specs"][spec])
        engine.set_scoring(sc)
        cs = [c for c in cs if len(c["a"]) <= 512 and len(c["a"]) > 0 and len(c["b"]) > 0]
        try:
            engine.submit(SW, MODE_HITS, [c["a"].encode() for c in cs], [c["b"].encode() for c in cs])
        except seqalign.SeqAlignError as e:
            assert e.code == seqalign.ERR_ARG   # scoring shapes only the general kernel handles
            continue
        for i, c in enumerate(cs):
            got = [(h.score, h.result_a.decode(), h.result_b.decode(), h.pos_a, h.pos_b, h.len_a, h.len_b)
                   for h in engine.hits(i)]
            exp = [(e["score"], e["result_a"], e["result_b"], e["pos_a"], e["pos_b"], e["len_a"], e["len_b"])
                   for e in c["sw"]]
            assert got == exp, c
            checked += 1
    assert checked > 100
    engine.set_hit_limits(8, 1)
