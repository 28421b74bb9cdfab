"""Shared test plumbing: oracle bindings, scoring specs, seeded generators.

The oracle (oracle/liboracle.so, plus oracle/_ref/libalign_ref.so when it has
been built from /root/reference) is the checker.  The product is reached only
through the C-ABI (seqalign package -> libseqalign_b200.so).
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "seq-align_b200"))

import seqalign  # noqa: E402

ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libalign_ref.so")
REF_BATCH = os.path.join(ORACLE_DIR, "_ref", "ref_batch")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libseqalign_emu.so")


class OrcScoring(ctypes.Structure):
    """orc_scoring_t of oracle/sa_oracle.h"""

    _fields_ = [
        ("gap_open", ctypes.c_int), ("gap_extend", ctypes.c_int),
        ("no_start_gap_penalty", ctypes.c_int), ("no_end_gap_penalty", ctypes.c_int),
        ("no_gaps_in_a", ctypes.c_int), ("no_gaps_in_b", ctypes.c_int), ("no_mismatches", ctypes.c_int),
        ("use_match_mismatch", ctypes.c_int), ("match", ctypes.c_int), ("mismatch", ctypes.c_int),
        ("case_sensitive", ctypes.c_int),
        ("min_penalty", ctypes.c_int), ("max_penalty", ctypes.c_int),
        ("is_wild", ctypes.c_ubyte * 256),
        ("wild_score", ctypes.c_int * 256),
        ("has_swap", (ctypes.c_ubyte * 256) * 256),
        ("swap_score", (ctypes.c_int * 256) * 256),
    ]


class OrcAlignment(ctypes.Structure):
    _fields_ = [
        ("result_a", ctypes.c_char_p), ("result_b", ctypes.c_char_p),
        ("length", ctypes.c_size_t),
        ("pos_a", ctypes.c_size_t), ("pos_b", ctypes.c_size_t),
        ("len_a", ctypes.c_size_t), ("len_b", ctypes.c_size_t),
        ("score", ctypes.c_int),
    ]


_orc = None


def oracle():
    global _orc
    if _orc is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"])
        _orc = ctypes.CDLL(path)
        _orc.orc_scoring_sizeof.restype = ctypes.c_size_t
        assert _orc.orc_scoring_sizeof() == ctypes.sizeof(OrcScoring)
        _orc.orc_sw_hits.restype = ctypes.c_long
    return _orc


def orc_from_scoring(sc):
    """Copy a product scoring_t (seqalign.Scoring) field by field into the
    oracle's own struct, so both sides score with the same model."""
    s = sc.s
    o = OrcScoring()
    for f in ("gap_open", "gap_extend", "no_start_gap_penalty", "no_end_gap_penalty", "no_gaps_in_a",
              "no_gaps_in_b", "no_mismatches", "use_match_mismatch", "match", "mismatch",
              "case_sensitive", "min_penalty", "max_penalty"):
        setattr(o, f, int(getattr(s, f)))
    wild = np.frombuffer(s.wildcards, dtype=np.uint32)
    swap = np.frombuffer(s.swap_set, dtype=np.uint32).reshape(256, 8)
    wsc = np.frombuffer(s.wildscores, dtype=np.int32)
    ssc = np.frombuffer(s.swap_scores, dtype=np.int32).reshape(256, 256)
    bits = np.arange(256)
    is_wild = ((wild[bits >> 5] >> (bits & 31)) & 1).astype(np.uint8)
    has_swap = ((swap[:, bits >> 5] >> (bits & 31)) & 1).astype(np.uint8)
    np.frombuffer(o.is_wild, dtype=np.uint8)[:] = is_wild
    np.frombuffer(o.wild_score, dtype=np.int32)[:] = np.where(is_wild, wsc, 0)
    np.frombuffer(o.has_swap, dtype=np.uint8).reshape(256, 256)[:] = has_swap
    np.frombuffer(o.swap_score, dtype=np.int32).reshape(256, 256)[:] = np.where(has_swap, ssc, 0)
    return o


def orc_nw(o, a, b):
    """oracle Needleman-Wunsch -> (rc, score, result_a, result_b)"""
    cap = len(a) + len(b) + 1
    ra, rb = ctypes.create_string_buffer(cap), ctypes.create_string_buffer(cap)
    out = OrcAlignment()
    out.result_a = ctypes.cast(ra, ctypes.c_char_p)
    out.result_b = ctypes.cast(rb, ctypes.c_char_p)
    rc = oracle().orc_nw_align(ctypes.byref(o), a, ctypes.c_size_t(len(a)), b, ctypes.c_size_t(len(b)),
                               ctypes.byref(out))
    return rc, out.score, ra.raw[:out.length], rb.raw[:out.length]


def orc_sw_hits(o, a, b, max_hits):
    """oracle Smith-Waterman hit list on a fresh mask"""
    stride = len(a) + len(b) + 1
    pa = ctypes.create_string_buffer(stride * max_hits)
    pb = ctypes.create_string_buffer(stride * max_hits)
    hits = (OrcAlignment * max_hits)()
    n = oracle().orc_sw_hits(ctypes.byref(o), a, ctypes.c_size_t(len(a)), b, ctypes.c_size_t(len(b)),
                             ctypes.c_size_t(max_hits), hits, pa, pb, ctypes.c_size_t(stride))
    out = []
    for i in range(max(n, 0)):
        h = hits[i]
        out.append(dict(result_a=pa.raw[i * stride:i * stride + h.length],
                        result_b=pb.raw[i * stride:i * stride + h.length],
                        score=h.score, pos_a=h.pos_a, pos_b=h.pos_b, len_a=h.len_a, len_b=h.len_b))
    return n, out


def orc_fill(o, a, b, is_sw):
    cells = (len(a) + 1) * (len(b) + 1)
    m, ga, gb = (np.zeros(cells, dtype=np.int32) for _ in range(3))
    rc = oracle().orc_fill(ctypes.byref(o), a, ctypes.c_size_t(len(a)), b, ctypes.c_size_t(len(b)),
                           int(is_sw), m.ctypes.data_as(ctypes.c_void_p), ga.ctypes.data_as(ctypes.c_void_p),
                           gb.ctypes.data_as(ctypes.c_void_p))
    shape = (len(b) + 1, len(a) + 1)
    return rc, m.reshape(shape), ga.reshape(shape), gb.reshape(shape)


def orc_batch_sw(o, seq_a, off_a, seq_b, off_b):
    n = len(off_a) - 1
    s, x, y = (np.zeros(n, dtype=np.int32) for _ in range(3))
    rc = oracle().orc_batch_sw_best(ctypes.byref(o), ctypes.c_size_t(n),
                                    seq_a.ctypes.data_as(ctypes.c_void_p), off_a.ctypes.data_as(ctypes.c_void_p),
                                    seq_b.ctypes.data_as(ctypes.c_void_p), off_b.ctypes.data_as(ctypes.c_void_p),
                                    s.ctypes.data_as(ctypes.c_void_p), x.ctypes.data_as(ctypes.c_void_p),
                                    y.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return s, x, y


def orc_batch_nw(o, seq_a, off_a, seq_b, off_b):
    n = len(off_a) - 1
    s = np.zeros(n, dtype=np.int32)
    rc = oracle().orc_batch_nw_score(ctypes.byref(o), ctypes.c_size_t(n),
                                     seq_a.ctypes.data_as(ctypes.c_void_p), off_a.ctypes.data_as(ctypes.c_void_p),
                                     seq_b.ctypes.data_as(ctypes.c_void_p), off_b.ctypes.data_as(ctypes.c_void_p),
                                     s.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return s


# ---------------------------------------------------------------------------
# seeded synthetic batches: b is a mutated copy of a

DNA = np.frombuffer(b"ACGT", dtype=np.uint8)
PROTEIN = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)


def related_block(rng, npairs, len_a, len_b, alphabet, p_sub, p_indel):
    """npairs x (len_a, len_b) uint8 matrices; b = a with substitutions,
    insertions and deletions, cut / padded with fresh letters to len_b."""
    k = len(alphabet)
    ext = len_b + len_a + 64
    src = rng.integers(0, k, size=(npairs, ext), dtype=np.int64)
    a = src[:, :len_a]
    b = np.empty((npairs, len_b), dtype=np.int64)
    ia = np.zeros(npairs, dtype=np.int64)
    rows = np.arange(npairs)
    for j in range(len_b):
        r = rng.random(npairs)
        ins = r < p_indel
        dele = (r >= p_indel) & (r < 2 * p_indel)
        ia += dele
        ia = np.minimum(ia, ext - 1)
        base = src[rows, ia]
        sub = rng.random(npairs) < p_sub
        base = np.where(sub, (base + 1 + rng.integers(0, k - 1, size=npairs)) % k, base)
        b[:, j] = np.where(ins, rng.integers(0, k, size=npairs), base)
        ia += ~ins
        ia = np.minimum(ia, ext - 1)
    return alphabet[a], alphabet[b]


def synthetic_batch(seed, npairs, len_a, len_b, kind="dna", related=True, block=4096):
    """Packed batch (seq_a, off_a, seq_b, off_b) of fixed-length pairs.
    Block b of pairs depends only on (seed, b): shards can be generated
    independently on any rank."""
    alphabet, p_sub, p_indel = (DNA, 0.05, 0.01) if kind == "dna" else (PROTEIN, 0.15, 0.02)
    A = np.empty((npairs, len_a), dtype=np.uint8)
    B = np.empty((npairs, len_b), dtype=np.uint8)
    for b0 in range(0, npairs, block):
        m = min(block, npairs - b0)
        rng = np.random.Generator(np.random.Philox(key=[seed, b0 // block]))
        if related:
            a, b = related_block(rng, m, len_a, len_b, alphabet, p_sub, p_indel)
        else:
            a = alphabet[rng.integers(0, len(alphabet), size=(m, len_a))]
            b = alphabet[rng.integers(0, len(alphabet), size=(m, len_b))]
        A[b0:b0 + m] = a
        B[b0:b0 + m] = b
    off_a = np.arange(npairs + 1, dtype=np.int64) * len_a
    off_b = np.arange(npairs + 1, dtype=np.int64) * len_b
    return A.reshape(-1), off_a, B.reshape(-1), off_b


def ragged_batch(seed, npairs, max_a, max_b, alphabet=b"ACGT", min_len=0, related=True):
    """Pairs with random lengths in [min_len, max]; half of them related."""
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(alphabet, dtype=np.uint8)
    sa, sb = [], []
    for _ in range(npairs):
        la = int(rng.integers(min_len, max_a + 1))
        lb = int(rng.integers(min_len, max_b + 1))
        a = letters[rng.integers(0, len(letters), size=la)]
        if related and rng.random() < 0.5 and la > 0:
            src = np.resize(a, lb) if lb else a[:0]
            mut = rng.random(lb) < 0.1
            b = np.where(mut, letters[rng.integers(0, len(letters), size=lb)], src)
        else:
            b = letters[rng.integers(0, len(letters), size=lb)]
        sa.append(a.astype(np.uint8).tobytes())
        sb.append(b.astype(np.uint8).tobytes())
    return sa, sb


# ---------------------------------------------------------------------------
# scoring specs: plain data, so the same spec can drive the product
# (seqalign.Scoring), the oracle (via orc_from_scoring) and the compiled
# reference (tools/gen_golden.py)
#   init: match, mismatch, gap_open, gap_extend, no_start, no_end,
#         no_gaps_in_a, no_gaps_in_b, no_mismatches, case_sensitive
SPECS = {
    "nw_default": dict(system="default"),
    "sw_cli": dict(system="default", poke=dict(match=2, mismatch=-2, gap_open=-2, gap_extend=-1)),
    "free_ends": dict(init=[1, -2, -4, -1, 1, 1, 0, 0, 0, 0]),
    "free_start": dict(init=[1, -1, -4, -1, 1, 0, 0, 0, 0, 0]),
    "free_end": dict(init=[1, -1, -4, -1, 0, 1, 0, 0, 0, 0]),
    "linear_gap": dict(init=[2, -3, 0, -2, 0, 0, 0, 0, 0, 0]),
    "free_gaps": dict(init=[3, -2, 0, 0, 0, 0, 0, 0, 0, 0]),      # gaps cost nothing (found by tools/gpu_fuzz.py: open == 0)
    "no_gaps_a": dict(init=[1, -2, -4, -1, 0, 0, 1, 0, 0, 0]),
    "no_gaps_b": dict(init=[1, -2, -4, -1, 0, 0, 0, 1, 0, 0]),
    "no_mismatch": dict(init=[1, -2, -4, -1, 0, 0, 0, 0, 1, 0]),
    "case_sens": dict(init=[1, -2, -4, -1, 0, 0, 0, 0, 0, 1]),
    "wild_n": dict(init=[1, -2, -4, -1, 0, 0, 0, 0, 0, 0], wildcards=[["N", 0]]),
    "no_mismatch_wild": dict(init=[1, -2, -4, -1, 0, 0, 0, 0, 1, 0], wildcards=[["N", -1]]),
    "mutations": dict(init=[3, -3, -5, -2, 0, 0, 0, 0, 0, 0],
                      mutations=[["a", "g", -1], ["g", "a", -1], ["c", "t", 1]]),
    "big_scores": dict(init=[300, -400, -500, -100, 0, 0, 0, 0, 0, 0]),
    "blosum62": dict(system="BLOSUM62"),
    "pam30": dict(system="PAM30"),
    "pam70": dict(system="PAM70"),
    "blosum80": dict(system="BLOSUM80"),
    "dna_hyb": dict(system="DNA_hybridization"),
}


def scoring_from_spec(spec):
    S = seqalign.Scoring
    if "system" in spec:
        sc = S.system(spec["system"])
    else:
        i = spec["init"]
        sc = S(i[0], i[1], i[2], i[3], *[bool(v) for v in i[4:]])
    if "poke" in spec:
        sc.poke(**spec["poke"])
    for c, v in spec.get("wildcards", []):
        sc.add_wildcard(c, v)
    for a, b, v in spec.get("mutations", []):
        sc.add_mutation(a, b, v)
    return sc


def scoring_specs():
    return {k: (lambda spec=v: scoring_from_spec(spec)) for k, v in SPECS.items()}
