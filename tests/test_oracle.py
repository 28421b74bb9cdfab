"""The oracle is pinned before it is trusted (CPU only).

1. against the golden vectors the reference's own tests and README hold
   (src/tools/tests.c:65-268, README.md:71-74, README.md:118-145), typed in
   here literally;
2. against tests/golden/reference_vectors.json, produced by the unmodified
   reference library (tools/gen_golden.py);
3. differentially against oracle/_ref/libalign_ref.so when that library is
   present (build container, or shipped prebuilt to the GPU box).
"""
import ctypes
import json
import os

import numpy as np
import pytest

from helpers import (ROOT, REF_LIB, SPECS, orc_fill, orc_from_scoring, orc_nw, orc_sw_hits,
                     ragged_batch, scoring_from_spec)

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))


def _spec(name):
    return GOLD["specs"][name]


def test_readme_nw_example():
    # reference README.md:71-74, BASELINE config 1
    o = orc_from_scoring(scoring_from_spec(SPECS["nw_default"]))
    rc, score, ra, rb = orc_nw(o, b"CAGACGT", b"CGATA")
    assert (rc, score, ra, rb) == (0, -11, b"C-AGACGT", b"CGATA---")


def test_readme_matrices():
    # reference README.md:118-145: ACAGGT vs AAGGT, sentinel INT_MIN+5
    o = orc_from_scoring(scoring_from_spec(SPECS["nw_default"]))
    rc, m, ga, gb = orc_fill(o, b"ACAGGT", b"AAGGT", 0)
    assert rc == 0
    S = -2147483643
    assert m.tolist() == [[0, S, S, S, S, S, S],
                          [S, 1, -7, -5, -9, -10, -11],
                          [S, -4, -1, -3, -7, -8, -9],
                          [S, -8, -6, -3, -2, -6, -10],
                          [S, -9, -7, -8, -2, -1, -8],
                          [S, -10, -8, -9, -10, -4, 0]]
    assert ga.tolist() == [[0, S, S, S, S, S, S],
                           [-5, -10, -11, -12, -13, -14, -15],
                           [-6, -4, -9, -10, -11, -12, -13],
                           [-7, -5, -6, -8, -12, -13, -14],
                           [-8, -6, -7, -8, -7, -11, -13],
                           [-9, -7, -8, -9, -7, -6, -11]]
    assert gb.tolist() == [[0, -5, -6, -7, -8, -9, -10],
                           [S, -10, -4, -5, -6, -7, -8],
                           [S, -11, -9, -6, -7, -8, -9],
                           [S, -12, -10, -11, -8, -7, -8],
                           [S, -13, -11, -12, -13, -7, -6],
                           [S, -14, -12, -13, -14, -12, -9]]
    rc, score, ra, rb = orc_nw(o, b"ACAGGT", b"AAGGT")
    assert (ra, rb) == (b"ACAGGT", b"A-AGGT")


def test_reference_unit_vectors():
    # src/tools/tests.c:65-98, 102-131, 133-163, 233-268
    cs = dict(init=[1, -2, -4, -1, 0, 0, 1, 0, 0, 1])
    rc, _, ra, rb = orc_nw(orc_from_scoring(scoring_from_spec(cs)), b"aaaaacg", b"acgt")
    assert (rc, ra, rb) == (0, b"aaaaacg-", b"a----cgt")
    fe = dict(init=[1, -1, -4, -1, 1, 1, 0, 0, 0, 0])
    rc, score, ra, rb = orc_nw(orc_from_scoring(scoring_from_spec(fe)), b"acg", b"tttacgttt")
    assert (rc, score, ra, rb) == (0, 3, b"---acg---", b"tttacgttt")
    nm = orc_from_scoring(scoring_from_spec(SPECS["no_mismatch"]))
    assert orc_nw(nm, b"atc", b"ac")[2:] == (b"atc", b"a-c")
    assert orc_nw(nm, b"cgatcga", b"catcctcga")[2:] == (b"cgatc---ga", b"c-atcctcga")
    # SW: gacag vs tgaagt without gaps -> "ga"/"ga" then "ag"/"ag"
    ng = dict(init=[1, -2, -4, -1, 0, 0, 1, 1, 0, 0])
    n, hits = orc_sw_hits(orc_from_scoring(scoring_from_spec(ng)), b"gacag", b"tgaagt", 4)
    assert n >= 2
    assert (hits[0]["result_a"], hits[0]["result_b"]) == (b"ga", b"ga")
    assert (hits[1]["result_a"], hits[1]["result_b"]) == (b"ag", b"ag")


def test_no_mismatch_property():
    # src/tools/tests.c:176-218: with no_mismatches no column may pair two different letters
    o = orc_from_scoring(scoring_from_spec(SPECS["no_mismatch"]))
    sa, sb = ragged_batch(31, 50, 98, 98, alphabet=b"acgt", min_len=1)
    for a, b in zip(sa, sb):
        rc, _, ra, rb = orc_nw(o, a, b)
        assert rc == 0
        for x, y in zip(ra, rb):
            assert x == y or x == ord("-") or y == ord("-")


@pytest.mark.parametrize("chunk", range(4))
def test_golden_vectors(chunk):
    cases = GOLD["cases"][chunk::4]
    for c in cases:
        o = orc_from_scoring(scoring_from_spec(_spec(c["spec"])))
        a, b = c["a"].encode(), c["b"].encode()
        rc, score, ra, rb = orc_nw(o, a, b)
        assert (rc, score, ra.decode(), rb.decode()) == (0, c["nw"]["score"], c["nw"]["result_a"], c["nw"]["result_b"]), c
        n, hits = orc_sw_hits(o, a, b, 6)
        assert n == len(c["sw"]), c
        for h, e in zip(hits, c["sw"]):
            got = dict(h, result_a=h["result_a"].decode(), result_b=h["result_b"].decode())
            assert got == e, (c["spec"], c["a"], c["b"])
        for key, is_sw in (("nw_mats", 0), ("sw_mats", 1)):
            if key in c:
                rc, m, ga, gb = orc_fill(o, a, b, is_sw)
                assert rc == 0
                assert m.ravel().tolist() == c[key][0]
                assert ga.ravel().tolist() == c[key][1]
                assert gb.ravel().tolist() == c[key][2]


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built")
def test_differential_vs_compiled_reference():
    """seeded random pairs through the unmodified reference's aligner_align"""
    ref = ctypes.CDLL(REF_LIB)

    class RefAligner(ctypes.Structure):
        _fields_ = [("scoring", ctypes.c_void_p), ("seq_a", ctypes.c_void_p), ("seq_b", ctypes.c_void_p),
                    ("w", ctypes.c_size_t), ("h", ctypes.c_size_t),
                    ("m", ctypes.POINTER(ctypes.c_int)), ("ga", ctypes.POINTER(ctypes.c_int)),
                    ("gb", ctypes.POINTER(ctypes.c_int)), ("cap", ctypes.c_size_t)]

    ref.aligner_align.argtypes = [ctypes.POINTER(RefAligner), ctypes.c_char_p, ctypes.c_char_p,
                                  ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_char]
    for k, name in enumerate(("nw_default", "sw_cli", "free_ends", "wild_n", "blosum62", "no_mismatch_wild")):
        sc = scoring_from_spec(SPECS[name])
        o = orc_from_scoring(sc)
        alpha = b"ARNDCQEGHILKMFPSTWYV" if name == "blosum62" else b"ACGTNacgt"
        sa, sb = ragged_batch(500 + k, 40, 60, 60, alphabet=alpha)
        for a, b in zip(sa, sb):
            for is_sw in (0, 1):
                al = RefAligner()
                # the product's scoring_t has the reference's layout (test_abi), so it can be passed as is
                ref.aligner_align(ctypes.byref(al), a, b, len(a), len(b), ctypes.addressof(sc.s),
                                  ctypes.c_char(bytes([is_sw])))
                n = (len(a) + 1) * (len(b) + 1)
                rc, m, ga, gb = orc_fill(o, a, b, is_sw)
                assert rc == 0
                assert np.array_equal(m.ravel(), np.ctypeslib.as_array(al.m, (n,)))
                assert np.array_equal(ga.ravel(), np.ctypeslib.as_array(al.ga, (n,)))
                assert np.array_equal(gb.ravel(), np.ctypeslib.as_array(al.gb, (n,)))
                ref.aligner_destroy(ctypes.byref(al))
