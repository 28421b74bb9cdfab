#!/usr/bin/env python3
"""Generate tests/golden/*.json from the UNMODIFIED reference library.

Build-container only: needs oracle/_ref/libalign_ref.so (compiled by
oracle/Makefile from the sources under /root/reference).  Every case is
computed through the reference's own public API -- needleman_wunsch_align2,
smith_waterman_align2 + smith_waterman_fetch on a FRESH sw_aligner_t per pair
(the reused-aligner mask is stale upstream), aligner_align
for matrices -- and written as plain JSON so the GPU box, which has no
/root/reference, can check both the oracle and the CUDA path against it.

    python tools/gen_golden.py
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from helpers import SPECS, ragged_batch, REF_LIB  # noqa: E402

ref = ctypes.CDLL(REF_LIB)
SCORING_BYTES = 271428


class RefAlignment(ctypes.Structure):
    _fields_ = [("result_a", ctypes.c_char_p), ("result_b", ctypes.c_char_p),
                ("capacity", ctypes.c_size_t), ("length", ctypes.c_size_t),
                ("pos_a", ctypes.c_size_t), ("pos_b", ctypes.c_size_t),
                ("len_a", ctypes.c_size_t), ("len_b", ctypes.c_size_t), ("score", ctypes.c_int)]


class RefAligner(ctypes.Structure):
    _fields_ = [("scoring", ctypes.c_void_p), ("seq_a", ctypes.c_void_p), ("seq_b", ctypes.c_void_p),
                ("score_width", ctypes.c_size_t), ("score_height", ctypes.c_size_t),
                ("match_scores", ctypes.POINTER(ctypes.c_int)), ("gap_a_scores", ctypes.POINTER(ctypes.c_int)),
                ("gap_b_scores", ctypes.POINTER(ctypes.c_int)), ("capacity", ctypes.c_size_t)]


ref.alignment_create.restype = ctypes.POINTER(RefAlignment)
ref.needleman_wunsch_new.restype = ctypes.POINTER(RefAligner)
ref.smith_waterman_new.restype = ctypes.c_void_p
ref.smith_waterman_free.argtypes = [ctypes.c_void_p]
ref.smith_waterman_align2.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t,
                                      ctypes.c_void_p, ctypes.c_void_p]
ref.smith_waterman_fetch.argtypes = [ctypes.c_void_p, ctypes.POINTER(RefAlignment)]
ref.needleman_wunsch_align2.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t,
                                        ctypes.c_void_p, ctypes.POINTER(RefAligner), ctypes.POINTER(RefAlignment)]
ref.aligner_align.argtypes = [ctypes.POINTER(RefAligner), ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t,
                              ctypes.c_size_t, ctypes.c_void_p, ctypes.c_char]


def ref_scoring(spec):
    buf = ctypes.create_string_buffer(SCORING_BYTES)
    if "system" in spec:
        getattr(ref, "scoring_system_" + spec["system"])(buf)
    else:
        i = spec["init"]
        ref.scoring_init(buf, i[0], i[1], i[2], i[3], *[ctypes.c_bool(bool(v)) for v in i[4:]])
    if "poke" in spec:
        ints = (ctypes.c_int * 6).from_buffer(buf)  # gap_open, gap_extend, flags(8 bytes), match, mismatch
        for k, v in spec["poke"].items():
            ints[{"gap_open": 0, "gap_extend": 1, "match": 4, "mismatch": 5}[k]] = v
    for c, v in spec.get("wildcards", []):
        ref.scoring_add_wildcard(buf, ctypes.c_char(c.encode()), v)
    for a, b, v in spec.get("mutations", []):
        ref.scoring_add_mutation(buf, ctypes.c_char(a.encode()), ctypes.c_char(b.encode()), v)
    return buf


def ref_nw(buf, a, b):
    nw = ref.needleman_wunsch_new()
    res = ref.alignment_create(256)
    ref.needleman_wunsch_align2(a, b, len(a), len(b), buf, nw, res)
    r = res.contents
    out = dict(score=r.score, result_a=r.result_a.decode(), result_b=r.result_b.decode())
    ref.alignment_free(res)
    ref.needleman_wunsch_free(nw)
    return out


def ref_sw(buf, a, b, max_hits):
    sw = ref.smith_waterman_new()
    res = ref.alignment_create(256)
    ref.smith_waterman_align2(a, b, len(a), len(b), buf, sw)
    hits = []
    while len(hits) < max_hits and ref.smith_waterman_fetch(sw, res):
        r = res.contents
        hits.append(dict(score=r.score, result_a=r.result_a.decode(), result_b=r.result_b.decode(),
                         pos_a=r.pos_a, pos_b=r.pos_b, len_a=r.len_a, len_b=r.len_b))
    ref.alignment_free(res)
    ref.smith_waterman_free(sw)
    return hits


def ref_mats(buf, a, b, is_sw):
    al = RefAligner()
    ref.aligner_align(ctypes.byref(al), a, b, len(a), len(b), buf, ctypes.c_char(bytes([is_sw])))
    n = (len(a) + 1) * (len(b) + 1)
    out = [list(al.match_scores[:n]), list(al.gap_a_scores[:n]), list(al.gap_b_scores[:n])]
    ref.aligner_destroy(ctypes.byref(al))
    return out


def main():
    cases = []
    # the reference's own vectors, recomputed through its library
    fixed = [
        ("nw_default", "CAGACGT", "CGATA"),                       # README.md:71-74 / BASELINE config 1
        ("nw_default", "ACAGGT", "AAGGT"),                        # README.md:118-145 (matrices)
        ("free_ends_11", "acg", "tttacgttt"),                     # tests.c:102-131
        ("no_mismatch", "atc", "ac"),                             # tests.c:153-155
        ("no_mismatch", "cgatcga", "catcctcga"),                  # tests.c:157-159
        ("blosum62", "HEAGAWGHEE", "PAWHEAE"),
        ("sw_cli", "gacag", "tgaagt"),
        ("no_gaps_a_cs", "aaaaacg", "acgt"),                      # tests.c:65-98
        ("nw_default", "", "ACGT"), ("nw_default", "ACGT", ""), ("nw_default", "", ""),
        ("free_ends", "", "ACGT"), ("nw_default", "A", "A"), ("nw_default", "A", "C"),
    ]
    extra = dict(SPECS)
    extra["free_ends_11"] = dict(init=[1, -1, -4, -1, 1, 1, 0, 0, 0, 0])
    extra["no_gaps_a_cs"] = dict(init=[1, -2, -4, -1, 0, 0, 1, 0, 0, 1])
    for name, a, b in fixed:
        cases.append(dict(spec=name, a=a, b=b, small=True))
    # restricted-gap flags, one at a time (both together overflow upstream: SURVEY 8c H3)
    rng = np.random.default_rng(11)
    for name in ("no_gaps_a", "no_gaps_b"):
        for _ in range(8):
            la, lb = int(rng.integers(1, 24)), int(rng.integers(1, 24))
            a = "".join(rng.choice(list("acgt"), la))
            b = "".join(rng.choice(list("acgt"), lb))
            cases.append(dict(spec=name, a=a, b=b, small=True))
    # random sweeps per spec
    for k, name in enumerate(sorted(SPECS)):
        if name in ("no_gaps_a", "no_gaps_b"):
            continue
        alpha = b"ARNDCQEGHILKMFPSTWYVBZX" if name in ("blosum62", "pam30", "pam70", "blosum80") else (
            b"ACGT" if name == "dna_hyb" else b"ACGTNacgtn")
        sa, sb = ragged_batch(100 + k, 10, 48, 48, alphabet=alpha)
        for a, b in zip(sa, sb):
            cases.append(dict(spec=name, a=a.decode(), b=b.decode(), small=len(a) * len(b) <= 400))
    # a few wider pairs (strip-mining, > 256 columns)
    for k, name in enumerate(("nw_default", "sw_cli", "free_ends", "blosum62")):
        alpha = b"ARNDCQEGHILKMFPSTWYV" if name == "blosum62" else b"ACGT"
        sa, sb = ragged_batch(200 + k, 2, 620, 330, alphabet=alpha, min_len=280)
        for a, b in zip(sa, sb):
            cases.append(dict(spec=name, a=a.decode(), b=b.decode(), small=False))

    for c in cases:
        buf = ref_scoring(extra[c["spec"]])
        a, b = c["a"].encode(), c["b"].encode()
        c["nw"] = ref_nw(buf, a, b)
        c["sw"] = ref_sw(buf, a, b, 6)
        if c.pop("small"):
            c["nw_mats"] = ref_mats(buf, a, b, 0)
            c["sw_mats"] = ref_mats(buf, a, b, 1)
    out = dict(generator="tools/gen_golden.py", reference="noporpoise/seq-align @ dc41988",
               glibc=os.confstr("CS_GNU_LIBC_VERSION"), specs=extra, cases=cases)
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote %s: %d cases, %d bytes" % (path, len(cases), os.path.getsize(path)))


if __name__ == "__main__":
    main()
