"""ABI and host-logic checks that need no GPU.

* struct layouts of scoring_t / aligner_t / alignment_t match the reference
  (callers stack-allocate and poke these structs);
* the product library loads without a device and exports every symbol that
  include/*.h declares;
* it refuses to run without a device instead of falling back to a CPU path;
* the host scoring functions (scoring_init, add_*, scoring_lookup, built-in
  systems) agree with the compiled reference.
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import seqalign
from helpers import ROOT, REF_LIB, SPECS, scoring_from_spec

REAL_LIB = os.path.join(ROOT, "seq-align_b200", "lib", "libseqalign_b200.so")


@pytest.fixture(scope="module")
def real_lib():
    subprocess.check_call(["make", "-s", "-C", ROOT])   # no-op when up to date
    return ctypes.CDLL(REAL_LIB)


def test_struct_layout():
    S = seqalign.ScoringT
    assert ctypes.sizeof(S) == 271428
    assert (S.gap_open.offset, S.gap_extend.offset) == (0, 4)
    assert S.no_start_gap_penalty.offset == 8 and S.use_match_mismatch.offset == 13
    assert (S.match.offset, S.mismatch.offset, S.case_sensitive.offset) == (16, 20, 24)
    assert (S.wildcards.offset, S.swap_set.offset) == (28, 60)
    assert (S.wildscores.offset, S.swap_scores.offset) == (8252, 9276)
    assert (S.min_penalty.offset, S.max_penalty.offset) == (271420, 271424)
    assert ctypes.sizeof(seqalign.AlignerT) == 9 * 8
    assert ctypes.sizeof(seqalign.AlignmentT) == 8 * 8 + 8


def test_c_sizeof_matches(tmp_path):
    """compile a probe against include/*.h: C's view of the structs"""
    src = tmp_path / "probe.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "smith_waterman.h"\n#include "needleman_wunsch.h"\n'
        '#include "seqalign_b200.h"\n'
        'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(scoring_t), offsetof(scoring_t, swap_scores),'
        'offsetof(scoring_t, min_penalty), sizeof(aligner_t), sizeof(alignment_t), offsetof(alignment_t, score));return 0;}\n')
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).split()
    assert [int(v) for v in out] == [271428, 9276, 271420, 72, 72, 64]


def test_cmdline_struct_layout(tmp_path):
    """cmdline_t / read_t are read field by field by the reference's tool mains (src/tools/nw_cmdline.c:81-143,
    sw_cmdline.c:192-217, 316-322).  The numbers are those of the reference's own headers
    (src/alignment_cmdline.h:24-56, libs/seq_file/seq_file.h:61-73), probed with the same program."""
    src = tmp_path / "probe.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "alignment_cmdline.h"\n#include "alignment_scoring_load.h"\n'
        '#include "alignment_macros.h"\n'
        'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %d\\n", sizeof(cmdline_t), offsetof(cmdline_t, file_paths2),'
        'offsetof(cmdline_t, min_score), offsetof(cmdline_t, print_seq), offsetof(cmdline_t, zam_stle_output),'
        'offsetof(cmdline_t, interactive), offsetof(cmdline_t, no_mismatches), offsetof(cmdline_t, seq1), sizeof(read_t),'
        '(int)ARR_2D_INDEX(7, 3, 2) + MAX4(1, 9, 3, 2) + MIN3(4, 2, 8) + ABSDIFF(3, 10));return 0;}\n')
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-std=gnu99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).split()
    assert [int(v) for v in out] == [96, 24, 52, 66, 71, 72, 78, 80, 96, 17 + 9 + 2 + 7]


def _declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        text = open(os.path.join(inc, fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        text = re.sub(r"#[^\n]*(\\\n[^\n]*)*", "", text)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text):
            name = m.group(1)
            if name.startswith(("seqalign_", "scoring_", "aligner_", "alignment_", "needleman_", "smith_", "cmdline_",
                                "parse_entire_", "align_from_", "align_scoring_")):
                names.add(name)
        for m in re.finditer(r"\b(align_col_[a-z]+)\b", text):
            names.add(m.group(1))
    names.discard("aligner_init")  # macro
    return sorted(names)


def test_library_exports_every_declared_symbol(real_lib):
    syms = _declared_symbols()
    assert len(syms) > 50
    # the command-line layer of the reference's libalign.a (src/alignment_cmdline.h:58-72, alignment_scoring_load.h:14-18)
    for name in ("cmdline_new", "cmdline_free", "cmdline_add_files", "cmdline_get_num_of_file_pairs", "cmdline_get_file1",
                 "cmdline_get_file2", "align_from_file", "parse_entire_int", "parse_entire_uint",
                 "align_scoring_load_matrix", "align_scoring_load_pairwise"):
        assert name in syms, name
    missing = [s for s in syms if not hasattr(real_lib, s)]
    assert not missing, missing


def test_no_cpu_fallback(real_lib):
    """without a device the engine must refuse, loudly"""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is visible")
    except ImportError:
        pass
    real_lib.seqalign_batch_create.restype = ctypes.c_void_p
    real_lib.seqalign_last_create_error.restype = ctypes.c_char_p
    assert real_lib.seqalign_device_count() == 0
    assert real_lib.seqalign_batch_create(0) is None
    assert b"no CPU path" in real_lib.seqalign_last_create_error()


def test_product_does_not_link_oracle(real_lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", REAL_LIB]).decode()
    assert "orc_" not in out
    src = ""
    for d in ("csrc", "host", "seqalign"):
        for fn in os.listdir(os.path.join(ROOT, "seq-align_b200", d)):
            if fn.endswith((".c", ".cu", ".cuh", ".h", ".py")):
                src += open(os.path.join(ROOT, "seq-align_b200", d, fn)).read()
    assert "sa_oracle" not in src and "liboracle" not in src


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built")
def test_host_scoring_matches_reference():
    ref = ctypes.CDLL(REF_LIB)
    letters = b"ARNDCQEGHILKMFPSTWYVBZX*acgtnACGTN-xyz"
    for name in ("PAM30", "PAM70", "BLOSUM80", "BLOSUM62", "DNA_hybridization", "default"):
        mine = seqalign.Scoring.system(name)
        buf = ctypes.create_string_buffer(271428)
        getattr(ref, "scoring_system_" + name)(buf)
        theirs = seqalign.ScoringT.from_buffer(buf)
        for f in ("gap_open", "gap_extend", "no_start_gap_penalty", "no_end_gap_penalty", "no_gaps_in_a",
                  "no_gaps_in_b", "no_mismatches", "use_match_mismatch", "match", "mismatch",
                  "case_sensitive", "min_penalty", "max_penalty"):
            assert getattr(mine.s, f) == getattr(theirs, f), (name, f)
        assert bytes(mine.s.wildcards) == bytes(theirs.wildcards)
        assert bytes(mine.s.swap_set) == bytes(theirs.swap_set)
        if name == "DNA_hybridization":
            continue  # lookups of unknown pairs exit(1) by contract
        sc, im = ctypes.c_int(), ctypes.c_bool()
        for a in letters:
            for b in letters:
                ref.scoring_lookup(buf, ctypes.c_char(bytes([a])), ctypes.c_char(bytes([b])),
                                   ctypes.byref(sc), ctypes.byref(im))
                assert mine.lookup(bytes([a]), bytes([b])) == (sc.value, im.value), (name, a, b)


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built")
def test_host_scoring_specs_match_reference():
    ref = ctypes.CDLL(REF_LIB)
    for name, spec in SPECS.items():
        if "system" in spec:
            continue
        mine = scoring_from_spec(spec)
        buf = ctypes.create_string_buffer(271428)
        i = spec["init"]
        ref.scoring_init(buf, i[0], i[1], i[2], i[3], *[ctypes.c_bool(bool(v)) for v in i[4:]])
        for c, v in spec.get("wildcards", []):
            ref.scoring_add_wildcard(buf, ctypes.c_char(c.encode()), v)
        for a, b, v in spec.get("mutations", []):
            ref.scoring_add_mutation(buf, ctypes.c_char(a.encode()), ctypes.c_char(b.encode()), v)
        theirs = seqalign.ScoringT.from_buffer(buf)
        assert (mine.s.min_penalty, mine.s.max_penalty) == (theirs.min_penalty, theirs.max_penalty), name
        assert bytes(mine.s.wildcards) == bytes(theirs.wildcards)
        assert bytes(mine.s.swap_set) == bytes(theirs.swap_set)
        sc, im = ctypes.c_int(), ctypes.c_bool()
        for a in b"acgtnACGTN":
            for b in b"acgtnACGTN":
                ref.scoring_lookup(buf, ctypes.c_char(bytes([a])), ctypes.c_char(bytes([b])),
                                   ctypes.byref(sc), ctypes.byref(im))
                assert mine.lookup(bytes([a]), bytes([b])) == (sc.value, im.value), (name, a, b)


@pytest.mark.parity
def test_shared_buffer_entry_points(backend):
    """seqalign_shared_alloc / open / close / free and seqalign_enable_peer_access: a buffer the
    owner allocates and a peer maps.  Single process here: opening one's own handle is the case
    CUDA IPC forbids, so on the GPU only alloc/free and the self-peer no-op are exercised; the
    cross-process path runs in tools/gpu_config5.py --peer."""
    import ctypes
    import seqalign
    L = seqalign.load()
    ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
    assert L.seqalign_shared_alloc(0, 4096, ctypes.byref(ptr), handle) == 0 and ptr.value
    assert any(handle.raw)
    if backend == "emu":
        peer = ctypes.c_void_p()
        assert L.seqalign_shared_open(0, handle, ctypes.byref(peer)) == 0 and peer.value == ptr.value
        assert L.seqalign_shared_close(0, peer) == 0
    assert L.seqalign_shared_free(0, ptr) == 0
    assert L.seqalign_enable_peer_access(0, 0) == 0
    assert L.seqalign_shared_alloc(0, 16, None, handle) == seqalign.ERR_ARG

