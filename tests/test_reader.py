"""Sequence-file reading: FASTA / FASTQ / one-per-line text -> records.

tests/golden/reader_vectors.json holds texts with the records the REFERENCE's own reader
(libs/seq_file/seq_file.h, opened as align_from_file() opens it) returns for them, recorded by
tools/gen_reader_golden.py through oracle/_ref/ref_reader.  Checked against them:
  * the oracle's restatement orc_read_records (oracle/sa_oracle.c)              -- CPU
  * the library's host reader sa_reader_* (seq-align_b200/host/sa_cli.c)        -- CPU
  * the device decoder seqalign_reads_* (csrc/sa_decode.cuh)                    -- parity backends
"""
import ctypes
import json
import os
import tempfile

import numpy as np
import pytest

from helpers import ROOT, oracle

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "reader_vectors.json")))["cases"]


def orc_records(text):
    """(records [(name, seq)], last status, rec_pos, fmt) from the oracle"""
    o = oracle()
    n = len(text)
    cap = n + 2
    seq = ctypes.create_string_buffer(n + 1)
    arr = lambda: (ctypes.c_longlong * (cap + 1))()
    seq_off, name_pos, name_len, rec_pos = arr(), arr(), arr(), arr()
    fmt = (ctypes.c_int * (cap + 1))()
    last = ctypes.c_int(0)
    o.orc_read_records.restype = ctypes.c_long
    cnt = o.orc_read_records(text, ctypes.c_size_t(n), ctypes.c_size_t(cap), seq, seq_off, name_pos, name_len, rec_pos, fmt,
                             ctypes.byref(last))
    recs = [(text[name_pos[i]:name_pos[i] + name_len[i]], seq.raw[seq_off[i]:seq_off[i + 1]]) for i in range(cnt)]
    return recs, last.value, [rec_pos[i] for i in range(cnt)], [fmt[i] for i in range(cnt)]


@pytest.mark.parametrize("i", range(len(GOLD)))
def test_oracle_reader_matches_reference(i):
    case = GOLD[i]
    text = case["text"].encode("latin1")
    recs, last, _, _ = orc_records(text)
    want = [(n.encode("latin1"), s.encode("latin1")) for n, s in case["records"]]
    assert recs == want
    assert last == case["last"]


class _Str(ctypes.Structure):
    _fields_ = [("b", ctypes.c_char_p), ("len", ctypes.c_size_t), ("cap", ctypes.c_size_t)]


class _Rec(ctypes.Structure):
    _fields_ = [("name", _Str), ("seq", _Str)]


def host_records(lib, text):
    """the library's host reader (buffered, as the tools open files)"""
    with tempfile.NamedTemporaryFile(delete=False) as f:
        f.write(text)
    try:
        lib.sa_reader_open.restype = ctypes.c_void_p
        r = lib.sa_reader_open(f.name.encode(), 1)
        assert r
        rec = _Rec()
        out = []
        while True:
            s = lib.sa_reader_next(ctypes.c_void_p(r), ctypes.byref(rec))
            if s <= 0:
                break
            out.append((ctypes.string_at(rec.name.b, rec.name.len), ctypes.string_at(rec.seq.b, rec.seq.len)))
        lib.sa_reader_close(ctypes.c_void_p(r))
        return out, s
    finally:
        os.unlink(f.name)


@pytest.fixture(scope="module")
def hostlib():
    import seqalign
    return seqalign.load()


@pytest.mark.parametrize("i", range(len(GOLD)))
def test_host_reader_matches_reference(hostlib, i):
    case = GOLD[i]
    recs, last = host_records(hostlib, case["text"].encode("latin1"))
    want = [(n.encode("latin1"), s.encode("latin1")) for n, s in case["records"]]
    assert recs == want
    assert last == case["last"]


# ---------------------------------------------------------------------------
# the device decoder (parity backends: B200 / lane emulator)

import seqalign
from seqalign import NW, SW, MODE_SCORE, MODE_ALIGN


@pytest.fixture(scope="module")
def reads(backend):
    r = seqalign.Reads(0)
    yield r
    r.close()


def _want(case):
    return [(n.encode("latin1"), s.encode("latin1")) for n, s in case["records"]]


@pytest.mark.parity
@pytest.mark.parametrize("split", [False, True])
def test_device_decoder_matches_reference(reads, split):
    """every golden text through seqalign_reads_decode as a whole input (final = 1): either the records
    of the reference's reader, bit for bit (names, sequences, order), or a decline -- and declines only
    where the text is outside the documented device grammar"""
    taken = declined = 0
    for ci, case in enumerate(GOLD):
        text = case["text"].encode("latin1")
        want = _want(case)
        if not reads.decode(text, final=True, split=split):
            declined += 1
            continue
        taken += 1
        n = reads.records
        got_names = [reads.name(i) for i in range(n)]
        if split:
            sa, sb = reads.sequences(0), reads.sequences(1)
            got_seqs = [(sb if i & 1 else sa)[i >> 1] for i in range(n)]
            assert len(sa) == (n + 1) // 2 and len(sb) == n // 2
        else:
            got_seqs = reads.sequences(0)
        assert list(zip(got_names, got_seqs)) == want, (ci, case["text"][:80])
        # record starts: the oracle's positions (first character of the record's line)
        _, _, pos, fmt = orc_records(text)
        for i in range(n):
            p = reads.record_start(i)
            assert p <= pos[i] and text[p:pos[i]].strip(b" \t\r\x0b\x0c") == b"", (ci, i)
    assert taken >= 150 and declined <= 45, (taken, declined)


@pytest.mark.parity
def test_device_decoder_declines_only_outside_its_grammar(reads):
    regular = [b">a\nACGT\n>b\nTTGA\n", b">a\r\nAC\r\nGT\r\n>b\r\nTT\r\n", b"ACGT\nTTGA\n", b"\n\nACGT\n\n\nTT\n\n",
               b"@a\nACGT\n+\nIIII\n@b\nTTGA\n+\nIIII\n", b"@a\nACGT\n+a\nIIII\n@b\nTTGA\n+b\nIIII", b">a\n AC\n\tGT \n",
               b"@a\nACGT\n+\n@III\n@b\nTT\n+\n+I\n", b"  lead\nACGT\n", b">a\nAC\r\r\n\rGT\n"]
    irregular = [b"ACGT\n>x\nAAA\nCCC\n", b"ACGT\n@q\nAAA\n+\nIII\nGG\n", b"@a\nAC\nGT\n+\nII\nII\n@b\nTT\n+\nII\n",
                 b"@a\nACGT\n+\nII\n@b\nTT\n+\nII\n", b"@a\nACGT\n+\nIIII\n\n\n@b\nTT\n+\nII\n\n", b"@a\nACGT\n", b"@a\nACGT\n+\nIIII\n@c\nA\n"]
    for t in regular:
        assert reads.decode(t), t
    for t in irregular:
        assert not reads.decode(t), t


def _chunked_records(reads, text, chunk, split):
    """feed `text` in pieces of `chunk` bytes the way the tools do: decode what is there, keep the complete
    records (an even number of them when records pair up), carry the rest in front of the next piece"""
    out, pos, carry = [], 0, b""
    while True:
        piece = text[pos:pos + chunk]
        pos += len(piece)
        final = pos >= len(text)
        buf = carry + piece
        ok = reads.decode(buf, final=final, split=split)
        assert ok
        n = reads.records
        if split and not final:
            n -= n & 1
        names = [reads.name(i) for i in range(n)]
        if split:
            sa, sb = reads.sequences(0), reads.sequences(1)
            seqs = [(sb if i & 1 else sa)[i >> 1] for i in range(n)]
        else:
            seqs = reads.sequences(0)[:n]
        out += list(zip(names, seqs))
        if final:
            return out
        carry = buf[reads.record_start(n):]


@pytest.mark.parity
@pytest.mark.parametrize("kind", ["fasta", "fasta_wrapped", "fastq", "plain", "plain_crlf", "plain_long_lines"])
def test_device_decoder_chunked(reads, kind):
    """a longer file cut at arbitrary byte positions: the held-back tail + carry protocol loses and
    duplicates nothing (records equal to the oracle's on the whole text)"""
    rng = np.random.default_rng(7)
    recs = []
    text = b""
    for i in range(40):
        L = int(rng.integers(1, 90)) if kind != "plain_long_lines" else int(rng.integers(600, 1500))   # long lines: a warp per line
        s = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=L))
        nl = b"\r\n" if kind == "plain_crlf" else b"\n"
        if kind == "fasta":
            text += b">r%d some text" % i + nl + s + nl
        elif kind == "fasta_wrapped":
            text += b">r%d" % i + nl + b"".join(s[k:k + 17] + nl for k in range(0, L, 17))
        elif kind == "fastq":
            text += b"@r%d" % i + nl + s + nl + b"+" + nl + b"I" * L + nl
        else:
            text += s + nl + (nl if i % 7 == 3 else b"")
    if kind != "fastq":
        text = text[:-1]          # no newline at the end of the input
    want, last, _, _ = orc_records(text)
    assert last == 0 and len(want) == 40
    for chunk in ((len(text), 997, 311, 150) if kind != "plain_long_lines" else (len(text), 7001)):
        for split in (False, True):
            assert _chunked_records(reads, text, chunk, split) == want, (kind, chunk, split)


@pytest.mark.parity
def test_decoded_reads_align_in_place(reads, engine, big):
    """decode -> seqalign_batch_submit_reads: the DP kernels read the decoder's packed buffers in HBM;
    scores and strings equal the host-packed submit of the same records"""
    from helpers import ragged_batch, scoring_from_spec, SPECS
    sa, sb = ragged_batch(21, 64 if big else 20, 140 if big else 50, 140 if big else 50, min_len=1)
    text = b"".join(b">a%d\n%s\n>b%d\n%s\n" % (i, sa[i], i, sb[i]) for i in range(len(sa)))
    assert reads.decode(text, final=True, split=True)
    n = reads.records // 2
    assert n == len(sa)
    engine.set_scoring(scoring_from_spec(SPECS["sw_cli"]))
    from seqalign import MODE_HITS, MODE_MATS, MODE_SCORE_ONLY
    for first in (0, 7):
        m = n - first
        a, oa = seqalign.pack(sa[first:])
        b, ob = seqalign.pack(sb[first:])
        probe = (0, 5, m - 1)
        for algo, mode in ((SW, MODE_SCORE), (SW, MODE_SCORE_ONLY), (NW, MODE_SCORE), (SW, MODE_ALIGN), (NW, MODE_ALIGN),
                           (SW, MODE_HITS), (NW, MODE_MATS)):
            if mode == MODE_HITS:
                engine.set_hit_limits(4, 5)
            engine.submit_packed(algo, mode, a, oa, b, ob)
            want_scores = engine.scores().copy() if mode not in (MODE_HITS, MODE_MATS) else None
            want_al = [engine.alignment(i) for i in probe] if mode == MODE_ALIGN else []
            want_hits = [engine.hits(i) for i in probe] if mode == MODE_HITS else []
            want_mats = [engine.matrices(i, len(sa[first + i]), len(sb[first + i])) for i in probe] if mode == MODE_MATS else []
            engine.submit_reads(algo, mode, reads, 0, reads, 1, m, first=first)
            if want_scores is not None:
                assert np.array_equal(engine.scores(), want_scores), (first, algo, mode)
            for i, w in zip(probe, want_al):
                g = engine.alignment(i)
                assert (g.score, g.result_a, g.result_b) == (w.score, w.result_a, w.result_b)
            for i, w in zip(probe, want_hits):
                got = engine.hits(i)
                assert [(x.score, x.result_a, x.result_b, x.pos_a, x.pos_b) for x in got] == \
                       [(x.score, x.result_a, x.result_b, x.pos_a, x.pos_b) for x in w]
            for i, w in zip(probe, want_mats):
                g = engine.matrices(i, len(sa[first + i]), len(sb[first + i]))
                assert all(np.array_equal(x, y) for x, y in zip(g, w))
