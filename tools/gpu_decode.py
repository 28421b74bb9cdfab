#!/usr/bin/env python3
"""Device decoder (csrc/sa_decode.cu) on the GPU box: a FASTA / FASTQ / one-per-line file of n pairs of 150 bp
decoded in one chunk; records checked against the oracle's reader on the whole text; decode time (CUDA events
around the H2D of the text + every decode kernel) and the host reader's time on the same text beside it.

    python tools/gpu_decode.py [pairs]   >> profiles/decode_r02.jsonl"""
import ctypes, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import *
from test_reader import orc_records, host_records

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
L = 150
rng = np.random.default_rng(11)
seqs = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=(2 * n, L), dtype=np.uint8)]
nl = np.full((2 * n, 1), 10, np.uint8)


def fasta():
    names = np.frombuffer(b"".join(b">read%09d\n" % i for i in range(2 * n)), np.uint8).reshape(2 * n, 15)
    return np.concatenate([names, seqs, nl], axis=1).tobytes()


def fastq():
    names = np.frombuffer(b"".join(b"@read%09d\n" % i for i in range(2 * n)), np.uint8).reshape(2 * n, 15)
    plus = np.tile(np.frombuffer(b"+\n", np.uint8), (2 * n, 1))
    qual = np.full((2 * n, L), ord("I"), np.uint8)
    return np.concatenate([names, seqs, nl, plus, qual, nl], axis=1).tobytes()


def plain():
    return np.concatenate([seqs, nl], axis=1).tobytes()


lib = seqalign.load()
reads = seqalign.Reads(0)
pin = lib.seqalign_host_alloc
for kind, make in (("fasta", fasta), ("fastq", fastq), ("plain", plain)):
    text = make()
    nb = len(text)
    hp = pin(nb + 64)
    ctypes.memmove(hp, text, nb)
    best = 1e9
    for rep in range(4):
        rc = lib.seqalign_reads_decode(reads._h, hp, nb, 1, 1)
        assert rc == 0, (rc, lib.seqalign_reads_error(reads._h))
        best = min(best, reads.last_ms)
    reads._text = text
    R = reads.records
    sa, sb = reads.sequences(0), reads.sequences(1)
    ok_seq = all(sa[i] == seqs[2 * i].tobytes() and sb[i] == seqs[2 * i + 1].tobytes() for i in range(0, n, max(1, n // 5000)))
    ok_name = kind == "plain" or all(reads.name(i) == b"read%09d" % i for i in range(0, 2 * n, max(1, n // 2000)))
    # oracle reader on a prefix (it is a sequential C loop; the whole text takes a while)
    m = min(2 * n, 20000)
    per = nb // (2 * n)
    want, last, _, _ = orc_records(text[:m * per])
    got = [(reads.name(i), (sb if i & 1 else sa)[i >> 1]) for i in range(m)]
    t0 = time.time(); hrec, _ = host_records(lib, text[:min(nb, 64 << 20)]); host_s = time.time() - t0
    print(json.dumps(dict(what="device decode, %s, %d pairs of %d bp, one chunk, split into sides" % (kind, n, L), text_bytes=nb, records=R,
                          decode_ms_with_h2d=round(best, 3), text_gbs=round(nb / best / 1e6, 2),
                          pcie_only_ms_at_55gbs=round(nb / 55e6, 3),
                          sequences_equal_sampled=bool(ok_seq), names_equal_sampled=bool(ok_name),
                          oracle_prefix_records=m, oracle_prefix_equal=bool(got == want and last == 0),
                          host_reader_mb_s=round(min(nb, 64 << 20) / host_s / 1e6, 1), host_reader_records=len(hrec))), flush=True)
    lib.seqalign_host_free(hp)
