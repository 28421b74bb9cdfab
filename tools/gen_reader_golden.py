#!/usr/bin/env python3
"""Records tests/golden/reader_vectors.json: texts of sequence files (hand-written edge cases + seeded
random ones) with the records the REFERENCE's own reader returns for them (oracle/_ref/ref_reader, which
includes the unmodified libs/seq_file/seq_file.h and opens the file as align_from_file() does).

    python tools/gen_reader_golden.py          (build container only: needs /root/reference)

The fixtures travel; tests/test_reader.py checks the oracle's restatement (orc_read_records), the
library's host reader and the device decoder against them."""
import json, os, random, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_reader")


def ref_records(text):
    with tempfile.NamedTemporaryFile(delete=False) as f:
        f.write(text)
    try:
        out = subprocess.run([REF, f.name], capture_output=True, check=True).stdout
    finally:
        os.unlink(f.name)
    recs, pos = [], 0
    while True:
        nl = out.index(b"\n", pos)
        head = out[pos:nl].split()
        pos = nl + 1
        if head[0] == b"end":
            return recs, int(head[1])
        ln, ls = int(head[1]), int(head[2])
        name = out[pos:pos + ln]; pos += ln + 1
        seq = out[pos:pos + ls]; pos += ls + 1
        recs.append([name.decode("latin1"), seq.decode("latin1")])


HAND = [
    b"", b"\n", b"\n\n\n", b"ACGT", b"ACGT\n", b"ACGT\nTTGA\n", b"ACGT\r\nTTGA\r\n", b"\n\nACGT\n\n\nTT\n\n",
    b"  lead\nACGT\n", b"ACGT\n  skipme TTT\nGGGG\n\t\nCC CC \r\n", b"AC GT\nA\n", b"ACGT\n>x\nAAA\nCCC\n", b"ACGT\n@q\nAAA\n+\nIII\nGG\n",
    b">a\nACGT\n>b\nTTGA\n", b">a\nAC\nGT\n>b\nTT\nGA\n", b">a\r\nAC\r\nGT\r\n>b\r\nTT\r\n", b">a\nACGT", b">a\nACGT\n>b", b">a\nACGT\n>",
    b">a\n>b\nAC\n", b">a\n\n\nAC\n\nGT\n\n>b\n\n", b">a desc here\nAC>GT\nA>C\n>b\n>\nGG\n", b"\n \n>a\nAC\n", b">a\n AC\n\tGT \n",
    b">a\nAC\r\r\n\rGT\n", b">\nAC\n>\nGT\n", b">a\nACGT\n>b\nTTGA\n>c\nGG\n",
    b"@a\nACGT\n+\nIIII\n@b\nTTGA\n+\nIIII\n", b"@a\nACGT\n+a\nIIII\n@b\nTTGA\n+b\nIIII", b"@a\r\nACGT\r\n+\r\nIIII\r\n@b\r\nTT\r\n+\r\nII\r\n",
    b"@a\nAC\nGT\n+\nII\nII\n@b\nTT\n+\nII\n", b"@a\nACGT\n+\n@III\n@b\nTT\n+\n+I\n", b"@a\nACGT\n+\nII\n@b\nTT\n+\nII\n", b"@a\nACGT\n+\nIIIIII\n@b\nTT\n+\nII\n",
    b"@a\nACGT\n+\nIIII\n\n\n@b\nTT\n+\nII\n\n", b"@a\nACGT\n+\nIIII\njunk\n@b\nTT\n+\nII\n", b"@a\nACGT\n", b"@a\nACGT\n+\n", b"@a\nACGT\n+\nII", b"@a\n\n+\n\n@b\nA\n+\nI\n",
    b"@a\n+\n\n@b\nA\n+\nI\n", b"@a\nACGT\n+\nIIII\n@b\n+CGT\n+\nIIII\n", b"@a\nACGT\n+\nIIII\n@b\nTTGA\n+\nIIII\n@c\nA\n",
]


def rand_text(rng):
    kind = rng.choice(["plain", "fasta", "fastq", "fastq4", "fasta1"])
    nl = rng.choice([b"\n", b"\n", b"\r\n"])
    out = b""
    def seq(n): return bytes(rng.choice(b"ACGTacgtN") for _ in range(n))
    nrec = rng.randint(0, 9)
    if rng.random() < 0.2: out += rng.choice([b"\n", b"\n\n", nl])
    for r in range(nrec):
        L = rng.choice([0, 1, 2, 5, 17, 40, 80])
        if kind == "plain":
            out += seq(max(L, 1)) + nl
            if rng.random() < 0.15: out += nl
        elif kind in ("fasta", "fasta1"):
            out += b">r%d" % r + (b" d" if rng.random() < 0.3 else b"") + nl
            s = seq(L)
            w = 10 ** 6 if kind == "fasta1" else rng.choice([3, 7, 60])
            for i in range(0, len(s), w): out += s[i:i + w] + nl
            if rng.random() < 0.1: out += nl
        else:
            s = seq(max(L, 1))
            out += b"@r%d" % r + nl
            if kind == "fastq4" or rng.random() < 0.6: out += s + nl
            else:
                h = len(s) // 2
                out += s[:h] + nl + s[h:] + nl
            out += b"+" + (b"r%d" % r if rng.random() < 0.3 else b"") + nl
            q = bytes(rng.choice(b"IJ#@+5") for _ in range(len(s)))
            if kind == "fastq4" or rng.random() < 0.7: out += q + nl
            else:
                h = len(q) // 2
                out += q[:h] + nl + q[h:] + nl
    if out.endswith(nl) and rng.random() < 0.3: out = out[:-len(nl)]
    return out


def main():
    rng = random.Random(20261017)
    texts = list(HAND) + [rand_text(rng) for _ in range(160)]
    cases = []
    for t in texts:
        recs, last = ref_records(t)
        cases.append(dict(text=t.decode("latin1"), records=recs, last=last))
    path = os.path.join(ROOT, "tests", "golden", "reader_vectors.json")
    json.dump(dict(source="oracle/_ref/ref_reader = reference libs/seq_file/seq_file.h, seq_open() as in src/alignment_cmdline.c:570-596",
                   cases=cases), open(path, "w"), indent=0)
    print(len(cases), "cases ->", path)


if __name__ == "__main__":
    main()
