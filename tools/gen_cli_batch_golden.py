#!/usr/bin/env python3
"""Golden vectors for the batching command-line tools (tests/test_cli_batch.py).

Runs the reference's own tools (tests/integration/_ref_own: the unmodified
sources with the reference's own DP, built by tests/integration/Makefile) on
multi-pair inputs -- FASTA / FASTQ / plain / gzip, two-file input, stdin,
scoring files, error cases -- and records rc, stdout and the first stderr
line.  Build container only (needs /root/reference):

    make -C tests/integration && python tools/gen_cli_batch_golden.py

Smith-Waterman over several pairs: the reference reuses one sw_aligner_t whose
visited mask is only partly cleared between pairs, so its
multi-pair output is not a valid golden beyond pair 0.  Those cases are marked
"stitch": the expectation is built from one reference process per pair, with
the alignment counter rewritten.
"""
import gzip
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from helpers import ragged_batch, synthetic_batch  # noqa: E402

OWN = os.path.join(ROOT, "tests", "integration", "_ref_own")


def wrap(s, w=60):
    return "\n".join(s[i:i + w] for i in range(0, len(s), w))


def dna_pairs(seed, n, lo, hi):
    sa, sb = ragged_batch(seed, n, hi, hi, min_len=lo)
    return [(a.decode(), b.decode()) for a, b in zip(sa, sb)]


def prot_pairs(seed, n, lo, hi):
    sa, sb = ragged_batch(seed, n, hi, hi, alphabet=b"ARNDCQEGHILKMFPSTWYV", min_len=lo)
    return [(a.decode(), b.decode()) for a, b in zip(sa, sb)]


def fasta(pairs, width=60, tag=""):
    return "".join(">%sa%d some description\n%s\n>%sb%d\n%s\n" % (tag, i, wrap(a, width), tag, i, wrap(b, width))
                   for i, (a, b) in enumerate(pairs))


def fastq(pairs):
    out = []
    for i, (a, b) in enumerate(pairs):
        for nm, s in (("ra%d" % i, a), ("rb%d/2" % i, b)):
            out.append("@%s\n%s\n+\n%s\n" % (nm, s, "I" * len(s)))
    return "".join(out)


def plain(pairs, crlf=False):
    nl = "\r\n" if crlf else "\n"
    return "".join(a + nl + b + nl + ("" if i % 2 else nl) for i, (a, b) in enumerate(pairs))


P7 = dna_pairs(11, 7, 20, 75)
P5 = dna_pairs(12, 5, 30, 60)
P40 = dna_pairs(13, 40, 8, 48)
PROT6 = prot_pairs(14, 6, 25, 70)
SW8 = dna_pairs(15, 8, 25, 60)
SWP4 = prot_pairs(16, 4, 30, 60)
MIXED = [(a.lower() if i % 2 else a, b if i % 3 else b.lower()) for i, (a, b) in enumerate(P5)]

DNA_MATRIX = """# toy nucleotide matrix, NCBI layout
    A   C   G   T   N
A   5  -4  -4  -4  -2
C  -4   5  -4  -4  -2
G  -4  -4   5  -4  -2
T  -4  -4  -4   5  -2
N  -2  -2  -2  -2  -1
"""
PAIRS_WS = "# pairs, white space\na c -3\na g -1\nc a -3\ng a -1\nc t -1\nt c -1\n"
PAIRS_SEP = "a,c,-3\na,g,-1\nc,a,-3\ng,a,-1\n"

T = lambda text, gz=False: dict(text=text, gz=gz)  # noqa: E731

# (tool, argv, files, stdin, stitch)
CASES = [
    ("needleman_wunsch", ["--printscores", "--file", "@fa"], {"@fa": T(fasta(P7))}, None, False),
    ("needleman_wunsch", ["--printscores", "--printfasta", "--pretty", "--file", "@fq"], {"@fq": T(fastq(P5))}, None, False),
    ("needleman_wunsch", ["--printscores", "--file", "@txt"], {"@txt": T(plain(P5, crlf=True))}, None, False),
    ("needleman_wunsch", ["--printscores", "--printfasta", "--file", "@gz"], {"@gz": T(fasta(P7, 33), gz=True)}, None, False),
    ("needleman_wunsch", ["--printscores", "--files", "@f1", "@f2"],
     {"@f1": T("".join(">x%d\n%s\n" % (i, a) for i, (a, b) in enumerate(P5))), "@f2": T("".join(b + "\n" for a, b in P5))}, None, False),
    ("needleman_wunsch", ["--printscores", "--pretty", "--stdin"], {}, plain(P5), False),
    ("needleman_wunsch", ["--printscores", "--file", "-"], {}, fasta(P5), False),
    # lines that start with white space: skipped on --stdin, read (minus the white space) through --file
    ("needleman_wunsch", ["--printscores", "--file", "@ws"], {"@ws": T("AAAA\n  indented line\nCCCC\n\tACGT x\nGGGG\nTTTT\n")}, None, False),
    ("needleman_wunsch", ["--printscores", "--stdin"], {}, "AAAA\n  indented line\nCCCC\n\tACGT x\nGGGG\nTTTT\n", False),
    ("needleman_wunsch", ["--printscores", "--file", "@ws"], {"@ws": T("\n\n  ACGTAC\nACGTTC\n")}, None, False),
    ("needleman_wunsch", ["--printscores", "--file", "@a", "--file", "@b", "ACGTAC", "ACTTAC"],
     {"@a": T(fasta(P5[:2])), "@b": T(plain(P5[2:]))}, None, False),
    ("needleman_wunsch", ["--substitution_matrix", "@m", "--printscores", "--file", "@fa"],
     {"@m": T(DNA_MATRIX), "@fa": T(fasta(P5))}, None, False),
    ("needleman_wunsch", ["--substitution_matrix", "@m", "--gapopen", "-6", "--gapextend", "-2", "--printscores", "ACGNNTACGT", "ACGTTTACG"],
     {"@m": T(DNA_MATRIX, gz=True)}, None, False),
    ("needleman_wunsch", ["--substitution_pairs", "@p", "--match", "2", "--mismatch", "-4", "--printscores", "--file", "@fa"],
     {"@p": T(PAIRS_WS), "@fa": T(fasta(P5))}, None, False),
    ("needleman_wunsch", ["--substitution_pairs", "@p", "--match", "2", "--mismatch", "-4", "--printscores", "--pretty", "ACGGTCA", "ACAGTTA"],
     {"@p": T(PAIRS_SEP)}, None, False),
    ("needleman_wunsch", ["--scoring", "BLOSUM62", "--printscores", "--file", "@fa"], {"@fa": T(fasta(PROT6))}, None, False),
    # the reference's own NCBI matrix files (scoring/*.txt), embedded as fixtures
    ("needleman_wunsch", ["--substitution_matrix", "@b62", "--printscores", "--pretty", "--file", "@fa"],
     {"@b62": T(open("/root/reference/scoring/BLOSUM62.txt").read()), "@fa": T(fasta(PROT6[:3]))}, None, False),
    ("smith_waterman", ["--substitution_matrix", "@nuc", "--gapopen", "-6", "--minscore", "8", "--maxhits", "2", "ACGTNRYACGTTAGC", "ACGTACGTCAGC"],
     {"@nuc": T(open("/root/reference/scoring/NUC.4.4.txt").read())}, None, False),
    ("needleman_wunsch", ["--substitution_matrix", "@p250", "--printscores", "HEAGAWGHEE", "PAWHEAE"],
     {"@p250": T(open("/root/reference/scoring/PAM250.txt").read(), gz=True)}, None, False),
    ("needleman_wunsch", ["--scoring", "PAM70", "--freestartgap", "--freeendgap", "--printscores", "--pretty", "--file", "@fa"],
     {"@fa": T(fasta(PROT6))}, None, False),
    ("needleman_wunsch", ["--printscores", "--file", "@odd"], {"@odd": T(fasta(P5[:2]) + ">lonely\nACGT\n")}, None, False),
    ("needleman_wunsch", ["--printscores", "--file", "@empty"], {"@empty": T("")}, None, False),
    ("needleman_wunsch", ["--wildcard", "N", "0", "--printscores", "--pretty", "--file", "@fa"],
     {"@fa": T(fasta([(a[:10] + "NN" + a[12:], b) for a, b in P5]))}, None, False),
    ("needleman_wunsch", ["--printscores", "--colour", "--file", "@fa"], {"@fa": T(fasta(MIXED))}, None, False),
    ("needleman_wunsch", ["--case_sensitive", "--printscores", "--file", "@fa"], {"@fa": T(fasta(MIXED))}, None, False),
    ("needleman_wunsch", ["--zam", "--file", "@fa"], {"@fa": T(fasta(P5))}, None, False),
    ("needleman_wunsch", ["--nomismatches", "--printscores", "--file", "@fa"], {"@fa": T(fasta(P5))}, None, False),
    ("needleman_wunsch", ["--nogapsin1", "--printscores", "--file", "@fa"], {"@fa": T(fasta([(a, b) for a, b in P7 if len(a) <= len(b)]))}, None, False),
    ("needleman_wunsch", ["--gapopen", "0", "--gapextend", "-3", "--printscores", "--file", "@fa"], {"@fa": T(fasta(P40, 1000))}, None, False),
    ("needleman_wunsch", ["--printmatrices", "--printscores", "--file", "@fa"], {"@fa": T(fasta([(a[:9], b[:7]) for a, b in P5[:2]]))}, None, False),
    # --printmatrices over several pairs: the batch materialise mode (NW rows) behind alignment_print_matrices;
    # free start gaps: borders of 0, same mode
    ("needleman_wunsch", ["--printmatrices", "--scoring", "BLOSUM62", "--printscores", "--file", "@fa"],
     {"@fa": T(fasta([(a[:8], b[:11]) for a, b in PROT6[:3]]) + ">e1\n\n>e2\nHEAG\n")}, None, False),
    ("needleman_wunsch", ["--printmatrices", "--freestartgap", "--printscores", "--file", "@fa"],
     {"@fa": T(fasta([(a[:6], b[:8]) for a, b in P5[:2]]))}, None, False),
    # a character outside the loaded table: pairs before it are printed, then the reference exits
    ("needleman_wunsch", ["--substitution_matrix", "@m", "--printscores", "--file", "@fa"],
     {"@m": T(DNA_MATRIX), "@fa": T(fasta(P5[:2] + [("ACGTXACGT", "ACGTACGT")] + P5[2:]))}, None, False),
    # parse errors: rc and first stderr line
    ("needleman_wunsch", ["--bogus", "ACGT", "ACGT"], {}, None, False),
    ("needleman_wunsch", ["--match", "1", "ACGT", "ACGT"], {}, None, False),
    ("needleman_wunsch", ["--nomismatches", "--nogaps", "ACGT", "ACGT"], {}, None, False),
    ("needleman_wunsch", ["--minscore", "3", "ACGT", "ACGT"], {}, None, False),
    ("needleman_wunsch", ["--printscores"], {}, None, False),
    ("needleman_wunsch", ["--zam", "--pretty", "ACGT", "ACGT"], {}, None, False),
    ("needleman_wunsch", ["--match", "-5", "--mismatch", "2", "ACGT", "ACGT"], {}, None, False),
    ("smith_waterman", ["--printscores", "ACGT", "ACGT"], {}, None, False),
    ("smith_waterman", ["--maxhits", "x", "ACGT", "ACGT"], {}, None, False),
    # Smith-Waterman, several pairs
    ("smith_waterman", ["--file", "@fa"], {"@fa": T(fasta(SW8))}, None, True),
    ("smith_waterman", ["--maxhits", "3", "--minscore", "4", "--printfasta", "--file", "@fa"], {"@fa": T(fasta(SW8))}, None, True),
    ("smith_waterman", ["--minscore", "1", "--file", "@fa"], {"@fa": T(fasta(SW8[:4]))}, None, True),
    ("smith_waterman", ["--maxhits", "1", "--file", "@fq"], {"@fq": T(fastq(SW8))}, None, True),
    ("smith_waterman", ["--scoring", "BLOSUM62", "--minscore", "8", "--context", "5", "--pretty", "--printseq", "--printfasta", "--colour", "--file", "@fa"],
     {"@fa": T(fasta(SWP4))}, None, True),
    ("smith_waterman", ["--maxhits", "2", "--files", "@f1", "@f2"],
     {"@f1": T("".join(a + "\n" for a, b in SW8)), "@f2": T("".join(">t%d\n%s\n" % (i, b) for i, (a, b) in enumerate(SW8)), gz=True)}, None, True),
    ("smith_waterman", ["--minscore", "3", "--maxhits", "4", "--file", "@fa"],
     {"@fa": T(fasta(SW8[:2]) + ">e1\n\n>e2\nACGT\n" + fasta(SW8[2:4], tag="z"))}, None, True),
    ("smith_waterman", ["--nogaps", "--minscore", "3", "--maxhits", "3", "--file", "@fa"], {"@fa": T(fasta(SW8[:3]))}, None, True),
    ("smith_waterman", ["--gapopen", "0", "--minscore", "6", "--file", "@fa"], {"@fa": T(fasta(SW8[:5]))}, None, True),
    # --printmatrices over several pairs: the batch materialise mode behind alignment_print_matrices
    ("smith_waterman", ["--printmatrices", "--minscore", "3", "--maxhits", "2", "--file", "@fa"],
     {"@fa": T(fasta([(a[:14], b[:11]) for a, b in SW8[:4]]))}, None, True),
    ("smith_waterman", ["--printmatrices", "--maxhits", "1", "--minscore", "1", "--scoring", "BLOSUM62", "--file", "@fa"],
     {"@fa": T(fasta([(a[:9], b[:12]) for a, b in SWP4[:3]]))}, None, True),
    ("lcs", ["abcabcdabcdexabcd"], {}, None, False),
]


def materialise(td, argv, files):
    args = []
    for x in argv:
        if x in files:
            f = files[x]
            path = os.path.join(td, x[1:] + (".gz" if f["gz"] else ".txt"))
            if f["gz"]:
                with gzip.open(path, "wt") as fh:
                    fh.write(f["text"])
            else:
                with open(path, "w", newline="") as fh:
                    fh.write(f["text"])
            args.append(path)
        else:
            args.append(x)
    return args


def run(tool_dir, tool, argv, files, stdin):
    with tempfile.TemporaryDirectory() as td:
        args = materialise(td, argv, files)
        p = subprocess.run([os.path.join(tool_dir, tool)] + args, input=stdin, capture_output=True, text=True, timeout=300)
        err = p.stderr.replace(td, "<tmp>").split("\n")[0]
        return p.returncode, p.stdout, err


def read_records(text):
    """(name, seq) records of FASTA / FASTQ / plain text, the way the tools read them"""
    lines = text.split("\n")
    recs = []
    first = next((l for l in lines if l.strip()), "")
    if first.startswith(">"):
        for l in lines:
            l = l.rstrip("\r")
            if l.startswith(">"):
                recs.append([l[1:], ""])
            elif recs:
                recs[-1][1] += l
    elif first.startswith("@"):
        i = 0
        while i < len(lines):
            if lines[i].startswith("@"):
                recs.append([lines[i][1:], lines[i + 1]])
                i += 4
            else:
                i += 1
    else:
        recs = [["", l.rstrip("\r")] for l in lines if l.strip()]
    return recs


def stitch(tool, argv, files):
    """expected multi-pair SW output: one reference process per pair"""
    spots = [x for x in argv if x in files]
    recs = [read_records(files[x]["text"]) for x in spots]
    if len(spots) == 1:
        pairs = [(recs[0][i], recs[0][i + 1]) for i in range(0, len(recs[0]) - 1, 2)]
    else:
        pairs = list(zip(recs[0], recs[1]))
    out, idx = [], 0
    for (na, a), (nb, b) in pairs:
        if not a or not b:
            continue
        one = {"@one": T("%s%s\n%s%s\n" % (">%s\n" % na if na else "", a, ">%s\n" % nb if nb else "", b))}
        if (na == "") != (nb == ""):   # mixed named / unnamed: two files again
            one = {"@o1": T("%s%s\n" % (">%s\n" % na if na else "", a)), "@o2": T("%s%s\n" % (">%s\n" % nb if nb else "", b))}
            argv1 = [x for x in argv if x not in spots and x not in ("--file", "--files")] + ["--files", "@o1", "@o2"]
        else:
            argv1 = [x for x in argv if x not in spots and x not in ("--file", "--files")] + ["--file", "@one"]
        rc, so, _ = run(OWN, tool, argv1, one, None)
        assert rc == 0
        so = re.sub(r"^== Alignment 0 lengths", "== Alignment %d lengths" % idx, so, flags=re.M)
        so = re.sub(r"^hit 0\.(\d+) score", lambda m: "hit %d.%s score" % (idx, m.group(1)), so, flags=re.M)
        out.append(so)
        idx += 1
    return 0, "".join(out)


def main():
    out = []
    for tool, argv, files, stdin, st in CASES:
        if st:
            rc, so = stitch(tool, argv, files)
            err = None
        else:
            rc, so, err = run(OWN, tool, argv, files, stdin)
        out.append(dict(tool=tool, argv=argv, files=files, stdin=stdin, rc=rc, stdout=so,
                        stderr_first=err if rc != 0 else None, stitched=st))
    path = os.path.join(ROOT, "tests", "golden", "cli_batch_vectors.json")
    json.dump(dict(generator="tools/gen_cli_batch_golden.py", cases=out), open(path, "w"), indent=0)
    print("wrote %s: %d invocations, %d bytes, %d with rc != 0" % (path, len(out), os.path.getsize(path), sum(1 for c in out if c["rc"])))


if __name__ == "__main__":
    main()
