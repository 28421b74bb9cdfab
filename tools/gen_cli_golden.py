#!/usr/bin/env python3
"""Record stdout of the reference's own command-line tools (tests/integration/_ref_own, built
from the unmodified sources with the reference's own DP) for a list of invocations.
tests/test_cli_dropin.py replays them against the same tools linked with this repository's
library (tests/integration/_ref_cli).  Build container only.

    make -C tests/integration && python tools/gen_cli_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import synthetic_batch  # noqa: E402

OWN = os.path.join(ROOT, "tests", "integration", "_ref_own")

a, oa, b, ob = synthetic_batch(2, 3, 150, 150)
R = [(a[i * 150:(i + 1) * 150].tobytes().decode(), b[i * 150:(i + 1) * 150].tobytes().decode()) for i in range(3)]
pa, _, pb, _ = synthetic_batch(4, 1, 120, 110, kind="protein")
P = (pa.tobytes().decode(), pb.tobytes().decode())
FASTA3 = "".join(">a%d\n%s\n>b%d\n%s\n" % (i, x, i, y) for i, (x, y) in enumerate(R))
FASTA1 = ">r1\n%s\n>r2\n%s\n" % R[0]
PLAIN = "CAGACGT\nCGATA\nACAGGT\nAAGGT\n"

# (tool, argv, files {placeholder: content}, stdin)
CASES = [
    ("needleman_wunsch", ["--printscores", "CAGACGT", "CGATA"], {}, None),
    ("needleman_wunsch", ["--printscores", "--pretty", "CAGACGT", "CGATA"], {}, None),
    ("needleman_wunsch", ["--printmatrices", "ACAGGT", "AAGGT"], {}, None),
    ("needleman_wunsch", ["--freestartgap", "--freeendgap", "--match", "1", "--mismatch", "-1", "--printscores", "acg", "tttacgttt"], {}, None),
    ("needleman_wunsch", ["--scoring", "BLOSUM62", "--printscores", "--pretty", "HEAGAWGHEE", "PAWHEAE"], {}, None),
    ("needleman_wunsch", ["--scoring", "PAM30", "--printscores", P[0], P[1]], {}, None),
    ("needleman_wunsch", ["--nomismatches", "--printscores", "cgatcga", "catcctcga"], {}, None),
    ("needleman_wunsch", ["--nogapsin1", "--case_sensitive", "--printscores", "aaaaacg", "acgt"], {}, None),
    ("needleman_wunsch", ["--wildcard", "N", "0", "--printscores", "--pretty", "ACGNNTAC", "ACGTTTAC"], {}, None),
    ("needleman_wunsch", ["--gapopen", "0", "--gapextend", "-2", "--printscores", "--colour", R[0][0][:60], R[0][1][:50]], {}, None),
    ("needleman_wunsch", ["--zam", R[1][0][:40], R[1][1][:45]], {}, None),
    ("needleman_wunsch", ["--printscores", "--printfasta", "--file", "@F3"], {"@F3": FASTA3}, None),
    ("needleman_wunsch", ["--printscores", "--pretty", "--stdin"], {}, PLAIN),
    ("needleman_wunsch", ["--printscores", "--freestartgap", "--freeendgap", R[2][0], R[2][1]], {}, None),
    ("smith_waterman", ["--maxhits", "1", "--minscore", "1", "--scoring", "BLOSUM62", "HEAGAWGHEE", "PAWHEAE"], {}, None),
    ("smith_waterman", ["--minscore", "1", "gacag", "tgaagt"], {}, None),
    ("smith_waterman", ["--minscore", "2", "--nogaps", "gacag", "tgaagt"], {}, None),
    ("smith_waterman", ["--printseq", "--context", "3", "--minscore", "4", "--pretty", "ACGTACGTTTGACCA", "TTACGTACGAAGACC"], {}, None),
    ("smith_waterman", ["--maxhits", "3", R[0][0], R[0][1]], {}, None),
    ("smith_waterman", [R[1][0], R[1][1]], {}, None),
    ("smith_waterman", ["--scoring", "BLOSUM62", "--maxhits", "1", "--minscore", "1", P[0], P[1]], {}, None),
    ("smith_waterman", ["--maxhits", "2", "--colour", "--printfasta", "--file", "@F1"], {"@F1": FASTA1}, None),
    ("smith_waterman", ["--printmatrices", "--minscore", "1", "ACAGGT", "AAGGT"], {}, None),
    ("lcs", ["abcabcdabcdexabcd"], {}, None),
    ("lcs", [R[0][0][:70]], {}, None),
    ("nw_example", ["CAGACGT", "CGATA"], {}, None),
    ("sw_example", ["gacagtttacg", "tgaagtacgttt"], {}, None),
]


def run(tool_dir, tool, argv, files, stdin):
    with tempfile.TemporaryDirectory() as td:
        args = []
        for x in argv:
            if x in files:
                path = os.path.join(td, x[1:] + ".fa")
                open(path, "w").write(files[x])
                args.append(path)
            else:
                args.append(x)
        p = subprocess.run([os.path.join(tool_dir, tool)] + args, input=stdin, capture_output=True, text=True, timeout=120)
        return p.returncode, p.stdout


def main():
    out = []
    for tool, argv, files, stdin in CASES:
        rc, stdout = run(OWN, tool, argv, files, stdin)
        assert rc == 0, (tool, argv, rc)
        out.append(dict(tool=tool, argv=argv, files=files, stdin=stdin, rc=rc, stdout=stdout))
    path = os.path.join(ROOT, "tests", "golden", "cli_vectors.json")
    json.dump(dict(generator="tools/gen_cli_golden.py", cases=out), open(path, "w"), indent=0)
    print("wrote %s: %d invocations, %d bytes" % (path, len(out), os.path.getsize(path)))


if __name__ == "__main__":
    main()
