#!/usr/bin/env python3
"""Summarise `ncu --page source --csv --print-source sass` output (optionally .gz):
per kernel section, total stall-reason shares and the hottest instructions.
    python tools/ncu_source_summary.py file.csv[.gz] [section-index] [top-n]"""
import csv, gzip, sys
path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else None
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
sections = []
for row in csv.reader(f):
    if row and row[0] == "Kernel Name":
        sections.append({"name": row[1], "hdr": None, "rows": []})
    elif sections and sections[-1]["hdr"] is None:
        sections[-1]["hdr"] = row
    elif sections:
        sections[-1]["rows"].append(row)
for si, s in enumerate(sections):
    if which is None:
        tot = sum(int(r[s["hdr"].index("# Samples")] or 0) for r in s["rows"] if len(r) > 5)
        print(si, len(s["rows"]), tot, s["name"][:110])
        continue
    if si != which:
        continue
    h = s["hdr"]; ix = {k: i for i, k in enumerate(h)}
    rows = [r for r in s["rows"] if len(r) == len(h)]
    samp = lambda r: int(r[ix["# Samples"]] or 0)
    tot = sum(samp(r) for r in rows) or 1
    print(s["name"], "instructions", len(rows), "samples", tot)
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(r[ix[k]] or 0) for r in rows) for k in stalls}
    print("stall shares:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    ex = sum(int(r[ix["Instructions Executed"]] or 0) for r in rows)
    print("warp instructions executed:", ex)
    op = {}
    for r in rows:
        m = r[ix["Source"]].split()
        if not m: continue
        k = m[1] if m[0].startswith("@") else m[0]
        k = k.split(".")[0] if not k.startswith(("VIADDMNMX", "VIMNMX", "LDS", "STS", "STG", "LDG")) else k
        op[k] = op.get(k, 0) + int(r[ix["Instructions Executed"]] or 0)
    print("opcode mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / ex) for k, v in sorted(op.items(), key=lambda kv: -kv[1])[:18]))
    print("hottest instructions:")
    for r in sorted(rows, key=samp, reverse=True)[:topn]:
        top = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:3]
        print("  %6d %5.1f%%  %-60s %s" % (samp(r), 100.0 * samp(r) / tot, r[ix["Source"]].strip()[:60], " ".join("%s:%d" % (k, v) for v, k in top if v)))
