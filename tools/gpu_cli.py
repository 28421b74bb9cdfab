#!/usr/bin/env python3
"""Throughput of the batching command-line tools next to the reference's own
tools (tests/integration/_ref_own, unmodified sources + the reference's DP, one pair per call)
on the same FASTA input.  Wall clock of the whole process, GPU context creation included."""
import json, os, subprocess, sys, tempfile, time, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import synthetic_batch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
NREF = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
a, oa, b, ob = synthetic_batch(2, N, 150, 150)
td = tempfile.mkdtemp()
def write(path, n):
    with open(path, "w") as f:
        for i in range(n):
            f.write(">a%d\n%s\n>b%d\n%s\n" % (i, a[i * 150:(i + 1) * 150].tobytes().decode(), i, b[i * 150:(i + 1) * 150].tobytes().decode()))
big, small = os.path.join(td, "big.fa"), os.path.join(td, "small.fa")
write(big, N); write(small, NREF)
def run(exe, args, path):
    t = time.time()
    p = subprocess.run([exe] + args + ["--file", path], capture_output=True, env=dict(os.environ, SEQALIGN_CLI_TIMING="1"))
    dt = time.time() - t
    assert p.returncode == 0, p.stderr[:500]
    global last_timing
    last_timing = p.stderr.decode().strip().split("\n")[-1] if p.stderr else ""
    return dt, p.stdout
rows = []
for tool, args in (("needleman_wunsch", ["--printscores"]), ("smith_waterman", ["--maxhits", "1"]), ("smith_waterman", [])):
    ours, ref = os.path.join(ROOT, "bin", tool), os.path.join(ROOT, "tests", "integration", "_ref_own", tool)
    run(ours, args, small)                       # warm the driver / page cache
    t_small, out_small = run(ours, args, small)
    t_big, out_big = run(ours, args, big)
    row = dict(tool=tool, args=args, pairs=N, seconds=round(t_big, 3), pairs_per_s=round(N / t_big), gcups=round(N * 22500 / t_big / 1e9, 1),
               small_pairs=NREF, small_seconds=round(t_small, 3), stdout_mb=round(len(out_big) / 1e6, 1),
               phases=last_timing)
    try:
        ph = dict((k, float(v)) for k, v in (x.strip().split(" ")[:2] for x in last_timing.replace("timing: ", "").replace(" s", "").split(",")))
        work = ph["read"] + ph["align"] + ph["print"]
        row.update(work_seconds=round(work, 3), work_pairs_per_s=round(N / work), work_gcups=round(N * 22500 / work / 1e9, 1), init_seconds=ph["init"])
    except Exception as e:
        row["phase_parse_error"] = str(e)
    if os.path.exists(ref):
        t_ref, out_ref = run(ref, args, small)
        # NW output must be identical; SW beyond pair 0 differs by the reference's stale-mask defect
        row.update(ref_seconds=round(t_ref, 3), ref_pairs_per_s=round(NREF / t_ref), speedup_whole_process=round(row["pairs_per_s"] / (NREF / t_ref), 1),
                   speedup_work_only=round(row.get("work_pairs_per_s", 0) / (NREF / t_ref), 1),
                   same_stdout_as_reference=out_small == out_ref, first_pair_same=out_small.split(b"\n\n")[0] == out_ref.split(b"\n\n")[0])
    print(json.dumps(row), flush=True)
    rows.append(row)
# BASELINE config 4 through the tool: 50k protein pairs 400x400, BLOSUM62, first hit only
NP, NPREF, LP = 50000, 300, 400
pa, _, pb, _ = synthetic_batch(4, NP, LP, LP, kind="protein")
def writep(path, n):
    with open(path, "w") as f:
        for i in range(n):
            f.write(">a%d\n%s\n>b%d\n%s\n" % (i, pa[i * LP:(i + 1) * LP].tobytes().decode(), i, pb[i * LP:(i + 1) * LP].tobytes().decode()))
bigp, smallp = os.path.join(td, "bigp.fa"), os.path.join(td, "smallp.fa")
writep(bigp, NP); writep(smallp, NPREF)
args = ["--scoring", "BLOSUM62", "--maxhits", "1", "--minscore", "1"]
ours, ref = os.path.join(ROOT, "bin", "smith_waterman"), os.path.join(ROOT, "tests", "integration", "_ref_own", "smith_waterman")
run(ours, args, smallp)
t_big, out_big = run(ours, args, bigp)
row = dict(tool="smith_waterman", args=args, config="BASELINE config 4", pairs=NP, seconds=round(t_big, 3), pairs_per_s=round(NP / t_big),
           gcups=round(NP * LP * LP / t_big / 1e9, 1), stdout_mb=round(len(out_big) / 1e6, 1), phases=last_timing)
if os.path.exists(ref):
    t_ref, out_ref = run(ref, args, smallp)
    t_s, out_s = run(ours, args, smallp)
    row.update(ref_pairs=NPREF, ref_seconds=round(t_ref, 3), ref_pairs_per_s=round(NPREF / t_ref, 1), ref_gcups=round(NPREF * LP * LP / t_ref / 1e9, 3),
               speedup_whole_process=round((NP / t_big) / (NPREF / t_ref), 1), first_pair_same=out_s.split(b"==\n")[0] == out_ref.split(b"==\n")[0])
print(json.dumps(row), flush=True)
rows.append(row)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "gpu_cli.json"), "w"), indent=1)
