#!/usr/bin/env python3
"""A large FASTA file through the batching tools on the GPU box: N pairs of 150 bp,
`smith_waterman --maxhits 1` and `needleman_wunsch --printscores`, read by the device decoder and by the host
reader, on 1 GPU and on every GPU of the box; stdout must be byte-identical across all of them (md5), the
phases the tools report (SEQALIGN_CLI_TIMING) and the wall clock go to the record.

    python tools/gpu_cli_big.py [pairs] [gpus]    >> profiles/cli_big_r02.jsonl"""
import hashlib, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import *
from seqalign.synth import synth_batch

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ONLY = sys.argv[3] if len(sys.argv) > 3 else ""        # "sw" / "nw": one tool only
L = 150
td = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
path = os.path.join(td, "big.fa")
t0 = time.time()
with open(path, "wb") as f:
    step = 200000
    for p0 in range(0, N, step):
        m = min(step, N - p0)
        a, _, b, _ = synth_batch(5, p0, m, L, L)
        a = a.reshape(m, L); b = b.reshape(m, L)
        na = np.frombuffer(b"".join(b">a%09d\n" % (p0 + i) for i in range(m)), np.uint8).reshape(m, 12)
        nb = np.frombuffer(b"".join(b">b%09d\n" % (p0 + i) for i in range(m)), np.uint8).reshape(m, 12)
        nl = np.full((m, 1), 10, np.uint8)
        f.write(np.concatenate([na, a, nl, nb, b, nl], axis=1).tobytes())
size = os.path.getsize(path)
print(json.dumps(dict(what="input", pairs=N, bytes=size, write_s=round(time.time() - t0, 2))), flush=True)


def run(tool, args, env):
    exe = os.path.join(ROOT, "bin", tool)
    t = time.time()
    p = subprocess.Popen([exe] + args + ["--file", path], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         env=dict(os.environ, SEQALIGN_CLI_TIMING="1", **env))
    h, nbytes = hashlib.md5(), 0
    while True:
        blk = p.stdout.read(1 << 24)
        if not blk:
            break
        h.update(blk); nbytes += len(blk)
    err = p.stderr.read().decode()
    p.wait()
    dt = time.time() - t
    assert p.returncode == 0, err[-500:]
    return dt, h.hexdigest(), nbytes, err.strip().split("\n")[-1]


for tool, args in (("smith_waterman", ["--maxhits", "1"]), ("needleman_wunsch", ["--printscores"])):
    if ONLY and not tool.startswith({"sw": "smith", "nw": "needle"}[ONLY]):
        continue
    ref_md5 = None
    variants = [("device decoder, 1 GPU", [], {}), ("host reader, 1 GPU", [], {"SEQALIGN_CLI_DECODE": "host"})]
    if G > 1:
        variants += [("device decoder, %d GPUs" % G, ["--gpus", str(G)], {}), ("host reader, %d GPUs" % G, ["--gpus", str(G)], {"SEQALIGN_CLI_DECODE": "host"})]
    # CLI_BIG_BATCHES=16384,65536,262144: the same run with other numbers of pairs per engine submit
    for bsz in [x for x in os.environ.get("CLI_BIG_BATCHES", "").split(",") if x]:
        variants.append(("device decoder, 1 GPU, %s pairs per submit" % bsz, [], {"SEQALIGN_CLI_BATCH_PAIRS": bsz}))
    run(tool, args, {})   # page cache, driver
    for name, extra, env in variants:
        dt, md5, nbytes, phases = run(tool, extra + args, env)
        ref_md5 = ref_md5 or md5
        print(json.dumps(dict(tool=tool, args=extra + args, reader=name, pairs=N, input_mb=round(size / 1e6, 1), stdout_mb=round(nbytes / 1e6, 1),
                              seconds=round(dt, 3), pairs_per_s=round(N / dt), gcups_whole_process=round(N * L * L / dt / 1e9, 1),
                              phases=phases, same_stdout_as_first_variant=md5 == ref_md5, md5=md5)), flush=True)
os.unlink(path)
