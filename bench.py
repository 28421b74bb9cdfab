#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native seq-align hot path.

Metric (BASELINE.json): DP cell updates per second (GCUPS = sum(len_a*len_b)
/ seconds / 1e9) for batched Smith-Waterman with affine gaps, score-only.
Workload at every N: BASELINE configs[1] per GPU -- 100,000 synthetic DNA
pairs of 150x150, smith_waterman CLI default scoring 2/-2/-2/-1 (weak
scaling: every rank aligns its own 100k-pair shards, generated locally from
the counter-based generator; no collective inside the timed region, scores
are gathered to rank 0 after it for the checksum).

A step = one pass of the hot path over one batch:
  value : inputs resident in HBM (seqalign_batch_run_device_async / _wait:
          alphabet scan + DP kernel, two steps enqueued behind the one being
          completed so the GPU does not idle across the host's launch and
          synchronisation latency; every step's plan is verified against its
          own scan), timed with CUDA events on the launching stream, max over
          ranks;
  e2e   : the same batch through the host-buffer C-ABI call
          (seqalign_batch_submit_packed from pinned host memory: H2D copies,
          kernels, D2H of the scores inside the timed region).
          E2E_DEPTH batches are in flight (one engine and host thread each), so
          the PCIe copy of one step overlaps the kernel of another.
Between timed steps the input rotates over NB distinct batches whose total
size exceeds L2 (126 MB), so no step re-reads a cached batch.

roofline : HBM, as the contract asks (algorithmic bytes of the DP kernel /
           its CUDA-event time / measured copy bandwidth) -- 0.9 %: score-only
           moves 0.0135 B/cell.  roofline.issue is the roof that binds this
           kernel (ALU-pipe issue, from the DPX rate microbenchmarked live);
           roofline.materialise is the HBM roofline of the one mode of the path
           that HBM does bind (all three matrices of every pair, 12 B/cell),
           measured on half of the step's pairs.

--impl reference times the reference's own CPU fill (oracle/_ref/ref_batch,
the unmodified seq-align sources compiled by oracle/Makefile; aligner_align +
best cell) on the same workload with all host threads.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "seq-align_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

PAIRS = 100000
LEN = 150
SEED = 2
E2E_DEPTH = 4   # batches in flight in the end-to-end arm
NCU_TRAFFIC_BYTES = 31644160   # dram read + write of one fast16 launch on this workload (ncu --set full)
MATCH, MISMATCH, GAP_OPEN, GAP_EXTEND = 2, -2, -2, -1
WORKLOAD = "SW score-only, %d synthetic DNA pairs %dx%d per GPU per step, scoring 2/-2/-2/-1" % (PAIRS, LEN, LEN)
REF_BATCH = os.path.join(ROOT, "oracle", "_ref", "ref_batch")
RB_MAGIC = 0x5345514252454631


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def gen_batch(block_index):
    """batch number `block_index` of the global synthetic stream (seed 2)"""
    from helpers import synthetic_batch
    # every batch has its own seed, so any rank can generate any shard locally
    a, oa, b, ob = synthetic_batch(SEED + 1000 * block_index, PAIRS, LEN, LEN)
    return a, oa, b, ob


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                  "--format=csv,noheader,nounits", "-lms", "100"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        while not self.stop_flag:
            line = p.stdout.readline()
            if not line:
                break
            self.rows.append([c.strip() for c in line.split(",")])
        p.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(v for v in sm if v > 0.5 * mx) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def write_ref_input(path, a, oa, b, ob, n):
    hdr = np.zeros(16, dtype=np.int64)
    hdr[0], hdr[1], hdr[2], hdr[3] = RB_MAGIC, n, 1, 0
    hdr[4:8] = (MATCH, MISMATCH, GAP_OPEN, GAP_EXTEND)
    hdr[14] = 1  # scoring_system_default poked in place, as sw_cmdline.c:37-46 does
    with open(path, "wb") as f:
        f.write(hdr.tobytes())
        f.write(oa[:n + 1].astype(np.int64).tobytes())
        f.write(ob[:n + 1].astype(np.int64).tobytes())
        f.write(a[:oa[n]].tobytes())
        f.write(b[:ob[n]].tobytes())


def run_ref_batch(mode, threads, a, oa, b, ob, n):
    """returns (gcups, seconds, scores) of the compiled reference on n pairs"""
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        write_ref_input(fin, a, oa, b, ob, n)
        out = subprocess.check_output([REF_BATCH, mode, str(threads), fin, fout, "nostrings"], text=True)
        info = json.loads(out.strip().splitlines()[-1])
        scores = np.fromfile(fout, dtype=np.int32, count=n)
    return info["cells"] / info["seconds"] / 1e9, info["seconds"], scores


def cpu_port_gcups(a, oa, b, ob, n):
    """fallback when oracle/_ref is absent: the C restatement, one thread"""
    import seqalign
    from helpers import orc_batch_sw, orc_from_scoring
    o = orc_from_scoring(seqalign.Scoring.sw_cli_default())
    t = time.perf_counter()
    s, _, _ = orc_batch_sw(o, a[:oa[n]], oa[:n + 1], b[:ob[n]], ob[:n + 1])
    dt = time.perf_counter() - t
    return n * LEN * LEN / dt / 1e9, dt, s


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    a, oa, b, ob = gen_batch(0)
    sample = 20000  # pairs per step: ~1 s of CPU work on 16 cores
    have_ref = os.path.exists(REF_BATCH)
    vals = []
    for i in range(args.warmup + args.steps):
        if have_ref:
            g, sec, _ = run_ref_batch("fill", cores, a, oa, b, ob, sample)
        else:
            g, sec, _ = cpu_port_gcups(a, oa, b, ob, sample // 20)
        if i >= args.warmup:
            vals.append((g, sec))
    cells = sample * LEN * LEN if have_ref else (sample // 20) * LEN * LEN
    total_s = sum(s for _, s in vals)
    value = cells * len(vals) / total_s / 1e9
    line = {
        "impl": "reference", "metric": "DP cell updates/s (GCUPS)", "value": value, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_s / len(vals), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step_sample_pairs": sample if have_ref else sample // 20},
        "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": cores if have_ref else 1,
                         "kind": "reference" if have_ref else "port",
                         "sample": "aligner_align (fill) + best cell on %d pairs of the workload per step, %d threads"
                                   % (sample if have_ref else sample // 20, cores if have_ref else 1)},
        "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def issue_peak_gcups(kernel_name):
    """ALU-pipe issue ceiling (Gcells/s) of the kernel that ran, from tools/microbench.cu run on
    this GPU just now: (measured lane-ops/clk/SM of the kernel's DPX instruction) x SMs x clock
    / (ALU-pipe instructions per cell).  Also returns the event-timed rate of the whole cell's
    instruction mix as a cross-check."""
    s16 = kernel_name.startswith("fast16")
    want = "s16x2 mix" if s16 else ("int32 end-cell mix" if kernel_name.endswith("_end") else "int32 score-only mix")
    # ALU-pipe instructions per cell: packed kernel 5.5 per cell PAIR (PRMT, 3 x VIADDMNMX.S16x2,
    # VIMNMX3.S16x2, half a VIMNMX3 of the running best); int32 kernels 4.5 / 5.5 per cell
    alu_per_cell = 2.75 if s16 else (5.5 if kernel_name.endswith("_end") else 4.5)
    op = "VIADDMNMX.S16x2.RELU" if s16 else "VIADDMNMX"
    exe = os.path.join(ROOT, "bin", "microbench")
    rate, mix, dev = None, None, None
    try:
        out = subprocess.check_output([exe], text=True, timeout=60)
        for ln in out.splitlines():
            d = json.loads(ln)
            if "sms" in d:
                dev = d
            elif d.get("op") == op:
                rate = d["lane_ops_per_clk_per_sm_at_max_clock"]
            elif want in d.get("op", ""):
                mix = d["cells_gcups"]
    except Exception:
        pass
    if rate is None or dev is None:
        return (6560.0 if s16 else 3900.0), None, "recorded (profiles/microbench_r01b.jsonl)"
    peak = dev["sms"] * rate * dev["clock_khz"] * 1e3 / alu_per_cell / 1e9
    return peak, mix, "live tools/microbench.cu: %s at %.1f lane-ops/clk/SM / %.2f ALU-pipe instr per cell" % (op, rate, alu_per_cell)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import seqalign

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    # stdout carries exactly one JSON line: everything else that libraries print there
    # (NCCL announces its version on stdout) is sent to stderr until the line is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    eng = seqalign.BatchAligner(local, seqalign.Scoring.sw_cli_default())

    # NB distinct batches per rank, > L2 in total (6 x 30 MB = 180 MB)
    NB = 6
    host, devb = [], []
    for k in range(NB):
        a, oa, b, ob = gen_batch(rank * NB + k)
        pa, pb = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
        poa, pob = torch.from_numpy(oa).pin_memory(), torch.from_numpy(ob).pin_memory()
        host.append((pa, poa, pb, pob))
        devb.append(tuple(t.to(dev) for t in (pa, poa, pb, pob)))
    DEV_DEPTH = 2   # device-resident arm: runs enqueued ahead of their verification (run_device_async)
    d_scores = [torch.zeros(PAIRS, dtype=torch.int32, device=dev) for _ in range(DEV_DEPTH + 1)]
    d_score = d_scores[0]
    cells_step = PAIRS * LEN * LEN
    # a dedicated (non-default) stream: the engine launches on it and the timing events are recorded on it
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    def step_device(i):
        """enqueue step i (its scores go to d_scores[i % (DEV_DEPTH+1)]); the engine launches the DP kernel
        with the previous step's plan and verifies it against this batch's scan in run_device_wait()"""
        a, oa, b, ob = devb[i % NB]
        # score-only: no end-cell buffers, so the engine may pick its packed 16-bit kernel
        eng.run_device_async(seqalign.SW, a.data_ptr(), oa.data_ptr(), b.data_ptr(), ob.data_ptr(), PAIRS,
                             d_scores[i % (DEV_DEPTH + 1)].data_ptr(), 0, 0, stream)

    # end-to-end arm: seqalign.PipelinedAligner keeps E2E_DEPTH batches in flight (one engine and
    # host thread each), so the PCIe copy of one step overlaps the kernel of another
    pipe = seqalign.PipelinedAligner(local, seqalign.Scoring.sw_cli_default(), depth=E2E_DEPTH)

    def submit_host(i):
        a, oa, b, ob = host[i % NB]
        return pipe.submit_ptrs(seqalign.SW, seqalign.MODE_SCORE_ONLY, a.data_ptr(), oa.data_ptr(), b.data_ptr(),
                                ob.data_ptr(), PAIRS)

    def run_host_steps(first, count):
        """count steps through the pipeline; returns the score arrays' checksum"""
        import collections
        pending, total = collections.deque(), 0
        for i in range(count):
            pending.append(submit_host(first + i))
            if len(pending) > E2E_DEPTH:
                total += int(pending.popleft().result().sum())
        while pending:
            total += int(pending.popleft().result().sum())
        return total

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm -------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
        eng.run_device_wait()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, outstanding = [], 0
    e0.record()
    for i in range(args.steps):
        step_device(args.warmup + i)
        outstanding += 1
        if outstanding > DEV_DEPTH:        # keep DEV_DEPTH steps enqueued behind the one being completed
            eng.run_device_wait()
            kernel_ms.append(eng.last_kernel_ms)
            outstanding -= 1
    while outstanding:
        eng.run_device_wait()
        kernel_ms.append(eng.last_kernel_ms)
        outstanding -= 1
    e1.record()
    barrier()
    dt_ms = e0.elapsed_time(e1)
    kernel_name = eng.last_kernel
    launches = 2 * args.steps              # per step: the DP kernel and the alphabet / shape scan next to it
    d_score = d_scores[(args.warmup + args.steps - 1) % (DEV_DEPTH + 1)]
    checksum = int(d_score.sum().item())

    # ---- end-to-end arm (host buffers through the C-ABI) ----------------------
    # warm-up: enough submissions that every engine of the pipeline has made its first call
    # (device / pinned allocations happen there), at least the W the caller asked for
    run_host_steps(0, max(args.warmup, 2 * E2E_DEPTH))
    barrier()
    t0 = time.perf_counter()
    e2e_checksum = run_host_steps(args.warmup, args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    sampler.stop_flag = True
    time.sleep(0.15)

    t = torch.tensor([dt_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(cells_step * args.steps), float(checksum)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        # gather the last step's scores to rank 0 (what a caller of a sharded job gets back)
        gathered = [torch.zeros_like(d_score) for _ in range(world)] if rank == 0 else None
        dist.gather(d_score, gathered, dst=0)
    dt_ms, e2e_ms = t.tolist()
    total_cells = tot[0].item()

    if rank == 0:
        value = total_cells / (dt_ms * 1e-3) / 1e9
        e2e_val = total_cells / (e2e_ms * 1e-3) / 1e9
        hbm_peak, peak_src = peaks()
        clocks = sampler.summary()
        k_ms = float(np.mean(kernel_ms))
        # algorithmic traffic, score mode: both sequences read once, score + end cell written
        alg_bytes = PAIRS * (LEN + LEN + 4)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        issue_peak, issue_mix, issue_src = issue_peak_gcups(kernel_name)
        f_mhz = clocks["sm_mhz"] or 1965.0
        line = {
            "metric": "DP cell updates/s (GCUPS)", "value": value, "unit": "GCUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": PAIRS, "cells_per_step": cells_step * world,
                       "l2": "input rotates over %d distinct batches per GPU (%.0f MB > 126 MB L2)" % (NB, NB * 2 * PAIRS * LEN / 1e6),
                       "parallelism": "pairs sharded by rank, no collective in the timed region",
                       "kernel": kernel_name, "score_checksum": int(tot[1].item())},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "GCUPS", "h2d_bytes_per_step": int(2 * PAIRS * LEN),   # the sequences; the offset arrays of a
                    # uniform batch are not shipped, the engine makes them on the device
                    "d2h_bytes_per_step": int(4 * PAIRS), "ms_per_step": e2e_ms / args.steps,
                    "api": "seqalign.PipelinedAligner(depth=%d).submit_ptrs -> seqalign_batch_submit_packed; pinned host "
                           "buffers in, int32 scores out on the host, every step" % E2E_DEPTH,
                    "score_checksum_rank0": e2e_checksum},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel on this
                         # workload, from profiles/ncu_fast16_r01f_raw.csv (ncu --set full); not re-measured live
                         "traffic": NCU_TRAFFIC_BYTES if kernel_name == "fast16_sw_score" else None,
                         "traffic_source": "profiles/ncu_fast16_r01f_raw.csv",
                         "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                         "kernel": kernel_name, "kernel_ms": k_ms,
                         "kernel_gcups": cells_step / (k_ms * 1e-3) / 1e9,
                         "note": "score-only moves 0.0139 B/cell, so HBM is not the binding roof; the "
                                 "binding one is INT32/DPX issue, reported under 'issue'",
                         "issue": {"achieved_gcups": cells_step / (k_ms * 1e-3) / 1e9, "peak_gcups": issue_peak,
                                   "frac": cells_step / (k_ms * 1e-3) / 1e9 / issue_peak,
                                   "mix_gcups": issue_mix,   # the whole cell's instruction mix, microbenchmarked
                                   "sm_mhz": f_mhz, "source": issue_src}},
        }
        # The one mode of the path that HBM binds (SURVEY 8d mode M, the literal aligner_align contract):
        # all three int32 matrices of every pair written out, 12 B/cell.  Measured here next to the
        # headline so that the HBM roofline the contract asks for has a kernel it applies to.
        try:
            n_m = 50000
            a, oa, b, ob = [t.numpy() for t in host[0]]
            ms_m = []
            for _ in range(3):
                eng.submit_packed(seqalign.SW, seqalign.MODE_MATS, a[: n_m * LEN], oa[: n_m + 1], b[: n_m * LEN], ob[: n_m + 1])
                ms_m.append(eng.last_kernel_ms)
            bytes_m = 12 * n_m * (LEN + 1) * (LEN + 1)
            gbs_m = bytes_m / (min(ms_m) * 1e-3) / 1e9
            line["roofline"]["materialise"] = {
                "workload": "SW, %d of the step's pairs, match/gap_a/gap_b matrices of every pair (SEQALIGN_MODE_MATS)" % n_m,
                "kernel": eng.last_kernel, "kernel_ms": min(ms_m), "bound": "hbm", "algorithmic_bytes": bytes_m,
                "achieved": gbs_m, "peak": hbm_peak, "unit": "GB/s", "frac": gbs_m / hbm_peak,
                "gcups": n_m * LEN * LEN / (min(ms_m) * 1e-3) / 1e9,
                "traffic": 5419337000 * n_m // 20000, "traffic_source": "profiles/ncu_mats_r01f_raw.csv (20k pairs: 5.42 GB written, 0.04 GB read), scaled"}
        except Exception as e:   # never lose the headline line over the side measurement
            line["roofline"]["materialise"] = {"error": str(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            a, oa, b, ob = [t.numpy() for t in host[0]]
            cores = os.cpu_count() or 1
            if os.path.exists(REF_BATCH):
                g, sec, ref_scores = run_ref_batch("fill", cores, a, oa, b, ob, PAIRS)
                gf, secf, _ = run_ref_batch("full", cores, a, oa, b, ob, 10000)
                eng.submit_ptrs(seqalign.SW, seqalign.MODE_SCORE_ONLY, *[t.data_ptr() for t in host[0]], PAIRS)
                line["cpu_baseline"] = {
                    "value": g, "unit": "GCUPS", "cores": cores, "kind": "reference",
                    "sample": "unmodified reference aligner_align (fill) + best cell on all %d pairs of one step, "
                              "%d threads, %.1f s" % (PAIRS, cores, sec),
                    "full_call_value": gf,
                    "full_call_sample": "smith_waterman_align2 + first fetch (fresh aligner) on 10000 pairs, %.1f s" % secf,
                    "scores_match_gpu": bool(np.array_equal(ref_scores, eng.scores()))}
            else:
                g, sec, s = cpu_port_gcups(a, oa, b, ob, 2000)
                line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": 1, "kind": "port",
                                        "sample": "oracle C restatement on 2000 pairs of one step, 1 thread, %.1f s" % sec}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
