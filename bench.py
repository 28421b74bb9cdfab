#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native seq-align hot path.

Metric (BASELINE.json): DP cell updates per second (GCUPS = sum(len_a*len_b)
/ seconds / 1e9) for batched Smith-Waterman with affine gaps, score-only,
smith_waterman CLI default scoring 2/-2/-2/-1.  Inputs come from the
counter-based generator of SURVEY.md 8d (seqalign.synth / csrc/sa_synth.cuh).

N = 1 (BASELINE configs[1]): 100,000 synthetic DNA pairs of 150x150 per step.
  value : inputs resident in HBM (seqalign_batch_run_device_async / _wait:
          alphabet scan + DP kernel, two steps enqueued behind the one being
          completed), CUDA events on the launching stream.
  e2e   : the same batch through the host-buffer C-ABI call
          (seqalign_batch_submit_packed from pinned host memory: H2D copies,
          kernels, D2H of the scores inside the timed region), E2E_DEPTH
          batches in flight.
  The input rotates over NB distinct batches whose total size exceeds L2.
  Extras on the line: `sustained` (the same step looped for >= 2 s with clocks
  and utilisation sampled during it), `value_int32_kernel` (the pure-int32
  kernel on the same step), `config5` (the N>1 workload on this one GPU, so
  the strong-scaling curve has its own N=1 point).

N > 1 (BASELINE configs[4]): ONE job of 10,000,000 pairs of 150x150, strong
scaling: rank r owns pairs [r*P/N, (r+1)*P/N); a step aligns the whole job
and brings every score to rank 0 over NCCL (gather) inside the timed region.
  value       : every rank's shard resident in its HBM.
  e2e         : every rank's shard in ITS OWN pinned host memory, pulled over
                its own PCIe link by seqalign_batch_submit_packed; scores
                gathered to rank 0's host.
  e2e_scatter : the whole job in rank 0's host memory only; chunks go host ->
                GPU 0 -> NCCL isend over NVLink -> run on arrival
                (seqalign.distributed.align_sharded_stream).  Bound by rank
                0's PCIe link by construction.
  The score checksum of the whole job does not depend on N; rank 0's first
  pairs are checked against the compiled reference.

roofline : HBM, as the contract asks (algorithmic bytes of the DP kernel /
           its CUDA-event time / measured copy bandwidth).  Score-only moves
           0.0135 B/cell, so roofline.issue (ALU-pipe issue, from the DPX rate
           microbenchmarked live) is the roof that binds this kernel;
           roofline.materialise is the HBM roofline of the one mode of the path
           that HBM does bind (all three matrices of every pair, 12 B/cell).

--impl reference times the reference's own CPU fill (oracle/_ref/ref_batch,
the unmodified seq-align sources compiled by oracle/Makefile; aligner_align +
best cell) on a bounded sample of the same workload with all host threads.
"""
import argparse
import collections
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "seq-align_b200"))

import numpy as np  # noqa: E402

PAIRS = 100000
LEN = 150
SEED = 2
C5_PAIRS = 10_000_000
C5_SEED = 5
E2E_DEPTH = 4   # batches in flight in the N=1 end-to-end arm
MATCH, MISMATCH, GAP_OPEN, GAP_EXTEND = 2, -2, -2, -1
DTYPE = "int32 results; int16x2 DPX arithmetic where every score provably fits (bit-exact, range checked per batch)"
WORKLOAD_1 = "SW score-only, %d synthetic DNA pairs %dx%d per step, scoring 2/-2/-2/-1" % (PAIRS, LEN, LEN)
WORKLOAD_N = ("SW score-only, ONE job of %d synthetic DNA pairs %dx%d sharded over the GPUs, scoring 2/-2/-2/-1 "
              "(BASELINE config 5)" % (C5_PAIRS, LEN, LEN))
REF_BATCH = os.path.join(ROOT, "oracle", "_ref", "ref_batch")
RB_MAGIC = 0x5345514252454631


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_name):
    """dram bytes (read + write) per launch of `kernel_name` on the N=1 workload, from the committed
    ncu --set full capture (profiles/ncu_traffic.json, taken from the ncu capture it names); None if that
    kernel was not captured"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = t.get(kernel_name)
        return (e["bytes"], e["source"]) if e else (None, None)
    except Exception:
        return None, None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks, utilisation and throttle reasons, every 100 ms while the bench runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,utilization.gpu,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
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
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))
        p.kill()

    def summary(self, t0=None, t1=None):
        """over the samples taken in [t0, t1] (all if None); `sm_mhz` is the median over the samples
        that saw the GPU busy (utilisation >= 50 %), which is what 'under load' means here"""
        sm, busy, mx, reasons, util = [], [], 0, set(), []
        for t, r in self.rows:
            if (t0 is not None and t < t0) or (t1 is not None and t > t1):
                continue
            try:
                f, m, u = float(r[0]), float(r[1]), float(r[2])
            except Exception:
                continue
            sm.append(f); mx = max(mx, m); util.append(u)
            if u >= 50:
                busy.append(f)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        pick = sorted(busy) or sorted(sm)
        return {"sm_mhz": pick[len(pick) // 2] if pick else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(busy),
                "util_max": max(util) if util else None}


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
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import seqalign
    from helpers import orc_batch_sw, orc_from_scoring
    o = orc_from_scoring(seqalign.Scoring.sw_cli_default())
    t = time.perf_counter()
    s, _, _ = orc_batch_sw(o, a[:oa[n]], oa[:n + 1], b[:ob[n]], ob[:n + 1])
    dt = time.perf_counter() - t
    return n * LEN * LEN / dt / 1e9, dt, s


def host_batch(seed, first_pair, n):
    from seqalign.synth import synth_batch
    return synth_batch(seed, first_pair, n, LEN, LEN, "dna")


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    multi = args.gpus > 1
    sample = 20000  # pairs per step: ~1 s of CPU work on 16 cores
    have_ref = os.path.exists(REF_BATCH)
    if not have_ref:
        sample //= 20
    # the first pairs of the arm's workload (N=1: step 0 of config 2; N>1: the head of the config-5 job)
    a, oa, b, ob = host_batch(C5_SEED if multi else SEED, 0, sample)
    vals = []
    for i in range(args.warmup + args.steps):
        if have_ref:
            g, sec, _ = run_ref_batch("fill", cores, a, oa, b, ob, sample)
        else:
            g, sec, _ = cpu_port_gcups(a, oa, b, ob, sample)
        if i >= args.warmup:
            vals.append((g, sec))
    cells = sample * LEN * LEN
    total_s = sum(s for _, s in vals)
    value = cells * len(vals) / total_s / 1e9
    line = {
        "impl": "reference", "metric": "DP cell updates/s (GCUPS)", "value": value, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_s / len(vals), "higher_is_better": True, "scaling": "strong" if multi else "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD_N if multi else WORKLOAD_1, "step_sample_pairs": sample},
        "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": cores if have_ref else 1,
                         "kind": "reference" if have_ref else "port",
                         "sample": "aligner_align (fill) + best cell on the first %d pairs of the workload per step, %d threads"
                                   % (sample, cores if have_ref else 1)},
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
    return peak, mix, ("live tools/microbench.cu: %s at %.1f lane-ops/clk/SM / %.2f ALU-pipe instr per cell "
                       "(instruction count from cuobjdump -sass of the kernel's row loop, DESIGN.md 3 K1)" % (op, rate, alu_per_cell))


class Ctx:
    """everything both workloads share: device, stream, engine, generator, timing helpers"""

    def __init__(self):
        import torch
        import seqalign
        self.torch, self.sa = torch, seqalign
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            # a rank's host threads stay on its own cores: 8 ranks x (pipeline threads + main) otherwise
            # migrate over the whole socket while they feed PCIe
            try:
                cores = sorted(os.sched_getaffinity(0))
                per = max(1, len(cores) // self.world)
                os.sched_setaffinity(0, cores[self.local * per:(self.local + 1) * per] or cores)
            except Exception:
                pass
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.scoring = seqalign.Scoring.sw_cli_default()
        self.eng = seqalign.BatchAligner(self.local, self.scoring)
        # a dedicated (non-default) stream: the engine launches on it and the timing events are recorded on it
        self.tstream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.tstream)
        self.stream = self.tstream.cuda_stream

    def device_batch(self, seed, first_pair, n):
        """pairs [first_pair, first_pair+n) of stream `seed`, generated on the device"""
        torch = self.torch
        a = torch.empty(n * LEN + 32, dtype=torch.uint8, device=self.dev)[: n * LEN]
        b = torch.empty(n * LEN + 32, dtype=torch.uint8, device=self.dev)[: n * LEN]
        self.sa.synth_device(self.local, "dna", seed, first_pair, n, LEN, LEN, a.data_ptr(), b.data_ptr(), self.stream)
        off = torch.arange(n + 1, dtype=torch.int64, device=self.dev) * LEN
        return a, off, b, off.clone()

    def pinned(self, t):
        out = self.torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        out.copy_(t)
        return out

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op="max"):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return t.tolist()


def device_steps(cx, batches, d_scores, nsteps, first=0, depth=2, after=None):
    """nsteps device-resident steps through run_device_async with `depth` steps enqueued behind the one
    being completed; step i aligns batches[i % len] into d_scores[i % len(d_scores)].  after(i): called
    when step i has been verified (its kernel may still be followed by later ones on the stream).
    Returns the per-step kernel times (ms, CUDA events inside the engine)."""
    eng, sa = cx.eng, cx.sa
    kernel_ms, q = [], collections.deque()
    for i in range(first, first + nsteps):
        a, oa, b, ob = batches[i % len(batches)]
        eng.run_device_async(sa.SW, a.data_ptr(), oa.data_ptr(), b.data_ptr(), ob.data_ptr(), oa.numel() - 1,
                             d_scores[i % len(d_scores)].data_ptr(), 0, 0, cx.stream)
        q.append(i)
        if len(q) > depth:
            eng.run_device_wait()
            kernel_ms.append(eng.last_kernel_ms)
            j = q.popleft()
            if after:
                after(j)
    while q:
        eng.run_device_wait()
        kernel_ms.append(eng.last_kernel_ms)
        j = q.popleft()
        if after:
            after(j)
    return kernel_ms


def roofline_block(cx, kernel_name, k_ms, pairs_per_launch, clocks):
    hbm_peak, peak_src = peaks()
    # algorithmic traffic, score mode: both sequences read once, one int32 score written
    alg_bytes = pairs_per_launch * (LEN + LEN + 4)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    cells = pairs_per_launch * LEN * LEN
    issue_peak, issue_mix, issue_src = issue_peak_gcups(kernel_name)
    traffic, traffic_src = ncu_traffic(kernel_name) if pairs_per_launch == PAIRS else (None, None)
    return {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes": alg_bytes, "peak_source": peak_src, "kernel": kernel_name, "kernel_ms": k_ms,
            "kernel_gcups": cells / (k_ms * 1e-3) / 1e9,
            "note": "score-only moves 0.0135 B/cell, so HBM is not the binding roof; the binding one is "
                    "INT32/DPX issue, reported under 'issue'",
            "issue": {"achieved_gcups": cells / (k_ms * 1e-3) / 1e9, "peak_gcups": issue_peak,
                      "frac": cells / (k_ms * 1e-3) / 1e9 / issue_peak, "mix_gcups": issue_mix,
                      "sm_mhz": clocks.get("sm_mhz") or 1965.0, "source": issue_src}}


# ---------------------------------------------------------------------------------------------
# N = 1: BASELINE config 2

def run_single(cx, args):
    torch, sa, eng, dev = cx.torch, cx.sa, cx.eng, cx.dev
    NB = 6   # distinct batches, > L2 in total (6 x 30 MB = 180 MB)
    devb = [cx.device_batch(SEED, k * PAIRS, PAIRS) for k in range(NB)]
    torch.cuda.synchronize()
    host = [tuple(cx.pinned(t) for t in bt) for bt in devb]
    DEV_DEPTH = 2
    d_scores = [torch.zeros(PAIRS, dtype=torch.int32, device=dev) for _ in range(DEV_DEPTH + 1)]
    cells_step = PAIRS * LEN * LEN

    pipe = sa.PipelinedAligner(cx.local, cx.scoring, depth=E2E_DEPTH)

    def run_host_steps(first, count):
        pending, total = collections.deque(), 0
        for i in range(count):
            a, oa, b, ob = host[(first + i) % NB]
            pending.append(pipe.submit_ptrs(sa.SW, sa.MODE_SCORE_ONLY, a.data_ptr(), oa.data_ptr(), b.data_ptr(),
                                            ob.data_ptr(), PAIRS))
            if len(pending) > E2E_DEPTH:
                total += int(pending.popleft().result().sum())
        while pending:
            total += int(pending.popleft().result().sum())
        return total

    # ---- device-resident arm: warm-up, a sustained loop (clocks sampled under load), then the timed steps
    device_steps(cx, devb, d_scores, args.warmup, depth=0)
    sampler = ClockSampler(cx.local)
    sampler.start()
    time.sleep(0.3)
    cx.barrier()
    ts0 = time.perf_counter()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    sus_steps = 0
    while time.perf_counter() - ts0 < args.sustain:
        device_steps(cx, devb, d_scores, 200, first=sus_steps, depth=DEV_DEPTH)
        sus_steps += 200
    s1.record()
    torch.cuda.synchronize()
    ts1 = time.perf_counter()
    sus_ms = s0.elapsed_time(s1)
    # the timed region follows the sustained loop without a pause: same clocks, same thermal state
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tt0 = time.perf_counter()
    e0.record()
    kernel_ms = device_steps(cx, devb, d_scores, args.steps, first=args.warmup, depth=DEV_DEPTH)
    e1.record()
    cx.barrier()
    tt1 = time.perf_counter()
    dt_ms = e0.elapsed_time(e1)
    kernel_name = eng.last_kernel
    checksum = int(d_scores[(args.warmup + args.steps - 1) % (DEV_DEPTH + 1)].sum().item())

    # ---- end-to-end arm (host buffers through the C-ABI)
    run_host_steps(0, max(args.warmup, 2 * E2E_DEPTH))
    cx.barrier()
    t0 = time.perf_counter()
    e2e_checksum = run_host_steps(args.warmup, args.steps)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag = True

    clocks = sampler.summary(ts0, tt1 + 0.05)
    clocks["covers"] = "the %.1f s sustained loop and the timed steps that follow it without a pause" % (ts1 - ts0)
    k_ms = float(np.mean(kernel_ms))
    value = cells_step * args.steps / (dt_ms * 1e-3) / 1e9
    line = {
        "metric": "DP cell updates/s (GCUPS)", "value": value, "unit": "GCUPS", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": WORKLOAD_1, "pairs_per_step": PAIRS, "cells_per_step": cells_step,
                   "generator": "splitmix64 counter generator of SURVEY.md 8d (seqalign.synth, csrc/sa_synth.cuh), seed %d, "
                                "step k = pairs [k*%d, (k+1)*%d)" % (SEED, PAIRS, PAIRS),
                   "l2": "input rotates over %d distinct batches (%.0f MB > 126 MB L2)" % (NB, NB * 2 * PAIRS * LEN / 1e6),
                   "kernel": kernel_name, "score_checksum": checksum},
        "clocks": clocks,
        "sustained": {"seconds": sus_ms * 1e-3, "steps": sus_steps, "value": cells_step * sus_steps / (sus_ms * 1e-3) / 1e9,
                      "unit": "GCUPS", "clocks": sampler.summary(ts0, ts1)},
        "e2e": {"value": cells_step * args.steps / (e2e_ms * 1e-3) / 1e9, "unit": "GCUPS",
                "h2d_bytes_per_step": int(2 * PAIRS * LEN),   # the offset arrays of a uniform batch are made on the device
                "d2h_bytes_per_step": int(4 * PAIRS), "ms_per_step": e2e_ms / args.steps,
                "api": "seqalign.PipelinedAligner(depth=%d).submit_ptrs -> seqalign_batch_submit_packed; pinned host "
                       "buffers in, int32 scores out on the host, every step" % E2E_DEPTH,
                "score_checksum_last": e2e_checksum},
        "gpu_launches": 2 * args.steps,   # per step: the DP kernel and the alphabet / shape scan next to it
        "roofline": roofline_block(cx, kernel_name, k_ms, PAIRS, clocks),
    }

    # the same step on the pure int32 kernel (what runs when a batch's scores do not provably fit 16 bits)
    try:
        eng.force_general(5)
        ms32 = []
        for i in range(4):
            a, oa, b, ob = devb[i % NB]
            eng.run_device(sa.SW, a.data_ptr(), oa.data_ptr(), b.data_ptr(), ob.data_ptr(), PAIRS, d_scores[0].data_ptr(), 0, 0, cx.stream)
            ms32.append(eng.last_kernel_ms)
        line["value_int32_kernel"] = {"value": cells_step / (min(ms32[1:]) * 1e-3) / 1e9, "unit": "GCUPS", "kernel": eng.last_kernel,
                                      "kernel_ms": min(ms32[1:])}
        a, oa, b, ob = devb[3 % NB]
        eng.force_general(0)
        chk = torch.zeros(PAIRS, dtype=torch.int32, device=dev)
        eng.run_device(sa.SW, a.data_ptr(), oa.data_ptr(), b.data_ptr(), ob.data_ptr(), PAIRS, chk.data_ptr(), 0, 0, cx.stream)
        line["value_int32_kernel"]["scores_equal"] = bool(torch.equal(chk, d_scores[0]))
    except Exception as e:
        line["value_int32_kernel"] = {"error": str(e)[:200]}
    finally:
        eng.force_general(0)

    # The one mode of the path that HBM binds (SURVEY 8d mode M, the literal aligner_align contract):
    # all three int32 matrices of every pair written out, 12 B/cell.
    try:
        n_m = 50000
        a, oa, b, ob = [t.numpy() for t in host[0]]
        ms_m = []
        for _ in range(3):
            eng.submit_packed(sa.SW, sa.MODE_MATS, a[: n_m * LEN], oa[: n_m + 1], b[: n_m * LEN], ob[: n_m + 1])
            ms_m.append(eng.last_kernel_ms)
        bytes_m = 12 * n_m * (LEN + 1) * (LEN + 1)
        gbs_m = bytes_m / (min(ms_m) * 1e-3) / 1e9
        hbm_peak, _ = peaks()
        tr, tr_src = ncu_traffic(eng.last_kernel)
        line["roofline"]["materialise"] = {
            "workload": "SW, %d of the step's pairs, match/gap_a/gap_b matrices of every pair (SEQALIGN_MODE_MATS)" % n_m,
            "kernel": eng.last_kernel, "kernel_ms": min(ms_m), "bound": "hbm", "algorithmic_bytes": bytes_m,
            "achieved": gbs_m, "peak": hbm_peak, "unit": "GB/s", "frac": gbs_m / hbm_peak,
            "gcups": n_m * LEN * LEN / (min(ms_m) * 1e-3) / 1e9, "traffic": tr, "traffic_source": tr_src}
    except Exception as e:   # never lose the headline line over the side measurement
        line["roofline"]["materialise"] = {"error": str(e)[:200]}

    if not args.no_cpu_baseline:
        a, oa, b, ob = [t.numpy() for t in host[0]]
        ha, hoa, hb, hob = host_batch(SEED, 0, PAIRS)
        same_input = bool(np.array_equal(a, ha) and np.array_equal(b, hb))
        cores = os.cpu_count() or 1
        eng.submit_ptrs(sa.SW, sa.MODE_SCORE_ONLY, *[t.data_ptr() for t in host[0]], PAIRS)
        gpu_scores = eng.scores()
        if os.path.exists(REF_BATCH):
            g, sec, ref_scores = run_ref_batch("fill", cores, ha, hoa, hb, hob, PAIRS)
            gf, secf, _ = run_ref_batch("full", cores, ha, hoa, hb, hob, 10000)
            line["cpu_baseline"] = {
                "value": g, "unit": "GCUPS", "cores": cores, "kind": "reference",
                "sample": "unmodified reference aligner_align (fill) + best cell on all %d pairs of one step, "
                          "%d threads, %.1f s" % (PAIRS, cores, sec),
                "full_call_value": gf,
                "full_call_sample": "smith_waterman_align2 + first fetch (fresh aligner) on 10000 pairs, %.1f s" % secf,
                "scores_match_gpu": bool(np.array_equal(ref_scores, gpu_scores)),
                "device_generator_matches_host_generator": same_input}
        else:
            g, sec, s = cpu_port_gcups(ha, hoa, hb, hob, 2000)
            line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": 1, "kind": "port",
                                    "sample": "oracle C restatement on 2000 pairs of one step, 1 thread, %.1f s" % sec,
                                    "scores_match_gpu": bool(np.array_equal(s, gpu_scores[:2000]))}

    if not args.no_config5:
        try:
            del devb, host
            pipe.close()
            line["config5"] = run_job(cx, args, steps=max(3, args.steps // 4))
        except Exception as e:
            line["config5"] = {"error": str(e)[:300]}
    return line


# ---------------------------------------------------------------------------------------------
# BASELINE config 5: one job of 10 M pairs, strong scaling over the ranks (N=1: the same job on one GPU)

def run_job(cx, args, steps):
    torch, sa, eng, dev, dist = cx.torch, cx.sa, cx.eng, cx.dev, cx.dist
    world, rank = cx.world, cx.rank
    P = args.job_pairs
    first, last = P * rank // world, P * (rank + 1) // world
    n_loc = last - first
    a, oa, b, ob = cx.device_batch(C5_SEED, first, n_loc)
    torch.cuda.synchronize()
    uniform_n = (P % world == 0)
    cells_job = P * LEN * LEN
    d_scores = [torch.zeros(n_loc, dtype=torch.int32, device=dev) for _ in range(3)]
    all_scores = torch.zeros(P, dtype=torch.int32, device=dev) if rank == 0 else None
    gstream = torch.cuda.Stream(device=dev)
    counts = [P * (r + 1) // world - P * r // world for r in range(world)]

    def gather_to_root(src_scores, stream_ctx=True):
        """every rank's scores to rank 0 over NCCL, in pair order (all_scores)"""
        if dist is None:
            all_scores.copy_(src_scores)
            return
        if uniform_n:
            dist.gather(src_scores, list(all_scores.split(n_loc)) if rank == 0 else None, dst=0)
        else:
            width = max(counts)
            pad = torch.zeros(width, dtype=torch.int32, device=dev)
            pad[:n_loc] = src_scores
            bucket = [torch.zeros(width, dtype=torch.int32, device=dev) for _ in range(world)] if rank == 0 else None
            dist.gather(pad, bucket, dst=0)
            if rank == 0:
                at = 0
                for r in range(world):
                    all_scores[at:at + counts[r]] = bucket[r][:counts[r]]
                    at += counts[r]

    # ---- value: shards resident in HBM; a step = the DP over every shard + the gather of the scores to rank 0
    evs = {}

    def after(i):
        # step i is complete on the compute stream up to the event recorded right after its enqueue;
        # its gather runs on a side stream so that step i+1's kernel is not held up behind it
        gstream.wait_event(evs.pop(i))
        with torch.cuda.stream(gstream):
            gather_to_root(d_scores[i % 3])

    def value_steps(n, first_step):
        eng_ms, q = [], collections.deque()
        for i in range(first_step, first_step + n):
            eng.run_device_async(sa.SW, a.data_ptr(), oa.data_ptr(), b.data_ptr(), ob.data_ptr(), n_loc,
                                 d_scores[i % 3].data_ptr(), 0, 0, cx.stream)
            ev = torch.cuda.Event()
            ev.record(cx.tstream)
            evs[i] = ev
            q.append(i)
            if len(q) > 1:
                eng.run_device_wait()
                eng_ms.append(eng.last_kernel_ms)
                after(q.popleft())
        while q:
            eng.run_device_wait()
            eng_ms.append(eng.last_kernel_ms)
            after(q.popleft())
        cx.tstream.wait_stream(gstream)
        return eng_ms

    value_steps(max(args.warmup, 3), 0)
    sampler = ClockSampler(cx.local)
    sampler.start()
    cx.barrier()
    ts0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    kernel_ms = value_steps(steps, 100)
    e1.record()
    cx.barrier()
    ts1 = time.perf_counter()
    dt_ms = e0.elapsed_time(e1)
    kernel_name = eng.last_kernel
    checksum = int(all_scores.to(torch.int64).sum().item()) if rank == 0 else 0
    head_scores = all_scores[:20000].cpu().numpy() if rank == 0 else None

    # ---- e2e: every rank's shard in its own pinned host memory, through the host-buffer C-ABI call;
    #      scores back to the host of their rank, then over NCCL to rank 0 and onto its host
    host = tuple(cx.pinned(t) for t in (a, oa, b, ob))
    del a, b
    torch.cuda.empty_cache()
    depth = 2
    pipe = sa.PipelinedAligner(cx.local, cx.scoring, depth=depth)
    h_scores = [torch.empty(n_loc, dtype=torch.int32, pin_memory=True) for _ in range(depth + 1)]
    h_all = torch.empty(P, dtype=torch.int32, pin_memory=True) if rank == 0 else None

    def e2e_steps(n):
        pending, last_sum = collections.deque(), 0

        def finish(k, fut):
            hs = h_scores[k % (depth + 1)]
            fut.result()            # the scores are in hs: written there by the engine's device->host copies
            d = d_scores[k % 3]
            d.copy_(hs, non_blocking=True)
            gather_to_root(d)
            if rank == 0:
                h_all.copy_(all_scores, non_blocking=True)

        for k in range(n):
            pending.append((k, pipe.submit_uniform_ptrs(sa.SW, sa.MODE_SCORE_ONLY, host[0].data_ptr(), LEN, host[2].data_ptr(), LEN,
                                                        n_loc, h_scores[k % (depth + 1)].data_ptr())))
            if len(pending) > depth - 1:
                finish(*pending.popleft())
        while pending:
            finish(*pending.popleft())
        torch.cuda.synchronize()
        return int(h_all.numpy().astype(np.int64).sum()) if rank == 0 else 0

    e2e_steps(2 * depth)
    cx.barrier()
    t0 = time.perf_counter()
    e2e_checksum = e2e_steps(steps)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    cx.barrier()
    pipe.close()
    te1 = time.perf_counter()

    # ---- e2e_scatter: the whole job in rank 0's host memory; chunked host -> GPU 0 -> NCCL -> run on arrival
    scat = None
    if world > 1 and not args.no_scatter:
        from seqalign.distributed import align_sharded_stream
        job = None
        if rank == 0:
            # the job is generated shard by shard on the device and parked in rank 0's pinned memory
            ja = torch.empty(P * LEN, dtype=torch.uint8, pin_memory=True)
            jb = torch.empty(P * LEN, dtype=torch.uint8, pin_memory=True)
            step_n = 1_000_000
            for p0 in range(0, P, step_n):
                m = min(step_n, P - p0)
                ta, _, tb, _ = cx.device_batch(C5_SEED, p0, m)
                ja[p0 * LEN:(p0 + m) * LEN].copy_(ta)
                jb[p0 * LEN:(p0 + m) * LEN].copy_(tb)
            job = (ja, jb)
        torch.cuda.empty_cache()
        sc_ms, sc_sum = [], 0
        for it in range(2 + max(2, steps // 2)):
            cx.barrier()
            t0 = time.perf_counter()
            res = align_sharded_stream(eng, sa.SW, job[0] if job else None, None, job[1] if job else None, None, src=0,
                                       device=dev, chunk_pairs=args.scatter_chunk, uniform=(LEN, LEN), ring=4)
            if rank == 0:
                h_all.copy_(res, non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) * 1e3
            if it >= 2:
                sc_ms.append(dt)
            if rank == 0:
                sc_sum = int(h_all.numpy().astype(np.int64).sum())
        ms = cx.reduce([float(np.mean(sc_ms))])[0]
        scat = {"value": cells_job / (ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": ms,
                "h2d_bytes_per_step": int(2 * P * LEN), "nccl_bytes_per_step": int(2 * (P - counts[0]) * LEN + 4 * (P - counts[0])),
                "d2h_bytes_per_step": int(4 * P), "chunk_pairs": args.scatter_chunk, "score_checksum": sc_sum,
                "pcie_floor_note": "rank 0's link carries the whole job: %.2f GB per step" % (2 * P * LEN / 1e9),
                "api": "seqalign.distributed.align_sharded_stream: pinned host (rank 0) -> GPU 0 -> NCCL isend/irecv per "
                       "chunk -> seqalign_batch_run_device_async on arrival -> NCCL gather"}
        del job
    sampler.stop_flag = True

    dt_ms, e2e_ms = cx.reduce([dt_ms, e2e_ms])
    k_ms = float(np.mean(kernel_ms))
    clocks = sampler.summary(ts0, te1)
    rec = {"workload": WORKLOAD_N, "n_gpus": world, "steps": steps, "pairs": P, "pairs_per_gpu": n_loc,
           "value": cells_job / (dt_ms / steps * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": dt_ms / steps,
           "kernel": kernel_name, "kernel_ms_rank0": k_ms,
           "kernel_gcups_rank0": n_loc * LEN * LEN / (k_ms * 1e-3) / 1e9,
           "score_checksum": checksum,
           "e2e": {"value": cells_job / (e2e_ms / steps * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": e2e_ms / steps,
                   "h2d_bytes_per_step": int(2 * P * LEN), "d2h_bytes_per_step": int(4 * P + 4 * P),
                   "h2d_bytes_per_step_per_gpu": int(2 * n_loc * LEN),
                   "nccl_bytes_per_step": int(4 * (P - counts[0])) if world > 1 else 0,
                   "score_checksum": e2e_checksum,
                   "api": "per rank: seqalign.PipelinedAligner(depth=%d) -> seqalign_batch_submit_uniform on the rank's own pinned "
                          "shard, scores into a pinned result sink; then host -> device -> NCCL gather -> rank 0's host" % depth},
           "clocks": clocks}
    if scat:
        rec["e2e_scatter"] = scat
    if rank == 0 and not args.no_cpu_baseline and os.path.exists(REF_BATCH):
        n_chk = 20000
        ha, hoa, hb, hob = host_batch(C5_SEED, 0, n_chk)
        _, _, ref_scores = run_ref_batch("fill", os.cpu_count() or 1, ha, hoa, hb, hob, n_chk)
        rec["first_pairs_match_reference"] = bool(np.array_equal(ref_scores, head_scores[:n_chk]))
        rec["first_pairs_checked"] = n_chk
    rec["_roofline_args"] = (kernel_name, k_ms, n_loc)
    return rec


def run_sharded(cx, args):
    rec = run_job(cx, args, steps=args.steps)
    if cx.rank != 0:
        return None
    kernel_name, k_ms, n_loc = rec.pop("_roofline_args")
    line = {
        "metric": "DP cell updates/s (GCUPS)", "value": rec["value"], "unit": "GCUPS", "n_gpus": cx.world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": WORKLOAD_N, "pairs": rec["pairs"], "pairs_per_gpu": rec["pairs_per_gpu"],
                   "cells_per_step": rec["pairs"] * LEN * LEN,
                   "generator": "splitmix64 counter generator of SURVEY.md 8d, seed %d; rank r makes pairs [r*P/N, (r+1)*P/N) "
                                "of the job on its own device" % C5_SEED,
                   "l2": "every step re-reads a shard of %.0f MB (> 126 MB L2)" % (2 * n_loc * LEN / 1e6),
                   "parallelism": "pairs sharded by rank (strong scaling); NCCL gather of all scores to rank 0 inside every "
                                  "timed step; e2e_scatter adds the NCCL input scatter from rank 0",
                   "kernel": kernel_name, "score_checksum": rec["score_checksum"],
                   "first_pairs_match_reference": rec.get("first_pairs_match_reference")},
        "clocks": rec["clocks"],
        "e2e": rec["e2e"],
        "gpu_launches": 2 * args.steps * cx.world,
        "roofline": roofline_block(cx, kernel_name, k_ms, n_loc, rec["clocks"]),
    }
    if "e2e_scatter" in rec:
        line["e2e_scatter"] = rec["e2e_scatter"]
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="N=1: skip the side record of the 10 M-pair job")
    ap.add_argument("--no-scatter", action="store_true", help="N>1: skip the rank-0 NCCL scatter arm")
    ap.add_argument("--job-pairs", type=int, default=C5_PAIRS)
    ap.add_argument("--scatter-chunk", type=int, default=131072)
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of the sustained loop (N=1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        reference_arm(args)
        return

    # stdout carries exactly one JSON line: everything else that libraries print there
    # (NCCL announces its version on stdout) is sent to stderr until the line is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    cx = Ctx()
    if cx.world == 1:
        line = run_single(cx, args)
        if "config5" in line and isinstance(line["config5"], dict):
            line["config5"].pop("_roofline_args", None)
    else:
        line = run_sharded(cx, args)
    if cx.rank == 0:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if cx.dist is not None:
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
