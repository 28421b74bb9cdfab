#!/usr/bin/env python3
"""Floor of the multi-GPU end-to-end arm: how fast can N ranks pull pinned host memory over their own
PCIe links AT THE SAME TIME, with no kernels at all?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \\
        tools/h2d_floor.py  >> profiles/h2d_concurrent_r02.jsonl

For k = 1, 2, 4, .. N active ranks (the others idle at the barrier): every active rank copies its
shard of the BASELINE config 5 job (10 M pairs x 300 B / N, and a fixed 375 MB for comparison) host ->
device `reps` times on one stream; aggregate GB/s = bytes of all active ranks / slowest rank's time.
Variants: ordinary pinned memory, write-combined pinned memory (cudaHostAllocWriteCombined), and
ranks pinned to disjoint core sets or not.  The D2H direction is measured the same way (scores)."""
import ctypes, json, os, sys, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cudart = ctypes.CDLL("libcudart.so.12")


def wc_pinned(nbytes):
    p = ctypes.c_void_p()
    rc = cudart.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(4))   # cudaHostAllocWriteCombined
    assert rc == 0, rc
    return p


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed_copy(active, fn, reps):
    barrier()
    t0 = time.perf_counter()
    if active:
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0 if active else 0.0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


rows = []
for pin_cores in (False, True):
    if pin_cores:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // world)
        os.sched_setaffinity(0, cores[local * per:(local + 1) * per] or cores)
    for nbytes_label, nbytes in (("config5 shard", 3_000_000_000 // world), ("375 MB", 375_000_000)):
        d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        h.fill_(65)
        wc = wc_pinned(nbytes)
        ctypes.memset(wc, 65, nbytes)
        st = torch.cuda.current_stream().cuda_stream
        reps = max(3, int(3e9 // nbytes))
        k = 1
        while k <= world:
            active = rank < k
            for label, fn in (("pinned h2d", lambda: d.copy_(h, non_blocking=True)),
                              ("write-combined h2d", lambda: cudart.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), wc, ctypes.c_size_t(nbytes), 1, ctypes.c_void_p(st))),
                              ("pinned d2h", lambda: h.copy_(d, non_blocking=True))):
                timed_copy(active, fn, 1)
                dt = timed_copy(active, fn, reps)
                rows.append(dict(what=label, active_ranks=k, world=world, bytes_per_rank=nbytes, size=nbytes_label, reps=reps,
                                 cores_pinned=pin_cores, seconds=dt, aggregate_gbs=k * reps * nbytes / dt / 1e9,
                                 per_rank_gbs=reps * nbytes / dt / 1e9))
            k *= 2
        cudart.cudaFreeHost(wc)
        del d, h
if rank == 0:
    import subprocess
    info = dict(what="host", nproc=os.cpu_count(),
                cpu=subprocess.run("lscpu | grep -E 'Model name|Socket|NUMA node\\(s\\)|Thread' | tr -s ' ' | tr '\\n' ';'", shell=True, capture_output=True, text=True).stdout,
                mem=subprocess.run("free -g | sed -n 2p", shell=True, capture_output=True, text=True).stdout.strip())
    print(json.dumps(info))
    for r in rows:
        print(json.dumps(r))
if world > 1:
    dist.destroy_process_group()
