#!/usr/bin/env python3
"""BASELINE config 5: SW score-only over N x B200, the batch living on rank 0 and sharded over
NCCL (seqalign.distributed.align_sharded: broadcast of the shard table, grouped isend/recv of the
packed sequences over NVLink, local run on the received device buffers, gather of the scores).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port 29511 tools/gpu_config5.py [total_pairs] [unique_pairs]

total_pairs (default 10,000,000) are built from `unique_pairs` generated pairs (default 1,000,000,
seed 5) repeated: generation is single-threaded numpy, the aligner does the full work either way.
Parity: the first 100,000 scores against a one-GPU host-buffer run of the same pairs, and a
checksum over all scores that must not depend on N."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
from helpers import *
from seqalign.distributed import align_sharded, align_sharded_peer, SharedBatch

PEER = "--peer" in sys.argv   # zero-copy: ranks read rank 0's HBM over NVLink instead of receiving a scatter
argv = [x for x in sys.argv[1:] if not x.startswith("--")]
TOTAL = int(argv[0]) if len(argv) > 0 else 10_000_000
UNIQUE = int(argv[1]) if len(argv) > 1 else 1_000_000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = seqalign.BatchAligner(local, seqalign.Scoring.sw_cli_default())
L = 150
seq_a = off_a = seq_b = off_b = None
if rank == 0:
    t = time.time()
    from concurrent.futures import ProcessPoolExecutor
    blocks = [(5 + 1000 * k, 100000, L, L) for k in range(UNIQUE // 100000)]
    with ProcessPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        parts = list(ex.map(synthetic_batch, *zip(*blocks)))
    ua = np.concatenate([p[0] for p in parts]); ub = np.concatenate([p[2] for p in parts])
    reps = TOTAL // UNIQUE
    seq_a = torch.from_numpy(np.tile(ua, reps)).pin_memory(); seq_b = torch.from_numpy(np.tile(ub, reps)).pin_memory()
    off_a = torch.from_numpy(np.arange(0, (TOTAL + 1) * L, L, dtype=np.int64)).pin_memory(); off_b = off_a.clone().pin_memory()
    print("rank 0: %d pairs (%d unique) built in %.1f s, %.2f GB" % (TOTAL, UNIQUE, time.time() - t, 2 * seq_a.numel() / 1e9), file=sys.stderr, flush=True)
rows = []
for rep in range(3):
    dist.barrier(); torch.cuda.synchronize()
    tm = {}
    t0 = time.perf_counter()
    if rank == 0:
        if PEER:
            batch = SharedBatch.create(local, seq_a, off_a, seq_b, off_b)    # host -> GPU 0 over PCIe, straight into the shared buffer
            torch.cuda.synchronize()
            t_h2d = time.perf_counter() - t0
            res = align_sharded_peer(eng, seqalign.SW, batch, src=0, want_ends=False, timings=tm)
            batch.close()
        else:
            da, db = seq_a.to(dev, non_blocking=True), seq_b.to(dev, non_blocking=True)   # host -> GPU 0 over PCIe
            torch.cuda.synchronize()
            t_h2d = time.perf_counter() - t0
            res = align_sharded(eng, seqalign.SW, da, off_a, db, off_b, src=0, device=dev, want_ends=False, timings=tm)
    else:
        t_h2d = 0.0
        if PEER:
            res = align_sharded_peer(eng, seqalign.SW, src=0, want_ends=False, timings=tm)
        else:
            res = align_sharded(eng, seqalign.SW, src=0, device=dev, want_ends=False, timings=tm)
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    stats = torch.tensor([tm["scatter"], tm["align"], tm["gather"], tm["kernel_ms"] / 1e3, t_all], device=dev, dtype=torch.float64)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    if rank == 0:
        scores = res[0]
        cells = TOTAL * L * L
        sc, al, ga, km, ta = [float(x) for x in stats.tolist()]
        rows.append(dict(rep=rep, mode="peer" if PEER else "scatter", n_gpus=world, pairs=TOTAL, h2d_rank0_s=round(t_h2d, 4), scatter_s=round(sc, 4), align_s=round(al, 4),
                         gather_s=round(ga, 4), kernel_s_max=round(km, 4), total_s=round(ta, 4), kernel=tm["kernel"],
                         gcups_align=round(cells / al / 1e9, 1), gcups_from_gpu0=round(cells / (sc + al + ga) / 1e9, 1),
                         gcups_from_host=round(cells / ta / 1e9, 1), checksum=int(scores.to(torch.int64).sum().item())))
        print(json.dumps(rows[-1]), flush=True)
if rank == 0:
    # parity of the first 100k pairs against a plain one-GPU host-buffer run
    n = 100000
    eng.submit_packed(seqalign.SW, seqalign.MODE_SCORE, seq_a[: n * L].numpy(), off_a[: n + 1].numpy(), seq_b[: n * L].numpy(), off_b[: n + 1].numpy())
    same = bool(np.array_equal(eng.scores(), scores[:n].cpu().numpy()))
    o = orc_from_scoring(seqalign.Scoring.sw_cli_default())
    es, _, _ = orc_batch_sw(o, seq_a[: 2000 * L].numpy(), off_a[:2001].numpy(), seq_b[: 2000 * L].numpy(), off_b[:2001].numpy())
    oracle_same = bool(np.array_equal(es, scores[:2000].cpu().numpy()))
    print(json.dumps(dict(first_100k_same_as_one_gpu=same, first_2000_same_as_oracle=oracle_same)), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(rows=rows, first_100k_same_as_one_gpu=same, first_2000_same_as_oracle=oracle_same),
              open(os.path.join(ROOT, "gpurun_out", "config5_%s_n%d.json" % ("peer" if PEER else "scatter", world)), "w"), indent=1)
eng.close()
dist.destroy_process_group()
