"""Sharding a batch of pairs over one process per GPU (torch.distributed).

Pairs are independent (the reference itself loops pair by pair,
src/alignment_cmdline.c:611-622), so the DP needs no exchange step: a job is
cut into contiguous pair ranges balanced by cell count, every rank aligns its
range on its own device, and the only communication is moving inputs out from
and results back to the rank that owns them.  With the NCCL backend the
tensors are CUDA tensors and travel over NVLink; the engine then runs on the
received device buffers directly (seqalign_batch_run_device).  The same code
runs on the gloo backend with CPU tensors (host-buffer submit), which is how
the logic is tested without GPUs.

Ranks that can load or generate their own shard should do that instead and
skip scatter_pairs (bench.py does): the input scatter from one rank is
bounded by that rank's PCIe link, not by NVLink.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import MODE_SCORE


def shard_bounds(off_a, off_b, world):
    """Contiguous pair ranges [b[r], b[r+1]) with near-equal sum(len_a*len_b)."""
    la = np.diff(np.asarray(off_a, dtype=np.int64))
    lb = np.diff(np.asarray(off_b, dtype=np.int64))
    n = len(la)
    # +1 so that empty pairs still spread out
    w = np.cumsum(la * lb + 1)
    total = int(w[-1]) if n else 0
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(w, total * r / world, side="right")) if n else 0)
    bounds.append(n)
    for r in range(1, world + 1):
        bounds[r] = max(bounds[r], bounds[r - 1])
    return bounds


def _as_tensor(x, device):
    t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    return t.to(device, non_blocking=True)


def shard_bounds_tensor(toa, tob, world):
    """shard_bounds() on int64 offset tensors, computed where they live (for a
    batch that is already on the GPU this avoids a host pass over every pair)"""
    n = toa.numel() - 1
    if n <= 0:
        return [0] * (world + 1)
    w = torch.cumsum(toa.diff() * tob.diff() + 1, 0)
    total = int(w[-1].item())
    # same cut points as the numpy version: first index with w > total*r/world (float64 targets)
    targets = torch.tensor([total * r / world for r in range(1, world)], dtype=torch.float64, device=w.device)
    cuts = torch.searchsorted(w.to(torch.float64), targets, right=True).tolist() if world > 1 else []
    bounds = [0] + [int(c) for c in cuts] + [n]
    for r in range(1, world + 1):
        bounds[r] = max(bounds[r], bounds[r - 1])
    return bounds


def scatter_pairs(seq_a, off_a, seq_b, off_b, src=0, device="cpu", group=None):
    """Rank `src` passes the packed batch (numpy arrays or tensors; other ranks
    pass None) and every rank gets back its shard as tensors on `device`:
    (seq_a, off_a, seq_b, off_b, first_pair, bounds), offsets rebased to 0.
    The shard table is computed on `device`, the shards travel as ONE grouped
    send/recv (dist.batch_isend_irecv: NCCL over NVLink for CUDA tensors)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    meta = torch.zeros(3 * (world + 1), dtype=torch.int64, device=device)
    if rank == src:
        toa = _as_tensor(off_a, device).to(torch.int64)
        tob = _as_tensor(off_b, device).to(torch.int64)
        bounds = shard_bounds_tensor(toa, tob, world)
        bt = torch.tensor(bounds, device=device)
        meta[: world + 1] = bt
        meta[world + 1: 2 * (world + 1)] = toa[bt]
        meta[2 * (world + 1):] = tob[bt]
    if world > 1:
        dist.broadcast(meta, src, group=group)
    m = meta.cpu().numpy()
    bounds, ba, bb = m[: world + 1], m[world + 1: 2 * (world + 1)], m[2 * (world + 1):]
    n_local = int(bounds[rank + 1] - bounds[rank])
    if rank == src:
        ta, tb = _as_tensor(seq_a, device), _as_tensor(seq_b, device)
        ops, keep, mine = [], [], None
        for r in range(world):
            parts = [ta[ba[r]: ba[r + 1]], toa[bounds[r]: bounds[r + 1] + 1] - int(ba[r]),
                     tb[bb[r]: bb[r + 1]], tob[bounds[r]: bounds[r + 1] + 1] - int(bb[r])]
            if r == src:
                mine = parts            # views of the source buffers: no copy for the local shard
            else:
                keep += parts
                ops += [dist.P2POp(dist.isend, p, r, group) for p in parts]
        if ops:
            for q in dist.batch_isend_irecv(ops):
                q.wait()
    else:
        mine = [torch.empty(int(ba[rank + 1] - ba[rank]), dtype=torch.uint8, device=device),
                torch.empty(n_local + 1, dtype=torch.int64, device=device),
                torch.empty(int(bb[rank + 1] - bb[rank]), dtype=torch.uint8, device=device),
                torch.empty(n_local + 1, dtype=torch.int64, device=device)]
        for q in dist.batch_isend_irecv([dist.P2POp(dist.irecv, t, src, group) for t in mine]):
            q.wait()
    return mine[0], mine[1], mine[2], mine[3], int(bounds[rank]), [int(v) for v in bounds]


def gather_results(local, bounds, dst=0, group=None):
    """local: int32 tensor [k, n_local] (e.g. score/x_end/y_end rows).  Rank
    `dst` gets [k, n_total] in the original pair order, others None."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    counts = [bounds[r + 1] - bounds[r] for r in range(world)]
    width = max(counts) if counts else 0
    k = local.shape[0]
    padded = torch.zeros((k, width), dtype=local.dtype, device=local.device)
    padded[:, : local.shape[1]] = local
    bucket = [torch.zeros_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bucket[r][:, : counts[r]] for r in range(world)], dim=1)


def align_sharded(engine, algo, seq_a=None, off_a=None, seq_b=None, off_b=None, src=0, device="cpu", group=None,
                  want_ends=True, timings=None):
    """Score mode over all ranks: scatter from `src`, align locally, gather
    (score, x_end, y_end) back to `src` as an int32 tensor [3, n].
    want_ends=False (CUDA tensors): scores only, rows 1-2 stay 0 and the engine
    may use its packed 16-bit kernel.  timings: dict that receives the seconds
    spent in scatter / align / gather on this rank (device synchronised)."""
    import time

    def tick():
        if timings is not None and str(device) != "cpu":
            torch.cuda.synchronize()
        return time.perf_counter()

    t0 = tick()
    a, oa, b, ob, _, bounds = scatter_pairs(seq_a, off_a, seq_b, off_b, src, device, group)
    t1 = tick()
    n_local = oa.numel() - 1
    if a.is_cuda:
        out = torch.zeros((3, n_local), dtype=torch.int32, device=a.device)
        if n_local:
            engine.run_device(algo, a.data_ptr(), oa.data_ptr(), b.data_ptr(), ob.data_ptr(), n_local,
                              out[0].data_ptr(), out[1].data_ptr() if want_ends else 0,
                              out[2].data_ptr() if want_ends else 0,
                              torch.cuda.current_stream().cuda_stream)
        t2 = tick()
        res = gather_results(out, bounds, src, group)
        t3 = tick()
        if timings is not None:
            timings.update(scatter=t1 - t0, align=t2 - t1, gather=t3 - t2, kernel_ms=engine.last_kernel_ms,
                           kernel=engine.last_kernel, n_local=n_local)
        return res
    else:
        engine.submit_packed(algo, MODE_SCORE, a.numpy(), oa.numpy(), b.numpy(), ob.numpy())
        out = torch.from_numpy(np.stack(engine.ends()).astype(np.int32))
    return gather_results(out, bounds, src, group)


class SharedBatch:
    """A packed batch [seq_a | seq_b | off_a | off_b] in ONE device buffer that
    the owner's peers can map (seqalign.SharedBuffer, CUDA IPC).  The owner
    fills it once (host -> its GPU); nobody copies it again."""

    def __init__(self, buf, n, ta, tb, pos):
        self.buf, self.n, self.ta, self.tb, self.pos = buf, n, ta, tb, pos
        t = buf.tensor()
        self.seq_a = t[pos[0]: pos[0] + ta]
        self.seq_b = t[pos[1]: pos[1] + tb]
        self.off_a = t[pos[2]: pos[2] + 8 * (n + 1)].view(torch.int64)
        self.off_b = t[pos[3]: pos[3] + 8 * (n + 1)].view(torch.int64)

    @staticmethod
    def _layout(n, ta, tb):
        pos, at = [], 0
        for size in (ta + 32, tb + 32, 8 * (n + 1), 8 * (n + 1)):   # +32: the kernels' 16-byte bulk loads may run past the end
            pos.append(at)
            at = (at + size + 255) & ~255
        return pos, at

    @classmethod
    def create(cls, device, seq_a, off_a, seq_b, off_b):
        """owner side: numpy arrays or (pinned) host / device tensors"""
        from . import SharedBuffer
        ts = [torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x for x in (seq_a, off_a, seq_b, off_b)]
        n, ta, tb = ts[1].numel() - 1, ts[0].numel(), ts[2].numel()
        pos, total = cls._layout(n, ta, tb)
        sb = cls(SharedBuffer(device, total), n, ta, tb, pos)
        sb.seq_a.copy_(ts[0].view(torch.uint8), non_blocking=True)
        sb.seq_b.copy_(ts[2].view(torch.uint8), non_blocking=True)
        sb.off_a.copy_(ts[1].to(torch.int64), non_blocking=True)
        sb.off_b.copy_(ts[3].to(torch.int64), non_blocking=True)
        return sb

    def meta(self):
        return dict(handle=self.buf.handle, nbytes=self.buf.nbytes, n=self.n, ta=self.ta, tb=self.tb, pos=self.pos)

    @classmethod
    def open(cls, device, meta):
        from . import SharedBuffer
        return cls(SharedBuffer(device, meta["nbytes"], meta["handle"]), meta["n"], meta["ta"], meta["tb"], meta["pos"])

    def close(self):
        del self.seq_a, self.seq_b, self.off_a, self.off_b
        self.buf.close()


def align_sharded_peer(engine, algo, batch=None, src=0, group=None, want_ends=False, timings=None):
    """align_sharded() without the scatter: the batch stays in the HBM of rank
    `src` (a SharedBatch there, None elsewhere); every other rank maps it onto
    its own device through CUDA IPC and its DP kernel pulls its pair range over
    NVLink while it computes (TMA bulk loads from peer memory: 0.0135 B/cell,
    ~65 GB/s per GPU at full speed, far below a link).  Only the scores travel
    back (gather).  Returns [3, n] int32 on `src` like align_sharded."""
    import time
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device())

    def tick():
        torch.cuda.synchronize()
        return time.perf_counter()

    t0 = tick()
    box = [(batch.meta(), shard_bounds_tensor(batch.off_a, batch.off_b, world)) if rank == src else None]
    dist.broadcast_object_list(box, src, group=group)
    meta, bounds = box[0]
    mine = batch if rank == src else SharedBatch.open(dev.index, meta)
    first, n_local = bounds[rank], bounds[rank + 1] - bounds[rank]
    t1 = tick()
    out = torch.zeros((3, n_local), dtype=torch.int32, device=dev)
    if n_local:
        base = mine.buf.address
        # offsets stay absolute: the kernels address seq + off[p], so a shard is just a window of the offset arrays
        engine.run_device(algo, base + mine.pos[0], base + mine.pos[2] + 8 * first, base + mine.pos[1],
                          base + mine.pos[3] + 8 * first, n_local, out[0].data_ptr(),
                          out[1].data_ptr() if want_ends else 0, out[2].data_ptr() if want_ends else 0,
                          torch.cuda.current_stream().cuda_stream)
    t2 = tick()
    res = gather_results(out, bounds, src, group)
    t3 = tick()
    if timings is not None:
        timings.update(scatter=t1 - t0, align=t2 - t1, gather=t3 - t2, kernel_ms=engine.last_kernel_ms,
                       kernel=engine.last_kernel, n_local=n_local)
    if rank != src:
        mine.close()
    dist.barrier(group=group)   # the owner may reuse or free the buffer only after every reader is done
    return res
