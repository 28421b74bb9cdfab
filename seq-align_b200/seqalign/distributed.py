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


def _cut_points(toa, tob, world):
    """tensor [world+1] of shard cut points, computed where the offsets live, no host sync"""
    n = toa.numel() - 1
    dev = toa.device
    if n <= 0:
        return torch.zeros(world + 1, dtype=torch.int64, device=dev)
    w = torch.cumsum(toa.diff() * tob.diff() + 1, 0)
    # same cut points as shard_bounds(): first index with w > total*r/world (float64 targets)
    frac = torch.arange(1, world, dtype=torch.float64, device=dev) / world
    cuts = torch.searchsorted(w.to(torch.float64), w[-1].to(torch.float64) * frac, right=True)
    bounds = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), cuts.to(torch.int64),
                        torch.full((1,), n, dtype=torch.int64, device=dev)])
    return torch.cummax(bounds, 0).values


def shard_bounds_tensor(toa, tob, world):
    """shard_bounds() on int64 offset tensors, computed where they live (for a
    batch that is already on the GPU this avoids a host pass over every pair);
    one device->host read of world+1 numbers at the end"""
    return [int(v) for v in _cut_points(toa, tob, world).tolist()]


def _slack_empty(nbytes, device):
    """uint8 buffer for a received shard: the DP kernels stage sequences with 16-byte bulk copies
    aligned on the OFFSET, so the base must be 16-byte aligned (torch allocations are 512-byte
    aligned) and readable up to the next 16-byte boundary past the end"""
    return torch.empty(int(nbytes) + 32, dtype=torch.uint8, device=device)[: int(nbytes)]


def scatter_pairs(seq_a, off_a, seq_b, off_b, src=0, device="cpu", group=None):
    """Rank `src` passes the packed batch (numpy arrays or tensors; other ranks
    pass None) and every rank gets back its shard as tensors on `device`:
    (seq_a, off_a, seq_b, off_b, first_pair, bounds).  off_* index into the
    returned seq_* (rebased to 0 for received shards; for `src` on a CUDA device
    the returned sequences are the source buffers themselves and the offsets
    stay absolute, so that no shard starts on an unaligned address).  The shard
    table is computed on `device` and read back once; the shards travel as ONE
    grouped send/recv (dist.batch_isend_irecv: NCCL over NVLink for CUDA
    tensors).  For a job that starts in host memory use align_sharded_stream():
    it overlaps the host->device copy, the sends and the alignment chunk by chunk."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    meta = torch.zeros(3 * (world + 1), dtype=torch.int64, device=device)
    if rank == src:
        toa = _as_tensor(off_a, device).to(torch.int64)
        tob = _as_tensor(off_b, device).to(torch.int64)
        bt = _cut_points(toa, tob, world)
        meta = torch.cat([bt, toa[bt], tob[bt]])
    if world > 1:
        dist.broadcast(meta, src, group=group)
    m = meta.cpu().numpy()          # the one host read: every rank needs its sizes to allocate
    bounds, ba, bb = m[: world + 1], m[world + 1: 2 * (world + 1)], m[2 * (world + 1):]
    n_local = int(bounds[rank + 1] - bounds[rank])
    if rank == src:
        ta, tb = _as_tensor(seq_a, device), _as_tensor(seq_b, device)
        ops, keep, mine = [], [], None
        for r in range(world):
            oa_r, ob_r = toa[bounds[r]: bounds[r + 1] + 1], tob[bounds[r]: bounds[r + 1] + 1]
            if r == src:
                if ta.is_cuda:
                    mine = [ta, oa_r, tb, ob_r]       # in place: absolute offsets into the source buffers
                else:
                    mine = [ta[ba[r]: ba[r + 1]], oa_r - int(ba[r]), tb[bb[r]: bb[r + 1]], ob_r - int(bb[r])]
            else:
                parts = [ta[ba[r]: ba[r + 1]], oa_r - meta[world + 1 + r], tb[bb[r]: bb[r + 1]], ob_r - meta[2 * (world + 1) + r]]
                keep += parts
                ops += [dist.P2POp(dist.isend, p, r, group) for p in parts]
        if ops:
            for q in dist.batch_isend_irecv(ops):
                q.wait()
    else:
        mine = [_slack_empty(ba[rank + 1] - ba[rank], device),
                torch.empty(n_local + 1, dtype=torch.int64, device=device),
                _slack_empty(bb[rank + 1] - bb[rank], device),
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


def _chunk_table(off_a, off_b, world, chunk_pairs):
    """host-side plan of a chunked scatter: shard bounds (balanced by cells) and, per rank, the
    chunks [p0, p1) of at most chunk_pairs pairs with their byte ranges in seq_a / seq_b"""
    oa, ob = np.asarray(off_a, dtype=np.int64), np.asarray(off_b, dtype=np.int64)
    bounds = shard_bounds(oa, ob, world)
    chunks = []
    for r in range(world):
        mine = []
        for p0 in range(bounds[r], bounds[r + 1], chunk_pairs):
            p1 = min(p0 + chunk_pairs, bounds[r + 1])
            mine.append((p0, p1, int(oa[p0]), int(oa[p1]), int(ob[p0]), int(ob[p1])))
        chunks.append(mine)
    return bounds, chunks


def align_sharded_stream(engine, algo, seq_a=None, off_a=None, seq_b=None, off_b=None, src=0, device="cpu",
                         group=None, chunk_pairs=65536, uniform=None, timings=None, ring=4):
    """Score-only job that starts in HOST memory of rank `src` (pinned tensors or numpy arrays;
    other ranks pass None): the batch is cut into shards balanced by cells and every shard into
    chunks of `chunk_pairs` pairs; `src` pushes the chunks host -> its GPU on a copy stream and
    straight on to their rank (NCCL isend over NVLink), round-robin over the ranks, while every
    rank -- `src` included -- aligns the chunks it already holds (seqalign_batch_run_device_async).
    The host->device copy, the sends and the DP kernels of different chunks overlap; the job is
    bound by `src`'s PCIe link, not by the sum of the three.  Scores are gathered
    to `src` (int32 tensor [n] there, None elsewhere).
    uniform=(len_a, len_b): every pair has that shape, so no offsets are shipped -- each rank
    makes its own.  Otherwise the offset slices travel with their chunk (16 B per pair).
    The same code runs on gloo with CPU tensors (tests)."""
    import time
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    cuda = str(device) != "cpu"
    t_start = time.perf_counter()

    box = [None]
    if rank == src:
        if uniform is not None:
            n = (len(off_a) - 1) if off_a is not None else int(len(seq_a) // max(uniform[0], 1))
            la, lb = uniform
            bounds = [n * r // world for r in range(world + 1)]
            chunks = [[(p0, min(p0 + chunk_pairs, bounds[r + 1]), p0 * la, min(p0 + chunk_pairs, bounds[r + 1]) * la,
                        p0 * lb, min(p0 + chunk_pairs, bounds[r + 1]) * lb)
                       for p0 in range(bounds[r], bounds[r + 1], chunk_pairs)] for r in range(world)]
        else:
            bounds, chunks = _chunk_table(off_a, off_b, world, chunk_pairs)
        box = [(bounds, chunks)]
    if world > 1:
        dist.broadcast_object_list(box, src, group=group)
    bounds, chunks = box[0]
    mine = chunks[rank]
    first, n_local = bounds[rank], bounds[rank + 1] - bounds[rank]
    a_lo = mine[0][2] if mine else 0
    b_lo = mine[0][4] if mine else 0
    a_hi = mine[-1][3] if mine else 0
    b_hi = mine[-1][5] if mine else 0

    # this rank's shard buffers (16-byte aligned base, slack for the kernels' bulk loads)
    buf_a, buf_b = _slack_empty(a_hi - a_lo, device), _slack_empty(b_hi - b_lo, device)
    if uniform is not None:
        idx = torch.arange(n_local + 1, dtype=torch.int64, device=device)
        loc_oa, loc_ob = idx * uniform[0], idx * uniform[1]
    else:
        loc_oa = torch.empty(n_local + 1, dtype=torch.int64, device=device)
        loc_ob = torch.empty(n_local + 1, dtype=torch.int64, device=device)
    scores = torch.zeros(n_local, dtype=torch.int32, device=device)

    if cuda:
        compute = torch.cuda.current_stream()
        copy = torch.cuda.Stream(device=device)
        cstream = compute.cuda_stream
    else:
        cstream = 0

    def to_host_tensor(x):
        return torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x

    outstanding = 0

    def run_chunk(c):
        """align chunk c of this rank (its bytes are in buf_a / buf_b)"""
        nonlocal outstanding
        p0, p1 = c[0] - first, c[1] - first
        if p1 <= p0:
            return
        engine.run_device_async(algo, buf_a.data_ptr(), loc_oa.data_ptr() + 8 * p0, buf_b.data_ptr(),
                                loc_ob.data_ptr() + 8 * p0, p1 - p0, scores.data_ptr() + 4 * p0, 0, 0, cstream)
        outstanding += 1
        if outstanding > 3:
            engine.run_device_wait()
            outstanding -= 1

    def on(stream):
        return torch.cuda.stream(stream) if cuda else _null_ctx()

    if rank == src:
        ha, hb = to_host_tensor(seq_a).view(torch.uint8), to_host_tensor(seq_b).view(torch.uint8)
        hoa = to_host_tensor(off_a).to(torch.int64) if uniform is None else None
        hob = to_host_tensor(off_b).to(torch.int64) if uniform is None else None
        others = [c for r, cs in enumerate(chunks) if r != rank for c in cs]
        max_a = max([c[3] - c[2] for c in others], default=0)
        max_b = max([c[5] - c[4] for c in others], default=0)
        max_n = max([c[1] - c[0] for c in others], default=0)
        # staging ring for the chunks that travel on: a slot is reused once its sends have completed
        stage = [(torch.empty(max_a, dtype=torch.uint8, device=device), torch.empty(max_b, dtype=torch.uint8, device=device),
                  torch.empty(max_n + 1, dtype=torch.int64, device=device), torch.empty(max_n + 1, dtype=torch.int64, device=device))
                 for _ in range(ring if others else 0)]
        pending = [[] for _ in range(ring)]
        # round-robin over the ranks, so that every GPU has work from the start
        order = [(r, cs[k]) for k in range(max(len(cs) for cs in chunks)) for r, cs in enumerate(chunks) if k < len(cs)]
        sent = 0
        for r, c in order:
            p0, p1, a0, a1, b0, b1 = c
            if r == rank:
                with on(copy if cuda else None):
                    buf_a[a0 - a_lo: a1 - a_lo].copy_(ha[a0:a1], non_blocking=True)
                    buf_b[b0 - b_lo: b1 - b_lo].copy_(hb[b0:b1], non_blocking=True)
                    if uniform is None:
                        loc_oa[p0 - first: p1 - first + 1].copy_(hoa[p0:p1 + 1] - a_lo, non_blocking=True)
                        loc_ob[p0 - first: p1 - first + 1].copy_(hob[p0:p1 + 1] - b_lo, non_blocking=True)
                if cuda:
                    compute.wait_stream(copy)
                run_chunk(c)
                continue
            slot = sent % ring
            sent += 1
            with on(copy if cuda else None):
                for w in pending[slot]:
                    w.wait()
                sa, sb, soa, sob = stage[slot]
                parts = [(sa[: a1 - a0], ha[a0:a1]), (sb[: b1 - b0], hb[b0:b1])]
                if uniform is None:
                    parts += [(soa[: p1 - p0 + 1], hoa[p0:p1 + 1]), (sob[: p1 - p0 + 1], hob[p0:p1 + 1])]
                for d, h in parts:
                    d.copy_(h, non_blocking=True)
                pending[slot] = [dist.isend(d, r, group=group) for d, _ in parts]
        for ws in pending:
            for w in ws:
                w.wait()
    else:
        recv = torch.cuda.Stream(device=device) if cuda else None
        for k, c in enumerate(mine):
            p0, p1, a0, a1, b0, b1 = c
            parts = [buf_a[a0 - a_lo: a1 - a_lo], buf_b[b0 - b_lo: b1 - b_lo]]
            if uniform is None:
                tmp_oa = torch.empty(p1 - p0 + 1, dtype=torch.int64, device=device)
                tmp_ob = torch.empty(p1 - p0 + 1, dtype=torch.int64, device=device)
                parts += [tmp_oa, tmp_ob]
            # receives are queued on their own stream: chunk k+1 arrives while chunk k is being aligned
            with on(recv):
                ws = [dist.irecv(t, src, group=group) for t in parts]
            for w in ws:
                w.wait()
            if uniform is None:
                # offsets arrive absolute: rebase onto this rank's shard buffers.  Entry p0 was written by the
                # previous chunk (same value) and may be in use by its kernel: only the first chunk writes it
                lo = 0 if k == 0 else 1
                loc_oa[p0 - first + lo: p1 - first + 1] = tmp_oa[lo:] - a_lo
                loc_ob[p0 - first + lo: p1 - first + 1] = tmp_ob[lo:] - b_lo
            run_chunk(c)
    while outstanding:
        engine.run_device_wait()
        outstanding -= 1
    if cuda:
        torch.cuda.current_stream().synchronize()
    t_align = time.perf_counter()
    res = gather_results(scores[None, :], bounds, src, group)
    if cuda:
        torch.cuda.current_stream().synchronize()
    if timings is not None:
        timings.update(stream=t_align - t_start, gather=time.perf_counter() - t_align, n_local=n_local,
                       chunks=len(mine), kernel=engine.last_kernel)
    return res[0] if res is not None else None


class _null_ctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


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
