"""world_size-2 sharding over torch.distributed (gloo, CPU): bounds, scatter,
local alignment, gather.  The engine behind it is the parity backend of the
session (lane emulator here; on the GPU box the NCCL path is exercised by
bench.py --gpus N and tests/test_parity.py::test_device_resident_api)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from helpers import ROOT, SPECS, orc_batch_sw, orc_from_scoring, ragged_batch, scoring_from_spec

pytestmark = pytest.mark.parity


def test_shard_bounds_balanced():
    sys.path.insert(0, os.path.join(ROOT, "seq-align_b200"))
    from seqalign.distributed import shard_bounds
    off = np.arange(0, 1001 * 150, 150)
    assert shard_bounds(off, off, 4) == [0, 250, 500, 750, 1000]
    assert shard_bounds(off[:1], off[:1], 3) == [0, 0, 0, 0]
    # ragged: heavier pairs at the end -> later shards get fewer pairs
    la = np.concatenate([np.full(100, 10), np.full(100, 100)])
    off = np.concatenate([[0], np.cumsum(la)])
    b = shard_bounds(off, off, 2)
    assert b[0] == 0 and b[2] == 200 and 140 < b[1] < 160
    cells = la * la
    assert abs(cells[:b[1]].sum() - cells[b[1]:].sum()) < 0.05 * cells.sum()
    # the tensor version (used by scatter_pairs, runs where the offsets live) cuts at the same places
    from seqalign.distributed import shard_bounds_tensor
    rng = np.random.default_rng(5)
    for world in (1, 2, 3, 8):
        la = rng.integers(0, 300, 5000); lb = rng.integers(0, 300, 5000)
        oa = np.concatenate([[0], np.cumsum(la)]); ob = np.concatenate([[0], np.cumsum(lb)])
        assert shard_bounds_tensor(torch.from_numpy(oa), torch.from_numpy(ob), world) == shard_bounds(oa, ob, world)
    assert shard_bounds_tensor(torch.zeros(1, dtype=torch.int64), torch.zeros(1, dtype=torch.int64), 3) == [0, 0, 0, 0]


def _worker(rank, world, port, lib, seed, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if lib:
        os.environ["SEQALIGN_LIB"] = lib
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "seq-align_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import seqalign
    from seqalign.distributed import align_sharded
    eng = seqalign.BatchAligner(0, seqalign.Scoring.sw_cli_default())
    if rank == 0:
        sa, sb = ragged_batch(seed, 37, 50, 50)
        a, oa = seqalign.pack(sa)
        b, ob = seqalign.pack(sb)
        res = align_sharded(eng, seqalign.SW, a, oa, b, ob, src=0)
        np.save(out_path, res.numpy())
    else:
        assert align_sharded(eng, seqalign.SW, src=0) is None
    eng.close()
    dist.destroy_process_group()


def test_align_sharded_world2(backend, tmp_path):
    if backend == "gpu":
        pytest.skip("CPU/gloo logic test; the GPU box runs the NCCL path through bench.py --gpus N")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, port, os.environ.get("SEQALIGN_LIB", ""), 321, out), nprocs=2, join=True)
    got = np.load(out)
    import seqalign
    sa, sb = ragged_batch(321, 37, 50, 50)
    a, oa = seqalign.pack(sa)
    b, ob = seqalign.pack(sb)
    es, ex, ey = orc_batch_sw(orc_from_scoring(scoring_from_spec(SPECS["sw_cli"])), a, oa, b, ob)
    assert np.array_equal(got[0], es) and np.array_equal(got[1], ex) and np.array_equal(got[2], ey)


def _stream_worker(rank, world, port, lib, seed, uniform, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if lib:
        os.environ["SEQALIGN_LIB"] = lib
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "seq-align_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import seqalign
    from seqalign.distributed import align_sharded_stream
    eng = seqalign.BatchAligner(0, seqalign.Scoring.sw_cli_default())
    a, oa, b, ob = _stream_job(seed, uniform)
    src = world - 1          # not rank 0: the source is a parameter, not an assumption
    if rank == src:
        res = align_sharded_stream(eng, seqalign.SW, a, oa, b, ob, src=src, chunk_pairs=7,
                                   uniform=(24, 30) if uniform else None, ring=2)
        np.save(out_path, res.numpy())
    else:
        assert align_sharded_stream(eng, seqalign.SW, src=src, chunk_pairs=7,
                                    uniform=(24, 30) if uniform else None, ring=2) is None
    eng.close()
    dist.destroy_process_group()


def _stream_job(seed, uniform):
    import seqalign
    if uniform:
        from seqalign.synth import synth_batch
        return synth_batch(seed, 1000, 45, 24, 30)
    sa, sb = ragged_batch(seed, 45, 40, 40)
    a, oa = seqalign.pack(sa)
    b, ob = seqalign.pack(sb)
    return a, oa, b, ob


@pytest.mark.parametrize("uniform", [False, True])
def test_align_sharded_stream_world2(backend, tmp_path, uniform):
    """chunk-overlapped scatter from host memory (the BASELINE config 5 data plane): chunks travel
    rank by rank while earlier ones are aligned; scores come back in pair order"""
    if backend == "gpu":
        pytest.skip("CPU/gloo logic test; the GPU box runs the NCCL path through bench.py --gpus N")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "res.npy")
    mp.spawn(_stream_worker, args=(2, port, os.environ.get("SEQALIGN_LIB", ""), 99, uniform, out), nprocs=2, join=True)
    got = np.load(out)
    a, oa, b, ob = _stream_job(99, uniform)
    es, _, _ = orc_batch_sw(orc_from_scoring(scoring_from_spec(SPECS["sw_cli"])), a, oa, b, ob)
    assert np.array_equal(got, es)
