"""The counter-based generator: the CUDA kernel (csrc/sa_synth.cuh, through the C-ABI)
against its numpy mirror (seqalign/synth.py), plus the properties the benchmark relies on."""
import numpy as np
import pytest

import seqalign
from seqalign.synth import DNA, PROTEIN, synth_batch, splitmix64


def test_splitmix64_known_answers():
    # first outputs of the splitmix64 stream seeded with 0 and 1234567 (Vigna's reference values)
    assert int(splitmix64(np.uint64(0))) == 0xE220A8397B1DCDAF
    assert int(splitmix64(np.uint64(0x9E3779B97F4A7C15))) == 0x6E789E6AA1B965F4
    assert int(splitmix64(np.uint64(1234567))) == 6457827717110365317


def test_shards_are_windows_of_one_stream():
    a, _, b, _ = synth_batch(5, 0, 64, 30, 41)
    a2, _, b2, _ = synth_batch(5, 17, 20, 30, 41)
    assert np.array_equal(a.reshape(64, 30)[17:37], a2.reshape(20, 30))
    assert np.array_equal(b.reshape(64, 41)[17:37], b2.reshape(20, 41))
    a3, _, _, _ = synth_batch(6, 0, 64, 30, 41)
    assert not np.array_equal(a, a3)


def test_mutation_rates():
    a, _, b, _ = synth_batch(2, 0, 4000, 150, 150)
    assert set(np.unique(a)) <= set(DNA) and set(np.unique(b)) <= set(DNA)
    # column identity decays with position as indels shift the frame; the first columns see only substitutions
    first = (a.reshape(-1, 150)[:, :5] == b.reshape(-1, 150)[:, :5]).mean()
    assert 0.90 < first < 0.97
    pa, _, pb, _ = synth_batch(4, 0, 500, 60, 60, "protein")
    assert set(np.unique(pa)) <= set(PROTEIN) and len(np.unique(pa)) == 20


@pytest.mark.parity
@pytest.mark.parametrize("kind,la,lb,first", [("dna", 150, 150, 0), ("dna", 37, 61, 12345678), ("protein", 40, 33, 7), ("dna", 5, 40, 3)])
def test_device_generator_matches_numpy(engine, backend, kind, la, lb, first):
    n = 257
    a, _, b, _ = synth_batch(5, first, n, la, lb, kind)
    if backend == "gpu":
        import torch
        da = torch.zeros(n * la, dtype=torch.uint8, device="cuda:0")
        db = torch.zeros(n * lb, dtype=torch.uint8, device="cuda:0")
        seqalign.synth_device(0, kind, 5, first, n, la, lb, da.data_ptr(), db.data_ptr())
        da, db = da.cpu().numpy(), db.cpu().numpy()
    else:
        da, db = np.zeros(n * la, np.uint8), np.zeros(n * lb, np.uint8)
        seqalign.synth_device(0, kind, 5, first, n, la, lb, da.ctypes.data, db.ctypes.data)
    assert np.array_equal(a, da) and np.array_equal(b, db)
