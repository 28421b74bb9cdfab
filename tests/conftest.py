"""pytest configuration.

Markers
  gpu     needs a B200: runs the product library (libseqalign_b200.so) on cuda:0
  parity  CUDA-path-vs-oracle tests.  They are written once and run on one of
          two backends, chosen per session:
            * a CUDA device is visible  -> the real library; the tests are
              marked `gpu` (so `-m gpu` selects them on the GPU box);
            * no device (build container) -> the same kernel/engine sources
              compiled against the lane emulator in tests/emu (small sizes),
              not marked `gpu`, so `-m "not gpu"` exercises the kernel logic.
          SEQALIGN_TEST_BACKEND=gpu|emu overrides the choice.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "seq-align_b200"))


def _cuda_visible():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _backend():
    b = os.environ.get("SEQALIGN_TEST_BACKEND")
    if b in ("gpu", "emu"):
        return b
    return "gpu" if _cuda_visible() else "emu"


BACKEND = _backend()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
    config.addinivalue_line("markers", "parity: CUDA path vs oracle; gpu-marked when a device is visible")
    if BACKEND == "emu":
        emu = os.path.join(ROOT, "tests", "emu", "libseqalign_emu.so")
        subprocess.check_call(["make", "-s", "-C", ROOT, "emu"])
        os.environ["SEQALIGN_LIB"] = emu
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])


def pytest_collection_modifyitems(config, items):
    for item in items:
        if item.get_closest_marker("parity") and BACKEND == "gpu":
            item.add_marker(pytest.mark.gpu)


@pytest.fixture(scope="session")
def backend():
    return BACKEND


@pytest.fixture(scope="session")
def big(backend):
    """True on the GPU: full-size batches; False in the emulator: small ones.
    SEQALIGN_TEST_SMALL=1 keeps the small sizes on the GPU too (runs under compute-sanitizer)."""
    return backend == "gpu" and os.environ.get("SEQALIGN_TEST_SMALL") != "1"


@pytest.fixture(scope="session")
def engine(backend):
    import seqalign
    eng = seqalign.BatchAligner(0)
    yield eng
    eng.close()
