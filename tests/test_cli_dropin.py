"""Drop-in check at the reference's own user surface (GPU only).

tests/integration/Makefile compiles the reference's UNMODIFIED command-line
tools (src/tools/{nw,sw,lcs}_cmdline.c + alignment_cmdline.c + loaders), its
unit-test program (src/tools/tests.c, 4243 assertions) and its C examples
against this repository's library; none of the reference's DP sources are in
those binaries.  Their stdout must equal, byte for byte, what the same tools
print when built with the reference's own DP (tests/golden/cli_vectors.json,
recorded by tools/gen_cli_golden.py).  The binaries are built in the build
container (they need /root/reference) and travel prebuilt to the GPU box.
"""
import json
import os
import subprocess
import tempfile

import pytest

from helpers import ROOT

CLI = os.path.join(ROOT, "tests", "integration", "_ref_cli")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "cli_vectors.json")))["cases"]

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(CLI, "needleman_wunsch")),
                                 reason="tests/integration/_ref_cli not built (needs /root/reference)")]


def _run(case):
    with tempfile.TemporaryDirectory() as td:
        args = []
        for x in case["argv"]:
            if x in case["files"]:
                path = os.path.join(td, x[1:] + ".fa")
                open(path, "w").write(case["files"][x])
                args.append(path)
            else:
                args.append(x)
        return subprocess.run([os.path.join(CLI, case["tool"])] + args, input=case["stdin"],
                              capture_output=True, text=True, timeout=300)


@pytest.mark.parametrize("i", range(len(GOLD)))
def test_reference_cli_on_our_library(i):
    case = GOLD[i]
    p = _run(case)
    assert p.returncode == case["rc"], p.stderr
    assert p.stdout == case["stdout"], (case["tool"], case["argv"])


def test_reference_unit_tests_on_our_library():
    """src/tools/tests.c, unmodified: 4 NW tests (incl. 50 random pairs) + 1 SW test"""
    p = subprocess.run([os.path.join(CLI, "seq_align_tests")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert " 0 / " in p.stdout or "0/" in p.stdout.replace(" ", ""), p.stdout[-400:]
