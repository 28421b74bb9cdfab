"""The batching command-line tools (seq-align_b200/tools -> bin/needleman_wunsch,
bin/smith_waterman, bin/lcs) against stdout recorded from the reference's own tools.

Two golden sets:
  tests/golden/cli_vectors.json        single-pair invocations (tools/gen_cli_golden.py)
  tests/golden/cli_batch_vectors.json  multi-pair files in every input format, scoring files,
                                       stdin protocols, error cases (tools/gen_cli_batch_golden.py)
Backends as for the parity tests: on a GPU box the real binaries in bin/ (gpu-marked); in the
build container the same tool sources linked against the lane emulator (tests/emu/bin).
"""
import gzip
import json
import os
import subprocess
import tempfile

import pytest

from conftest import BACKEND
from helpers import ROOT

GOLD1 = json.load(open(os.path.join(ROOT, "tests", "golden", "cli_vectors.json")))["cases"]
GOLD2 = json.load(open(os.path.join(ROOT, "tests", "golden", "cli_batch_vectors.json")))["cases"]
GOLD1 = [c for c in GOLD1 if c["tool"] in ("needleman_wunsch", "smith_waterman", "lcs")]

pytestmark = [pytest.mark.parity]


REFMAIN = os.path.join(ROOT, "tests", "integration", "_ref_main" if BACKEND == "gpu" else "_ref_main_emu")


@pytest.fixture(scope="module", params=["batching_tools", "reference_mains"])
def tool_dir(request):
    """batching_tools: this repository's tools.  reference_mains: ONLY the reference's own
    src/tools/{nw,sw,lcs}_cmdline.c mains, linked against libalign.a (GPU) / the emulator build
    alone -- cmdline_new, align_from_file and the scoring loaders come from the library
    (tests/integration/Makefile; built where /root/reference exists, travels prebuilt)."""
    if request.param == "reference_mains":
        if os.path.isdir("/root/reference/src"):
            if BACKEND != "gpu":
                subprocess.check_call(["make", "-s", "-C", ROOT, "emu"])
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "integration")], stdout=subprocess.DEVNULL)
        if not os.path.exists(os.path.join(REFMAIN, "needleman_wunsch")):
            pytest.skip("tests/integration/%s not built (needs /root/reference)" % os.path.basename(REFMAIN))
        return REFMAIN
    if BACKEND == "gpu":
        d = os.path.join(ROOT, "bin")
        if not os.path.exists(os.path.join(d, "needleman_wunsch")):
            subprocess.check_call(["make", "-s", "-C", ROOT, "tools"])
        return d
    subprocess.check_call(["make", "-s", "-C", ROOT, "emu-tools"])
    return os.path.join(ROOT, "tests", "emu", "bin")


def _run(tool_dir, case, env=None):
    with tempfile.TemporaryDirectory() as td:
        args = []
        for x in case["argv"]:
            if x in case["files"]:
                f = case["files"][x]
                if isinstance(f, str):
                    f = dict(text=f, gz=False)
                path = os.path.join(td, x[1:] + (".gz" if f["gz"] else ".txt"))
                if f["gz"]:
                    with gzip.open(path, "wt") as fh:
                        fh.write(f["text"])
                else:
                    with open(path, "w", newline="") as fh:
                        fh.write(f["text"])
                args.append(path)
            else:
                args.append(x)
        p = subprocess.run([os.path.join(tool_dir, case["tool"])] + args, input=case["stdin"],
                           capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
        return p.returncode, p.stdout, p.stderr.replace(td, "<tmp>")


@pytest.mark.parametrize("i", range(len(GOLD1)))
def test_single_pair_invocations(tool_dir, i):
    case = GOLD1[i]
    rc, out, err = _run(tool_dir, case)
    assert rc == case["rc"], err
    assert out == case["stdout"], (case["tool"], case["argv"])


@pytest.mark.parametrize("i", range(len(GOLD2)))
def test_batched_invocations(tool_dir, i):
    case = GOLD2[i]
    rc, out, err = _run(tool_dir, case)
    assert rc == case["rc"], (case["argv"], err)
    if case["rc"] == 0 or not case["stderr_first"].startswith("Error: "):
        assert out == case["stdout"], (case["tool"], case["argv"])
    if case["stderr_first"] is not None:
        # usage errors print the usage text on stdout/stderr in their own words; the message is the contract
        assert err.split("\n")[0] == case["stderr_first"], (case["argv"], err[:300])


# how the tools read files: records decoded on the device from whole chunks of text (default; 64 MB chunks),
# the same with chunks so small that every record straddles one (carry + buffer growth), the host reader only
READERS = {"device": {}, "device_tiny_chunks": {"SEQALIGN_CLI_CHUNK_MB": "-48"}, "host": {"SEQALIGN_CLI_DECODE": "host"},
           # --gpus 3 (the flag is added to argv below): batches cut over three engines -- the box's GPUs, or three
           # engines on the one (emulated) device
           "three_engines": {"SEQALIGN_CLI_DEVICES": "0,0,0"}, "three_engines_host_reader": {"SEQALIGN_CLI_DEVICES": "0,0,0", "SEQALIGN_CLI_DECODE": "host"},
           # the print loop's buffer starts at 16 bytes: every sync / grow path of tools/sa_batch.h sa_out_*
           "tiny_outbuf": {"SEQALIGN_CLI_OUTBUF": "16"}}


@pytest.mark.parametrize("reader", ["device_tiny_chunks", "host", "three_engines", "three_engines_host_reader", "tiny_outbuf"])
@pytest.mark.parametrize("i", range(len(GOLD2)))
def test_batched_invocations_other_readers(tool_dir, i, reader, backend):
    """the recorded multi-pair invocations again, through the other two ways of reading the input"""
    if tool_dir == REFMAIN:
        pytest.skip("the reference's mains read through align_from_file only")
    if reader == "tiny_outbuf" and backend == "gpu":
        pytest.skip("host-side buffer logic: covered in the build container, not worth 54 more processes on the GPU box")
    case = GOLD2[i]
    if reader.startswith("three_engines"):
        if case["tool"] == "lcs":
            pytest.skip("lcs takes no options")
        case = dict(case, argv=["--gpus", "3"] + case["argv"])
    rc, out, err = _run(tool_dir, case, READERS[reader])
    assert rc == case["rc"], (case["argv"], err)
    if case["rc"] == 0 or not case["stderr_first"].startswith("Error: "):
        assert out == case["stdout"], (case["tool"], case["argv"])
    if case["stderr_first"] is not None:
        assert err.split("\n")[0] == case["stderr_first"], (case["argv"], err[:300])


def test_device_reader_is_the_one_that_runs(tool_dir):
    """SEQALIGN_CLI_DECODE=device turns a decline into an error: a FASTA file must not need the host reader,
    a wrapped FASTQ file must"""
    if tool_dir == REFMAIN:
        pytest.skip("batching tools only")
    with tempfile.TemporaryDirectory() as td:
        fa, fq = os.path.join(td, "x.fa"), os.path.join(td, "x.fq")
        open(fa, "w").write(">a\nACGTACGT\n>b\nACGAACGT\n>c\nTTTT\n>d\nTTAT\n")
        open(fq, "w").write("@a\nACGT\nACGT\n+\nIIII\nIIII\n@b\nACGAACGT\n+\nIIIIIIII\n")
        env = dict(os.environ, SEQALIGN_CLI_DECODE="device")
        exe = os.path.join(tool_dir, "needleman_wunsch")
        p = subprocess.run([exe, "--printscores", "--file", fa], capture_output=True, text=True, env=env, timeout=300)
        assert p.returncode == 0 and p.stdout.count("score:") == 2, p.stderr
        q = subprocess.run([exe, "--printscores", "--file", fq], capture_output=True, text=True, env=env, timeout=300)
        assert q.returncode != 0 and "declined" in q.stderr
        r = subprocess.run([exe, "--printscores", "--file", fq], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and r.stdout.count("score:") == 1 and "ACGTACGT" in r.stdout


def test_interactive_smith_waterman_prompt(tool_dir):
    """--stdin: hits are handed out one keystroke at a time (reference sw_cmdline.c:84-122);
    'h' = next hit, 'a' = next alignment, EOF ends the session"""
    if tool_dir == REFMAIN:
        # the reference's main reads its keystrokes with getc(stdin) while the records come through raw
        # read()s of fd 0: on a pipe its stdio buffer swallows the records that follow (its own binary
        # prints one prompt for this input).  Only a terminal drives that main; nothing to compare.
        pytest.skip("the reference's main needs a terminal for its prompt")
    stdin = "ACGTACGTTTGACCA\nTTACGTACGAAGACC\nh\nh\na\ngacag\ntgaagt\nh\n"
    p = subprocess.run([os.path.join(tool_dir, "smith_waterman"), "--stdin"], input=stdin, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0
    out = p.stdout
    assert out.count("next [h]it or [a]lignment: ") == 5
    assert "== Alignment 0 lengths (15, 15):" in out and "== Alignment 1 lengths (5, 6):" in out
    assert "hit 0.0 score:" in out and "hit 0.1 score:" in out and "hit 1.0 score:" in out
