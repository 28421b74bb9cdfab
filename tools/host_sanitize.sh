#!/bin/bash
# ASan + UBSan over the host C of the library and the command-line tools.
# Everything -- host/*.c, tools/*.c, and the engine + decoder compiled against the lane emulator so that
# the tools run without a GPU -- is built with -fsanitize=address,undefined into tests/emu/bin_asan/, and
# every recorded CLI invocation (tests/golden/cli_vectors.json, cli_batch_vectors.json) is replayed through
# those binaries with each way of reading the input.  Any sanitizer report fails the run.
#   tools/host_sanitize.sh [outfile]        (CPU only; system gcc: the one with libasan / libubsan)
set -e
cd "$(dirname "$0")/.."
out=${1:-profiles/sanitizer_host_r02.txt}
CC=/usr/bin/gcc; CXX=/usr/bin/g++
B=tests/emu/bin_asan; mkdir -p $B
SAN="-fsanitize=address,undefined -fno-omit-frame-pointer -g -O1"
INC="-Iinclude -Iseq-align_b200/host -Iseq-align_b200/csrc -Itests/emu -Iseq-align_b200/tools"
for f in sa_engine sa_decode; do $CXX $SAN -std=c++17 -DSA_EMU $INC -c -x c++ seq-align_b200/csrc/$f.cu -o $B/$f.o; done
$CXX $SAN -std=c++17 -Itests/emu -c tests/emu/cuda_emu.cpp -o $B/cuda_emu.o
for f in seq-align_b200/host/*.c; do $CC $SAN -std=gnu99 $INC -c $f -o $B/$(basename $f .c).o; done
LIBO="$B/sa_engine.o $B/sa_decode.o $B/cuda_emu.o $B/sa_scoring.o $B/sa_alignment.o $B/sa_nw.o $B/sa_sw.o $B/sa_multi.o $B/sa_cli.o $B/sa_cmdline.o"
for t in nw:needleman_wunsch sw:smith_waterman lcs:lcs; do
  $CC $SAN -std=gnu99 $INC -c seq-align_b200/tools/${t%%:*}_main.c -o $B/${t%%:*}_main.o
  $CXX $SAN -o $B/${t##*:} $B/${t%%:*}_main.o $LIBO -lz -lpthread
done
export ASAN_OPTIONS=detect_leaks=1:abort_on_error=0:exitcode=97 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1:exitcode=98
python - "$B" > "$out" <<'PY'
import gzip, json, os, subprocess, sys, tempfile
B = sys.argv[1]
cases = [c for c in json.load(open("tests/golden/cli_vectors.json"))["cases"] if c["tool"] in ("needleman_wunsch", "smith_waterman", "lcs")]
cases += json.load(open("tests/golden/cli_batch_vectors.json"))["cases"]
readers = {"device": {}, "device_tiny_chunks": {"SEQALIGN_CLI_CHUNK_MB": "-48"}, "host": {"SEQALIGN_CLI_DECODE": "host"},
           "three_engines": {"SEQALIGN_CLI_DEVICES": "0,0,0"}}
runs = bad = mism = 0
for rname, env in readers.items():
    for c in cases:
        argv = list(c["argv"])
        if rname == "three_engines":
            if c["tool"] == "lcs": continue
            argv = ["--gpus", "3"] + argv
        with tempfile.TemporaryDirectory() as td:
            args = []
            for x in argv:
                if x in c["files"]:
                    f = c["files"][x]
                    if isinstance(f, str): f = dict(text=f, gz=False)
                    path = os.path.join(td, x[1:] + (".gz" if f["gz"] else ".txt"))
                    (gzip.open(path, "wt") if f["gz"] else open(path, "w", newline="")).write(f["text"])
                    args.append(path)
                else:
                    args.append(x)
            p = subprocess.run([os.path.join(B, c["tool"])] + args, input=c["stdin"], capture_output=True, text=True, timeout=600,
                               env=dict(os.environ, **env))
        runs += 1
        if p.returncode in (97, 98) or "AddressSanitizer" in p.stderr or "runtime error" in p.stderr or "LeakSanitizer" in p.stderr:
            bad += 1
            print("SANITIZER REPORT [%s] %s %s\n%s" % (rname, c["tool"], argv, p.stderr[-3000:]))
        elif p.returncode == c["rc"] and c["rc"] == 0 and p.stdout != c["stdout"]:
            mism += 1
            print("STDOUT MISMATCH [%s] %s %s" % (rname, c["tool"], argv))
print("host sanitizers (ASan + UBSan, leak detection on): %d invocations of the tools (%d recorded cases x %d readers), "
      "%d with a sanitizer report, %d with other stdout than recorded" % (runs, len(cases), len(readers), bad, mism))
sys.exit(1 if bad or mism else 0)
PY
rm -rf $B   # 200 MB of instrumented binaries: the report is what stays
tail -1 "$out"
