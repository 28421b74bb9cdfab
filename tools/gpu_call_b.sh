#!/bin/bash
# round 2, GPU call B (1 GPU): tests after the CKPT traceback / ragged fast16 commits, bench, config 3, perf survey, sanitizers
out=gpurun_out/r02b
mkdir -p $out
( time timeout 700 python -m pytest tests -m gpu -x -q ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $out/pytest.log)"
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$? $(head -c 300 $out/bench_n1.json)"
( time timeout 400 python tools/gpu_config3.py ) > $out/config3.jsonl 2> $out/config3.err
echo "config3 rc=$? $(head -c 900 $out/config3.jsonl)"
SEQALIGN_LONG_FLAGS=1 timeout 400 python tools/gpu_config3.py 2000 10000 4 > $out/config3_flags.jsonl 2> $out/config3_flags.err
timeout 200 python tools/gpu_ragged.py > $out/ragged.jsonl 2> $out/ragged.err
echo "ragged rc=$?"; cat $out/ragged.jsonl
timeout 400 python tools/gpu_perf.py > $out/perf.jsonl 2> $out/perf.err
echo "perf rc=$?"
SAN_TIMEOUT=300 tools/gpu_sanitize.sh $out/san > $out/san.log 2>&1
cat $out/san/summary.txt
