#!/bin/bash
# round 2, GPU call M (1 GPU): full GPU suite on the final code, fuzz (wide SW traceback included), perf survey, bench both arms, config 3
out=gpurun_out/r02m
mkdir -p $out
( time timeout 1200 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $out/pytest.log)"
timeout 200 python tools/gpu_fuzz.py 150 401 > $out/fuzz_401.log 2>&1; echo "fuzz rc=$? $(tail -1 $out/fuzz_401.log | cut -c1-900)"
timeout 500 python tools/gpu_perf.py > $out/perf.jsonl 2> $out/perf.err; echo "perf rc=$?"; grep -i "wide" $out/perf.jsonl | cut -c1-300
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$? $(head -c 200 $out/bench_n1.json)"
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > $out/bench_ref.json 2> $out/bench_ref.err
echo "ref rc=$? $(head -c 200 $out/bench_ref.json)"
( time timeout 300 python tools/gpu_config3.py ) > $out/config3.jsonl 2> $out/config3.err; echo "config3 rc=$? $(head -c 700 $out/config3.jsonl)"
timeout 120 python tools/gpu_cli.py > $out/cli.log 2>&1; echo "cli rc=$?"; grep -o '"tool[^}]*seconds": [0-9.]*' $out/cli.log | head -5
