#!/bin/bash
# round 2, GPU call Q (1 GPU): the final code -- smoke(), full GPU suite, bench both arms, ragged check, tools on a 2 M-pair file
out=gpurun_out/r02q
mkdir -p $out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $out/smoke.log 2>&1; echo "smoke rc=$? $(grep 'smoke ok' $out/smoke.log | cut -c1-160)"
( time timeout 1500 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(grep -E 'passed|failed' $out/pytest.log | tail -1)"
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$? $(head -c 200 $out/bench_n1.json)"
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > $out/bench_ref.json 2> $out/bench_ref.err
echo "ref rc=$? $(head -c 160 $out/bench_ref.json)"
timeout 200 python tools/gpu_ragged.py 100000 2> $out/ragged.err | head -3 > $out/ragged.jsonl; cut -c50-260 $out/ragged.jsonl
timeout 300 python tools/gpu_cli_big.py 2000000 1 > $out/cli_big.jsonl 2> $out/cli_big.err; echo "cli rc=$?"; cut -c1-330 $out/cli_big.jsonl
