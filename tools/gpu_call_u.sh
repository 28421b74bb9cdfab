#!/bin/bash
# round 2, GPU call U (1 GPU): last full pass on the final commit -- smoke(), the whole GPU suite, the bench line
out=gpurun_out/r02u
mkdir -p $out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $out/smoke.log 2>&1; echo "smoke rc=$? $(grep 'smoke ok' $out/smoke.log | cut -c1-160)"
( time timeout 1500 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(grep -E 'passed|failed' $out/pytest.log | tail -1)"
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$? $(head -c 200 $out/bench_n1.json)"
timeout 200 python tools/gpu_cli_big.py 2000000 1 sw > $out/cli_big.jsonl 2> $out/cli_big.err; echo "cli rc=$?"; cut -c1-330 $out/cli_big.jsonl
