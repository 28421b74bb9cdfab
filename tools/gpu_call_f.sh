#!/bin/bash
# round 2, GPU call F (1 GPU): decoder + CLI tests on hardware after the rebuild, full suite, large-file CLI run, decode ncu list
out=gpurun_out/r02f
mkdir -p $out
( time timeout 900 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $out/pytest.log)"
( time timeout 600 python tools/gpu_cli_big.py 2000000 1 ) > $out/cli_big.jsonl 2> $out/cli_big.err
echo "cli_big rc=$?"; cut -c1-500 $out/cli_big.jsonl
( time timeout 300 python tools/gpu_decode.py 500000 ) > $out/decode.jsonl 2> $out/decode.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file $out/launches_decode.csv python tools/gpu_decode.py 200000 > $out/ncu_decode.log 2>&1
echo "ncu rc=$?"
