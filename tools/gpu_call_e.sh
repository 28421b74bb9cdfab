#!/bin/bash
# round 2, GPU call E (1 GPU): full GPU suite (decoder, reference mains, pad-row change), decode throughput, bench
out=gpurun_out/r02e
mkdir -p $out
( time timeout 900 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $out/pytest.log)"
( time timeout 300 python tools/gpu_decode.py 500000 ) > $out/decode.jsonl 2> $out/decode.err
echo "decode rc=$?"; cut -c1-600 $out/decode.jsonl
( time timeout 400 python bench.py --steps 20 --warmup 3 --no-config5 ) > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$? $(head -c 300 $out/bench_n1.json)"
timeout 200 python tools/gpu_ragged.py > $out/ragged.jsonl 2> $out/ragged.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file $out/launches_decode.csv python tools/gpu_decode.py 200000 > $out/ncu_decode.log 2>&1
echo "ncu rc=$?"
