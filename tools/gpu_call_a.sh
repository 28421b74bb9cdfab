#!/bin/bash
# round 2, GPU call A (1 GPU): tests, bench (N=1 incl. config5 side record), config 3 at full size, sanitizers, launch list
out=gpurun_out/r02a
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $out/gpu.txt 2>&1
nproc > $out/host.txt; lscpu | grep -E "Model name|Socket|NUMA node\(s\)" >> $out/host.txt; free -g | sed -n 2p >> $out/host.txt
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $out/pytest.log)"
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$? $(head -c 300 $out/bench_n1.json)"
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > $out/bench_ref.json 2> $out/bench_ref.err
( time timeout 400 python tools/gpu_config3.py ) > $out/config3.jsonl 2> $out/config3.err
echo "config3 rc=$? $(head -c 600 $out/config3.jsonl)"
timeout 200 python tools/h2d_floor.py > $out/h2d_n1.jsonl 2> $out/h2d_n1.err
SAN_TIMEOUT=300 tools/gpu_sanitize.sh $out/san > $out/san.log 2>&1
cat $out/san/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config5 --sustain 0.05 > $out/ncu_bench.log 2>&1
echo "ncu rc=$?"
