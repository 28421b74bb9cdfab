#!/bin/bash
# round 2, GPU call D (N GPUs, default 2): the N>1 bench (BASELINE config 5, strong scaling, both e2e arms), the reference arm
# as the driver launches it, the C-level multi-device entry points on N devices, the concurrent-H2D floor
N=${1:-2}
out=gpurun_out/r02d_n$N
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
( time timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 ) > $out/bench.json 2> $out/bench.err
echo "bench rc=$? $(head -c 600 $out/bench.json)"
( time timeout 300 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 ) > $out/bench_ref.json 2> $out/bench_ref.err
echo "ref rc=$? $(head -c 300 $out/bench_ref.json)"
( time timeout 300 python -m pytest tests/test_multi.py tests/test_distributed.py -m gpu -x -q ) > $out/pytest_multi.log 2>&1
echo "multi rc=$? $(tail -1 $out/pytest_multi.log)"
timeout 300 $TR tools/h2d_floor.py > $out/h2d.jsonl 2> $out/h2d.err
echo "h2d rc=$?"; tail -3 $out/h2d.jsonl | cut -c1-400
