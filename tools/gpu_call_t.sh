#!/bin/bash
# round 2, GPU call T (2 GPUs): the final code through bench.py exactly as the driver launches it at N = 2 (both arms)
N=${1:-2}
out=gpurun_out/r02t_n$N
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
( time timeout 300 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 ) > $out/bench_ref.json 2> $out/bench_ref.err
echo "ref rc=$? $(head -c 200 $out/bench_ref.json)"
( time timeout 500 $TR bench.py --gpus $N --steps 5 --warmup 3 ) > $out/bench.json 2> $out/bench.err
echo "bench rc=$? $(head -c 300 $out/bench.json)"
timeout 200 python -m pytest tests/test_multi.py -m gpu -q > $out/pytest_multi.log 2>&1; echo "multi rc=$? $(tail -1 $out/pytest_multi.log)"
