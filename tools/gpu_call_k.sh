#!/bin/bash
# round 2, GPU call K (8 GPUs, charged 8x: keep it short): BASELINE config 5 on 8 x B200 (bench.py as the driver launches
# it), the concurrent-H2D floor, and a 2 M-pair FASTA through `smith_waterman --maxhits 1 --gpus 8` against 1 GPU
N=${1:-8}
out=gpurun_out/r02k_n$N
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1; nproc > $out/host.txt; lscpu | grep -E "Model name|Socket|NUMA node" >> $out/host.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
( time timeout 420 $TR bench.py --gpus $N --steps 5 --warmup 3 ) > $out/bench.json 2> $out/bench.err
echo "bench rc=$? $(head -c 400 $out/bench.json)"
timeout 240 $TR tools/h2d_floor.py > $out/h2d.jsonl 2> $out/h2d.err
echo "h2d rc=$?"; grep '"pinned h2d"' $out/h2d.jsonl | tail -4 | cut -c1-300
( time timeout 300 python tools/gpu_cli_big.py 2000000 $N sw ) > $out/cli_big.jsonl 2> $out/cli_big.err
echo "cli rc=$?"; cut -c1-420 $out/cli_big.jsonl
