#!/bin/bash
# round 2, GPU call O (1 GPU): one ncu --set full capture over every kernel family at its BASELINE shape (round-2 kernels
# included), and the GPU suite + bench on the final code
out=gpurun_out/r02o
mkdir -p $out
SEQALIGN_CHUNKS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fast|long|walk|hits|mats|emit|classify|nl_|scan3|contrib" -c 60 -o $out/ncu_all python tools/gpu_prof_all.py > $out/ncu_all.log 2>&1
echo "ncu rc=$?"; tail -12 $out/ncu_all.log | cut -c1-200; ls -la $out/*.ncu-rep
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(tail -4 $out/pytest.log | head -1)"
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$? $(head -c 200 $out/bench_n1.json)"
