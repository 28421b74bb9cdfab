#!/bin/bash
# round 2, GPU call Y (1 GPU): the two score/end-cell fuzzers on the final packed kernel (three-warp CTAs for protein profiles)
out=gpurun_out/r02y
mkdir -p $out
timeout 200 python tools/gpu_fuzz_ends.py 9 120 512 40 > $out/fuzz_ends.json 2> $out/fuzz_ends.err; echo "fuzz_ends rc=$? $(cut -c1-300 $out/fuzz_ends.json)"
FUZZ_MODES=0,4 timeout 200 python tools/gpu_fuzz.py 70 99 > $out/fuzz.json 2> $out/fuzz.err; echo "fuzz rc=$? $(tail -c 500 $out/fuzz.json)"
