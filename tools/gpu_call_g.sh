#!/bin/bash
# round 2, GPU call G (1 GPU): full GPU suite after the opt-in fix + single-pair API change; classic-API timings; sanitizers
out=gpurun_out/r02g
mkdir -p $out
( time timeout 1200 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $out/pytest.log)"
( time timeout 300 python tools/gpu_classic.py ) > $out/classic.jsonl 2> $out/classic.err
echo "classic rc=$?"; cat $out/classic.jsonl
SAN_TIMEOUT=400 tools/gpu_sanitize.sh $out/san > $out/san.log 2>&1
cat $out/san/summary.txt
