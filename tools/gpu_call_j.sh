#!/bin/bash
# round 2, GPU call J (1 GPU): ragged throughput with device-side bucket ranges; full GPU suite
out=gpurun_out/r02j
mkdir -p $out
timeout 200 python tools/gpu_ragged.py > $out/ragged.jsonl 2> $out/ragged.err; echo "ragged rc=$?"; cat $out/ragged.jsonl
SEQALIGN_NO_BUCKETS=1 timeout 200 python tools/gpu_ragged.py 100000 > $out/ragged_nobuckets.jsonl 2> $out/ragged_nobuckets.err; head -3 $out/ragged_nobuckets.jsonl
( time timeout 1200 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $out/pytest.log)"
