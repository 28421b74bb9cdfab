#!/bin/bash
# round 2, GPU call C (1 GPU): full GPU suite after the checkpoint alignment fix, config 3 through checkpoints
out=gpurun_out/r02c
mkdir -p $out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $out/pytest.log)"
( time timeout 400 python tools/gpu_config3.py ) > $out/config3.jsonl 2> $out/config3.err
echo "config3 rc=$? $(head -c 1200 $out/config3.jsonl)"
