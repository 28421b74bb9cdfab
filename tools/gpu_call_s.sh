#!/bin/bash
# round 2, GPU call S (1 GPU): flat multi-hit walk with four candidates per scan pass -- parity, fuzz, timing
out=gpurun_out/r02s
mkdir -p $out
timeout 600 python -m pytest tests/test_parity.py -m gpu -q -k "multi_hit or classic or golden" > $out/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $out/pytest.log)"
FUZZ_MODES=3 timeout 150 python tools/gpu_fuzz.py 100 601 > $out/fuzz_hits.log 2>&1; echo "fuzz rc=$? $(tail -1 $out/fuzz_hits.log | cut -c1-300)"
timeout 200 python tools/gpu_hits.py > $out/hits_flat4.jsonl 2> $out/hits_flat4.err; cut -c1-250 $out/hits_flat4.jsonl
timeout 120 python tools/gpu_cli.py > $out/cli.log 2>&1; grep -o '"tool[^}]*total [0-9.]* s' $out/cli.log | cut -c1-330
