#!/bin/bash
# round 2, GPU call R (1 GPU): the flat multi-hit walk on hardware -- parity tests, fuzz of the hit mode, timing of both walks
out=gpurun_out/r02r
mkdir -p $out
timeout 600 python -m pytest tests/test_parity.py tests/test_cli_batch.py -m gpu -q -k "multi_hit or classic or golden or (batched_invocations and batching_tools and not other)" > $out/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $out/pytest.log)"
FUZZ_MODES=3 timeout 150 python tools/gpu_fuzz.py 100 501 > $out/fuzz_hits.log 2>&1; echo "fuzz rc=$? $(tail -1 $out/fuzz_hits.log | cut -c1-300)"
timeout 200 python tools/gpu_hits.py > $out/hits_flat.jsonl 2> $out/hits_flat.err; echo "flat:"; cut -c1-250 $out/hits_flat.jsonl
SEQALIGN_HITS_WALK=warp timeout 200 python tools/gpu_hits.py > $out/hits_warp.jsonl 2> $out/hits_warp.err; echo "warp:"; cut -c1-250 $out/hits_warp.jsonl
