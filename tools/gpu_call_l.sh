#!/bin/bash
# round 2, GPU call L (1 GPU): smoke(), ragged throughput in the three bucket modes, differential fuzz of every mode (also with
# tiny batches bucketed), materialise timings (SW / NW, packed NW scans), hits timing, the failing wave test again
out=gpurun_out/r02l
mkdir -p $out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $out/smoke.log 2>&1; echo "smoke rc=$? $(tail -2 $out/smoke.log | head -1)"
for m in 0 rows shapes; do SEQALIGN_BUCKETS=$m timeout 200 python tools/gpu_ragged.py 100000 2> $out/ragged_$m.err | head -3 > $out/ragged_$m.jsonl; echo "buckets=$m"; cut -c50-260 $out/ragged_$m.jsonl; done
timeout 200 python -m pytest tests/test_parity.py -m gpu -q -k "batch_matrices or length_buckets" > $out/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $out/pytest.log)"
timeout 260 python tools/gpu_fuzz.py 200 301 > $out/fuzz_301.log 2>&1; echo "fuzz rc=$? $(tail -1 $out/fuzz_301.log | cut -c1-700)"
SEQALIGN_BUCKETS=shapes SEQALIGN_BUCKET_MIN=8 FUZZ_MODES=0,2 timeout 120 python tools/gpu_fuzz.py 60 302 > $out/fuzz_302_buckets.log 2>&1; echo "fuzz buckets rc=$? $(tail -1 $out/fuzz_302_buckets.log | cut -c1-400)"
timeout 200 python tools/gpu_mats.py > $out/mats.log 2>&1; echo "mats rc=$?"; tail -4 $out/mats.log | cut -c1-300
timeout 200 python tools/gpu_mats_nw.py > $out/mats_nw.log 2>&1; echo "mats_nw rc=$?"; tail -6 $out/mats_nw.log | cut -c1-300
