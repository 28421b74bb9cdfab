#!/bin/bash
# round 2, GPU call X (1 GPU): packed kernel with 3 warps per CTA where that leaves more warps resident (protein profiles)
out=gpurun_out/r02x
mkdir -p $out
( time timeout 600 python -m pytest tests/test_parity.py -m gpu -q -k "relative_keys or fast16 or headline or gap_models or ragged or bucket or sweep" ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(grep -E 'passed|failed' $out/pytest.log | tail -1)"
timeout 300 python tools/gpu_perf.py > $out/perf_survey.jsonl 2> $out/perf.err; echo "perf rc=$?"; grep -E "prot400 (auto|score-only s16)|dna150 auto|dna150 score-only s16" $out/perf_survey.jsonl | cut -c1-230
SEQALIGN_FAST16_WARPS=4 timeout 300 python tools/gpu_perf.py > $out/perf_survey_4warps.jsonl 2> $out/perf4.err; echo "perf4 rc=$?"; grep -E "prot400 (auto|score-only s16)" $out/perf_survey_4warps.jsonl | cut -c1-230
