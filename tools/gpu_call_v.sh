#!/bin/bash
# round 2, GPU call V (1 GPU): packed end cells with relative keys (scores past 1024) -- parity subset, two fuzzers, the survey, the bench line
out=gpurun_out/r02v
mkdir -p $out
( time timeout 600 python -m pytest tests/test_parity.py -m gpu -q -k "relative_keys or fast16 or headline or gap_models or ragged or bucket or sweep" ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(grep -E 'passed|failed' $out/pytest.log | tail -1)"
timeout 300 python tools/gpu_fuzz_ends.py 7 150 512 40 > $out/fuzz_ends.json 2> $out/fuzz_ends.err; echo "fuzz_ends rc=$? $(cut -c1-300 $out/fuzz_ends.json)"
FUZZ_MODES=0 timeout 200 python tools/gpu_fuzz.py 60 77 > $out/fuzz.json 2> $out/fuzz.err; echo "fuzz rc=$? $(tail -c 400 $out/fuzz.json)"
timeout 400 python tools/gpu_perf.py > $out/perf_survey.jsonl 2> $out/perf.err; echo "perf rc=$?"; grep -E "prot400|dna512|dna150 auto" $out/perf_survey.jsonl | cut -c1-230
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$? $(head -c 200 $out/bench_n1.json)"
