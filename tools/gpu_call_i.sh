#!/bin/bash
# round 2, GPU call I (1 GPU): length buckets + 8-lane emit on hardware; ragged throughput; decode ncu list; bench line; initcheck
out=gpurun_out/r02i
mkdir -p $out
( time timeout 900 python -m pytest tests/test_parity.py tests/test_reader.py -m gpu -q -k "bucket or ragged or fast16 or reader or decoder or decoded or headline or multi_hit" ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $out/pytest.log)"
timeout 200 python tools/gpu_ragged.py > $out/ragged.jsonl 2> $out/ragged.err; echo "ragged rc=$?"; cat $out/ragged.jsonl
SEQALIGN_NO_BUCKETS=1 timeout 200 python tools/gpu_ragged.py 100000 > $out/ragged_nobuckets.jsonl 2> $out/ragged_nobuckets.err; head -1 $out/ragged_nobuckets.jsonl
timeout 300 python tools/gpu_decode.py 500000 > $out/decode.jsonl 2> $out/decode.err; echo "decode rc=$?"; cut -c1-230 $out/decode.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 45 --csv --log-file $out/launches_decode.csv python tools/gpu_decode.py 200000 > $out/ncu_decode.log 2>&1
grep emit_kernel $out/launches_decode.csv | grep time_duration | head -3
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$? $(head -c 200 $out/bench_n1.json)"
export SEQALIGN_TEST_SMALL=1
timeout 300 compute-sanitizer --tool initcheck --error-exitcode 86 --print-limit 10 python -m pytest tests/test_parity.py -m gpu -q -x -k "test_uniform_submit_and_result_sink or test_alignments_every_fill_shape or (test_multi_hit_on_device and sw_cli) or test_wide_pairs_sw_score or test_length_buckets" -p no:cacheprovider > $out/initcheck.log 2>&1
echo "initcheck rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' $out/initcheck.log | tr '\n' ' ')"
