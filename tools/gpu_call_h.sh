#!/bin/bash
# round 2, GPU call H (1 GPU): new kernels on hardware (SW wide, word-wise emit), perf survey, decode throughput + ncu list,
# initcheck after zeroing the walk output, one ncu --set full capture of the headline kernel and of the decoder's emit kernel
out=gpurun_out/r02h
mkdir -p $out
( time timeout 600 python -m pytest tests/test_parity.py tests/test_reader.py -m gpu -q -k "wide or reader or decoder or decoded or strip" ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $out/pytest.log)"
timeout 500 python tools/gpu_perf.py > $out/perf.jsonl 2> $out/perf.err; echo "perf rc=$?"
grep -i "wide\|general" $out/perf.jsonl | cut -c1-300
timeout 300 python tools/gpu_decode.py 500000 > $out/decode.jsonl 2> $out/decode.err; echo "decode rc=$?"; cut -c1-330 $out/decode.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $out/launches_decode.csv python tools/gpu_decode.py 200000 > $out/ncu_decode.log 2>&1
export SEQALIGN_TEST_SMALL=1
SEL='test_uniform_submit_and_result_sink or test_alignments_every_fill_shape or (test_multi_hit_on_device and sw_cli) or test_wide_pairs_sw_score'
timeout 300 compute-sanitizer --tool initcheck --error-exitcode 86 --print-limit 10 python -m pytest tests/test_parity.py -m gpu -q -x -k "$SEL" -p no:cacheprovider > $out/initcheck.log 2>&1
echo "initcheck rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' $out/initcheck.log | tr '\n' ' ')"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 86 --print-limit 10 python -m pytest tests/test_parity.py tests/test_reader.py -m gpu -q -x -k "test_wide_pairs_sw_score or decoder or decoded" -p no:cacheprovider > $out/memcheck.log 2>&1
echo "memcheck rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' $out/memcheck.log | tr '\n' ' ')"
unset SEQALIGN_TEST_SMALL
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast16_kernel -c 1 -o $out/ncu_fast16 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config5 --sustain 0.05 > $out/ncu_fast16.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:emit_kernel -c 1 -o $out/ncu_emit python tools/gpu_decode.py 200000 > $out/ncu_emit.log 2>&1
ls -la $out/*.ncu-rep 2>/dev/null
