#!/bin/bash
# compute-sanitizer over a reduced selection of the GPU parity tests that reaches every kernel family
# (fast, fast16, long, general, general_coop, mats, hits, walk, scan, synth), at the emulator's small sizes.
#   tools/gpu_sanitize.sh [outdir]     (run on the GPU box: gpurun -- tools/gpu_sanitize.sh gpurun_out/san)
# plus ASan/UBSan on the host C code of the library and the command-line tools (CPU only, lane emulator).
out=${1:-gpurun_out/san}
mkdir -p "$out"
export SEQALIGN_TEST_SMALL=1
SEL='test_headline_config_sample or test_alignments_every_fill_shape or test_wide_pairs_strip_pipeline or test_wide_pairs_cooperative or (test_batch_matrices and (sw_cli or nw_default or free_ends)) or (test_multi_hit_on_device and sw_cli) or (test_scores_ragged and (no_gaps_a or blosum62 or free_ends)) or test_uniform_submit_and_result_sink or test_device_generator_matches_numpy or test_fast16_tight_shapes or test_device_resident_async or test_alignments_wide_and_waves'
for tool in memcheck racecheck synccheck initcheck; do
  start=$(date +%s)
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 30 \
    python -m pytest tests/test_parity.py tests/test_synth.py -m gpu -q -x -k "$SEL" -p no:cacheprovider > "$out/$tool.log" 2>&1
  rc=$?
  echo "$tool rc=$rc seconds=$(( $(date +%s) - start )) $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' "$out/$tool.log" | tr '\n' ' ')" | tee -a "$out/summary.txt"
done
