#!/bin/bash
# round 2, GPU call N (1 GPU): wide-pair traceback with a row checkpoint every 64 / 32 / 16 rows (config 3 shape, 2000 pairs;
# the recompute walk is occupancy-bound by its flag tile in shared memory), parity of the variants on the wide-pair tests
out=gpurun_out/r02n
mkdir -p $out
for v in 64 32 16; do
  lib=seq-align_b200/lib/libseqalign_b200.so; [ $v != 64 ] && lib=seq-align_b200/lib_ck$v/libseqalign_b200.so
  SEQALIGN_LIB=$PWD/$lib timeout 300 python tools/gpu_config3.py 2000 10000 4 > $out/config3_ck$v.jsonl 2> $out/config3_ck$v.err
  echo "ck$v rc=$? $(python -c "
import json;d=json.load(open('$out/config3_ck$v.jsonl'));print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('align_kernel_ms','walk_ms','align_tcups_kernel','align_tcups_kernel_plus_walk','all_scores_equal_score_only_kernel','oracle_pairs_equal','launches')})")"
done
for v in 32 16; do
  SEQALIGN_LIB=$PWD/seq-align_b200/lib_ck$v/libseqalign_b200.so timeout 300 python -m pytest tests/test_parity.py -m gpu -q -k "wide_pairs or full_size_long" > $out/pytest_ck$v.log 2>&1
  echo "pytest ck$v rc=$? $(tail -1 $out/pytest_ck$v.log)"
done
