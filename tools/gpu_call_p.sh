#!/bin/bash
# round 2, GPU call P (1 GPU): ncu --set full over every kernel family (summarised on the box: the report itself is too big to
# travel), and the tools' pairs-per-submit experiment on a 2 M-pair FASTA
out=gpurun_out/r02p
mkdir -p $out
SEQALIGN_CHUNKS=1 timeout 900 ncu --set full --clock-control none -k regex:"fast|long|walk|hits|mats|emit|classify|nl_|scan3|contrib" -c 60 -o /tmp/ncu_all python tools/gpu_prof_all.py > $out/ncu_all.log 2>&1
echo "ncu rc=$?"
python tools/ncu_all_summary.py /tmp/ncu_all.ncu-rep $out/ncu_r02_all_kernels.csv "SEQALIGN_CHUNKS=1 ncu --set full --clock-control none -k regex:fast|long|walk|hits|mats|emit|classify|nl_|scan3|contrib python tools/gpu_prof_all.py (round 2, final kernels; one launch of every kernel family on its BASELINE shape)"
ncu -i /tmp/ncu_all.ncu-rep --page details --csv 2>/dev/null | grep -i "stall\|Kernel Name\|Issue Slots\|No Eligible\|Achieved Occupancy" | head -400 > $out/ncu_r02_details_excerpt.csv
rm -f /tmp/ncu_all.ncu-rep
CLI_BIG_BATCHES=16384,65536,262144 timeout 400 python tools/gpu_cli_big.py 2000000 1 > $out/cli_batches.jsonl 2> $out/cli_batches.err
echo "cli rc=$?"; cut -c1-330 $out/cli_batches.jsonl
du -sh $out
