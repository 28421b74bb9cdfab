#!/bin/bash
# round 2, GPU call W (1 GPU): ncu of the relative-key kernel, then the last full pass on the final commit
out=gpurun_out/r02w
mkdir -p $out
SEQALIGN_CHUNKS=1 timeout 300 ncu --set full --clock-control none -k regex:"fast16_kernel" -c 2 -o /tmp/ncu_endrel python tools/gpu_prof_endrel.py > $out/ncu_endrel.log 2>&1
echo "ncu rc=$?"
python tools/ncu_all_summary.py /tmp/ncu_endrel.ncu-rep $out/ncu_r02w_endrel.csv "SEQALIGN_CHUNKS=1 ncu --set full --clock-control none -k regex:fast16_kernel -c 2 python tools/gpu_prof_endrel.py (20k protein pairs 400x400, BLOSUM62, score + end cell)" > /dev/null 2>&1
ncu -i /tmp/ncu_endrel.ncu-rep --page details --csv 2>/dev/null | grep -i "stall\|Kernel Name\|Issue Slots\|No Eligible\|Achieved Occupancy\|Registers\|Theoretical Occ\|Duration\|ALU\|Executed Ipc" | head -120 > $out/ncu_r02w_endrel_details.csv
rm -f /tmp/ncu_endrel.ncu-rep
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $out/smoke.log 2>&1; echo "smoke rc=$? $(grep 'smoke ok' $out/smoke.log | cut -c1-160)"
( time timeout 1500 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$? $(grep -E 'passed|failed' $out/pytest.log | tail -1)"
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$? $(head -c 200 $out/bench_n1.json)"
