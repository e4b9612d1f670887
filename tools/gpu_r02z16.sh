#!/bin/bash
# sanitizer over every kernel family incl. the round-2 second-half kernels; ncu of the two-kernel deep update
O=gpurun_out/r02z16
mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_probe.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?" | tee $O/status.txt
tail -3 $O/memcheck.log
timeout 1800 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_probe.py > $O/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/status.txt
tail -3 $O/racecheck.log
NCU="ncu --set full --clock-control none --cache-control none"
timeout 900 $NCU -k regex:'rm_persist_kernel|eval_overlap_kernel|gather_poses|tile_key' -s 8 -c 4 -o $O/c5_twostep -f python tools/prof_r02.py c5 > $O/c5.log 2>&1
echo "ncu c5 rc=$?" | tee -a $O/status.txt
python tools/ncu_summary.py $O/c5_twostep.ncu-rep > $O/ncu_c5_twostep.txt 2>&1
rm -f $O/c5_twostep.ncu-rep
grep -E "^## |duration" $O/ncu_c5_twostep.txt
