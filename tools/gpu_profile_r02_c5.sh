#!/bin/bash
# ncu --set full (warm) of the two kernels of a deep fused update on BASELINE config 5's map (200000 x 1080)
O=gpurun_out/r02prof_c5
mkdir -p $O
NCU="ncu --set full --clock-control none --cache-control none --import-source on"
timeout 900 $NCU -k regex:'rm_persist_kernel|eval_overlap_kernel|gather_poses|tile_key' -s 8 -c 4 -o $O/c5_twostep -f python tools/prof_r02.py c5 > $O/c5.log 2>&1
echo "c5 rc=$?" | tee $O/status.txt
python tools/ncu_summary.py $O/c5_twostep.ncu-rep > $O/ncu_c5_twostep.txt 2>&1
ncu -i $O/c5_twostep.ncu-rep --page details --csv > $O/c5_details.csv 2>/dev/null
grep -E "^## |duration|dram__bytes|issue_active|hit_rate|warps_active" $O/ncu_c5_twostep.txt
ls -la $O
