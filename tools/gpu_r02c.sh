#!/bin/bash
O=gpurun_out/r02c
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/gputest.log 2>&1; echo "gputest rc=$?" > $O/status.txt
tail -5 $O/gputest.log
for bp in 3 4 6; do RL_BLOCK_BURST_PAIRS=$bp python tools/tune_fused.py 4 8 12 16 24 32 48 >> $O/tune_fused.log 2>&1; done
cat $O/tune_fused.log
for w in edt_1200 edt_8192; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:edt_pass --csv --log-file $O/$w.csv python tools/prof_r02.py $w 2 > /dev/null 2>&1
  grep edt_pass $O/$w.csv | awk -F'","' '{print $5, $(NF)}' | tail -4
done
cat $O/status.txt
