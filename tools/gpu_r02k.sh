#!/bin/bash
O=gpurun_out/r02k
mkdir -p $O
python -m pytest tests/test_gpu_scale_parity.py -m gpu -x -q -k "distance_transform or checkpoint" > $O/test.log 2>&1; echo "new tests rc=$?" > $O/status.txt
tail -15 $O/test.log
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "big_maps or structures or dynamic_map or degenerate or fuzz or golden or giant" >> $O/test.log 2>&1; echo "edt tests rc=$?" >> $O/status.txt
tail -3 $O/test.log
for w in edt_1200 edt_8192; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:edt_pass --csv --log-file $O/$w.csv python tools/prof_r02.py $w 2 > /dev/null 2>&1
  grep edt_pass $O/$w.csv | awk -F'","' '{print substr($5,1,44), $(NF)}' | tail -2
done
cat $O/status.txt
