#!/bin/bash
# round 2, session 4: next-window prefetch (scoreboard-safe) + smallest-parameter-first order of the parked rays
O=gpurun_out/r02z2
mkdir -p $O
for v in base nopf noorder pf0; do
  echo "== $v" | tee -a $O/tune.log
  RL_B200_LIB=tools/_trace/librangelib_b200_$v.so timeout 300 python tools/tune_fused.py 4 8 16 2>&1 | tee -a $O/tune.log
done
echo "== product" | tee -a $O/tune.log
timeout 300 python tools/tune_fused.py 4 8 16 2>&1 | tee -a $O/tune.log
timeout 300 python tools/trace_fused.py > $O/trace.log 2>&1; echo "trace rc=$?" | tee -a $O/status.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/status.txt
tail -3 $O/pytest.log
