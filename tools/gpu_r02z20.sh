#!/bin/bash
O=gpurun_out/r02z20
mkdir -p $O
for h in 0 2 4 8 16; do
  RL_PERSIST_HANDOFF=$h timeout 600 python tools/time_kernels.py handoff=$h 2>&1 | grep -v Warn | grep "RM" | tee -a $O/time.log
done
for h in 0 4; do
  RL_PERSIST_HANDOFF=$h timeout 600 python tools/fused_kinds.py 2>&1 | grep "RM " | sed "s/^/handoff=$h /" | tee -a $O/time.log
  RL_PERSIST_HANDOFF=$h timeout 600 python tools/c5_twostep.py 2>&1 | grep "tile order   calc_range\|tile order   fused" | sed "s/^/handoff=$h /" | tee -a $O/time.log
done
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee $O/status.txt
tail -4 $O/pytest.log
