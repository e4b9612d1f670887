#!/bin/bash
# round 2, session 4: deep fused launches with the product on a ninth warp (fused_overlap_kernel)
O=gpurun_out/r02z5
mkdir -p $O
RL_FUSED_OVERLAP=0 timeout 600 python tools/c5_threads.py 256 2>&1 | grep -v Warning | sed 's/^/overlap=0 /' | tee $O/c5.log
timeout 600 python tools/c5_threads.py 256 2>&1 | grep -v Warning | sed 's/^/overlap=1 /' | tee -a $O/c5.log
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee $O/status.txt
tail -5 $O/pytest.log
