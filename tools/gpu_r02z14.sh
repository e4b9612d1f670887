#!/bin/bash
O=gpurun_out/r02z14
mkdir -p $O
timeout 600 python tools/c5_twostep.py 2>&1 | grep -v Warn | sed "s/^/minb6 /" | tee $O/twostep.log
RL_B200_LIB=tools/_trace/librangelib_b200_minb5.so timeout 600 python tools/c5_twostep.py 2>&1 | grep -v Warn | sed "s/^/minb5 /" | tee -a $O/twostep.log
timeout 600 python tools/fused_kinds.py 2>&1 | grep -v Warn | tee $O/kinds.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee $O/status.txt
tail -4 $O/pytest.log
