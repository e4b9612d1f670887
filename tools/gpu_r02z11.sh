#!/bin/bash
O=gpurun_out/r02z11
mkdir -p $O
RL_FUSED_TWOSTEP=0 timeout 600 python tools/fused_kinds.py 2>&1 | grep -v Warn | sed "s/^/twostep=0  /" | tee $O/kinds.log
RL_FUSED_TWOSTEP=23 timeout 600 python tools/fused_kinds.py 2>&1 | grep -v Warn | sed "s/^/twostep=23 /" | tee -a $O/kinds.log
RL_FUSED_TWOSTEP=0 timeout 600 python tools/c5_threads.py 256 2>&1 | grep -v Warn | sed "s/^/twostep=0 /" | tee $O/c5.log
timeout 600 python tools/c5_threads.py 256 2>&1 | grep -v Warn | sed "s/^/twostep=dflt /" | tee -a $O/c5.log
timeout 600 python tools/c5_twostep.py 2>&1 | grep -v Warn | tee $O/twostep.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee $O/status.txt
tail -5 $O/pytest.log
