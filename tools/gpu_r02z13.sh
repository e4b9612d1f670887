#!/bin/bash
O=gpurun_out/r02z13
mkdir -p $O
RL_TWOSTEP_VALS=0 timeout 600 python tools/c5_threads.py 256 2>&1 | grep -v Warn | sed "s/^/vals=0 /" | tee $O/c5.log
timeout 600 python tools/c5_threads.py 256 2>&1 | grep -v Warn | sed "s/^/vals=1 /" | tee -a $O/c5.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee $O/status.txt
tail -4 $O/pytest.log
