#!/bin/bash
O=gpurun_out/r02z9
mkdir -p $O
timeout 600 python tools/c5_threads.py 256 2>&1 | grep -v Warning | sed "s/^/big=0 /" | tee -a $O/c5.log
RL_FUSED_OVERLAP_BIG=1 timeout 600 python tools/c5_threads.py 256 2>&1 | grep -v Warning | sed "s/^/big=1 ppb=auto /" | tee -a $O/c5.log
for p in 1 2 4; do
RL_FUSED_OVERLAP_BIG=1 RL_OVERLAP_BIG_PPB=$p timeout 600 python tools/c5_threads.py 256 2>&1 | grep -v Warning | sed "s/^/big=1 ppb=$p /" | tee -a $O/c5.log
done
