#!/bin/bash
O=gpurun_out/r02z6
mkdir -p $O
for f in 0 1 2 3; do
RL_OVERLAP_FLAGS=$f timeout 600 python tools/c5_threads.py 256 2>&1 | grep -v Warning | sed "s/^/flags=$f /" | tee -a $O/c5.log
done
