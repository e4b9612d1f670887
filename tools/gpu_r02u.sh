#!/bin/bash
O=gpurun_out/r02u
mkdir -p $O
for th in 256 192 128 64; do
  for bp in 6 4; do
    echo "== threads $th burst_pairs $bp" | tee -a $O/tune.log
    RL_FUSED_SMALL_THREADS=$th RL_BLOCK_BURST_PAIRS=$bp python tools/tune_fused.py 4 8 16 2>&1 | tee -a $O/tune.log
  done
done
