#!/bin/bash
O=gpurun_out/r02z18
mkdir -p $O
for v in literals ""; do
  if [ -z "$v" ]; then L=""; T=constmem; else L="tools/_trace/librangelib_b200_$v.so"; T=$v; fi
  RL_B200_LIB=$L timeout 600 python tools/time_kernels.py $T 2>&1 | grep -v Warn | tee -a $O/time.log
  RL_B200_LIB=$L timeout 600 python tools/c5_twostep.py 2>&1 | grep "tile order   calc_range\|tile order   fused" | sed "s/^/$T /" | tee -a $O/time.log
  RL_B200_LIB=$L timeout 300 python tools/tune_fused.py 8 2>&1 | grep -v Warn | sed "s/^/$T /" | tee -a $O/time.log
done
