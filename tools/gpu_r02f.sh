#!/bin/bash
O=gpurun_out/r02f
mkdir -p $O
python tools/trace_fused.py > $O/trace_p2.log 2>&1
RL_TRACE_LIB=tools/_trace/librangelib_b200_p4trace.so python tools/trace_fused.py > $O/trace_p4.log 2>&1
grep -A1 "longest ray\|cooperative tail" $O/trace_p2.log | head -40
echo ===== p4
grep -A1 "longest ray\|cooperative tail" $O/trace_p4.log | head -40
for v in p4 p3 p4w; do echo "== $v" >> $O/tune.log; RL_B200_LIB=tools/_trace/librangelib_b200_$v.so RL_BLOCK_BURST_PAIRS=6 python tools/tune_fused.py 4 8 16 >> $O/tune.log 2>&1; done
echo "== product" >> $O/tune.log; RL_BLOCK_BURST_PAIRS=6 python tools/tune_fused.py 4 8 16 >> $O/tune.log 2>&1
cat $O/tune.log
