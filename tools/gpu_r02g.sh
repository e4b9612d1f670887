#!/bin/bash
O=gpurun_out/r02g
mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rm or fused or golden or fuzz or rotated" > $O/rmtest.log 2>&1; echo "rm tests rc=$?" > $O/status.txt
tail -3 $O/rmtest.log
python tools/trace_fused.py > $O/trace.log 2>&1
grep -B2 -A1 "longest ray" $O/trace.log | head -60
echo "== product (unroll 4, 2 probes, spacing 0.75)" >> $O/tune.log; RL_BLOCK_BURST_PAIRS=6 python tools/tune_fused.py 4 8 16 >> $O/tune.log 2>&1
for v in u8 u2 s05 p3s05 p1 s1; do echo "== $v" >> $O/tune.log; RL_B200_LIB=tools/_trace/librangelib_b200_$v.so RL_BLOCK_BURST_PAIRS=6 python tools/tune_fused.py 4 8 >> $O/tune.log 2>&1; done
for bp in 3 4 5; do echo "== product bp=$bp" >> $O/tune.log; RL_BLOCK_BURST_PAIRS=$bp python tools/tune_fused.py 8 16 24 >> $O/tune.log 2>&1; done
cat $O/tune.log; cat $O/status.txt
