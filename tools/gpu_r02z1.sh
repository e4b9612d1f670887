#!/bin/bash
# round 2, session 4: compaction of live rays between bursts + next-window prefetch in the cooperative tail
O=gpurun_out/r02z1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee $O/status.txt
tail -3 $O/pytest.log
for v in base compact prefetch pf0 div1; do
  echo "== $v" | tee -a $O/tune.log
  RL_B200_LIB=tools/_trace/librangelib_b200_$v.so timeout 300 python tools/tune_fused.py 4 8 16 2>&1 | tee -a $O/tune.log
done
echo "== product" | tee -a $O/tune.log
timeout 300 python tools/tune_fused.py 4 8 16 2>&1 | tee -a $O/tune.log
for bp in 4 8; do
  echo "== product RL_BLOCK_BURST_PAIRS=$bp" | tee -a $O/tune.log
  RL_BLOCK_BURST_PAIRS=$bp timeout 300 python tools/tune_fused.py 4 8 16 2>&1 | tee -a $O/tune.log
done
timeout 300 python tools/trace_fused.py > $O/trace.log 2>&1; echo "trace rc=$?" | tee -a $O/status.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/status.txt
head -c 600 $O/bench.json
