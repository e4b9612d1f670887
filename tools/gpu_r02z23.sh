#!/bin/bash
O=gpurun_out/r02z23
mkdir -p $O
echo "== nounit" | tee -a $O/tune.log
RL_B200_LIB=tools/_trace/librangelib_b200_nounit.so timeout 300 python tools/tune_fused.py 4 8 2>&1 | grep -v Warn | tee -a $O/tune.log
echo "== product (unit runs)" | tee -a $O/tune.log
timeout 300 python tools/tune_fused.py 4 8 2>&1 | grep -v Warn | tee -a $O/tune.log
timeout 300 python tools/trace_fused.py > $O/trace.log 2>&1; echo "trace rc=$?" | tee -a $O/status.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/status.txt
tail -3 $O/pytest.log
