#!/bin/bash
O=gpurun_out/r02z21
mkdir -p $O
RL_PARAM_POSES_OFF=1 RL_B200_LIB=tools/_trace/librangelib_b200_htiming.so timeout 600 python tools/e2e_breakdown.py 2>&1 | grep -v Warn | sed "s/^/copy path:   /" | tee $O/e2e.log
RL_B200_LIB=tools/_trace/librangelib_b200_htiming.so timeout 600 python tools/e2e_breakdown.py 2>&1 | grep -v Warn | sed "s/^/pose params: /" | tee -a $O/e2e.log
RL_PARAM_POSES_OFF=1 timeout 600 python tests/probes/e2e_probe.py 2>&1 | grep -v Warn | sed "s/^/copy path:   /" | tee -a $O/e2e.log
timeout 600 python tests/probes/e2e_probe.py 2>&1 | grep -v Warn | sed "s/^/pose params: /" | tee -a $O/e2e.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee $O/status.txt
tail -3 $O/pytest.log
