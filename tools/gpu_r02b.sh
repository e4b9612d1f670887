#!/bin/bash
O=gpurun_out/r02b
mkdir -p $O
python -m pytest tests/test_gpu_scale_parity.py -m gpu -x -q > $O/scale_parity.log 2>&1; echo "scale parity rc=$?" > $O/status.txt
./tools/replay_bench > $O/replay_bench.log 2>&1; echo "replay rc=$?" >> $O/status.txt
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_probe.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/status.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_probe.py > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/status.txt
tail -3 $O/scale_parity.log; cat $O/replay_bench.log; tail -4 $O/memcheck.log; tail -4 $O/racecheck.log; cat $O/status.txt
