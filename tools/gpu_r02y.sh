#!/bin/bash
O=gpurun_out/r02y
mkdir -p $O
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_probe.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?" | tee $O/status.txt
tail -4 $O/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_probe.py > $O/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/status.txt
tail -4 $O/racecheck.log
