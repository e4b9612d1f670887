#!/bin/bash
O=gpurun_out/r02z8
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee $O/status.txt
tail -5 $O/pytest.log
timeout 900 python tools/sanitize_probe.py --help > /dev/null 2>&1
