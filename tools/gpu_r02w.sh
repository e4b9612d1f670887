#!/bin/bash
O=gpurun_out/r02w
mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/test.log 2>&1; echo "parity rc=$?" > $O/status.txt
tail -3 $O/test.log
python -m pytest tests/test_gpu_scale_parity.py -m gpu -x -q -k "c1 or c2 or c5 or reference" >> $O/test.log 2>&1; echo "scale subset rc=$?" >> $O/status.txt
tail -3 $O/test.log
python tools/time_r02.py divmod cast 2>&1 | tee $O/time.log
for bp in 2 4 5; do RL_RM_BURST_PAIRS=$bp python tools/time_r02.py burst_pairs_$bp cast 2>&1 | grep -v "CDDT" | tee -a $O/time.log; done
python bench.py --steps 20 --warmup 5 --no-extra > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" >> $O/status.txt
python - <<'PY'
import json
for l in open('gpurun_out/r02w/bench_n1.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print("value %.4g ms/step %.5f e2e %.4g (%.2f us)" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']*1e3))
PY
cat $O/status.txt
