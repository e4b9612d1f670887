#!/bin/bash
O=gpurun_out/r02t
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/test.log 2>&1; echo "gpu tests rc=$?" > $O/status.txt
tail -4 $O/test.log
python tools/time_r02.py predload cast,c5 2>&1 | tee $O/time.log
RL_FUSED_WARP=0 RL_FUSED_DEEP_THREADS=256 python tools/time_r02.py predload_cta256 c5 2>&1 | tee -a $O/time.log
python bench.py --steps 20 --warmup 5 --no-extra > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" >> $O/status.txt
python - <<'PY'
import json
for l in open('gpurun_out/r02t/bench_n1.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print("value %.4g ms/step %.5f e2e %.4g (%.2f us)" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']*1e3))
PY
cat $O/status.txt
