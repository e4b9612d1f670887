#!/bin/bash
O=gpurun_out/r02z
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/box.txt
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "gpu tests rc=$?" > $O/status.txt
tail -3 $O/pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err; echo "ref rc=$?" >> $O/status.txt
python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" >> $O/status.txt
python -c "
import __graft_entry__ as g
g.smoke()
" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
tail -1 $O/smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 20 --warmup 5 --no-extra > $O/b_ncu.log 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/r02z/bench_ref_n1.json','gpurun_out/r02z/bench_n1.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f, "value %.4g ms/step %.5f e2e %.4g" % (d['value'], d['ms_per_step'], d['e2e']['value']))
PY
cat $O/status.txt
