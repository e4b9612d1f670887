#!/bin/bash
# end of round 2 (second half): tests, both bench arms, smoke, launch list, sanitizer on one B200
O=gpurun_out/r02fin
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/box.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "gpu tests rc=$?" > $O/status.txt
tail -3 $O/pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err; echo "ref rc=$?" >> $O/status.txt
timeout 1500 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" >> $O/status.txt
timeout 600 python -c "
import __graft_entry__ as g
g.smoke()
" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
tail -1 $O/smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 20 --warmup 5 --no-extra > $O/b_ncu.log 2>&1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_probe.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/status.txt
timeout 1800 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_probe.py > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/status.txt
tail -2 $O/memcheck.log $O/racecheck.log
python - <<'PY'
import json
for f in ('gpurun_out/r02fin/bench_ref_n1.json','gpurun_out/r02fin/bench_n1.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f, "value %.4g ms/step %.5f e2e %.4g" % (d['value'], d['ms_per_step'], d['e2e']['value']))
PY
cat $O/status.txt
