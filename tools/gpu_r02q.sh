#!/bin/bash
O=gpurun_out/r02q
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/test.log 2>&1; echo "gpu tests rc=$?" > $O/status.txt
tail -4 $O/test.log
python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" >> $O/status.txt
python - <<'PY'
import json
for l in open('gpurun_out/r02q/bench_n1.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print("value %.3g ms/step %.4f e2e %.3g" % (d['value'], d['ms_per_step'], d['e2e']['value']))
        x=d['extra']
        for k in ('rm_random_rays_per_s','cddt_random_rays_per_s','pcddt_random_rays_per_s','bl_random_rays_per_s'):
            print(k, "%.3g" % x.get(k, 0))
        print(json.dumps(x.get('c3_gigantic_map'))[:900])
        print(json.dumps(x.get('c1_basement_10cm', {}).get('cddt'))[:300])
        print(json.dumps(x.get('c5_rm_fused_8192'))[:500])
        print(json.dumps(x.get('c4_dynamic_bl_4096'))[:500])
PY
cat $O/status.txt
