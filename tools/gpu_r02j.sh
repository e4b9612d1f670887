#!/bin/bash
O=gpurun_out/r02j
mkdir -p $O
python -m pytest tests/test_gpu_scale_parity.py -m gpu -x -q -k "c3_bin or c5_million" > $O/test.log 2>&1; echo "tests rc=$?" > $O/status.txt
tail -15 $O/test.log
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "deep_fused or spatial" >> $O/test.log 2>&1; echo "tests2 rc=$?" >> $O/status.txt
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "big_maps or structures or dynamic_map or degenerate or fuzz or golden" >> $O/test.log 2>&1; echo "edt tests rc=$?" >> $O/status.txt
tail -3 $O/test.log
for w in edt_1200 edt_8192; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:edt_pass --csv --log-file $O/$w.csv python tools/prof_r02.py $w 2 > /dev/null 2>&1
  grep edt_pass $O/$w.csv | awk -F'","' '{print substr($5,1,40), $(NF)}' | tail -2
done
python - > $O/time.log 2>&1 <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import range_libc_b200 as rl
from range_libc_b200 import workloads as wl
import bench
st = torch.cuda.current_stream()
def t(fn, it=5):
    for _ in range(2): fn()
    ts=[]
    for _ in range(it):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st); b.synchronize(); ts.append(a.elapsed_time(b))
    return np.median(ts)
big = wl.load_map("gigantic_map")
cd = rl.PyCDDTCast(rl.PyOMap(np.ascontiguousarray(big.T.astype(bool))), 500.0, 108)
cd.set_stream(st.cuda_stream)
n = 1 << 24
q = torch.from_numpy(wl.random_queries(big.shape[0], big.shape[1], n, seed=2)).cuda()
out = torch.empty(n, dtype=torch.float32, device="cuda")
for pr in (False, True):
    if pr: cd.prune()
    for sort in (True, False):
        cd.set_spatial_sort(sort)
        ms = t(lambda: cd.calc_range_many_grid(q, out))
        print("C3 pruned=%s partitioned=%s  %.3f ms  %.2f G rays/s" % (pr, sort, ms, n / ms / 1e6), flush=True)
del cd, q, out
occ5, p5_h, a5_h, o5_h = bench.c5_inputs()
rm5 = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ5.T.astype(bool))), 500.0)
rm5.set_sensor_model(wl.sensor_table(501)); rm5.set_stream(st.cuda_stream)
p5, a5, o5 = (torch.from_numpy(x).cuda() for x in (p5_h, a5_h, o5_h))
w5 = torch.empty(len(p5_h), dtype=torch.float64, device="cuda")
ms = t(lambda: rm5.calc_range_repeat_angles_eval_sensor_model(p5, a5, o5, w5), it=3)
print("C5 1M x 1080: %.2f ms  %.2f G rays/s" % (ms, len(p5_h) * 1080 / ms / 1e6))
for nn in (125000, 250000, 500000):
    ms = t(lambda: rm5.calc_range_repeat_angles_eval_sensor_model(p5[:nn], a5, o5, w5[:nn]), it=3)
    print("C5 %d x 1080: %.2f ms  %.2f G rays/s" % (nn, ms, nn * 1080 / ms / 1e6))
PY
cat $O/time.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:"cddt_" --csv --log-file $O/c3_launches.csv python tools/prof_r02.py c3_cddt 2 > /dev/null 2>&1
grep -E "cddt_" $O/c3_launches.csv | awk -F'","' '{print $5, $(NF-2), $(NF)}' | tail -12
cat $O/status.txt
