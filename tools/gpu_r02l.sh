#!/bin/bash
O=gpurun_out/r02l
mkdir -p $O
python -m pytest tests/test_gpu_scale_parity.py -m gpu -x -q -k "distance_transform or c5_million" > $O/test.log 2>&1; echo "new tests rc=$?" > $O/status.txt
tail -15 $O/test.log
python -m pytest tests/test_gpu_parity.py -m gpu -x -q >> $O/test.log 2>&1; echo "parity tests rc=$?" >> $O/status.txt
tail -3 $O/test.log
for w in edt_1200 edt_8192; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:edt_ --csv --log-file $O/$w.csv python tools/prof_r02.py $w 2 > /dev/null 2>&1
  grep edt_ $O/$w.csv | awk -F'","' '{print substr($5,1,44), $(NF)}' | tail -4
done
ncu --set full --clock-control none -k regex:edt_envelope -s 2 -c 2 -o $O/edt_env -f python tools/prof_r02.py edt_8192 2 > /dev/null 2>&1
python - > $O/time.log 2>&1 <<'PY'
import sys, os, time, numpy as np, torch
sys.path.insert(0, '.')
import range_libc_b200 as rl
from range_libc_b200 import workloads as wl
import bench
st = torch.cuda.current_stream()
def t(fn, it=3):
    for _ in range(2): fn()
    ts=[]
    for _ in range(it):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st); b.synchronize(); ts.append(a.elapsed_time(b))
    return np.median(ts)
occ5, p5_h, a5_h, o5_h = bench.c5_inputs()
rm5 = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ5.T.astype(bool))), 500.0)
rm5.set_sensor_model(wl.sensor_table(501)); rm5.set_stream(st.cuda_stream)
p5, a5, o5 = (torch.from_numpy(x).cuda() for x in (p5_h, a5_h, o5_h))
w5 = torch.empty(len(p5_h), dtype=torch.float64, device="cuda")
ms = t(lambda: rm5.calc_range_repeat_angles_eval_sensor_model(p5, a5, o5, w5))
print("C5 1M x 1080 (stream kernel): %.2f ms  %.2f G rays/s" % (ms, len(p5_h) * 1080 / ms / 1e6))
for nn in (125000, 250000):
    ms = t(lambda: rm5.calc_range_repeat_angles_eval_sensor_model(p5[:nn], a5, o5, w5[:nn]))
    print("C5 %d x 1080: %.2f ms  %.2f G rays/s" % (nn, ms, nn * 1080 / ms / 1e6))
# 5 cm map fan shapes through the fused call
occ = wl.load_map("basement_hallways_5cm")
rm = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), 500.0)
rm.set_sensor_model(wl.sensor_table(501)); rm.set_stream(st.cuda_stream)
for n_p, n_b in ((20000, 1080), (50000, 360), (100000, 128), (100000, 60)):
    parts = torch.from_numpy(wl.pf_particles_uniform(occ, n_p, seed=11)).cuda()
    ang = torch.from_numpy(wl.lidar_angles(n_b)).cuda()
    ob = torch.from_numpy(np.clip(120.0 + 80.0 * np.sin(np.linspace(0, 3.0, n_b)), 0, 500).astype(np.float32)).cuda()
    w = torch.empty(n_p, dtype=torch.float64, device="cuda")
    ms = t(lambda: rm.calc_range_repeat_angles_eval_sensor_model(parts, ang, ob, w))
    print("5cm fused %d x %d: %.3f ms  %.2f G rays/s" % (n_p, n_b, ms, n_p * n_b / ms / 1e6), flush=True)
PY
cat $O/time.log
RL_FUSED_STREAM=0 python - > $O/time0.log 2>&1 <<'PY'
import sys, os, time, numpy as np, torch
sys.path.insert(0, '.')
import range_libc_b200 as rl
from range_libc_b200 import workloads as wl
st = torch.cuda.current_stream()
def t(fn, it=3):
    for _ in range(2): fn()
    ts=[]
    for _ in range(it):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st); b.synchronize(); ts.append(a.elapsed_time(b))
    return np.median(ts)
occ = wl.load_map("basement_hallways_5cm")
rm = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), 500.0)
rm.set_sensor_model(wl.sensor_table(501)); rm.set_stream(st.cuda_stream)
for n_p, n_b in ((20000, 1080), (50000, 360), (100000, 128), (100000, 60)):
    parts = torch.from_numpy(wl.pf_particles_uniform(occ, n_p, seed=11)).cuda()
    ang = torch.from_numpy(wl.lidar_angles(n_b)).cuda()
    ob = torch.from_numpy(np.clip(120.0 + 80.0 * np.sin(np.linspace(0, 3.0, n_b)), 0, 500).astype(np.float32)).cuda()
    w = torch.empty(n_p, dtype=torch.float64, device="cuda")
    ms = t(lambda: rm.calc_range_repeat_angles_eval_sensor_model(parts, ang, ob, w))
    print("5cm fused (no stream kernel) %d x %d: %.3f ms  %.2f G rays/s" % (n_p, n_b, ms, n_p * n_b / ms / 1e6), flush=True)
PY
cat $O/time0.log; cat $O/status.txt
