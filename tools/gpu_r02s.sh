#!/bin/bash
O=gpurun_out/r02s
mkdir -p $O
python -m pytest tests/test_gpu_scale_parity.py -m gpu -x -q -k "c5" > $O/test.log 2>&1; echo "c5 tests rc=$?" > $O/status.txt
tail -3 $O/test.log
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "deep_fused or spatial or peer_store" >> $O/test.log 2>&1; echo "parity subset rc=$?" >> $O/status.txt
tail -3 $O/test.log
cat > /tmp/t_s.py <<'PY'
import sys, os, time, numpy as np, torch
sys.path.insert(0, '.')
import range_libc_b200 as rl
from range_libc_b200 import workloads as wl
import bench
st = torch.cuda.current_stream()
tag = sys.argv[1]
def t(fn, it=5):
    for _ in range(2): fn()
    ts=[]
    for _ in range(it):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st); b.synchronize(); ts.append(a.elapsed_time(b))
    return np.median(ts)
occ = wl.load_map("basement_hallways_5cm")
om = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
occ5, p5_h, a5_h, o5_h = bench.c5_inputs()
rm5 = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ5.T.astype(bool))), 500.0)
rm5.set_sensor_model(wl.sensor_table(501)); rm5.set_stream(st.cuda_stream)
p5, a5, o5 = (torch.from_numpy(x).cuda() for x in (p5_h, a5_h, o5_h))
w5 = torch.empty(len(p5_h), dtype=torch.float64, device="cuda")
ms = t(lambda: rm5.calc_range_repeat_angles_eval_sensor_model(p5, a5, o5, w5), it=3)
print("%s C5 1M x 1080: %.2f ms  %.2f G rays/s" % (tag, ms, len(p5_h) * 1080 / ms / 1e6), flush=True)
rm = rl.PyRayMarchingGPU(om, 500.0); rm.set_sensor_model(wl.sensor_table(501)); rm.set_stream(st.cuda_stream)
for n_p, n_b in ((20000, 1080), (50000, 360), (100000, 128)):
    parts = torch.from_numpy(wl.pf_particles_uniform(occ, n_p, seed=11)).cuda()
    ang = torch.from_numpy(wl.lidar_angles(n_b)).cuda()
    ob = torch.from_numpy(np.clip(120.0 + 80.0 * np.sin(np.linspace(0, 3.0, n_b)), 0, 500).astype(np.float32)).cuda()
    w = torch.empty(n_p, dtype=torch.float64, device="cuda")
    ms = t(lambda: rm.calc_range_repeat_angles_eval_sensor_model(parts, ang, ob, w))
    print("%s 5cm fused %d x %d: %.3f ms  %.2f G rays/s" % (tag, n_p, n_b, ms, n_p * n_b / ms / 1e6), flush=True)
PY
python /tmp/t_s.py warp_default 2>&1 | tee $O/time.log
RL_FUSED_WARP=0 RL_FUSED_DEEP_THREADS=256 python /tmp/t_s.py cta256 2>&1 | tee -a $O/time.log
RL_FUSED_WARP=0 python /tmp/t_s.py cta128 2>&1 | tee -a $O/time.log
RL_FUSED_WARP_MIN_M=100 RL_FUSED_PERSIST=0 python /tmp/t_s.py warp_all 2>&1 | tee -a $O/time.log
ncu --set full --cache-control none --clock-control none -k regex:fused_rm_warp -s 1 -c 1 -o $O/c5 -f python tools/prof_r02.py c5 2 > $O/c5.log 2>&1
python tools/ncu_summary.py $O/c5.ncu-rep > $O/c5.txt 2>&1
cat $O/c5.txt | head -40
cat $O/status.txt
