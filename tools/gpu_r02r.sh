#!/bin/bash
O=gpurun_out/r02r
mkdir -p $O
python -m pytest tests/test_gpu_scale_parity.py -m gpu -x -q -k "c5 or c3" > $O/test.log 2>&1; echo "c5/c3 tests rc=$?" > $O/status.txt
tail -3 $O/test.log
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "deep_fused or spatial or fresh or fuzz or rotated or golden or persistent" >> $O/test.log 2>&1; echo "parity subset rc=$?" >> $O/status.txt
tail -3 $O/test.log
cat > /tmp/t_r.py <<'PY'
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
n = 1 << 24
q = torch.from_numpy(wl.random_queries(occ.shape[0], occ.shape[1], n, seed=1)).cuda()
out = torch.empty(n, dtype=torch.float32, device="cuda")
if "cast" in sys.argv[2]:
    rm = rl.PyRayMarchingGPU(om, 500.0); rm.set_stream(st.cuda_stream)
    ms = t(lambda: rm.calc_range_many_grid(q, out)); print("%s RM 5cm random %.3f ms %.2f G rays/s" % (tag, ms, n/ms/1e6), flush=True)
    parts = torch.from_numpy(wl.pf_particles_tracking(occ, 262144, seed=11, dt=rm.distance_transform())[0]).cuda()
    ang = torch.from_numpy(wl.lidar_angles(60)).cuda()
    ms = t(lambda: rm.calc_range_repeat_angles(parts, ang, out)); print("%s RM fan tracking 262144x60 %.3f ms %.2f G rays/s" % (tag, ms, 262144*60/ms/1e6), flush=True)
    cd = rl.PyCDDTCast(om, 500.0, 108); cd.set_stream(st.cuda_stream)
    ms = t(lambda: cd.calc_range_many_grid(q, out)); print("%s CDDT 5cm random %.3f ms %.2f G rays/s" % (tag, ms, n/ms/1e6), flush=True)
    cd.prune()
    ms = t(lambda: cd.calc_range_many_grid(q, out)); print("%s PCDDT 5cm random %.3f ms %.2f G rays/s" % (tag, ms, n/ms/1e6), flush=True)
if "c5" in sys.argv[2]:
    occ5, p5_h, a5_h, o5_h = bench.c5_inputs()
    rm5 = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ5.T.astype(bool))), 500.0)
    rm5.set_sensor_model(wl.sensor_table(501)); rm5.set_stream(st.cuda_stream)
    p5, a5, o5 = (torch.from_numpy(x).cuda() for x in (p5_h, a5_h, o5_h))
    w5 = torch.empty(len(p5_h), dtype=torch.float64, device="cuda")
    ms = t(lambda: rm5.calc_range_repeat_angles_eval_sensor_model(p5, a5, o5, w5), it=3)
    print("%s C5 1M x 1080: %.2f ms  %.2f G rays/s" % (tag, ms, len(p5_h) * 1080 / ms / 1e6), flush=True)
    rm = rl.PyRayMarchingGPU(om, 500.0); rm.set_sensor_model(wl.sensor_table(501)); rm.set_stream(st.cuda_stream)
    parts = torch.from_numpy(wl.pf_particles_uniform(occ, 20000, seed=11)).cuda()
    w = torch.empty(20000, dtype=torch.float64, device="cuda")
    ms = t(lambda: rm.calc_range_repeat_angles_eval_sensor_model(parts, a5, o5, w))
    print("%s 5cm fused 20000 x 1080: %.3f ms  %.2f G rays/s" % (tag, ms, 20000 * 1080 / ms / 1e6), flush=True)
PY
python /tmp/t_r.py default cast,c5 2>&1 | tee $O/time.log
RL_B200_LIB=tools/_trace/librangelib_b200_rm8.so python /tmp/t_r.py rm8 cast 2>&1 | tee -a $O/time.log
RL_FUSED_DEEP_THREADS=256 python /tmp/t_r.py threads256 c5 2>&1 | tee -a $O/time.log
RL_FUSED_DEEP_THREADS=64 python /tmp/t_r.py threads64 c5 2>&1 | tee -a $O/time.log
RL_FUSED_GROUP_RAYS=6480 RL_FUSED_DEEP_THREADS=256 python /tmp/t_r.py group6480 c5 2>&1 | tee -a $O/time.log
cat $O/status.txt
