"""Timing of the large-batch kernels (RM random / lidar fans, BL) on device-resident inputs: python tools/time_kernels.py <tag>"""
import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import range_libc_b200 as rl
from range_libc_b200 import workloads as wl
st = torch.cuda.current_stream()
tag = sys.argv[1]
def t(fn, it=7):
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
out = torch.empty(max(n, 16384 * 1080), dtype=torch.float32, device="cuda")
rm = rl.PyRayMarchingGPU(om, 500.0); rm.set_stream(st.cuda_stream)
ms = t(lambda: rm.calc_range_many_grid(q, out[:n])); print("%s RM 5cm random %.3f ms %.2f G rays/s" % (tag, ms, n/ms/1e6), flush=True)
for label, parts_h in (("tracking", wl.pf_particles_tracking(occ, 262144, seed=11, dt=rm.distance_transform())[0]), ("uniform", wl.pf_particles_uniform(occ, 262144, seed=11))):
    parts = torch.from_numpy(parts_h).cuda()
    ang = torch.from_numpy(wl.lidar_angles(60)).cuda()
    ms = t(lambda: rm.calc_range_repeat_angles(parts, ang, out[:262144 * 60])); print("%s RM fan %s 262144x60 %.3f ms %.2f G rays/s" % (tag, label, ms, 262144*60/ms/1e6), flush=True)
parts = torch.from_numpy(wl.pf_particles_uniform(occ, 16384, seed=11)).cuda()
ang = torch.from_numpy(wl.lidar_angles(1080)).cuda()
ms = t(lambda: rm.calc_range_repeat_angles(parts, ang, out[:16384 * 1080])); print("%s RM fan uniform 16384x1080 %.3f ms %.2f G rays/s" % (tag, ms, 16384*1080/ms/1e6), flush=True)
bl = rl.PyBresenhamsLine(om, 500.0); bl.set_stream(st.cuda_stream)
nb = 1 << 22
ms = t(lambda: bl.calc_range_many_grid(q[:nb], out[:nb])); print("%s BL 5cm random 2^22 %.3f ms %.2f G rays/s" % (tag, ms, nb/ms/1e6), flush=True)
occ4 = wl.synthetic_map(4096, seed=2026)
bl4 = rl.PyBresenhamsLine(rl.PyOMap(np.ascontiguousarray(occ4.T.astype(bool))), 500.0); bl4.set_stream(st.cuda_stream)
q4 = torch.from_numpy(wl.random_queries(4096, 4096, 1 << 20, seed=3)).cuda()
ms = t(lambda: bl4.calc_range_many_grid(q4, out[:1 << 20])); print("%s BL 4096^2 2^20 %.3f ms %.2f G rays/s" % (tag, ms, (1<<20)/ms/1e6), flush=True)
