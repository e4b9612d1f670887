"""Kernel variant sweeps (development aid)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402
from tools.quick_bench import timeit  # noqa: E402

occ = wl.load_map("basement_hallways_5cm")
W, H = occ.shape
omap = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
rm = rl.PyRayMarchingGPU(omap, 500.0)
rm.set_stream(0)
rm.set_sensor_model(wl.sensor_table(501))
for N in (1 << 18, 1 << 22, 1 << 24, 1 << 26):
    q = torch.from_numpy(wl.random_queries(W, H, N, seed=1)).cuda()
    out = torch.empty(N, dtype=torch.float32, device="cuda")
    for pv in (0, 1):
        rm.set_persistent(pv)
        for coop in (0, 12):
            rm.set_coop_threshold(coop)
            med, mn = timeit(lambda: rm.calc_range_many_grid(q, out))
            print("RM random N=%9d persist=%d coop=%d  %8.3f ms %7.2f G rays/s" % (N, pv, coop, med, N / med / 1e6))
    del q, out
rm.set_persistent(1)
dt = rm.distance_transform()
for (n, M) in ((4000, 60), (100000, 60), (20000, 1080)):
    for tag in ("uniform", "tracking"):
        ph = wl.pf_particles_uniform(occ, n, seed=3) if tag == "uniform" else wl.pf_particles_tracking(occ, n, seed=3, dt=dt)[0]
        parts = torch.from_numpy(ph).cuda()
        angles = torch.from_numpy(wl.lidar_angles(M)).cuda()
        obs = torch.from_numpy(np.linspace(5, 450, M).astype(np.float32)).cuda()
        w = torch.empty(n, dtype=torch.float64, device="cuda")
        rng = torch.empty(n * M, dtype=torch.float32, device="cuda")
        for coop in (0, 2, 4, 6, 8, 12, 16, 24, 32):
            rm.set_coop_threshold(coop)
            med, mn = timeit(lambda: rm.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs, w), iters=20)
            med2, mn2 = timeit(lambda: rm.calc_range_repeat_angles(parts, angles, rng), iters=20)
            print("PF %7dx%4d %-8s coop=%2d fused %8.3f ms (min %.3f, %6.2f G rays/s) | angles %8.3f ms (%6.2f G rays/s)" % (
                n, M, tag, coop, med, mn, n * M / med / 1e6, med2, n * M / med2 / 1e6))
