"""Large fused launches: cooperative hand-off on/off (development aid)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402
from tools.quick_bench import timeit  # noqa: E402

for size, n5, mb in ((1200, 100000, 60), (1200, 20000, 1080), (8192, 200000, 1080)):
    occ5 = wl.load_map("basement_hallways_5cm") if size == 1200 else wl.synthetic_map(size, seed=2026)
    m5 = rl.PyOMap(np.ascontiguousarray(occ5.T.astype(bool)))
    rm5 = rl.PyRayMarchingGPU(m5, 500.0)
    rm5.set_stream(0)
    rm5.set_sensor_model(wl.sensor_table(501))
    p5 = torch.from_numpy(wl.pf_particles_uniform(occ5, n5, seed=4)).cuda()
    a5 = torch.from_numpy(wl.lidar_angles(mb)).cuda()
    o5 = torch.from_numpy(np.clip(150 + 100 * np.sin(np.linspace(0, 6, mb)), 0, 500).astype(np.float32)).cuda()
    w5 = torch.empty(n5, dtype=torch.float64, device="cuda")
    r5 = torch.empty(n5 * mb, dtype=torch.float32, device="cuda")
    for coop in (0, 16):
        rm5.set_coop_threshold(coop)
        med, mn = timeit(lambda: rm5.calc_range_repeat_angles_eval_sensor_model(p5, a5, o5, w5), iters=3, reps=2)
        print("%d^2 fused %dx%d coop=%2d: %8.3f ms %6.2f G rays/s" % (size, n5, mb, coop, med, n5 * mb / med / 1e6))
    for pers in (0, 1):
        rm5.set_persistent(pers)
        med, mn = timeit(lambda: rm5.calc_range_repeat_angles(p5, a5, r5), iters=3, reps=2)
        print("%d^2 angles %dx%d persist=%d: %8.3f ms %6.2f G rays/s" % (size, n5, mb, pers, med, n5 * mb / med / 1e6))
    del rm5, p5, w5, r5
