"""A/B of the DT layout (RL_DT_LINEAR env) on the three RM regimes (development aid)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402
from tools.quick_bench import timeit  # noqa: E402

tag = "linear" if os.environ.get("RL_DT_LINEAR") == "1" else "tiled"
occ = wl.load_map("basement_hallways_5cm")
W, H = occ.shape
omap = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
rm = rl.PyRayMarchingGPU(omap, 500.0)
rm.set_stream(0)
rm.set_sensor_model(wl.sensor_table(501))
for N in (1 << 22, 1 << 24):
    q = torch.from_numpy(wl.random_queries(W, H, N, seed=1)).cuda()
    out = torch.empty(N, dtype=torch.float32, device="cuda")
    med, mn = timeit(lambda: rm.calc_range_many_grid(q, out))
    print("%s RM random N=%9d %8.3f ms %7.2f G rays/s" % (tag, N, med, N / med / 1e6))
    del q, out
for (n, M) in ((100000, 60), (20000, 1080)):
    parts = torch.from_numpy(wl.pf_particles_uniform(occ, n, seed=3)).cuda()
    angles = torch.from_numpy(wl.lidar_angles(M)).cuda()
    obs = torch.from_numpy(np.linspace(5, 450, M).astype(np.float32)).cuda()
    w = torch.empty(n, dtype=torch.float64, device="cuda")
    rng = torch.empty(n * M, dtype=torch.float32, device="cuda")
    med, mn = timeit(lambda: rm.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs, w), iters=10)
    med2, mn2 = timeit(lambda: rm.calc_range_repeat_angles(parts, angles, rng), iters=10)
    print("%s PF %7dx%4d fused %8.3f ms (%6.2f G rays/s) | angles %8.3f ms (%6.2f G rays/s)" % (tag, n, M, med, n * M / med / 1e6, med2, n * M / med2 / 1e6))
occ5 = wl.synthetic_map(4096, seed=2026)
m5 = rl.PyOMap(np.ascontiguousarray(occ5.T.astype(bool)))
rm5 = rl.PyRayMarchingGPU(m5, 500.0)
rm5.set_stream(0)
rm5.set_sensor_model(wl.sensor_table(501))
n5, mb = 100000, 1080
p5 = torch.from_numpy(wl.pf_particles_uniform(occ5, n5, seed=4)).cuda()
a5 = torch.from_numpy(wl.lidar_angles(mb)).cuda()
o5 = torch.from_numpy(np.clip(150 + 100 * np.sin(np.linspace(0, 6, mb)), 0, 500).astype(np.float32)).cuda()
w5 = torch.empty(n5, dtype=torch.float64, device="cuda")
med, mn = timeit(lambda: rm5.calc_range_repeat_angles_eval_sensor_model(p5, a5, o5, w5), iters=3, reps=2)
print("%s 4096^2 fused 100k x 1080: %8.3f ms %6.2f G rays/s" % (tag, med, n5 * mb / med / 1e6))
q5 = torch.from_numpy(wl.random_queries(4096, 4096, 1 << 24, seed=1)).cuda()
out5 = torch.empty(1 << 24, dtype=torch.float32, device="cuda")
med, mn = timeit(lambda: rm5.calc_range_many_grid(q5, out5), iters=5, reps=3)
print("%s 4096^2 RM random 2^24: %8.3f ms %6.2f G rays/s" % (tag, med, (1 << 24) / med / 1e6))
