"""Config-5 shape on the 8192^2 map: the fused update against cast-to-memory + eval_sensor_model (two kernels), with the
particles in caller order and pre-ordered by 64x64-cell tile (what rl_sort.cu's processing order would give)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402
from tools.quick_bench import timeit  # noqa: E402

size, n, mb = 8192, 200000, 1080
occ = wl.synthetic_map(size, seed=2026)
rm = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), 500.0)
rm.set_stream(0)
rm.set_sensor_model(wl.sensor_table(501))
ph = wl.pf_particles_uniform(occ, n, seed=4)
# tile order (Morton of the 64x64 tile of the grid cell: grid x = world y, grid y = world x)
def morton(a, b):
    def spread(v):
        v = v.astype(np.uint64)
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v
    return spread(a) | (spread(b) << 1)
key = morton((ph[:, 1] // 64).astype(np.int64), (ph[:, 0] // 64).astype(np.int64))
ps = ph[np.argsort(key, kind="stable")]
a = torch.from_numpy(wl.lidar_angles(mb)).cuda()
o = torch.from_numpy(np.clip(150 + 100 * np.sin(np.linspace(0, 6, mb)), 0, 500).astype(np.float32)).cuda()
w = torch.empty(n, dtype=torch.float64, device="cuda")
r = torch.empty(n * mb, dtype=torch.float32, device="cuda")
for label, parts in (("caller order", ph), ("tile order  ", ps)):
    p = torch.from_numpy(np.ascontiguousarray(parts)).cuda()
    med, _ = timeit(lambda: rm.calc_range_repeat_angles_eval_sensor_model(p, a, o, w), iters=3, reps=3)
    print("%s fused                 %8.3f ms %6.2f G rays/s" % (label, med, n * mb / med / 1e6), flush=True)
    med1, _ = timeit(lambda: rm.calc_range_repeat_angles(p, a, r), iters=3, reps=3)
    print("%s calc_range_repeat_angles %8.3f ms %6.2f G rays/s" % (label, med1, n * mb / med1 / 1e6), flush=True)
    med2, _ = timeit(lambda: rm.eval_sensor_model(o, r, w, mb, n), iters=3, reps=3)
    print("%s eval_sensor_model     %8.3f ms %6.2f G evals/s -> two-step total %8.3f ms %6.2f G rays/s" % (
        label, med2, n * mb / med2 / 1e6, med1 + med2, n * mb / (med1 + med2) / 1e6), flush=True)
