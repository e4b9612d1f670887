"""Per-step latency of the RM kernels on single long rays (development aid)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402
from oracle import port  # noqa: E402
from tools.quick_bench import timeit  # noqa: E402

occ = wl.load_map("basement_hallways_5cm")
W, H = occ.shape
o = port.Oracle(port.RM, occ, 500.0)
q = wl.random_queries(W, H, 400000, seed=9)
steps = o.rm_step_counts(q)
order = np.argsort(-steps)
omap = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
rm = rl.PyRayMarchingGPU(omap, 500.0)
rm.set_stream(0)
for idx in list(order[:3]) + [int(np.argmin(np.abs(steps - 60))), int(np.argmin(np.abs(steps - 20)))]:
    ray = q[idx:idx + 1].copy()
    for nrep in (1, 32):
        rays = torch.from_numpy(np.repeat(ray, nrep, axis=0)).cuda()
        out = torch.empty(nrep, dtype=torch.float32, device="cuda")
        for coop in (0, 3):
            rm.set_coop_threshold(coop)
            med, mn = timeit(lambda: rm.calc_range_many_grid(rays, out), iters=5, reps=10)
            print("ray %s steps %3d x%2d coop=%d: %7.2f us  -> %6.1f ns/step" % (
                np.round(ray[0], 2), steps[idx], nrep, coop, mn * 1e3, mn * 1e6 / steps[idx]))
