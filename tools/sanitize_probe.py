"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402

occ = wl.load_map("basement_hallways_10cm")[150:406, 180:420].copy()
W, H = occ.shape
omap = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
q = wl.random_queries(W, H, 3000, seed=1)
parts = wl.pf_particles_uniform(occ, 40, seed=2)
angles = wl.lidar_angles(60)
obs = np.linspace(3, 200, 60).astype(np.float32)
table = wl.sensor_table(257)
out = np.empty(len(q), np.float32)
w = np.empty(len(parts), np.float64)
rng = np.empty(len(parts) * len(angles), np.float32)
for ctor in (lambda: rl.PyRayMarchingGPU(omap, 256.0), lambda: rl.PyBresenhamsLine(omap, 256.0),
             lambda: rl.PyCDDTCast(omap, 256.0, 24), lambda: rl.PyGiantLUTCast(omap, 256.0, 8)):
    m = ctor()
    if isinstance(m, rl.PyCDDTCast):
        m.prune()
    m.set_sensor_model(table)
    m.calc_range_many_grid(q, out)
    m.calc_range_many(q, out)
    m.calc_range_repeat_angles(parts, angles, rng)
    m.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs, w)
    m.eval_sensor_model(obs, rng, w, len(angles), len(parts))
    m.update_map_batch(np.ones(64, np.uint8), np.array([[8, 8, 8, 8]], np.int32))
    m.calc_range_many_grid(q, out)
    print(type(m).__name__, "ok", float(out.mean()))
