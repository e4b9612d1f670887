"""Deep fused updates of every kind (5 cm map, device resident): time + digest, for A/B of RL_FUSED_OVERLAP (read once
per process).  python tools/fused_kinds.py"""
import hashlib
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
om = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
table = wl.sensor_table(501)
tag = "overlap=%s" % os.environ.get("RL_FUSED_OVERLAP", "1")
for name, mk in (("RM", lambda: rl.PyRayMarchingGPU(om, 500.0)), ("CDDT", lambda: rl.PyCDDTCast(om, 500.0, 108)),
                 ("GLT", lambda: rl.PyGiantLUTCast(om, 500.0, 108)), ("BL", lambda: rl.PyBresenhamsLine(om, 500.0))):
    m = mk()
    m.set_stream(0)
    m.set_sensor_model(table)
    for n, mb in ((100000, 60), (20000, 1080), (50000, 360)) if name != "BL" else ((20000, 60), (2000, 1080)):
        p = torch.from_numpy(wl.pf_particles_uniform(occ, n, seed=4)).cuda()
        a = torch.from_numpy(wl.lidar_angles(mb)).cuda()
        o = torch.from_numpy(np.clip(150 + 100 * np.sin(np.linspace(0, 6, mb)), 0, 500).astype(np.float32)).cuda()
        w = torch.empty(n, dtype=torch.float64, device="cuda")
        med, mn = timeit(lambda: m.calc_range_repeat_angles_eval_sensor_model(p, a, o, w), iters=3, reps=3)
        print("%s %-4s fused %6dx%-4d %8.3f ms %7.2f G rays/s  sha %s" % (
            tag, name, n, mb, med, n * mb / med / 1e6, hashlib.sha256(w.cpu().numpy().tobytes()).hexdigest()[:12]), flush=True)
    del m
