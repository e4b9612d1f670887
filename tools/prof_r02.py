"""ncu driver, round 2: runs ONE named workload a few times so that `ncu -k regex:<kernel>` can capture it.

    python tools/prof_r02.py <workload> [reps]

workloads
  fan_tracking   rm_persist_kernel<MODE_ANGLES>: 262144 poses x 60 beams, tracking cloud, 5 cm map
  fan_uniform    the same, poses uniform over free cells
  rm_random      rm_persist_kernel<MODE_GRID>: 2^24 uniformly random rays, 5 cm map
  fused_deep     fused_rm_persist_kernel: 100000 x 60 fused sensor update, 5 cm map
  c2             fused_kernel<RM>: the judged 4000 x 60 update (alternating clouds), warm
  c3_cddt        cast_kernel<CDDT, GRID>: 2^24 random rays on gigantic_map (722 MB table)
  c3_pcddt       the same after prune()
  bl             bl_persist_kernel: 2^22 random rays, 5 cm map
  c4_bl          bl_persist_kernel on the synthetic 4096^2 grid, 2^20 rays
  c5             fused M=1080 path on the synthetic 8192^2 grid, 200000 particles (spatial order on)
  edt_1200 / edt_8192   the distance-transform build
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402

which = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
MAXR = 500.0


def omap_of(occ):
    return rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))


def run(fn):
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()


if which in ("fan_tracking", "fan_uniform", "rm_random", "fused_deep", "c2", "bl", "edt_1200"):
    occ = wl.load_map("basement_hallways_5cm")
    omap = omap_of(occ)
    if which == "edt_1200":
        for _ in range(reps):
            m = rl.PyRayMarchingGPU(omap, MAXR)
            del m
        torch.cuda.synchronize()
        sys.exit(0)
    if which == "bl":
        m = rl.PyBresenhamsLine(omap, MAXR)
        m.set_stream(0)
        n = 1 << 22
        q = torch.from_numpy(wl.random_queries(occ.shape[0], occ.shape[1], n, seed=1)).to(dev)
        out = torch.empty(n, dtype=torch.float32, device=dev)
        run(lambda: m.calc_range_many_grid(q, out))
        sys.exit(0)
    rm = rl.PyRayMarchingGPU(omap, MAXR)
    rm.set_stream(0)
    if which == "rm_random":
        n = 1 << 24
        q = torch.from_numpy(wl.random_queries(occ.shape[0], occ.shape[1], n, seed=1)).to(dev)
        out = torch.empty(n, dtype=torch.float32, device=dev)
        run(lambda: rm.calc_range_many_grid(q, out))
    elif which in ("fan_tracking", "fan_uniform"):
        n_p, n_b = 262144, 60
        if which == "fan_tracking":
            parts = wl.pf_particles_tracking(occ, n_p, seed=11, dt=rm.distance_transform())[0]
        else:
            parts = wl.pf_particles_uniform(occ, n_p, seed=11)
        parts = torch.from_numpy(parts).to(dev)
        ang = torch.from_numpy(wl.lidar_angles(n_b)).to(dev)
        out = torch.empty(n_p * n_b, dtype=torch.float32, device=dev)
        run(lambda: rm.calc_range_repeat_angles(parts, ang, out))
    else:
        rm.set_sensor_model(wl.sensor_table(501))
        n_b = 60
        ang = torch.from_numpy(wl.lidar_angles(n_b)).to(dev)
        obs = torch.from_numpy(np.clip(120.0 + 80.0 * np.sin(np.linspace(0, 3.0, n_b)), 0, MAXR).astype(np.float32)).to(dev)
        if which == "fused_deep":
            n_p = 100000
            parts = torch.from_numpy(wl.pf_particles_uniform(occ, n_p, seed=11)).to(dev)
            w = torch.empty(n_p, dtype=torch.float64, device=dev)
            run(lambda: rm.calc_range_repeat_angles_eval_sensor_model(parts, ang, obs, w))
        else:
            sys.path.insert(0, ROOT)
            import bench
            sets_h, _, _ = bench.make_inputs(occ, 8)
            sets = torch.from_numpy(sets_h).to(dev)
            w = torch.empty(4000, dtype=torch.float64, device=dev)
            for i in range(8 * reps):
                rm.calc_range_repeat_angles_eval_sensor_model(sets[i % 8], ang, obs, w)
            torch.cuda.synchronize()
elif which in ("c3_cddt", "c3_pcddt"):
    big = wl.load_map("gigantic_map")
    cd = rl.PyCDDTCast(omap_of(big), MAXR, 108)
    if which == "c3_pcddt":
        cd.prune()
    cd.set_stream(0)
    n = 1 << 24
    q = torch.from_numpy(wl.random_queries(big.shape[0], big.shape[1], n, seed=2)).to(dev)
    out = torch.empty(n, dtype=torch.float32, device=dev)
    run(lambda: cd.calc_range_many_grid(q, out))
elif which == "c4_bl":
    occ4 = wl.synthetic_map(4096, seed=2026)
    bl = rl.PyBresenhamsLine(omap_of(occ4), MAXR)
    bl.set_stream(0)
    n = 1 << 20
    q = torch.from_numpy(wl.random_queries(4096, 4096, n, seed=3)).to(dev)
    out = torch.empty(n, dtype=torch.float32, device=dev)
    run(lambda: bl.calc_range_many_grid(q, out))
elif which in ("c5", "edt_8192"):
    occ5 = wl.synthetic_map(8192, seed=2026)
    m5 = omap_of(occ5)
    if which == "edt_8192":
        for _ in range(reps):
            m = rl.PyRayMarchingGPU(m5, MAXR)
            del m
        torch.cuda.synchronize()
        sys.exit(0)
    rm5 = rl.PyRayMarchingGPU(m5, MAXR)
    rm5.set_sensor_model(wl.sensor_table(501))
    rm5.set_stream(0)
    n5, mb = 200000, 1080
    p5 = torch.from_numpy(wl.pf_particles_uniform(occ5, n5, seed=4)).to(dev)
    a5 = torch.from_numpy(wl.lidar_angles(mb)).to(dev)
    o5 = torch.from_numpy(np.clip(150 + 100 * np.sin(np.linspace(0, 6, mb)), 0, 500).astype(np.float32)).to(dev)
    w5 = torch.empty(n5, dtype=torch.float64, device=dev)
    run(lambda: rm5.calc_range_repeat_angles_eval_sensor_model(p5, a5, o5, w5))
else:
    raise SystemExit("unknown workload " + which)
