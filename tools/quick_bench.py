"""Quick device-resident timings of every kernel family (development aid, not the judged bench)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402


def timeit(fn, iters=10, warm=3, reps=20):
    """Median / min over `iters` batches of `reps` back-to-back asynchronous launches (per-launch ms):
    the launches queue up, so host-side call overhead overlaps with the previous kernel."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    return float(np.median(ts)), float(np.min(ts))


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "basement_hallways_5cm"
    occ = wl.load_map(name)
    W, H = occ.shape
    omap = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
    stream = 0
    t = time.time()
    rm = rl.PyRayMarchingGPU(omap, 500.0)
    t_rm = time.time() - t
    t = time.time()
    bl = rl.PyBresenhamsLine(omap, 500.0)
    t_bl = time.time() - t
    t = time.time()
    cd = rl.PyCDDTCast(omap, 500.0, 108)
    t_cd = time.time() - t
    t = time.time()
    pc = rl.PyCDDTCast(omap, 500.0, 108)
    pc.prune()
    t_pc = time.time() - t
    print("build s: rm %.3f bl %.3f cddt %.3f pcddt %.3f" % (t_rm, t_bl, t_cd, t_pc))
    for m in (rm, bl, cd, pc):
        m.set_stream(stream)
        m.set_sensor_model(wl.sensor_table(501))
    for N in (1 << 18, 1 << 22, 1 << 24):
        q = torch.from_numpy(wl.random_queries(W, H, N, seed=1)).cuda()
        out = torch.empty(N, dtype=torch.float32, device="cuda")
        for nm, m in (("rm", rm), ("bl", bl), ("cddt", cd), ("pcddt", pc)):
            if nm == "bl" and N > (1 << 22):
                continue
            med, mn = timeit(lambda: m.calc_range_many_grid(q, out))
            print("random N=%9d %-6s %9.3f ms  %8.2f G rays/s" % (N, nm, med, N / med / 1e6))
    for (n, M) in ((4000, 60), (100000, 60), (20000, 1080)):
        parts_h = wl.pf_particles_uniform(occ, n, seed=3)
        tr_h, _ = wl.pf_particles_tracking(occ, n, seed=3, dt=rm.distance_transform())
        for tag, ph in (("uniform", parts_h), ("tracking", tr_h)):
            parts = torch.from_numpy(ph).cuda()
            angles = torch.from_numpy(wl.lidar_angles(M)).cuda()
            obs = torch.from_numpy(np.linspace(5, 450, M).astype(np.float32)).cuda()
            w = torch.empty(n, dtype=torch.float64, device="cuda")
            rng = torch.empty(n * M, dtype=torch.float32, device="cuda")
            for nm, m in (("rm", rm), ("cddt", cd), ("bl", bl)):
                if nm == "bl" and n * M > 10_000_000:
                    continue
                med, mn = timeit(lambda: m.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs, w))
                med2, _ = timeit(lambda: m.calc_range_repeat_angles(parts, angles, rng))
                med3, _ = timeit(lambda: m.eval_sensor_model(obs, rng, w, M, n))
                print("PF %7dx%4d %-8s %-5s fused %8.3f ms (%7.2f G rays/s, min %.3f) | angles %8.3f ms | eval %7.3f ms" %
                      (n, M, tag, nm, med, n * M / med / 1e6, mn, med2, med3))


if __name__ == "__main__":
    main()
