"""Fixed cost of the fused PF kernel: same launch with max_range so small that every ray ends after one step."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402

occ = wl.load_map("basement_hallways_5cm")
omap = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
n, M = 4000, 60
parts = torch.from_numpy(wl.pf_particles_uniform(occ, n, seed=3)).cuda()
if len(sys.argv) > 1 and sys.argv[1] == "tracking":
    parts = torch.from_numpy(wl.pf_particles_tracking(occ, n, seed=3)[0]).cuda()
angles = torch.from_numpy(wl.lidar_angles(M)).cuda()
obs = torch.from_numpy(np.linspace(5, 450, M).astype(np.float32)).cuda()
w = torch.empty(n, dtype=torch.float64, device="cuda")
stream = torch.cuda.current_stream()
for mr in (0.5, 500.0):
    rm = rl.PyRayMarchingGPU(omap, mr)
    rm.set_sensor_model(wl.sensor_table(501))
    for coop in (0, 8, 16, 32):
        rm.set_coop_threshold(coop)
        side = torch.cuda.Stream()
        side.wait_stream(stream)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
            rm.set_stream(torch.cuda.current_stream().cuda_stream)
            for _ in range(200):
                rm.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs, w)
        rm.set_stream(stream.cuda_stream)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        print("max_range %6.1f coop=%d: %.2f us per fused 4000x60 launch (graph of 200)" % (mr, coop, e0.elapsed_time(e1) * 1e3 / 200))
