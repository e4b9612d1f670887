"""Sweep of the small fused launch (4000 x 60, bench.py's particle sets): cooperative hand-off threshold (runtime)
for the RL_BLOCK_BURST_PAIRS of this process.  Prints the mean launch time over alternating global / tracking sets."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402

occ = wl.load_map(bench.MAP)
sets_h, angles_h, obs_h = bench.make_inputs(occ, 128)
rm = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), bench.MAX_RANGE)
rm.set_sensor_model(wl.sensor_table(bench.K_TABLE))
st = torch.cuda.current_stream()
rm.set_stream(st.cuda_stream)
sets = torch.from_numpy(sets_h).cuda()
angles, obs = torch.from_numpy(angles_h).cuda(), torch.from_numpy(obs_h).cuda()
w = torch.empty(bench.N_PART, dtype=torch.float64, device="cuda")
for coop in [int(x) for x in (sys.argv[1:] or ["8", "12", "16", "24", "32"])]:
    rm.set_coop_threshold(coop)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(st)
    with torch.cuda.graph(g, stream=side):
        rm.set_stream(torch.cuda.current_stream().cuda_stream)
        for i in range(512):
            rm.calc_range_repeat_angles_eval_sensor_model(sets[i % 128], angles, obs, w)
    rm.set_stream(st.cuda_stream)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    g.replay()
    b.record(st)
    b.synchronize()
    print("burst_pairs=%s coop=%2d  %.2f us per launch" % (os.environ.get("RL_BLOCK_BURST_PAIRS", "dflt"), coop,
                                                          a.elapsed_time(b) / 512 * 1e3), flush=True)
