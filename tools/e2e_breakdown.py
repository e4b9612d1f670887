"""Host-side time of the blocking 4000 x 60 host-pointer update, phase by phase (library built with -DRL_HOST_TIMING:
python tools/build_variant.py htiming -DRL_HOST_TIMING; RL_B200_LIB=tools/_trace/librangelib_b200_htiming.so python tools/e2e_breakdown.py)"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import cabi, workloads as wl  # noqa: E402

occ = wl.load_map(bench.MAP)
sets, angles, obs = bench.make_inputs(occ, 64)
rm = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), bench.MAX_RANGE)
rm.set_sensor_model(wl.sensor_table(bench.K_TABLE))
pin = lambda a: torch.from_numpy(a.copy()).pin_memory().numpy()
hs = [pin(sets[i]) for i in range(64)]
ha, ho, hw = pin(angles), pin(obs), pin(np.zeros(bench.N_PART, np.float64))
lib = cabi.lib()
out = (C.c_double * 8)()
cnt = C.c_long()
for label, n_part in (("4000 particles", bench.N_PART), ("8 particles", 8)):
    for i in range(200):
        rm.calc_range_repeat_angles_eval_sensor_model(hs[i % 64][:n_part], ha, ho, hw[:n_part])
    lib.rl_debug_host_timing(out, C.byref(cnt))
    n = 3000
    t0 = time.perf_counter()
    for i in range(n):
        rm.calc_range_repeat_angles_eval_sensor_model(hs[i % 64][:n_part], ha, ho, hw[:n_part])
    dt = (time.perf_counter() - t0) / n * 1e6
    lib.rl_debug_host_timing(out, C.byref(cnt))
    names = ["pointer queries", "staging", "H2D enqueue", "kernel launch", "synchronise", "copy out"]
    print("%s: %.2f us per call (python loop); inside the C call: %s" % (
        label, dt, ", ".join("%s %.2f" % (names[i], out[i] / cnt.value) for i in range(6))), flush=True)
