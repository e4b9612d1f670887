"""Deep many-beam fused launches (BASELINE config 5 shape): CTA size sweep through RL_FUSED_DEEP_THREADS (one process
per value: the knob is read once).  python tools/c5_threads.py [threads ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, hashlib
import numpy as np, torch
sys.path.insert(0, %r)
import range_libc_b200 as rl
from range_libc_b200 import workloads as wl
from tools.quick_bench import timeit
for size, n5, mb in ((8192, 200000, 1080), (1200, 20000, 1080)):
    occ5 = wl.load_map("basement_hallways_5cm") if size == 1200 else wl.synthetic_map(size, seed=2026)
    rm5 = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ5.T.astype(bool))), 500.0)
    rm5.set_stream(0)
    rm5.set_sensor_model(wl.sensor_table(501))
    p5 = torch.from_numpy(wl.pf_particles_uniform(occ5, n5, seed=4)).cuda()
    a5 = torch.from_numpy(wl.lidar_angles(mb)).cuda()
    o5 = torch.from_numpy(np.clip(150 + 100 * np.sin(np.linspace(0, 6, mb)), 0, 500).astype(np.float32)).cuda()
    w5 = torch.empty(n5, dtype=torch.float64, device="cuda")
    med, mn = timeit(lambda: rm5.calc_range_repeat_angles_eval_sensor_model(p5, a5, o5, w5), iters=3, reps=3)
    print("threads=%%s %%d^2 fused %%dx%%d: %%8.3f ms %%6.2f G rays/s  sha %%s" %% (os.environ.get("RL_FUSED_DEEP_THREADS", "256"), size, n5, mb, med, n5 * mb / med / 1e6, hashlib.sha256(w5.cpu().numpy().tobytes()).hexdigest()[:12]), flush=True)
    del rm5, p5, w5
''' % ROOT
for th in (sys.argv[1:] or ["256", "216", "224", "192", "184"]):
    env = dict(os.environ, RL_FUSED_DEEP_THREADS=th)
    subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
