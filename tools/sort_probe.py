"""Spatial ordering of large fused updates (rl_sort.cu): 8192^2 synthetic map, uniform global-localisation cloud.
Run twice: RL_SPATIAL_SORT=0 and =1 (the switch is read once per process).  Prints time per update and a digest of
the weights, which must not depend on the switch."""
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

MB = int(sys.argv[1]) if len(sys.argv) > 1 else 360
for size, cases in ((8192, ((1000000, 60), (200000, MB))), (1200, ((100000, 60), (20000, MB)))):
  occ5 = wl.synthetic_map(size, seed=2026) if size != 1200 else wl.load_map("basement_hallways_5cm")
  rm5 = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ5.T.astype(bool))), 500.0)
  rm5.set_stream(0)
  rm5.set_sensor_model(wl.sensor_table(501))
  for n5, mb in cases:
      p5 = torch.from_numpy(wl.pf_particles_uniform(occ5, n5, seed=4)).cuda()
      a5 = torch.from_numpy(wl.lidar_angles(mb)).cuda()
      o5 = torch.from_numpy(np.clip(150 + 100 * np.sin(np.linspace(0, 6, mb)), 0, 500).astype(np.float32)).cuda()
      w5 = torch.empty(n5, dtype=torch.float64, device="cuda")
      med, mn = timeit(lambda: rm5.calc_range_repeat_angles_eval_sensor_model(p5, a5, o5, w5), iters=3, reps=2)
      torch.cuda.synchronize()
      print("RL_SPATIAL_SORT=%s RL_FUSED_PERSIST=%s %d^2 fused %dx%d: %8.3f ms %6.2f G rays/s  sha=%s" % (
          os.environ.get("RL_SPATIAL_SORT", "1"), os.environ.get("RL_FUSED_PERSIST", "1") + " G=" + os.environ.get("RL_FUSED_GROUP_RAYS", "dflt"), size, n5, mb, med, n5 * mb / med / 1e6,
          hashlib.sha256(w5.cpu().numpy().tobytes()).hexdigest()[:16]), flush=True)
