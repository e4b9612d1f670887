"""Small driver for ncu: a few launches of the RM/CDDT/BL random-query kernels (device resident)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "basement_hallways_5cm"
N = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 22)
occ = wl.load_map(name)
W, H = occ.shape
omap = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
q = torch.from_numpy(wl.random_queries(W, H, N, seed=1)).cuda()
out = torch.empty(N, dtype=torch.float32, device="cuda")
for ctor in (lambda: rl.PyRayMarchingGPU(omap, 500.0), lambda: rl.PyCDDTCast(omap, 500.0, 108),
             lambda: rl.PyBresenhamsLine(omap, 500.0)):
    m = ctor()
    m.set_stream(0)
    for _ in range(3):
        m.calc_range_many_grid(q, out)
    torch.cuda.synchronize()
