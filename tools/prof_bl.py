"""ncu driver: a few launches of the BL random-query kernel only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402

occ = wl.load_map("basement_hallways_5cm")
W, H = occ.shape
N = 1 << 22
omap = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
q = torch.from_numpy(wl.random_queries(W, H, N, seed=1)).cuda()
out = torch.empty(N, dtype=torch.float32, device="cuda")
m = rl.PyBresenhamsLine(omap, 500.0)
m.set_stream(0)
for _ in range(3):
    m.calc_range_many_grid(q, out)
torch.cuda.synchronize()
