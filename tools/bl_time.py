import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import range_libc_b200 as rl
from range_libc_b200 import workloads as wl
from tools.quick_bench import timeit
occ = wl.load_map("basement_hallways_5cm")
omap = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
bl = rl.PyBresenhamsLine(omap, 500.0)
bl.set_stream(0)
N = 1 << 22
q = torch.from_numpy(wl.random_queries(1200, 1200, N, seed=1)).cuda()
out = torch.empty(N, dtype=torch.float32, device="cuda")
med, mn = timeit(lambda: bl.calc_range_many_grid(q, out), iters=5, reps=5)
print("BL refill=%s burst=%s: %.3f ms %.2f G rays/s" % (os.environ.get("RL_BL_REFILL"), os.environ.get("RL_BL_BURST"), med, N / med / 1e6))
