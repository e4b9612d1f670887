import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import range_libc_b200 as rl
from range_libc_b200 import workloads as wl
from tools.quick_bench import timeit
occ = wl.load_map("basement_hallways_5cm")
m = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
t = time.time()
g = rl.PyGiantLUTCast(m, 500.0, 108)
print("GLT build 1200^2 x 108 (155.5 M casts): %.3f s, %.1f MB" % (time.time() - t, g.memory() / 1e6))
g.set_stream(0)
N = 1 << 24
q = torch.from_numpy(wl.random_queries(1200, 1200, N, seed=1)).cuda()
out = torch.empty(N, dtype=torch.float32, device="cuda")
med, mn = timeit(lambda: g.calc_range_many_grid(q, out))
print("GLT random N=2^24: %.3f ms %.2f G rays/s" % (med, N / med / 1e6))
