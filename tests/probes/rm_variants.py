"""A/B of the RM large-batch kernels.  The variant is chosen by the environment (read once per process):
    RL_RM_PERSIST=1 default (parked rays in registers) | 2 conditional load | 3 parked rays in shared memory | 4 = 3 with two rays per lane
    RL_RM_BURST_PAIRS=k   steps per refill round = 2k (variants 2, 3)
Prints the rate on 2^24 uniformly random rays (basement_hallways_5cm) and a digest of the ranges, which must be the
same for every variant (and is checked against the oracle on a prefix)."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from oracle import port  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402


def main():
    n = 1 << 24
    occ = wl.load_map("basement_hallways_5cm")
    q = wl.random_queries(occ.shape[0], occ.shape[1], n, seed=12345)
    rm = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), 500.0)
    dq = torch.from_numpy(q).cuda()
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream()
    rm.set_stream(st.cuda_stream)
    for _ in range(3):
        rm.calc_range_many_grid(dq, out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        rm.calc_range_many_grid(dq, out)
        b.record(st)
        b.synchronize()
        ts.append(a.elapsed_time(b))
    h = out.cpu().numpy()
    m = 200000
    want = port.Oracle(port.RM, occ, 500.0, threads=os.cpu_count()).calc_range_many(q[:m])
    ok = np.array_equal(h[:m].view(np.uint32), want.view(np.uint32))
    ms = float(np.median(ts))
    print("persist=%s burst_pairs=%s  %.3f ms  %.2f G rays/s  oracle-prefix-equal=%s  sha=%s" % (
        os.environ.get("RL_RM_PERSIST", "default"), os.environ.get("RL_RM_BURST_PAIRS", "default"), ms, n / ms / 1e6, ok,
        hashlib.sha256(h.tobytes()).hexdigest()[:16]), flush=True)


if __name__ == "__main__":
    main()
